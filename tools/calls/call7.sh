mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -s > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest7.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 2 --warmup 1 --tp --no-extras > gpurun_out/r2_bench_tp2b.json 2> gpurun_out/r2_bench_tp2b.err; echo "rc=$?" >> gpurun_out/r2_bench_tp2b.err
tail -4 gpurun_out/r2_pytest7.log; tail -2 gpurun_out/r2_bench_tp2b.err; cut -c1-300 gpurun_out/r2_bench_tp2b.json
