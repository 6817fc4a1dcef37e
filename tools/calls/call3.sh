mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_smi3.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -s > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "rc=$?" >> gpurun_out/r2_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 2 --warmup 1 --tp --no-extras > gpurun_out/r2_bench_tp2.json 2> gpurun_out/r2_bench_tp2.err; echo "rc=$?" >> gpurun_out/r2_bench_tp2.err
tail -6 gpurun_out/r2_pytest3.log; tail -3 gpurun_out/r2_bench_n2.err; cut -c1-300 gpurun_out/r2_bench_n2.json; tail -3 gpurun_out/r2_bench_tp2.err; cut -c1-300 gpurun_out/r2_bench_tp2.json
