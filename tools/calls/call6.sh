mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_smi6.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "rc=$?" >> gpurun_out/r2_bench_n8.err
tail -3 gpurun_out/r2_bench_n8.err; cut -c1-400 gpurun_out/r2_bench_n8.json
