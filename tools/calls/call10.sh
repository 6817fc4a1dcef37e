mkdir -p gpurun_out
timeout 600 python tools/gpu_check_attention.py > gpurun_out/r2_check_attn.log 2>&1; echo "check rc=$?" >> gpurun_out/r2_check_attn.log
MMDP_ATTN_VERSION=7 timeout 1000 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest10_v7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest10_v7.log
MMDP_ATTN_VERSION=7 timeout 600 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2_bench_v7.json 2> gpurun_out/r2_bench_v7.err; echo "bench rc=$?" >> gpurun_out/r2_bench_v7.err
MMDP_ATTN_VERSION=6 timeout 600 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2_bench_v6.json 2> gpurun_out/r2_bench_v6.err; echo "bench rc=$?" >> gpurun_out/r2_bench_v6.err
cut -c1-420 gpurun_out/r2_check_attn.log | tail -22; tail -5 gpurun_out/r2_pytest10_v7.log | cut -c1-300; cut -c1-200 gpurun_out/r2_bench_v7.json; cut -c1-200 gpurun_out/r2_bench_v6.json
