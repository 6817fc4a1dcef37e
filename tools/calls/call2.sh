mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -s > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
timeout 900 python tools/gpu_sweep_forward.py --variants r1 cur cur_nopdl cur_sk1 cur_noattn cur_pair r1_pair > gpurun_out/r2_sweep2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_sweep2.log
timeout 300 python tools/gpu_check_kernels.py --only pair_resid,pair_ffout,pair_swiglu,pair_plain,pair_qkv > gpurun_out/r2_kernels2.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench rc=$?" >> gpurun_out/r2_bench_a.err
tail -5 gpurun_out/r2_pytest2.log; tail -8 gpurun_out/r2_sweep2.log | cut -c1-400; tail -3 gpurun_out/r2_bench_a.err
