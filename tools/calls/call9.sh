mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest9.log
timeout 300 python tools/gpu_probe_gemm.py > gpurun_out/r2_probe_gemm.log 2>&1
timeout 300 python tools/gpu_sweep_forward.py --variants cur_pair cur_pair_nsplit --rounds 4 > gpurun_out/r2_sweep9.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; echo "bench rc=$?" >> gpurun_out/r2_bench_c.err
tail -4 gpurun_out/r2_pytest9.log; cat gpurun_out/r2_probe_gemm.log | cut -c1-400; grep -o '"variant": "[a-z_]*"\|"ms_per_forward_median": [0-9.]*\|"gemm_tflops": [0-9.]*' gpurun_out/r2_sweep9.log; cut -c1-300 gpurun_out/r2_bench_c.json
