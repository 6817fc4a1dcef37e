mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 500 -k "gemm" > gpurun_out/r2_pytest17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest17.log
timeout 300 python tools/gpu_probe_gemm.py > gpurun_out/r2_probe_gemm2.log 2>&1
timeout 300 python tools/gpu_sweep_forward.py --variants cur_pair_nsplit cur_all --rounds 4 > gpurun_out/r2_sweep17.log 2>&1
tail -6 gpurun_out/r2_pytest17.log | cut -c1-300; grep gate_up gpurun_out/r2_probe_gemm2.log | cut -c1-500; grep -o '"variant": "[a-z_]*"\|"ms_per_forward_median": [0-9.]*\|"gemm_tflops": [0-9.]*\|"max_abs_diff_text": [0-9.e-]*' gpurun_out/r2_sweep17.log
