mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
timeout 600 python tools/gpu_sweep_forward.py > gpurun_out/r2_sweep1.log 2>&1; echo "rc=$?" >> gpurun_out/r2_sweep1.log
timeout 300 python tools/gpu_check_kernels.py --only splitk,gemm_qkv,gemm_swiglu_b1,qkv_attn_real,pair_qkv,gemm_ffout > gpurun_out/r2_kernels1.log 2>&1
timeout 200 python tools/cublas_shapes.py > gpurun_out/r2_cublas.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --section LaunchStats --clock-control none -c 60 --csv --log-file gpurun_out/r2_cublas_ncu.csv python tools/cublas_shapes.py > /dev/null 2>&1
tail -5 gpurun_out/r2_pytest1.log; tail -15 gpurun_out/r2_sweep1.log
