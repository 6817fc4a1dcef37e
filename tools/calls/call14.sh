mkdir -p gpurun_out
timeout 120 tools/cuda/umma_rate > gpurun_out/r2_umma_rate.log 2>&1; echo "rc=$?" >> gpurun_out/r2_umma_rate.log
cat gpurun_out/r2_umma_rate.log
