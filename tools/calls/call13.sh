mkdir -p gpurun_out
timeout 300 python tools/gpu_probe_attention.py > gpurun_out/r2_probe_attn.log 2>&1; echo "rc=$?" >> gpurun_out/r2_probe_attn.log
cat gpurun_out/r2_probe_attn.log
