mkdir -p gpurun_out
timeout 600 python tools/gpu_check_attention.py > gpurun_out/r2_check_attn2.log 2>&1; echo "check rc=$?" >> gpurun_out/r2_check_attn2.log
MMDP_PROF_VERSIONS=7 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attention_v' --launch-skip 2 --launch-count 1 -o gpurun_out/r02_attn_v7b python tools/profile_attention.py > gpurun_out/r02_ncu_attn7b.log 2>&1
cut -c1-420 gpurun_out/r2_check_attn2.log | tail -8; tail -2 gpurun_out/r02_ncu_attn7b.log
