mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attention_v' --launch-skip 2 --launch-count 1 -o gpurun_out/r02_attn_v7 python tools/profile_attention.py > gpurun_out/r02_ncu_attn7.log 2>&1
MMDP_PROF_VERSIONS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attention_v' --launch-skip 2 --launch-count 1 -o gpurun_out/r02_attn_v6 python tools/profile_attention.py > gpurun_out/r02_ncu_attn6.log 2>&1
tail -3 gpurun_out/r02_ncu_attn7.log; tail -3 gpurun_out/r02_ncu_attn6.log; ls -la gpurun_out/*.ncu-rep | tail -3
