mkdir -p gpurun_out
MMDP_PROFILE_LAYERS=2 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'gemm_pair|gemm_bf16|attention_v6|attention_combine|rmsnorm|text_rows|text_commit|image_rows|image_remask|embed' -c 44 -o gpurun_out/r02_full_final python tools/profile_step.py > gpurun_out/r02_ncu_full_final.log 2>&1
tail -2 gpurun_out/r02_ncu_full_final.log; ls -la gpurun_out/r02_full_final.ncu-rep
