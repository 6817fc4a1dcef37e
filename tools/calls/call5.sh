mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -s > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest5.log
# launch list of one image-step iteration (32 layers)
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_image_step_raw.csv python tools/profile_step.py > gpurun_out/r02_profile_step.log 2>&1
# --set full on the kernels of a 2-layer iteration: pair GEMMs, attention (+combine), rmsnorm, the sampling kernels
MMDP_PROFILE_LAYERS=2 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'gemm_pair|gemm_bf16|attention_v6|attention_combine|rmsnorm|text_rows|text_commit|image_rows|image_remask' -c 40 -o gpurun_out/r02_full python tools/profile_step.py > gpurun_out/r02_ncu_full.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; echo "bench rc=$?" >> gpurun_out/r2_bench_b.err
tail -4 gpurun_out/r2_pytest5.log; tail -2 gpurun_out/r02_ncu_full.log; cut -c1-200 gpurun_out/r2_bench_b.json
