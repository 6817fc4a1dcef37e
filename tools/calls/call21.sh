mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest21.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke21.log 2>&1
# launch list of one image-step iteration (32 layers) of the final tree
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_final_raw.csv python tools/profile_step.py > gpurun_out/r02_profile_step_final.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err; echo "bench rc=$?" >> gpurun_out/r2_bench_f.err
tail -4 gpurun_out/r2_pytest21.log | cut -c1-300; tail -2 gpurun_out/r2_smoke21.log; tail -1 gpurun_out/r02_profile_step_final.log; wc -l gpurun_out/r02_launches_final_raw.csv; cut -c1-300 gpurun_out/r2_bench_f.json
