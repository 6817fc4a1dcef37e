mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest24.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest24.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke24.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err; echo "bench rc=$?" >> gpurun_out/r2_bench_g.err
tail -3 gpurun_out/r2_pytest24.log | cut -c1-300; tail -1 gpurun_out/r2_smoke24.log; cut -c1-260 gpurun_out/r2_bench_g.json
