mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest18.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke19.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err; echo "bench rc=$?" >> gpurun_out/r2_bench_e.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?" >> gpurun_out/r2_bench_ref.err
tail -4 gpurun_out/r2_pytest19.log | cut -c1-300; tail -2 gpurun_out/r2_smoke19.log; cut -c1-300 gpurun_out/r2_bench_e.json; cut -c1-400 gpurun_out/r2_bench_ref.json
