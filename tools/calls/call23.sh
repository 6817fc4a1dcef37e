mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_modes.py -m gpu -q --timeout 600 > gpurun_out/r2_pytest23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest23.log
timeout 600 python bench.py --variant m --steps 2 --warmup 1 > gpurun_out/r2_bench_m.json 2> gpurun_out/r2_bench_m.err; echo "rc=$?" >> gpurun_out/r2_bench_m.err
tail -4 gpurun_out/r2_pytest23.log | cut -c1-300; tail -2 gpurun_out/r2_bench_m.err | cut -c1-200; cut -c1-200 gpurun_out/r2_bench_m.json
