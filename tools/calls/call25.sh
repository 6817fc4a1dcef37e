mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 500 > gpurun_out/r2_pytest25.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest25.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke25.log 2>&1
tail -3 gpurun_out/r2_pytest25.log | cut -c1-200; tail -1 gpurun_out/r2_smoke25.log
