mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tp.py -m gpu -q --timeout 500 -s > gpurun_out/r2_pytest15_tp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest15_tp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err; echo "rc=$?" >> gpurun_out/r2_bench_n2b.err
tail -12 gpurun_out/r2_pytest15_tp.log | cut -c1-300; tail -3 gpurun_out/r2_bench_n2b.err; python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_bench_n2b.json') if l.startswith('{')][-1]); t=d['tp']
    print(d['value'], t['value'], t['row_chunks'], t['other_row_chunk_schedule'], t['nccl_allreduce_baseline_tokens_per_s'], t['tp_parity']['ok'], t['tp_parity']['ranks_final_ids_identical'], t['kernel_breakdown_one_sample_ms'])
except Exception as e: print('ERR', e)
PY
