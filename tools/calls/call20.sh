mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 4 --steps 2 --warmup 1 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "rc=$?" >> gpurun_out/r2_bench_n4.err
tail -3 gpurun_out/r2_bench_n4.err; python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_bench_n4.json') if l.startswith('{')][-1]); t=d['tp']
    print(d['value'], d['variant_m']['value'] if 'variant_m' in d else None, t['value'], t['row_chunks'], t['other_row_chunk_schedule'], t['nccl_allreduce_baseline_tokens_per_s'], t['tp_parity']['ok'], t['tp_parity']['ranks_final_ids_identical'])
except Exception as e: print('ERR', e)
PY
