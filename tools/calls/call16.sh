mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2_bench_n8b.json 2> gpurun_out/r2_bench_n8b.err; echo "rc=$?" >> gpurun_out/r2_bench_n8b.err
tail -3 gpurun_out/r2_bench_n8b.err; python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_bench_n8b.json') if l.startswith('{')][-1]); t=d['tp']
    print(d['value'], t['value'], t['row_chunks'], t['other_row_chunk_schedule'], t['nccl_allreduce_baseline_tokens_per_s'], t['tp_parity']['ok'], t['tp_parity']['ranks_final_ids_identical'], t['kernel_breakdown_one_sample_ms'])
except Exception as e: print('ERR', e)
PY
grep -m3 "mmdp:" gpurun_out/r2_bench_n8b.json
