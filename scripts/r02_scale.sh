# Round-2 scaling runs on one 8-GPU box, launched like the driver does: default workload (C4), weak scaling
# (4096 chains per GPU), the per-config runs and the fixed-wall-clock anneal included.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in ${NS:-8 4 2}; do
  $TR --nproc-per-node $N --master-port $((29520 + N)) bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 --anneal-budget ${ANNEAL:-60} \
      > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open('gpurun_out/r02_bench_${N}gpu.json') if l.startswith('{')][-1]
    print('N=${N}', 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], 'ms/step %.1f' % d['ms_per_step'],
          d['e2e'].get('last_step_ms'), d['clocks'], {k: '%.3e' % v['value'] for k, v in d.get('configs', {}).items()},
          (d.get('best_log2_at_60s') or {}).get('gpu'))
except Exception as e:
    print('N=${N} failed', e); print(open('gpurun_out/r02_bench_${N}gpu.err').read()[-1500:])
PY
done
