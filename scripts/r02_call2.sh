# Round-2 GPU call 2: smem layout correctness + measurement, new bench.py end to end (short), ncu of smem vs in-place on C1/C2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shared_memory or split_layout" > gpurun_out/pytest_smem.log 2>&1; tail -5 gpurun_out/pytest_smem.log
python scripts/gpu_probe.py SMEM > gpurun_out/r02_probe_smem.jsonl 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/r02_probe_smem.jsonl'):
    if l.startswith('{'):
        d = json.loads(l); print(d['cfg'], d['n_chains'], 'tile', d['tile'], 'layout', d['layout'], 'ms %.2f' % d['ms'], 'rate %.3e' % d['proposals_per_s'])
    else:
        print(l.rstrip()[:300])
PY
timeout 900 python bench.py --steps 3 --warmup 3 --anneal-budget 10 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; tail -3 gpurun_out/bench_short.err; cut -c1-3000 gpurun_out/bench_short.json
