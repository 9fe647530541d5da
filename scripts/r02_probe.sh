# quick kernel iteration: parity tests that touch the production kernels + throughput probes of every config
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_philox_replay.py tests/test_gpu_parity.py -m gpu -x -q -k "${PYTEST_K:-replay or finite or philox}" > gpurun_out/pytest_quick.log 2>&1; tail -3 gpurun_out/pytest_quick.log
python scripts/gpu_probe.py ${PROBE:-C1 C2 C3 C4 C5} > gpurun_out/probe.jsonl 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/probe.jsonl'):
    if l.startswith('{'):
        d = json.loads(l); print(d['cfg'], d['n_chains'], 'sweeps', d['n_sweeps'], 'tile', d['tile'], 'layout', d['layout'], 'ms %.2f' % d['ms'], 'rate %.3e' % d['proposals_per_s'], 'best %.2f mean %.2f' % (d['best_log2'], d['mean_best_log2']))
    else:
        print(l.rstrip()[:300])
PY
