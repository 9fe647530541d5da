mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "finite_width_chains_are_valid or split_layout or test_philox_chains or hyper_index_networks" 2>&1 | tail -4
python - <<'PY' > gpurun_out/probe20.log 2>&1
import sys, os; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
from gpu_probe import probe
probe('C2', 4096, 2000)
probe('C2', 16384, 1000)
probe('C1', 32768, 2000)
probe('C3', 8192, 1000)
probe('C4', 4096, 2000, max_width=32)
probe('C5', 4096, 500)
PY
python - <<'PY'
import json
for l in open('gpurun_out/probe20.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['cfg'], d['n_chains'], d['max_width'], 'ms=%.1f'%d['ms'], 'rate=%.3e'%d['proposals_per_s'])
    else: print(l.strip())
PY
