python - <<'PY'
import sys, os, json; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
import io, contextlib
from gpu_probe import probe
for a in [('C2', 4096, 2000), ('C2', 16384, 1000), ('C1', 32768, 2000), ('C3', 8192, 1000), ('C4', 4096, 2000, 32), ('C5', 4096, 500)]:
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        d = probe(a[0], a[1], a[2], max_width=a[3] if len(a) > 3 else None)
    print(d['cfg'], d['n_chains'], d['max_width'], 'ms=%.1f'%d['ms'], 'rate=%.3e'%d['proposals_per_s'], flush=True)
PY
