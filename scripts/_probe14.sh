mkdir -p gpurun_out
python - <<'PY' > gpurun_out/probe14.log 2>&1
import sys, os; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
from gpu_probe import probe
for pf in ('0','1'):
    os.environ['TNB_PREFETCH2']=pf
    print('PF', pf, flush=True)
    probe('C2', 4096, 2000)
    probe('C2', 16384, 1000)
    probe('C3', 8192, 1000)
    probe('C4', 4096, 2000, max_width=32)
    probe('C4', 4096, 2000)
    probe('C5', 4096, 500)
PY
cat gpurun_out/probe14.log
