mkdir -p gpurun_out
python - <<'PY' > gpurun_out/probe3.log 2>&1
import sys, os; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
from gpu_probe import probe
for mb in ('20','28'):
    os.environ['TNB_MINB']=mb
    print('MINB', mb, flush=True)
    probe('C2', 4096, 2000)
    probe('C2', 16384, 1000)
    probe('C3', 8192, 1000)
    probe('C1', 32768, 2000)
    probe('C5', 4096, 500)
PY
cat gpurun_out/probe3.log
