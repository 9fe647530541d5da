mkdir -p gpurun_out
python - <<'PY' > gpurun_out/probe16.log 2>&1
import sys, os; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
from gpu_probe import probe
for np_ in ('1', ''):
    if np_: os.environ['TNB_NO_PERSIST']='1'
    else: os.environ.pop('TNB_NO_PERSIST', None)
    print('NO_PERSIST', np_, flush=True)
    probe('C2', 4096, 2000)
    probe('C2', 16384, 1000)
    probe('C3', 8192, 1000)
    probe('C4', 4096, 2000, max_width=32)
    probe('C4', 4096, 2000)
    probe('C5', 4096, 500)
    probe('C5', 4096, 3000)
PY
cat gpurun_out/probe16.log
timeout 600 python -m pytest tests -m gpu -x -q -k "split or philox or golden" > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
