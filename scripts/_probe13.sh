mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python - <<'PY' > gpurun_out/probe13.log 2>&1
import sys, os; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
from gpu_probe import probe
probe('C2', 4096, 2000)
probe('C2', 16384, 1000)
probe('C1', 32768, 2000)
probe('C3', 8192, 1000)
probe('C4', 4096, 2000, max_width=32)
probe('C4', 4096, 2000)
probe('C5', 4096, 500)
PY
cat gpurun_out/probe13.log
