mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "finite or split or philox" > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python - <<'PY' > gpurun_out/probe10.log 2>&1
import sys, os; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
from gpu_probe import probe
probe('C4', 4096, 2000, max_width=32)
probe('C4', 1024, 2000, max_width=32)
probe('C4', 16384, 1000, max_width=32)
probe('C3', 4096, 1000, max_width=28)
PY
cat gpurun_out/probe10.log
