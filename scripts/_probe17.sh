mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python - <<'PY' > gpurun_out/probe17.log 2>&1
import sys, os; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
from gpu_probe import probe
probe('C2', 4096, 2000)
probe('C4', 4096, 2000, max_width=32)
probe('C5', 4096, 500)
PY
cat gpurun_out/probe17.log | cut -c1-50,330-460
python scripts/e2e_profile.py C2 4096 10000 2>&1 | head -4
