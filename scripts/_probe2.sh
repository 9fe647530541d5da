mkdir -p gpurun_out
python - <<'PY' > gpurun_out/probe2.log 2>&1
import sys; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
from gpu_probe import probe
probe('C2', 4096, 2000)
probe('C2', 4096, 2000, tile=32)
probe('C2', 16384, 1000)
probe('C2', 16384, 1000, tile=32)
probe('C1', 32768, 2000)
probe('C1', 32768, 2000, tile=8)
probe('C3', 8192, 1000)
PY
cat gpurun_out/probe2.log
