mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:sa_sweep_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_c4 python scripts/_prof.py 0 C4 1024 32 200 > gpurun_out/p_c4.log 2>&1
python - <<'PY' > gpurun_out/probe4.log 2>&1
import sys, os; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
from gpu_probe import probe
probe('C4', 4096, 500, max_width=32)
probe('C4', 4096, 500)
probe('C3', 4096, 1000, max_width=28)
probe('C2', 4096, 2000)
PY
cat gpurun_out/probe4.log
