# Round-2 GPU call 3: full GPU test suite, bench end to end (short anneal), ncu of every config + smem-vs-in-place
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --anneal-budget 10 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; tail -3 gpurun_out/bench_short.err; cut -c1-1500 gpurun_out/bench_short.json
bash scripts/ncu_all_configs.sh > gpurun_out/ncu_all.log 2>&1; tail -5 gpurun_out/ncu_all.log; cut -c1-600 gpurun_out/r02_sweep_kernel_ncu.json
# shared-memory-resident layout against the in-place layout under ncu (N1): C1 at tile 8, C2 at tile 32
rm -f gpurun_out/r02_smem_vs_inplace_ncu.txt
for spec in "C1 8 1" "C1 8 3" "C1 4 1" "C1 4 3" "C2 32 1" "C2 32 3"; do
  set -- $spec
  ncu --set full --clock-control none -k regex:sa_sweep --launch-skip 1 --launch-count 1 -f -o /tmp/prof_smem \
      python -c "
import sys; sys.path.insert(0, 'scripts'); sys.path.insert(0, '.')
from gpu_probe import probe
probe('$1', 32768 if '$1' == 'C1' else 4096, 2000, tile=$2, layout=$3)" > /tmp/p.log 2>&1
  echo "== $1 tile $2 layout $3 (1 = in place / L1-L2, 3 = shared-memory resident): $(grep -o '"proposals": [0-9]*' /tmp/p.log) proposals in the captured launch" >> gpurun_out/r02_smem_vs_inplace_ncu.txt
  python scripts/ncu_summary.py /tmp/prof_smem.ncu-rep >> gpurun_out/r02_smem_vs_inplace_ncu.txt
done
tail -30 gpurun_out/r02_smem_vs_inplace_ncu.txt
