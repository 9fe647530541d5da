# ncu --set full summary of the sweep kernel on every BASELINE config (second sweep launch of scripts/_prof.py)
mkdir -p gpurun_out
rm -f gpurun_out/all_configs_ncu_summary.txt
for cfg in "C1 32768 0 1000" "C2 4096 0 1000" "C3 8192 0 1000" "C4 4096 32 1000" "C5 4096 0 500"; do
set -- $cfg
ncu --set full --clock-control none -k regex:sa_sweep_kernel --launch-skip 1 --launch-count 1 -f -o /tmp/prof python scripts/_prof.py 0 $1 $2 $3 $4 > /tmp/p.log 2>&1
echo "== $1: $2 chains, max_width $3, $4 sweeps: $(grep -o '"proposals": [0-9]*' /tmp/p.log) proposals in the captured launch (9/10 of them)" >> gpurun_out/all_configs_ncu_summary.txt
python scripts/ncu_summary.py /tmp/prof.ncu-rep >> gpurun_out/all_configs_ncu_summary.txt
done
cat gpurun_out/all_configs_ncu_summary.txt
