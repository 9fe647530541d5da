mkdir -p gpurun_out
for lay in 1 2; do
for cfg in "C5 4096 0 500" "C2 4096 0 1000"; do
set -- $cfg
ncu --set full --clock-control none -k regex:sa_sweep_kernel --launch-skip 1 --launch-count 1 -f -o /tmp/prof python scripts/_prof.py 0 $1 $2 $3 $4 $lay > /tmp/p.log 2>&1
echo "== $1 layout $lay" >> gpurun_out/layout_compare.txt
python scripts/ncu_summary.py /tmp/prof.ncu-rep >> gpurun_out/layout_compare.txt
done
done
cat gpurun_out/layout_compare.txt
