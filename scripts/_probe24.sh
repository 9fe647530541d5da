export TNB_EVERY=1
ncu --set full --clock-control none --import-source on -k regex:sa_sweep_kernel --launch-skip 1 --launch-count 1 -f -o /tmp/prof python scripts/_prof.py 0 C4 4096 32 300 > /tmp/p.log 2>&1
python scripts/ncu_summary.py /tmp/prof.ncu-rep | tail -4
python scripts/ncu_lines.py /tmp/prof.ncu-rep | head -60
