mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:sa_sweep_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_c5 python scripts/_prof.py 0 C5 4096 0 500 > gpurun_out/p_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sa_sweep_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_c4d python scripts/_prof.py 0 C4 4096 32 1000 > gpurun_out/p_c4d.log 2>&1
