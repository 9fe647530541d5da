mkdir -p gpurun_out
python scripts/e2e_profile.py C2 4096 10000 > gpurun_out/e2e_c2.log 2>&1
python scripts/e2e_profile.py C4 4096 2000 32 > gpurun_out/e2e_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sa_sweep_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_c4b python scripts/_prof.py 0 C4 4096 32 500 > gpurun_out/p_c4b.log 2>&1
