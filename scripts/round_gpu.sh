# Round evidence run: GPU tests, smoke, bench (both arms), launch list, full ncu capture of the bench kernel, 60-s anneals.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sa_sweep_kernel --launch-skip 2 --launch-count 1 -f -o /tmp/prof_bench python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 --e2e-warmup 0 > gpurun_out/b_ncu2.log 2>&1
python scripts/ncu_summary.py /tmp/prof_bench.ncu-rep > gpurun_out/bench_kernel_ncu_summary.txt
python scripts/ncu_lines.py /tmp/prof_bench.ncu-rep 486500000 > gpurun_out/bench_kernel_ncu_lines.txt
python scripts/ncu_sass.py /tmp/prof_bench.ncu-rep 486500000 > gpurun_out/bench_kernel_ncu_sass.txt
python scripts/anneal60.py C4 --budget 60 --max-width 32 > gpurun_out/anneal60_c4.json 2> gpurun_out/anneal60_c4.err; cat gpurun_out/anneal60_c4.json
python scripts/anneal60.py C2 --budget 60 > gpurun_out/anneal60_c2.json 2> gpurun_out/anneal60_c2.err; cat gpurun_out/anneal60_c2.json
