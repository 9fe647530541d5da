# Round-2 evidence run on one B200: full GPU test suite, smoke, the default bench (both arms), launch list, ncu of the
# timed launch on every config.  Outputs -> gpurun_out/ (copy what is to be kept into profiles/).
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-300 gpurun_out/r02_bench_reference.json
timeout 1200 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -2 gpurun_out/r02_bench.err; cut -c1-300 gpurun_out/r02_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --anneal-budget 0 > gpurun_out/b_ncu.log 2>&1
KEEP_REP=none bash scripts/ncu_all_configs.sh > gpurun_out/ncu_all.log 2>&1; tail -3 gpurun_out/ncu_all.log
