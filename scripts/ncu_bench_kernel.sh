# Full ncu capture of one bench launch of the sweep kernel + launch list; summaries into gpurun_out/ (copy to profiles/).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sa_sweep_kernel --launch-skip 2 --launch-count 1 -f -o /tmp/prof_bench python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 --e2e-warmup 0 > gpurun_out/b_ncu2.log 2>&1
python scripts/ncu_summary.py /tmp/prof_bench.ncu-rep > gpurun_out/bench_kernel_ncu_summary.txt
python scripts/ncu_lines.py /tmp/prof_bench.ncu-rep 486500000 > gpurun_out/bench_kernel_ncu_lines.txt
python scripts/ncu_sass.py /tmp/prof_bench.ncu-rep 486500000 > gpurun_out/bench_kernel_ncu_sass.txt
cat gpurun_out/bench_kernel_ncu_summary.txt
