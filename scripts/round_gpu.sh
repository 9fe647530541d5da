# Round evidence run: GPU tests, smoke, full ncu capture of the bench kernel (-> profiles/sweep_kernel_traffic.json, which
# bench.py quotes), bench (both arms), probes of every BASELINE config, C4 bench, 60-s anneal (C4).
# Outputs under gpurun_out/ (copy what is judged into profiles/).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
bash scripts/ncu_bench_kernel.sh > /dev/null 2>&1
P=$(python -c "import json;print([json.loads(l) for l in open('gpurun_out/b_ncu2.log') if l.startswith('{')][-1]['proposals_per_step'])")
python scripts/ncu_traffic_json.py gpurun_out/bench_kernel_ncu_summary.txt $P > gpurun_out/sweep_kernel_traffic.json && cp gpurun_out/sweep_kernel_traffic.json profiles/sweep_kernel_traffic.json
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
python scripts/gpu_probe.py > gpurun_out/probe_all.jsonl 2>&1; cut -c1-330 gpurun_out/probe_all.jsonl
python bench.py --workload C4 --sweeps 4000 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; cat gpurun_out/bench_c4.json
python scripts/anneal60.py C4 --budget 60 --max-width 32 > gpurun_out/anneal60_c4.json 2> gpurun_out/anneal60_c4.err; cat gpurun_out/anneal60_c4.json
