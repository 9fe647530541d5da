# Round-2 GPU call 1: GPU test suite (incl. the production-kernel replay + equal-sweep statistics), measured INT / POPC /
# L1 / L2 peaks, probes of every config with the round-1 kernels, full ncu capture of the C4 production kernel.
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/peaks scripts/peaks_microbench.cu && /tmp/peaks > gpurun_out/r02_int_peaks.json 2> gpurun_out/r02_int_peaks.err; cat gpurun_out/r02_int_peaks.json
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log; grep -E "^c[1-4]_" gpurun_out/pytest_gpu.log
python scripts/gpu_probe.py > gpurun_out/r02_probe_base.jsonl 2>&1; cut -c1-400 gpurun_out/r02_probe_base.jsonl
ncu --set full --clock-control none --import-source on -k regex:sa_sweep_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_c4_base python scripts/gpu_probe.py C4 > gpurun_out/ncu_c4.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02_c4_base.ncu-rep > gpurun_out/r02_c4_base_summary.txt; head -60 gpurun_out/r02_c4_base_summary.txt
ls -la gpurun_out/*.ncu-rep
