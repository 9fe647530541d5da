# ncu --set full of the TIMED bench launch of the sweep kernel on every BASELINE config (the 4th sweep launch of
# `bench.py --workload Cx --steps 1 --warmup 3`), summaries -> gpurun_out/r02_all_configs_ncu_summary.txt and the
# per-proposal figures bench.py quotes (roofline.traffic / roofline.issue) -> gpurun_out/r02_sweep_kernel_ncu.json
mkdir -p gpurun_out
OUT=gpurun_out/r02_all_configs_ncu_summary.txt
rm -f $OUT
for cfg in ${CFGS:-C4 C2 C1 C3 C5}; do
  ncu --set full --clock-control none --import-source on -k regex:sa_sweep --launch-skip 3 --launch-count 1 -f -o /tmp/prof_$cfg \
      python bench.py --workload $cfg --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-configs --anneal-budget 0 > /tmp/bench_$cfg.log 2>&1
  echo "== $cfg: $(python -c "
import json
d=[json.loads(l) for l in open('/tmp/bench_$cfg.log') if l.startswith('{')][-1]
print(d['config']['workload'], '|', d['config']['chains_per_gpu'], 'chains x', d['config']['sweeps_per_step'], 'sweeps |', int(d['proposals_per_step']), 'proposals in the captured launch')")" >> $OUT
  python scripts/ncu_summary.py /tmp/prof_$cfg.ncu-rep >> $OUT
  [ "$cfg" = "${KEEP_REP:-C4}" ] && cp /tmp/prof_$cfg.ncu-rep gpurun_out/r02_${cfg}_bench.ncu-rep
done
python scripts/ncu_configs_json.py $OUT > gpurun_out/r02_sweep_kernel_ncu.json
cat $OUT
