# Per-source-line instruction / stall attribution of the sweep kernel on chosen configs (dev tool).
# usage: CFGS="C4 C5 C2" TAG=base bash scripts/ncu_probe_lines.sh
mkdir -p gpurun_out
for cfg in ${CFGS:-C4 C5 C2}; do
  case $cfg in C4) ARGS="'C4', 4096, 2000, max_width=32";; C5) ARGS="'C5', 4096, 1000";; C2) ARGS="'C2', 4096, 2000";; C3) ARGS="'C3', 8192, 500";; C1) ARGS="'C1', 32768, 1000";; esac
  ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__t_bytes.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum \
      --clock-control none --import-source on -k regex:sa_sweep --launch-skip 1 --launch-count 1 -f -o /tmp/pl_$cfg \
      python -c "
import sys; sys.path.insert(0, 'scripts'); sys.path.insert(0, '.')
from gpu_probe import probe
probe($ARGS)" > /tmp/pl_$cfg.log 2>&1
  PROPS=$(grep -o '"proposals": [0-9]*' /tmp/pl_$cfg.log | grep -o '[0-9]*$')
  OUT=gpurun_out/lines_${TAG:-x}_$cfg.txt
  echo "== $cfg proposals $PROPS" > $OUT
  python scripts/ncu_summary.py /tmp/pl_$cfg.ncu-rep >> $OUT 2>&1
  echo "-- by instructions" >> $OUT
  python scripts/ncu_lines.py /tmp/pl_$cfg.ncu-rep $PROPS 70 >> $OUT
  echo "-- by stall samples" >> $OUT
  python scripts/ncu_lines.py /tmp/pl_$cfg.ncu-rep $PROPS 50 stall >> $OUT
  python scripts/ncu_sass.py /tmp/pl_$cfg.ncu-rep $PROPS > gpurun_out/sass_${TAG:-x}_$cfg.txt
  python scripts/ncu_lines.py /tmp/pl_$cfg.ncu-rep $PROPS 2000 > gpurun_out/alllines_${TAG:-x}_$cfg.txt
  cp /tmp/pl_$cfg.ncu-rep gpurun_out/pl_${TAG:-x}_$cfg.ncu-rep
  head -30 $OUT
done
