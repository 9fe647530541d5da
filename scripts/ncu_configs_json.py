#!/usr/bin/env python3
"""profiles/r02_sweep_kernel_ncu.json from the summaries scripts/ncu_all_configs.sh writes: per BASELINE config the
per-proposal DRAM traffic and warp-instruction count of the timed bench launch (what bench.py quotes as
roofline.traffic / roofline.issue), plus the counters DESIGN.md cites."""
import json
import re
import sys


def main(path):
    out, cur = {}, None
    for line in open(path):
        m = re.match(r'== (C\d): (.*)\| (\d+) chains x (\d+) sweeps \| (\d+) proposals', line)
        if m:
            cur = out[m.group(1)] = dict(workload=m.group(2).strip(), chains=int(m.group(3)), sweeps=int(m.group(4)),
                                         proposals_per_launch=int(m.group(5)), counters={})
            continue
        if cur is None:
            continue
        if line.startswith('kernel:'):
            cur['kernel'] = line.split(':', 1)[1].strip().replace('void ', '')
        elif line.startswith('stall reasons'):
            cur['stall_reasons_pct'] = line.split(':', 1)[1].strip()
        else:
            m = re.match(r'(.+?)\s{2,}([-0-9.eE+,]+)\s*$', line.rstrip())
            if m:
                cur['counters'][m.group(1).strip()] = float(m.group(2).replace(',', ''))
    for name, c in out.items():
        k = c['counters']
        rd, wr = k.get('DRAM read', 0.0), k.get('DRAM write', 0.0)
        c['dram_bytes_per_launch'] = rd + wr   # (scripts/ncu_summary.py prints bytes)
        c['dram_bytes_per_proposal'] = (rd + wr) / c['proposals_per_launch']
        c['warp_instructions_per_proposal'] = k.get('warp instructions', 0.0) / c['proposals_per_launch']
        c['source'] = 'ncu --set full --clock-control none of the timed launch of `bench.py --workload %s --steps 1 --warmup 3`' % name
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main(sys.argv[1])
