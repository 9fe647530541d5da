#!/usr/bin/env python3
"""profiles/sweep_kernel_traffic.json (what bench.py quotes as roofline.traffic / roofline.issue) from the ncu
summary of one bench launch of the sweep kernel (scripts/ncu_bench_kernel.sh -> scripts/ncu_summary.py).
usage: ncu_traffic_json.py <bench_kernel_ncu_summary.txt> <proposals_per_launch> > profiles/sweep_kernel_traffic.json"""
import json
import re
import sys


def main(path, proposals):
    vals = {}
    for line in open(path):
        m = re.match(r'(.+?)\s{2,}([-0-9.eE+]+)\s*$', line.rstrip())
        if m:
            vals[m.group(1).strip()] = float(m.group(2))
    kernel = next((l.split(':', 1)[1].strip() for l in open(path) if l.startswith('kernel:')), '?')
    rd, wr = vals['DRAM read'], vals['DRAM write']
    # ncu_summary prints the raw-page values: MB for DRAM traffic when the unit column says so -- normalise to bytes
    scale = 1e6 if rd < 1e5 else 1.0
    rd, wr = rd * scale, wr * scale
    wi = vals['warp instructions']
    out = {
        'kernel': kernel.replace('void ', '').replace('(Params)', ''),
        'source': 'ncu --set full of `bench.py --steps 1 --warmup 3` (3rd sweep launch), ' + path,
        'workload': 'C2, 4096 chains x 10000 sweeps (one launch, %.1fe6 proposals)' % (proposals / 1e6),
        'dram_bytes_read': rd, 'dram_bytes_write': wr, 'dram_bytes_per_launch': rd + wr,
        'proposals_per_launch': proposals, 'dram_bytes_per_proposal': (rd + wr) / proposals,
        'warp_instructions_per_launch': wi, 'warp_instructions_per_proposal': wi / proposals,
        'ipc_active': vals.get('IPC active'), 'duration_ms_under_ncu': vals.get('duration (ns)'),
        'note': 'DRAM writes are mostly the write-back of lines dirtied by the L2 flush that precedes every timed step',
    }
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]))
