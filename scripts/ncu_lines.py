#!/usr/bin/env python3
"""Per-CUDA-source-line executed-instruction and stall-sample shares of the first kernel in an .ncu-rep."""
import csv
import subprocess
import sys


def main(path, per=None, top=40, by=0):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr, agg = None, None, {}
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            hdr = r
            ie, isamp = hdr.index('Instructions Executed'), hdr.index('# Samples')
            continue
        if hdr is None or len(r) <= ie or r[2] != '-':
            continue  # only the per-CUDA-line summary rows (Address == '-')
        try:
            v, s = int(r[ie]), int(r[isamp])
        except ValueError:
            continue
        agg[(cur_file, int(r[0]), r[1].strip()[:100])] = (v, s)
    tot = sum(v for v, _ in agg.values()) or 1
    tots = sum(s for _, s in agg.values()) or 1
    print(f'total warp instructions {tot}, samples {tots}')
    for (f, ln, src), (v, s) in sorted(agg.items(), key=lambda kv: -kv[1][by])[:top]:
        extra = f'{v / per:7.1f}/unit ' if per else ''
        print(f'{100 * v / tot:5.1f}% instr {extra}{100 * s / tots:5.1f}% stall  {f}:{ln:<4d} {src}')


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None, int(sys.argv[3]) if len(sys.argv) > 3 else 40,
         1 if len(sys.argv) > 4 and sys.argv[4] == 'stall' else 0)
