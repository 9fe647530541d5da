#!/usr/bin/env python3
"""Per-CUDA-source-line lane utilisation of the first kernel in an .ncu-rep: executed warp instructions, average
active threads, and the share of all idle lane-slots the line is responsible for."""
import csv
import subprocess
import sys


def main(path, top=35):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur, hdr, agg = None, None, {}
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            hdr = r
            ie, it = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
            continue
        if hdr is None or len(r) <= max(ie, it) or r[2] != '-':
            continue
        try:
            v, th = int(r[ie]), int(r[it])
        except ValueError:
            continue
        agg[(cur, int(r[0]), r[1].strip()[:90])] = (v, th)
    tot = sum(v for v, _ in agg.values()) or 1
    idle_tot = sum(32 * v - th for v, th in agg.values()) or 1
    print(f'warp instructions {tot}, average active lanes {sum(th for _, th in agg.values()) / tot:.2f}')
    for (f, ln, src), (v, th) in sorted(agg.items(), key=lambda kv: -(32 * kv[1][0] - kv[1][1]))[:top]:
        print(f'{100 * (32 * v - th) / idle_tot:5.1f}% of idle slots  {100 * v / tot:5.1f}% instr  {th / max(v, 1):5.1f} lanes  {f}:{ln:<4d} {src}')


if __name__ == '__main__':
    main(sys.argv[1])
