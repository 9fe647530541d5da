#!/usr/bin/env python3
"""Per-SASS-instruction executed counts (per unit) and stall samples of the first kernel in an .ncu-rep."""
import csv
import subprocess
import sys


def main(path, per):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    for r in rows:
        if not r:
            continue
        if 'Instructions Executed' in r and 'Source' in r:
            hdr = r
            ia, isrc = hdr.index('Address'), hdr.index('Source')
            ie, isamp = hdr.index('Instructions Executed'), hdr.index('# Samples')
            reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
            continue
        if hdr is None or len(r) <= max(ie, isamp):
            continue
        try:
            v, s = int(r[ie]), int(r[isamp])
        except ValueError:
            continue
        top = ''
        try:
            i, h = max(reasons, key=lambda ih: int(r[ih[0]] or 0))
            if int(r[i] or 0) > 0:
                top = f'{h[6:]}={r[i]}'
        except (ValueError, IndexError):
            pass
        print(f'{r[ia][-5:]} {v / per:7.3f} {s:7d}  {r[isrc][:90]:90s} {top}')


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]))
