#!/usr/bin/env python3
"""Decode the scheduling control fields of `cuobjdump -sass` output (dev tool): per instruction the scoreboard it
sets on completion (write barrier), on operand read (read barrier) and the scoreboards it waits for.
usage: cuobjdump -sass -fun NAME file.o | python scripts/sass_ctrl.py [lo hi]   (hex address range)"""
import re
import sys

lo = int(sys.argv[1], 16) if len(sys.argv) > 1 else 0
hi = int(sys.argv[2], 16) if len(sys.argv) > 2 else 1 << 60
lines = sys.stdin.read().splitlines()
i = 0
while i < len(lines):
    m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*?);\s*/\* 0x([0-9a-f]+) \*/', lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r'\s*/\* 0x([0-9a-f]+) \*/', lines[i + 1])
        if m2:
            a, ins, up = int(m.group(1), 16), m.group(2).strip(), int(m2.group(1), 16)
            c = (up >> 41) & ((1 << 21) - 1)
            wb, rb, wait = (c >> 5) & 7, (c >> 8) & 7, (c >> 11) & 0x3f
            if lo <= a <= hi:
                ws = ','.join(str(k) for k in range(6) if wait >> k & 1)
                print(f'{a:05x}  {"W%d" % wb if wb != 7 else "  "} {"R%d" % rb if rb != 7 else "  "} wait[{ws:<7}] st{c & 0xf:<2} {ins}')
            i += 2
            continue
    i += 1
