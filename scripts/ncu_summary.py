#!/usr/bin/env python3
"""Summarise an .ncu-rep (first kernel) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv
import subprocess
import sys


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2] if len(rows) > 2 else rows[1]
    d = dict(zip(hdr, vals))
    scale = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'usecond': 1e3, 'msecond': 1e6, 'second': 1e9}
    for k, u in zip(hdr, units):   # bytes and nanoseconds, whatever unit ncu chose to print
        if u in scale and k in d:
            try:
                d[k] = repr(float(d[k].replace(',', '')) * scale[u])
            except ValueError:
                pass
    return d


def f(d, k):
    try:
        return float(d[k].replace(',', ''))
    except Exception:
        return float('nan')


def main(path):
    d = raw(path)
    print('kernel:', d.get('Kernel Name', '?')[:120])
    keys = [
        ('gpu__time_duration.sum', 'duration (ns)'),
        ('launch__registers_per_thread', 'registers/thread'),
        ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
        ('launch__occupancy_limit_registers', 'block limit regs'),
        ('smsp__inst_executed.sum', 'warp instructions'),
        ('smsp__thread_inst_executed_per_inst_executed.ratio', 'avg active threads/warp instr'),
        ('sm__inst_executed.avg.per_cycle_active', 'IPC active'),
        ('smsp__issue_active.avg.pct', 'issue slots busy %'),
        ('sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active', 'ALU pipe %'),
        ('sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active', 'FP64 pipe %'),
        ('sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active', 'FMA pipe %'),
        ('sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active', 'LSU pipe %'),
        ('sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active', 'XU pipe %'),
        ('l1tex__t_sector_hit_rate.pct', 'L1 hit %'), ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
        ('l1tex__t_bytes.sum', 'L1 bytes'), ('lts__t_bytes.sum', 'L2 bytes'),
        ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem wavefronts'),
    ]
    for k, name in keys:
        for kk in d:
            if kk == k:
                print(f'{name:34s} {d[kk]}')
    st = {k.replace('smsp__pcsamp_warps_issue_stalled_', ''): f(d, k) for k in d
          if 'pcsamp_warps_issue_stalled' in k and 'not_issued' not in k}
    tot = sum(st.values()) or 1
    print('stall reasons (% of samples):', ', '.join(f'{k} {100 * v / tot:.1f}' for k, v in
                                                     sorted(st.items(), key=lambda kv: -kv[1])[:8]))


if __name__ == '__main__':
    main(sys.argv[1])
