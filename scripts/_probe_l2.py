"""A/B of the L2 carve-out for the hot block of SPLIT batches (dev tool): TNB_L2_PERSIST_MB = 0 (off), -1 (device
maximum) and fixed sizes; every setting in its own process, same probes."""
import json
import os
import subprocess
import sys

CODE = r'''
import sys; sys.path.insert(0, 'scripts'); sys.path.insert(0, '.')
from gpu_probe import probe
import torch
p = torch.cuda.get_device_properties(0)
print('L2', p.L2_cache_size, flush=True)
probe('C3', 12288, 1000)
probe('C4', 4096, 2000, max_width=32)
probe('C4', 8192, 1000, max_width=32)
probe('C5', 4096, 1000)
probe('C5', 8192, 500)
'''
settings = sys.argv[1:] or ['0', '-1', '32', '64', '0:32', '0:128', '-1:32']
for mb in settings:
    env = dict(os.environ, TNB_L2_PERSIST_MB=mb.split(':')[0])
    if ':' in mb:
        env['TNB_L2_FETCH'] = mb.split(':')[1]
    out = subprocess.run([sys.executable, '-c', CODE], env=env, capture_output=True, text=True)
    for l in out.stdout.splitlines():
        if l.startswith('{'):
            d = json.loads(l)
            print('persist_mb', mb, d['cfg'], d['n_chains'], 'x', d['n_sweeps'], 'layout', d['layout'],
                  '%.3e' % d['proposals_per_s'], 'best %.2f' % d['best_log2'], flush=True)
        else:
            print(mb, l[:200], flush=True)
    if out.returncode:
        print('persist_mb', mb, 'FAILED', out.stderr[-800:], flush=True)
