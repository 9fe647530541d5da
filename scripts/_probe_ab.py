import os, subprocess, sys, json
# A/B of library builds (dev tool): every variant in its own process, same probes
libs = sys.argv[1].split(',')
cfgs = sys.argv[2:]
for lib in libs:
    env = dict(os.environ)
    if lib != 'base':
        env['TNB_LIB'] = f'tnco_b200/libtnco_b200_{lib}.so'
    out = subprocess.run([sys.executable, 'scripts/gpu_probe.py'] + cfgs, env=env, capture_output=True, text=True).stdout
    for l in out.splitlines():
        if l.startswith('{'):
            d = json.loads(l)
            print(lib, d['cfg'], d['n_chains'], 'tile', d['tile'], '%.3e' % d['proposals_per_s'], flush=True)
