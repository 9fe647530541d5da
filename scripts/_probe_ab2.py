"""A/B of library builds on chosen probes (dev tool): python scripts/_probe_ab2.py base,nosfix 'C4:4096:2000:32' ..."""
import json, os, subprocess, sys
libs = sys.argv[1].split(',')
specs = sys.argv[2:]
code = "import sys; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')\nfrom gpu_probe import probe\n"
for s in specs:
    cfg, nc, ns, mw = (s.split(':') + ['0'])[:4]
    code += f"probe('{cfg}', {nc}, {ns}, max_width={mw if float(mw) > 0 else None})\n"
for rep in range(2):
    for lib in libs:
        env = dict(os.environ)
        if lib != 'base':
            env['TNB_LIB'] = f'tnco_b200/libtnco_b200_{lib}.so'
        out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True)
        for l in out.stdout.splitlines():
            if l.startswith('{'):
                d = json.loads(l)
                print(lib, d['cfg'], d['n_chains'], 'x', d['n_sweeps'], 'tile', d['tile'], '%.3e' % d['proposals_per_s'], 'best %.2f' % d['best_log2'], flush=True)
        if out.returncode:
            print(lib, 'FAILED', out.stderr[-500:])
