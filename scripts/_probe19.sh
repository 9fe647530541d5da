mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "generated or split or app" 2>&1 | tail -2
python - <<'PY'
import sys, os, time, json; sys.path.insert(0,'scripts'); sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from tnco_b200 import networks
from tnco_b200.engine import Engine, pack_leaf_bits, random_trees
for cfg, nc in (('C2', 4096), ('C4', 4096), ('C5', 4096)):
    ts, ni = networks.CONFIGS[cfg]['make'](); lb = pack_leaf_bits(ts, ni)
    seeds = np.arange(nc, dtype=np.uint64)+1
    e = Engine(); e.set_network(lb, ni).set_mode()
    e.generate_chains(seeds); e.costs()
    t0=time.time(); e.generate_chains(seeds); t,_ = e.costs(); dt=time.time()-t0
    p,a,b = random_trees(lb, ni, seeds[:256])
    e2 = Engine(); e2.set_network(lb, ni).set_mode(); e2.set_chains(p,a,b,seeds[:256]); th,_ = e2.costs()
    print(json.dumps(dict(cfg=cfg, chains=nc, gen_init_s=round(dt,4), device_tree_log2=float(np.log2(t).mean()), host_tree_log2=float(np.log2(th).mean()))), flush=True)
    e.close(); e2.close()
PY
python scripts/e2e_profile.py C4 4096 2000 32 2>&1 | head -3
