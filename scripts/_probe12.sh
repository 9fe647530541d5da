mkdir -p gpurun_out
python - <<'PY' > gpurun_out/probe12.log 2>&1
import sys, os, time, json; sys.path.insert(0,'scripts'); sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from tnco_b200 import networks
from tnco_b200.engine import Engine, pack_leaf_bits
for cfg, nc in (('C5', 4096), ('C3', 12500), ('C4', 4096), ('C1', 32768)):
    ts, ni = networks.CONFIGS[cfg]['make'](); lb = pack_leaf_bits(ts, ni)
    seeds = np.arange(nc, dtype=np.uint64)+1
    e = Engine(); e.set_network(lb, ni).set_mode()
    t0=time.time(); e.generate_chains(seeds); t,_ = e.costs(); dt=time.time()-t0
    e.set_betas(np.linspace(0,100,500,endpoint=False)); e.timing(); e.run(50); e.timing(); c0=e.counters(); e.run(500); ms,_=e.timing(); c1=e.counters()
    print(json.dumps(dict(cfg=cfg, chains=nc, gen_init_s=round(dt,3), init_log2=float(np.log2(t).mean()), rate=(c1['proposals']-c0['proposals'])/(ms*1e-3), best=float(np.log2(e.costs()[1]).min()), **e.config())), flush=True)
    e.close()
PY
cat gpurun_out/probe12.log
python scripts/e2e_profile.py C2 4096 10000 > gpurun_out/e2e_c2.log 2>&1; head -3 gpurun_out/e2e_c2.log
ncu --set full --clock-control none --import-source on -k regex:sa_sweep_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_c4c python scripts/_prof.py 0 C4 4096 32 500 > gpurun_out/p_c4c.log 2>&1
