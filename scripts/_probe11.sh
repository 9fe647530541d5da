mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python - <<'PY' > gpurun_out/probe11.log 2>&1
import sys, os, time, json; sys.path.insert(0,'scripts'); sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from gpu_probe import probe
probe('C2', 4096, 2000)
probe('C4', 4096, 2000, max_width=32)
# hyper-index network throughput (200 tensors, 3-regular + 20 hyper indices + 4 open)
from helpers import hyper_network, leaf_bits
from tnco_b200.engine import Engine, random_trees, pack_index_set
ts, ni, out = hyper_network(200, 3, n_hyper=20, n_open=4)
lb = leaf_bits(ts, ni); ob = pack_index_set(out, ni)
seeds = np.arange(4096, dtype=np.uint64)+1
t0=time.time(); p,a,b = random_trees(lb, ni, seeds, output_bits=ob); tg=time.time()-t0
for mw in (None, 24):
    e = Engine(); e.set_network(lb, ni, output_bits=ob).set_mode(max_width=mw); e.set_chains(p,a,b,seeds)
    e.set_betas(np.linspace(0,100,1000,endpoint=False)); t,_=e.costs(); e.timing(); e.run(100); e.timing(); c0=e.counters(); e.run(1000); ms,_=e.timing(); c1=e.counters()
    print(json.dumps(dict(cfg='hyper200', max_width=mw, hyper=e.hyper, tree_gen_s=round(tg,3), ms=round(ms,2), rate=(c1['proposals']-c0['proposals'])/(ms*1e-3), init=float(np.log2(t).mean()), best=float(np.log2(e.costs()[1]).min()), **e.config())), flush=True)
    e.close()
PY
cat gpurun_out/probe11.log
