"""ncu target: one probe run.  argv: tile|0 cfg n_chains [max_width|0] [n_sweeps]"""
import sys; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
from gpu_probe import probe
tile = int(sys.argv[1]) if len(sys.argv) > 1 and int(sys.argv[1]) else None
cfg = sys.argv[2] if len(sys.argv) > 2 else 'C2'
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
mw = float(sys.argv[4]) if len(sys.argv) > 4 and float(sys.argv[4]) > 0 else None
sw = int(sys.argv[5]) if len(sys.argv) > 5 else 1000
lay = int(sys.argv[6]) if len(sys.argv) > 6 else 0
probe(cfg, n, sw, tile=tile, max_width=mw, layout=lay)
