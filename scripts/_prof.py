import sys; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
from gpu_probe import probe
import os
tile = int(sys.argv[1]) if len(sys.argv) > 1 else None
cfg = sys.argv[2] if len(sys.argv) > 2 else 'C2'
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
probe(cfg, n, 1000, tile=tile)
