import sys; sys.path.insert(0, 'scripts'); sys.path.insert(0, '.')
from gpu_probe import probe
for nc in (1024, 2048, 3072, 4096, 8192):
    probe('C5', nc, 300)
for nc in (1024, 2048, 4096):
    probe('C4', nc, 500, max_width=32)
