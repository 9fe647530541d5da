# 8-GPU torchrun runs only (C2 and the north-star config C4); weak scaling, 4096 chains per GPU
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_8gpu_c2.json 2> gpurun_out/bench_8gpu_c2.err
$TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --workload C4 --sweeps 4000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_8gpu_c4.json 2> gpurun_out/bench_8gpu_c4.err
for f in bench_8gpu_c2 bench_8gpu_c4; do python -c "
import json,sys
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f', 'value %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], 'ms/step %.1f'%d['ms_per_step'], d['e2e'].get('last_step_ms'), d['clocks'])
"; done
