#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no-ops 2>gpurun_out/r02u.err | tee gpurun_out/r02u_bench_n2.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value']/1e6, 'e2e', d['e2e']['value']/1e6)
for k in ('cfg3_training','cfg3_training_amp'):
    print(k, json.dumps({a:b for a,b in d.get(k,{}).items() if a not in ('what','allreduce')}))
"
tail -3 gpurun_out/r02u.err
