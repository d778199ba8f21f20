#!/bin/bash
# usage: gpu_call_r02x.sh N   -- the driver's multi-GPU launch of bench.py on N GPUs
N=$1
mkdir -p gpurun_out
timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/r02x_n$N.err | tee gpurun_out/r02x_bench_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N', d['n_gpus'], 'value', d['value']/1e6, 'e2e', d['e2e']['value']/1e6)
for k in ('cfg3_training','cfg3_training_amp'):
    print(k, json.dumps({a:b for a,b in d.get(k,{}).items() if a not in ('what','allreduce','precision')}))
print(json.dumps(d.get('cfg5_sharded_knn'))[:1500])
"
tail -2 gpurun_out/r02x_n$N.err
