#!/bin/bash
# 8 GPUs: bench at N=8 with cfg3 / cfg5 entries (and the N=1 line of the same box for the efficiency denominator)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02h_bench_n8.json 2> gpurun_out/r02h_bench_n8.err; echo "bench n8 rc=$?"; tail -5 gpurun_out/r02h_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 4 --steps 20 --warmup 3 --no-ops > gpurun_out/r02h_bench_n4.json 2> gpurun_out/r02h_bench_n4.err; echo "bench n4 rc=$?"
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ops > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err; echo "bench n1 rc=$?"
python - <<'PY'
import json
for n in (1, 4, 8):
    try:
        d=json.load(open(f'gpurun_out/r02h_bench_n{n}.json'))
    except Exception as e:
        print(n, 'failed', e); continue
    print('N', n, 'value', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), 'ms/step', round(d['ms_per_step'],4), 'spread', d['windows']['spread_rel_max_over_ranks'])
    if 'cfg3_training' in d: print(json.dumps(d['cfg3_training'])[:900])
    if 'cfg5_sharded_knn' in d:
        for r in d['cfg5_sharded_knn'].get('rows', []): print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items()})
PY
