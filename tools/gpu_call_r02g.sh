#!/bin/bash
# 2 GPUs: multi-GPU tests (all-gather and fused P2P kNN), bench at N=2 with cfg3 / cfg5 entries
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short --timeout 600 -x > gpurun_out/r02g_multi_tests.txt 2>&1; echo "multi tests rc=$?"
tail -25 gpurun_out/r02g_multi_tests.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02g_bench_n2.json 2> gpurun_out/r02g_bench_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/r02g_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02g_bench_n2.json'))
print('value', d['value']/1e6, 'e2e', d['e2e']['value']/1e6, 'ms/step', d['ms_per_step'], d['windows']['spread_rel_max_over_ranks'])
print(json.dumps(d.get('cfg3_training'), indent=0)[:1500])
print(json.dumps(d.get('cfg5_sharded_knn'), indent=0)[:3000])
PY
