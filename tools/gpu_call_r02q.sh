#!/bin/bash
mkdir -p gpurun_out
B="timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ops"
run() { tag=$1; shift; "$@" 2>>gpurun_out/r02q.err | tee gpurun_out/r02q_$tag.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', round(d['value']/1e6,2), 'M pts/s  e2e', round(d['e2e']['value']/1e6,2), ' ms/room', round(d['ms_per_step'],4))"; }
run auto $B
run cublas $B --linear cublas
run pob $B --linear pob
POINTOPS_B200_FPS_CLUSTER=8 run c8_d12 $B
POINTOPS_B200_FPS_CLUSTER=8 run c8_d16 $B --depth 16
run auto_d16 $B --depth 16
