#!/bin/bash
mkdir -p gpurun_out
B="timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ops"
run() { tag=$1; shift; "$@" 2>>gpurun_out/r03b.err | tee gpurun_out/r03b_$tag.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', round(d['value']/1e6,2), 'M pts/s  e2e', round(d['e2e']['value']/1e6,2), ' ms/room', round(d['ms_per_step'],4))"; }
run base $B
POB_EXPERIMENT_HILBERT=1 run hilbert_input $B
tail -3 gpurun_out/r03b.err
