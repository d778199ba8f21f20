#!/bin/bash
mkdir -p gpurun_out
B="timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ops"
run() { tag=$1; shift; "$@" 2>>gpurun_out/r02r.err | tee gpurun_out/r02r_$tag.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', round(d['value']/1e6,2), 'M pts/s  e2e', round(d['e2e']['value']/1e6,2), ' ms/room', round(d['ms_per_step'],4), d['config'].get('l2','')[:60])"; }
run rotate $B
run flush $B --l2 flush
run rotate_d8 $B --depth 8
run rotate_d16 $B --depth 16
tail -3 gpurun_out/r02r.err
