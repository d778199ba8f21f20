#!/bin/bash
mkdir -p gpurun_out
for d in 6 8 16 24; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ops --depth $d > gpurun_out/r02n_depth$d.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r02n_depth$d.json')); print('depth $d value', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), 'ms', round(d['ms_per_step'],4))"
done
POINTOPS_B200_FPS=chain timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ops > gpurun_out/r02n_chain.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02n_chain.json')); print('round-1 FPS kernel (variant chain): value', round(d['value']/1e6,2), 'ms', round(d['ms_per_step'],4), d['config']['env'])"
