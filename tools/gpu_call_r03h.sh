#!/bin/bash
mkdir -p gpurun_out
for bits in 10 8 7 6 5 4; do
  echo "== curve bits $bits"
  POB_FPS_CURVE_BITS=$bits timeout 300 python tools/fps_time.py --sizes 80000,20000,5000 --variants merge 2>&1 | tail -3
done
for bits in 10 8 5; do
  POB_FPS_CURVE_BITS=$bits timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ops 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench bits $bits', round(d['value']/1e6,2), 'M pts/s', round(d['ms_per_step'],4))"
done
