#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --linear cublas --no-cpu-baseline --steps 200"
for d in 4 8 12; do
  timeout 200 $B --depth $d > gpurun_out/d_fps_depth$d.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/d_fps_depth$d.json'));print('FPS on  depth $d', round(d['ms_per_step'],3), 'ms', round(d['value']/1e6,2),'Mpts/s')"
done
for d in 1 2 4 8; do
  POINTOPS_B200_DIAG_SKIP_FPS=1 timeout 200 $B --depth $d > gpurun_out/d_nofps_depth$d.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/d_nofps_depth$d.json'));print('FPS off depth $d', round(d['ms_per_step'],3), 'ms', round(d['value']/1e6,2),'Mpts/s')"
done
timeout 200 python bench.py --linear cublas --no-cpu-baseline --steps 200 --depth 1 > gpurun_out/d_fps_depth1.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/d_fps_depth1.json'));print('FPS on  depth 1', round(d['ms_per_step'],3), 'ms')"
