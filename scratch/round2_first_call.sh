#!/bin/bash
# First GPU call of the next round: validate and time the experimental 32-group FPS layout (compiled in round 1,
# never run on a GPU), then the bench with it.   gpurun --timeout 900 -- 'bash scratch/round2_first_call.sh'
mkdir -p gpurun_out
POINTOPS_B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 180 -k "layouts_same" > gpurun_out/r02_fps_fine_tests.txt 2>&1; echo "fine-layout tests rc=$?"
tail -5 gpurun_out/r02_fps_fine_tests.txt
timeout 120 python scratch/fps_layout_time.py --fine 2>&1 | tee gpurun_out/r02_fps_layout_time.txt | tail -14
show() { python -c "import json,sys;d=json.load(open('$1'));print('$2', round(d['ms_per_step'],3), 'ms', round(d['value']/1e6,2),'Mpts/s e2e', round(d['e2e']['value']/1e6,2))"; }
B="python bench.py --no-cpu-baseline"
timeout 200 $B > gpurun_out/r02_wide.json 2>/dev/null; show gpurun_out/r02_wide.json "wide (default)"
POINTOPS_B200_FPS_LAYOUT=fine timeout 200 $B > gpurun_out/r02_fine.json 2>/dev/null; show gpurun_out/r02_fine.json "fine (32 groups)"
