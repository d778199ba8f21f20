#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scratch/fps_layout_time.py 2>&1 | tee gpurun_out/fps_layout_time.txt | tail -12
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ptv1.py tests/test_gpu_reference_ext.py -m gpu -q --tb=short --timeout 180 -k "fps or ptv1 or infer or golden" > gpurun_out/pytest_fps.txt 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_fps.txt
show() { python -c "import json,sys;d=json.load(open('$1'));print('$2', round(d['ms_per_step'],3), 'ms', round(d['value']/1e6,2),'Mpts/s e2e', round(d['e2e']['value']/1e6,2))"; }
B="python bench.py --no-cpu-baseline --steps 240"
POINTOPS_B200_FPS_LAYOUT=tall timeout 200 $B --depth 12 > gpurun_out/c5_tall_d12.json 2>/dev/null; show gpurun_out/c5_tall_d12.json "tall d12"
POINTOPS_B200_FPS_LAYOUT=tall timeout 200 $B --depth 16 > gpurun_out/c5_tall_d16.json 2>/dev/null; show gpurun_out/c5_tall_d16.json "tall d16"
POINTOPS_B200_FPS_LAYOUT=tall timeout 200 $B --depth 8 > gpurun_out/c5_tall_d8.json 2>/dev/null; show gpurun_out/c5_tall_d8.json "tall d8"
POINTOPS_B200_FPS_LAYOUT=wide timeout 200 $B --depth 12 > gpurun_out/c5_wide_d12.json 2>/dev/null; show gpurun_out/c5_wide_d12.json "wide d12"
