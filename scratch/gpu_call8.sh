#!/bin/bash
mkdir -p gpurun_out
show() { python -c "import json,sys;d=json.load(open('$1'));print('$2', round(d['ms_per_step'],3), 'ms', round(d['value']/1e6,2),'Mpts/s e2e', round(d['e2e']['value']/1e6,2))"; }
B="python bench.py --no-cpu-baseline"
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 200 $B --steps 240 > gpurun_out/c8_conn32.json 2>/dev/null; show gpurun_out/c8_conn32.json "connections 32 d12"
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 200 $B --steps 240 --depth 16 > gpurun_out/c8_conn32_d16.json 2>/dev/null; show gpurun_out/c8_conn32_d16.json "connections 32 d16"
timeout 200 $B --steps 240 > gpurun_out/c8_base.json 2>/dev/null; show gpurun_out/c8_base.json "default d12 K=240"
timeout 200 $B --steps 800 > gpurun_out/c8_k800.json 2>/dev/null; show gpurun_out/c8_k800.json "default d12 K=800"
