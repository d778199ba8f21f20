#!/bin/bash
mkdir -p gpurun_out
timeout 400 python scratch/linear_time.py > gpurun_out/linear_time2.txt 2>&1; echo "linear_time rc=$?"
tail -26 gpurun_out/linear_time2.txt | cut -c1-400
timeout 1000 python -m pytest tests -m gpu -q --tb=short --timeout 180 > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.txt
show() { python -c "import json,sys;d=json.load(open('$1'));print('$2', round(d['ms_per_step'],3), 'ms', round(d['value']/1e6,2),'Mpts/s e2e', round(d['e2e']['value']/1e6,2))"; }
B="python bench.py --no-cpu-baseline --steps 200"
POINTOPS_B200_FPS_POINTS=reg timeout 200 $B --linear cublas > gpurun_out/c3_cublas_reg.json 2>/dev/null; show gpurun_out/c3_cublas_reg.json "cublas + FPS reg "
POINTOPS_B200_FPS_POINTS=smem timeout 200 $B --linear cublas > gpurun_out/c3_cublas_smem.json 2>/dev/null; show gpurun_out/c3_cublas_smem.json "cublas + FPS smem"
POINTOPS_B200_FPS_POINTS=smem timeout 200 $B --linear cublas --depth 12 > gpurun_out/c3_cublas_smem_d12.json 2>/dev/null; show gpurun_out/c3_cublas_smem_d12.json "cublas + FPS smem d12"
POINTOPS_B200_FPS_POINTS=reg timeout 200 $B --linear pob > gpurun_out/c3_pob_reg.json 2>/dev/null; show gpurun_out/c3_pob_reg.json "pob + FPS reg    "
POINTOPS_B200_FPS_POINTS=smem timeout 200 $B --linear pob > gpurun_out/c3_pob_smem.json 2>/dev/null; show gpurun_out/c3_pob_smem.json "pob + FPS smem   "
POINTOPS_B200_FPS_POINTS=reg python scratch/fps_time.py 2>&1 | grep -E '^ *80000|80000' | tail -4; POINTOPS_B200_FPS_POINTS=smem python scratch/fps_time.py 2>&1 | grep -E '80000' | tail -4
