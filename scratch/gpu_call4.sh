#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
from pointcloudpdf_b200 import _lib
import torch
torch.cuda.init()
lib = _lib.load()
for P in (20, 12):
    for C in (16, 8, 4):
        for sp in (0, 1):
            print(f"max active clusters P={P} C={C} smem_points={sp}:", lib.pob_fps_max_active_clusters(P, C, sp))
PY
show() { python -c "import json,sys;d=json.load(open('$1'));print('$2', round(d['ms_per_step'],3), 'ms', round(d['value']/1e6,2),'Mpts/s e2e', round(d['e2e']['value']/1e6,2))"; }
B="python bench.py --no-cpu-baseline --steps 240"
POINTOPS_B200_FPS_POINTS=smem timeout 200 $B --linear cublas --depth 16 > gpurun_out/c4_smem_d16.json 2>/dev/null; show gpurun_out/c4_smem_d16.json "cublas smem d16"
POINTOPS_B200_FPS_POINTS=smem timeout 200 $B --linear cublas --depth 24 > gpurun_out/c4_smem_d24.json 2>/dev/null; show gpurun_out/c4_smem_d24.json "cublas smem d24"
POINTOPS_B200_FPS_POINTS=reg timeout 200 $B --linear cublas --depth 16 > gpurun_out/c4_reg_d16.json 2>/dev/null; show gpurun_out/c4_reg_d16.json "cublas reg d16"
POINTOPS_B200_FPS_POINTS=reg timeout 200 $B --linear auto --depth 12 > gpurun_out/c4_auto_reg_d12.json 2>/dev/null; show gpurun_out/c4_auto_reg_d12.json "auto reg d12"
POINTOPS_B200_FPS_POINTS=reg timeout 200 $B --linear cublas --depth 12 > gpurun_out/c4_cublas_reg_d12.json 2>/dev/null; show gpurun_out/c4_cublas_reg_d12.json "cublas reg d12"
