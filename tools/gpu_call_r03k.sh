#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_call_r02w.sh | tail -3
timeout 900 python tools/side_benches.py gpurun_out/r03k_side_benches.json 2>&1 | tail -12 | cut -c1-260
