#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_mma -s 1 -c 1 -o gpurun_out/r03_linear_mma_80k_32_96 -f python tools/linear_one.py 80000 32 96 0 > gpurun_out/r03l_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r03_linear_mma_80k_32_96.ncu-rep --page raw --csv > gpurun_out/r03_linear_mma_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:linear_mma -s 1 -c 1 -o gpurun_out/r03_linear_mma_20k_64_192 -f python tools/linear_one.py 20000 64 192 0 > gpurun_out/r03l_ncu2.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r03_linear_mma_20k_64_192.ncu-rep --page raw --csv > gpurun_out/r03_linear_mma_raw2.csv 2>/dev/null
ls -la gpurun_out/r03_linear_mma*
