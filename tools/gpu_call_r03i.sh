#!/bin/bash
# ncu --set full of the final FPS kernel (Hilbert-ordered input), source page for the stall breakdown
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_merge -s 1 -c 1 -o gpurun_out/r03_fps_merge_80k -f python tools/fps_one.py 80000 merge 2 > gpurun_out/r03i_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r03i_ncu.log
ncu -i gpurun_out/r03_fps_merge_80k.ncu-rep --page source --csv > gpurun_out/r03_fps_merge_source.csv 2>/dev/null
ncu -i gpurun_out/r03_fps_merge_80k.ncu-rep --page raw --csv > gpurun_out/r03_fps_merge_raw.csv 2>/dev/null
ls -la gpurun_out/r03_fps_merge_*
