#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_merge -s 1 -c 1 -o gpurun_out/r02b_fps_merge_80k -f python tools/fps_one.py 80000 merge 2 > gpurun_out/r02b_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02b_ncu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fps_merge -s 1 -c 1 -o gpurun_out/r02b_fps_merge_1250 -f python tools/fps_one.py 1250 merge 2 > gpurun_out/r02b_ncu2.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
