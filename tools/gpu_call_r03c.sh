#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/fps_time.py --sizes 1250,2500,5000,10000,20000 --variants merge,chain --clusters 1,2,4,8,16 2>&1 | tee gpurun_out/r03c_fps_small.txt | tail -52
