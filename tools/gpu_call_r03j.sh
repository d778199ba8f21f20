#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/fps_time.py --sizes 80000,20000,5000,150000,250000 --variants merge 2>&1 | tail -6
