#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/fps_order_experiment.py 2>&1 | tee gpurun_out/r02y_fps_order.txt | tail -30
