#!/bin/bash
# Regenerates the raw material of profiles/ on a B200 box (run through gpurun; ~6 GPU-minutes).
set -x
mkdir -p gpurun_out
TAG=${1:-r01c}
ncu --set full --import-source on --clock-control none -k regex:pt_layer_tile -o gpurun_out/${TAG}_ptlayer -f python scratch/ptl_ncu.py > gpurun_out/${TAG}_ptlayer.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fps_chain -o gpurun_out/${TAG}_fps -f python scratch/fps_ncu.py > gpurun_out/${TAG}_fps.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:knn_grid_kernel --launch-skip 3 -o gpurun_out/${TAG}_knn -f python scratch/knn_ncu.py > gpurun_out/${TAG}_knn.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --depth 1 > gpurun_out/${TAG}_launches_bench.log 2>&1
gzip -f gpurun_out/${TAG}_launches.csv
python scratch/ptl_time.py > gpurun_out/${TAG}_ptl_time.txt 2>&1
python scratch/fps_time.py > gpurun_out/${TAG}_fps_time.txt 2>&1
python scratch/ops_time.py > gpurun_out/${TAG}_ops_time.txt 2>&1
