#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ops_vs_reference.py gpurun_out/r02_reference_kernel_times.json 2>&1 | tee gpurun_out/r02i_ops_vs_reference.txt | tail -30
timeout 900 python tools/side_benches.py gpurun_out/r02_side_benches.json 2>&1 | tee gpurun_out/r02i_side.txt | tail -16
timeout 300 python tools/fps_time.py --json gpurun_out/r02i_fps_time.json --variants auto,chain,single --sizes 80000,20000,5000,1250 --reps 5 2>&1 | tee gpurun_out/r02i_fps_time.txt | tail -14
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/r02i_launches.csv python bench.py --steps 4 --warmup 3 --depth 1 --no-cpu-baseline --no-ops --windows 1 --min-window-s 0 > gpurun_out/r02i_launches_bench.log 2>&1; echo "launch list rc=$?"
gzip -f gpurun_out/r02i_launches.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"subtraction_fwd_fast|subtraction_bwd_fast|group_xyz_fwd_fast|score_fused" -c 4 -o gpurun_out/r02i_ops2 -f python tools/ops_one.py 1 > gpurun_out/r02i_ops2_ncu.log 2>&1; echo "ops2 ncu rc=$?"
