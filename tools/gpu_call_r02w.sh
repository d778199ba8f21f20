#!/bin/bash
# final single-GPU records: default bench line, reference arm line, launch list of the bench schedule
mkdir -p gpurun_out
timeout 1500 python bench.py 2>gpurun_out/r02w_bench.err | tee gpurun_out/r02w_bench_n1.json | cut -c1-600
timeout 1500 python bench.py --steps 20 --warmup 3 2>>gpurun_out/r02w_bench.err | tee gpurun_out/r02w_bench_n1_steps20.json | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/r02w_launches.csv python bench.py --steps 4 --warmup 3 --depth 1 --no-cpu-baseline --no-ops --windows 1 --min-window-s 0 > gpurun_out/r02w_ncu_bench.log 2>&1
gzip -f gpurun_out/r02w_launches.csv
ls -la gpurun_out/r02w_*
