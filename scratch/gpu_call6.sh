#!/bin/bash
mkdir -p gpurun_out
show() { python -c "import json,sys;d=json.load(open('$1'));print('$2', round(d['ms_per_step'],3), 'ms', round(d['value']/1e6,2),'Mpts/s e2e', round(d['e2e']['value']/1e6,2))"; }
B="python bench.py --no-cpu-baseline --steps 240"
POINTOPS_B200_GEO_PRIORITY=-1 timeout 200 $B > gpurun_out/c6_prio_d12.json 2>/dev/null; show gpurun_out/c6_prio_d12.json "geo priority -1 d12"
POINTOPS_B200_GEO_PRIORITY=-1 timeout 200 $B --depth 8 > gpurun_out/c6_prio_d8.json 2>/dev/null; show gpurun_out/c6_prio_d8.json "geo priority -1 d8"
POINTOPS_B200_GEO_PRIORITY=-1 POINTOPS_B200_FPS_LAYOUT=tall timeout 200 $B > gpurun_out/c6_prio_tall_d12.json 2>/dev/null; show gpurun_out/c6_prio_tall_d12.json "geo priority -1 tall d12"
timeout 300 python bench.py > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err; echo "bench rc=$?"; show gpurun_out/bench_r01d.json "default (prio 0, wide, auto, d12)"
timeout 240 ncu --set full --import-source on --clock-control none -k regex:linear_tile -o gpurun_out/r01d_linear -f python scratch/linear_ncu.py > gpurun_out/r01d_linear.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r01d_linear.ncu-rep --page raw --csv > gpurun_out/r01d_linear_raw.csv 2>/dev/null; gzip -f gpurun_out/r01d_linear_raw.csv; rm -f gpurun_out/r01d_linear.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --depth 1 > gpurun_out/r01d_launches_bench.log 2>&1; echo "launch list rc=$?"
gzip -f gpurun_out/r01d_launches.csv
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
