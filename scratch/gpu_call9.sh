#!/bin/bash
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -q --tb=short --timeout 180 > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.txt
timeout 300 python bench.py > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err; echo "bench rc=$?"
python scratch/show_bench.py gpurun_out/bench_r01d.json | cut -c1-400
timeout 300 python bench.py --impl reference --steps 20 --warmup 1 > gpurun_out/bench_r01d_reference.json 2>/dev/null; cut -c1-260 gpurun_out/bench_r01d_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --depth 1 > gpurun_out/r01d_launches_bench.log 2>&1; echo "launch list rc=$?"
gzip -f gpurun_out/r01d_launches.csv
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
