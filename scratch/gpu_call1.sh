#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
timeout 300 python scratch/linear_time.py > gpurun_out/linear_time.txt 2>&1; echo "linear_time rc=$?"
tail -26 gpurun_out/linear_time.txt
timeout 1000 python -m pytest tests -m gpu -q --tb=short --timeout 180 > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.txt
timeout 400 python bench.py > gpurun_out/bench_pob.json 2> gpurun_out/bench_pob.err; echo "bench pob rc=$?"
timeout 300 python bench.py --linear cublas --no-cpu-baseline > gpurun_out/bench_cublas.json 2> gpurun_out/bench_cublas.err; echo "bench cublas rc=$?"
python scratch/show_bench.py gpurun_out/bench_pob.json 2>&1 | tail -24; python scratch/show_bench.py gpurun_out/bench_cublas.json 2>&1 | head -2
timeout 240 ncu --set full --import-source on --clock-control none -k regex:linear_tile -o gpurun_out/r01d_linear -f python scratch/linear_ncu.py > gpurun_out/r01d_linear.log 2>&1; echo "ncu rc=$?"
timeout 300 python scratch/side_benches.py gpurun_out/r01d_side_benches.json > gpurun_out/side.log 2>&1; echo "side rc=$?"
tail -5 gpurun_out/side.log
