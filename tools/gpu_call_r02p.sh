#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_ptv1.py tests/test_gpu_dropin.py -m gpu -q --tb=short --timeout 300 > gpurun_out/r02p_tests.txt 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r02p_tests.txt
timeout 600 python tools/linear_time.py 2>&1 | tee gpurun_out/r02p_linear_time.txt | tail -24
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ops 2>gpurun_out/r02p_bench.err | tee gpurun_out/r02p_bench.json | cut -c1-400
POINTOPS_B200_LINEAR=cublas timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ops 2>>gpurun_out/r02p_bench.err | tee gpurun_out/r02p_bench_cublas.json | cut -c1-300
