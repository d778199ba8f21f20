#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_reference_ext.py -m gpu -q --tb=short --timeout 600 -k "fps or farthest or sampling" > gpurun_out/r02z_tests.txt 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02z_tests.txt
timeout 600 python tools/fps_time.py --sizes 80000,500000,1000000 --variants merge 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ops 2>gpurun_out/r02z_bench.err | tee gpurun_out/r02z_bench.json | cut -c1-260
