#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_ptv1.py tests/test_gpu_dropin.py -m gpu -q --tb=short --timeout 300 > gpurun_out/r02o_tests.txt 2>&1; echo "tests rc=$?"
tail -12 gpurun_out/r02o_tests.txt
timeout 600 python tools/linear_time.py 2>&1 | tee gpurun_out/r02o_linear_time.txt | tail -26
