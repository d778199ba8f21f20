#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_datapath.py tests/test_gpu_fused.py tests/test_gpu_linear.py -m gpu -q --tb=short --timeout 300 -x > gpurun_out/r02k_tests.txt 2>&1; echo "tests rc=$?"
tail -12 gpurun_out/r02k_tests.txt
timeout 600 python tools/datapath_time.py gpurun_out/r02_datapath_time.json 2>&1 | tail -10
