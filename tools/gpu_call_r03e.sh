#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_ext.py -m gpu -q --tb=short --timeout 600 -k "fps or farthest or sampling" > gpurun_out/r03e_tests.txt 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r03e_tests.txt
timeout 600 python tools/fps_time.py --sizes 312,1250,2048 --variants auto,chain 2>&1 | tail -6
