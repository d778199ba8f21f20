#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pseudo.py tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 300 -x -k "pseudo or ball or region or group" > gpurun_out/r02j_tests.txt 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r02j_tests.txt
timeout 600 python tools/ops_vs_reference.py gpurun_out/r02_reference_kernel_times.json 2>&1 | tee gpurun_out/r02j_ops_vs_reference.txt | grep "with_xyz"
