#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 300 -k "score or pseudo or auroc or msp or pdf" > gpurun_out/r02v_tests.txt 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02v_tests.txt
timeout 1200 python tools/side_benches.py gpurun_out/r02v_side_benches.json 2>&1 | tail -25
timeout 900 python tools/ops_vs_reference.py gpurun_out/r02v_reference_kernel_times.json 2>&1 | tail -26
