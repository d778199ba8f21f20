#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q --tb=short --timeout 900 > gpurun_out/r03g_gpu_tests.txt 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r03g_gpu_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
