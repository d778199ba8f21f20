#!/bin/bash
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_new_kernels.py > gpurun_out/r03m_memcheck.txt 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|linear|fps|done" gpurun_out/r03m_memcheck.txt | tail -16
