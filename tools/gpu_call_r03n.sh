#!/bin/bash
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_new_kernels.py > gpurun_out/r03n_racecheck.txt 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|hazard|Race reported|ERROR|done" gpurun_out/r03n_racecheck.txt | sort | uniq -c | sort -rn | head -12
grep -B2 -A12 "Race reported" gpurun_out/r03n_racecheck.txt | grep -E "Race reported|pob::|\.cu" | head -20
