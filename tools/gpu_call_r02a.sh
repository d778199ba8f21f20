#!/bin/bash
# round 2, call A: validate the merged-list FPS kernel and time it against the round-1 chain kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 300 -x -k "fps" > gpurun_out/r02a_fps_tests.txt 2>&1; echo "fps tests rc=$?"
tail -15 gpurun_out/r02a_fps_tests.txt
timeout 300 python tools/fps_time.py --json gpurun_out/r02a_fps_time.json 2>&1 | tee gpurun_out/r02a_fps_time.txt | tail -20
