#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_datapath.py -m gpu -q --tb=short --timeout 600 -x -k "fps or sphere" > gpurun_out/r02l_tests.txt 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/r02l_tests.txt
timeout 300 python tools/fps_time.py --json gpurun_out/r02l_fps_time.json --variants auto --sizes 131072,150000,190000,250000 --reps 3 2>&1 | tail -6
timeout 900 python tools/side_benches.py gpurun_out/r02_side_benches.json 2>&1 | grep "cfg4\|points_per_room\|ns_per_sample" | cut -c1-250
