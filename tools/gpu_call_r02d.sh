#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 300 -x -k "fps" > gpurun_out/r02d_fps_tests.txt 2>&1; echo "fps tests rc=$?"
tail -4 gpurun_out/r02d_fps_tests.txt
timeout 300 python tools/fps_time.py --json gpurun_out/r02d_fps_time.json --variants merge 2>&1 | tee gpurun_out/r02d_fps_time.txt | tail -8
timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -q --tb=short --timeout 600 > gpurun_out/r02d_dropin_tests.txt 2>&1; echo "dropin tests rc=$?"
tail -12 gpurun_out/r02d_dropin_tests.txt
timeout 1200 python -m pytest tests/test_gpu_scale.py -m gpu -q --tb=short --timeout 900 > gpurun_out/r02d_scale_tests.txt 2>&1; echo "scale tests rc=$?"
tail -12 gpurun_out/r02d_scale_tests.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_merge -s 1 -c 1 -o gpurun_out/r02d_fps_merge_80k -f python tools/fps_one.py 80000 merge 2 > gpurun_out/r02d_ncu.log 2>&1; echo "ncu rc=$?"
