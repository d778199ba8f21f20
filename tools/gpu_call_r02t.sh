#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ptv1.py -m gpu -q --tb=short --timeout 300 > gpurun_out/r02t_tests.txt 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02t_tests.txt
timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/r02t_ref.err | tee gpurun_out/r02t_reference.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps(d.get('gpu_reference_before'),indent=1)[:2500]); print('cpu', d['value'])"
tail -3 gpurun_out/r02t_ref.err
