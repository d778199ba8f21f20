#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q --tb=short --timeout 300 -x -k "fps or score or pseudo" > gpurun_out/r02e_fps_tests.txt 2>&1; echo "fps+score tests rc=$?"
tail -8 gpurun_out/r02e_fps_tests.txt
timeout 300 python tools/fps_time.py --json gpurun_out/r02e_fps_time.json --variants auto,chain --sizes 80000,20000,5000,1250,150000,250000,500000,1000000 --reps 3 2>&1 | tee gpurun_out/r02e_fps_time.txt | tail -20
timeout 600 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r02e_bench20.json 2> gpurun_out/r02e_bench20.err; echo "bench rc=$?"; tail -3 gpurun_out/r02e_bench20.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02e_bench20.json'))
print('value', d['value']/1e6, 'e2e', d['e2e']['value']/1e6, 'ms/step', d['ms_per_step'], 'windows', d['windows'])
print('roofline', json.dumps(d['roofline'])[:900])
for k,v in list(d['kernels'].items())[:8]: print(k, round(v['ms_per_step'],3), round(v['share_of_step'],3))
print(json.dumps(d.get('ops_cfg1'))[:3000])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gather_rows_fast|group_xyz_fwd_fast" -s 3 -c 3 -o gpurun_out/r02e_ops -f python tools/ops_one.py 2 > gpurun_out/r02e_ops_ncu.log 2>&1; echo "ops ncu rc=$?"
