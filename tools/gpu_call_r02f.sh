#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short --timeout 600 -x > gpurun_out/r02f_all_tests.txt 2>&1; echo "all gpu tests rc=$?"
tail -8 gpurun_out/r02f_all_tests.txt
timeout 300 python tools/fps_time.py --json gpurun_out/r02f_fps_time.json --variants auto --sizes 80000,20000,5000,1250,250000 --reps 3 2>&1 | tee gpurun_out/r02f_fps_time.txt | tail -8
timeout 600 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r02f_bench20.json 2> gpurun_out/r02f_bench20.err; echo "bench rc=$?"; tail -3 gpurun_out/r02f_bench20.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench20.json'))
print('value', d['value']/1e6, 'e2e', d['e2e']['value']/1e6, 'ms/step', d['ms_per_step'], 'windows', d['windows'])
for k,v in list(d['kernels'].items())[:6]: print(k, round(v['ms_per_step'],3), round(v['share_of_step'],3))
for k,v in d['ops_cfg1']['ops'].items(): print(k, {kk: round(vv,3) for kk,vv in v.items() if isinstance(vv,(int,float))})
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/r02f_bench_reference.json 2> gpurun_out/r02f_bench_reference.err; echo "ref arm rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02f_bench_reference.json')); print(d['value'], d['cpu_baseline']['cores'], json.dumps(d['gpu_reference_before'])[:1200])"
