#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --tb=short --timeout 900 -x > gpurun_out/r02m_all_tests.txt 2>&1; echo "all gpu tests rc=$?"
tail -6 gpurun_out/r02m_all_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02m_bench_n1.json 2> gpurun_out/r02m_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r02m_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02m_bench_n1.json'))
print('value', d['value']/1e6, 'e2e', d['e2e']['value']/1e6, 'ms/step', d['ms_per_step'], 'spread', d['windows']['spread_rel_max_over_ranks'])
print('cpu', d.get('cpu_baseline'))
print('roofline', {k:v for k,v in d['roofline'].items() if k in ('kernel','achieved','peak','frac','executed_TFLOPs','samples_per_exchange','share_of_step','avg_launch_ms')})
print('clocks', d['clocks'], 'launches', d['gpu_launches'])
PY
