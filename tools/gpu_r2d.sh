#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dequant.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'])
e=d['e2e']; print('e2e', e['ms_per_step'], e['value']); 
for k,v in e['variants'].items(): print('  ', k, round(v['ms_per_step'],3), 'ms', v['h2d_bytes_per_step'], v['d2h_bytes_per_step'])
print('latency', e.get('latency_us_per_call'))
for k,v in d['configs'].items(): print(k, round(v['ms_per_step'],4), round(v['roofline']['frac'],3), v['gpu_launches'])
PY
