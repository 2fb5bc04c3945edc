#!/bin/bash
# row_range with one load: parity (TNS cases) + A/B on config 4
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 180 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
run() {  # tag workload [env...]
  tag=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --steps 100 --warmup 3 --no-e2e --no-cpu --no-configs --workload $wl > gpurun_out/ab_$tag.json 2>> gpurun_out/bench.err
  python -c "import json;d=json.load(open('gpurun_out/ab_$tag.json'));print('%-28s %.4f ms  frac %.3f  launches %d' % ('$tag', d['ms_per_step'],d['roofline']['frac'],d['gpu_launches']))"
}
for rep in 1 2; do
run config4_new config4 A=1
run config4_prev config4 AACFB_LIB=$PWD/aac.js_b200/libaacfb_prev.so
done
