#!/bin/bash
# staggered short_load (odd windows read their row pairs in the opposite order): parity + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 180 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
run() {  # tag workload [env...]
  tag=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --steps 100 --warmup 3 --no-e2e --no-cpu --no-configs --workload $wl > gpurun_out/ab_$tag.json 2>> gpurun_out/bench.err
  python -c "import json;d=json.load(open('gpurun_out/ab_$tag.json'));print('%-28s %.4f ms  frac %.3f  launches %d' % ('$tag', d['ms_per_step'],d['roofline']['frac'],d['gpu_launches']))"
}
for rep in 1 2; do
run c3_stagger config3 A=1
run c3_plain config3 AACFB_LIB=$PWD/aac.js_b200/libaacfb_stag0.so
run c5_stagger config5 A=1
run c5_plain config5 AACFB_LIB=$PWD/aac.js_b200/libaacfb_stag0.so
done
