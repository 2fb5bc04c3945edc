#!/bin/bash
# scratch intervals through the async side-info ring: parity + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 180 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
AACFB_TNS_FUSED=1 timeout 600 python -m pytest tests -x -q -m gpu --timeout 180 --timeout-method=thread -k "tns or stereo" 2>&1 | tail -1
run() {  # tag workload [env...]
  tag=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --steps 100 --warmup 3 --no-e2e --no-cpu --no-configs --workload $wl > gpurun_out/ab_$tag.json 2>> gpurun_out/bench.err
  python -c "import json;d=json.load(open('gpurun_out/ab_$tag.json'));print('%-28s %.4f ms  frac %.3f  launches %d' % ('$tag', d['ms_per_step'],d['roofline']['frac'],d['gpu_launches']))"
}
for rep in 1 2; do
for wl in config4 config2 config5; do
run ${wl}_new $wl A=1
run ${wl}_prev $wl AACFB_LIB=$PWD/aac.js_b200/libaacfb_prev.so
done
done
