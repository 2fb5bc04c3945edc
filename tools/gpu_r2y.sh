#!/bin/bash
mkdir -p gpurun_out
run() {  # tag workload [env...]
  tag=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --steps 100 --warmup 3 --no-e2e --no-cpu --no-configs --workload $wl > gpurun_out/ab_$tag.json 2>> gpurun_out/bench.err
  python -c "import json;d=json.load(open('gpurun_out/ab_$tag.json'));print('%-28s %.4f ms  frac %.3f  launches %d' % ('$tag', d['ms_per_step'],d['roofline']['frac'],d['gpu_launches']))"
}
for rep in 1 2; do
for wl in config5 config3; do
run ${wl}_default $wl A=1
run ${wl}_generic_uniform $wl AACFB_LIB=$PWD/aac.js_b200/libaacfb_gu1.so
run ${wl}_no_park $wl AACFB_LIB=$PWD/aac.js_b200/libaacfb_np.so
done
done
