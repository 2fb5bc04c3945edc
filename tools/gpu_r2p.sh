#!/bin/bash
# captures half b + tns_kernel L2 prefetch A/B
bash tools/gpu_capture.sh b
run() {  # tag workload [env...]
  tag=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --steps 100 --warmup 3 --no-e2e --no-cpu --no-configs --workload $wl > gpurun_out/ab_$tag.json 2>> gpurun_out/bench.err
  python -c "import json;d=json.load(open('gpurun_out/ab_$tag.json'));print('%-28s %.4f ms  frac %.3f  launches %d' % ('$tag', d['ms_per_step'],d['roofline']['frac'],d['gpu_launches']))"
}
for rep in 1 2; do
run c4_pf0 config4 A=1
run c4_pf4 config4 AACFB_LIB=$PWD/aac.js_b200/libaacfb_pf4.so
run c4_pf8 config4 AACFB_LIB=$PWD/aac.js_b200/libaacfb_pf8.so
done
