#!/bin/bash
# tns_kernel: ring depth x resident CTAs A/B (config 4) + parity of the parametrised tile loop
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 180 --timeout-method=thread -k "tns or config or random or golden" > gpurun_out/pytest_gpu_tns.log 2>&1; tail -3 gpurun_out/pytest_gpu_tns.log
run() {  # tag workload [env...]
  tag=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --steps 100 --warmup 3 --no-e2e --no-cpu --no-configs --workload $wl > gpurun_out/ab_$tag.json 2>> gpurun_out/bench.err
  python -c "import json;d=json.load(open('gpurun_out/ab_$tag.json'));print('%-28s %.4f ms  frac %.3f  launches %d' % ('$tag', d['ms_per_step'],d['roofline']['frac'],d['gpu_launches']))"
}
for rep in 1 2; do
run c4_ring3_cta4 config4 A=1
run c4_ring4_cta3 config4 AACFB_LIB=$PWD/aac.js_b200/libaacfb_t34.so
run c4_ring3_cta3 config4 AACFB_LIB=$PWD/aac.js_b200/libaacfb_t33.so
run c4_ring5_cta2 config4 AACFB_LIB=$PWD/aac.js_b200/libaacfb_t52.so
done
for v in t34 t33; do AACFB_LIB=$PWD/aac.js_b200/libaacfb_$v.so timeout 600 python -m pytest tests -x -q -m gpu --timeout 180 --timeout-method=thread -k "tns" 2>&1 | tail -1; done
