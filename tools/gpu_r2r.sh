#!/bin/bash
# tns_kernel: 256-byte tile rows (AACFB_TNS_COLS=64) A/B on config 4 + parity of the variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 180 --timeout-method=thread -k "tns or config or random or golden" > gpurun_out/pytest_gpu_tns.log 2>&1; tail -2 gpurun_out/pytest_gpu_tns.log
run() {  # tag workload [env...]
  tag=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --steps 100 --warmup 3 --no-e2e --no-cpu --no-configs --workload $wl > gpurun_out/ab_$tag.json 2>> gpurun_out/bench.err
  python -c "import json;d=json.load(open('gpurun_out/ab_$tag.json'));print('%-28s %.4f ms  frac %.3f  launches %d' % ('$tag', d['ms_per_step'],d['roofline']['frac'],d['gpu_launches']))"
}
for rep in 1 2; do
run c4_cols32 config4 A=1
run c4_cols64_ring3_cta2 config4 AACFB_LIB=$PWD/aac.js_b200/libaacfb_c64.so
run c4_cols64_ring2_cta3 config4 AACFB_LIB=$PWD/aac.js_b200/libaacfb_c64r2.so
done
for v in c64 c64r2; do AACFB_LIB=$PWD/aac.js_b200/libaacfb_$v.so timeout 600 python -m pytest tests -x -q -m gpu --timeout 180 --timeout-method=thread -k "tns or random or golden" 2>&1 | tail -1; done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tns_kernel -s 3 -c 2 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --workload config4 2>&1 | grep -E "tns_kernel|duration|dram__" | head -8
AACFB_LIB=$PWD/aac.js_b200/libaacfb_c64.so timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tns_kernel -s 3 -c 2 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --workload config4 2>&1 | grep -E "tns_kernel|duration|dram__" | head -8
