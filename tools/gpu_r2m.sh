#!/bin/bash
# round-2 session-3 run 1: tests, run splitting on/off, long pair-load variant
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 180 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
run() {  # tag workload [env...]
  tag=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --steps 100 --warmup 3 --no-e2e --no-cpu --no-configs --workload $wl > gpurun_out/ab_$tag.json 2>> gpurun_out/bench.err
  python -c "import json;d=json.load(open('gpurun_out/ab_$tag.json'));print('%-28s %.4f ms  frac %.3f  launches %d' % ('$tag', d['ms_per_step'],d['roofline']['frac'],d['gpu_launches']))"
}
for rep in 1 2; do
run c5_split1 config5 AACFB_SPLIT_RUNS=1
run c5_split0 config5 AACFB_SPLIT_RUNS=0
run c3_split1 config3 AACFB_SPLIT_RUNS=1
run c3_split0 config3 AACFB_SPLIT_RUNS=0
run c2_default config2 A=1
run c2_lpl config2 AACFB_LIB=$PWD/aac.js_b200/libaacfb_lpl.so
run c5_lpl config5 AACFB_LIB=$PWD/aac.js_b200/libaacfb_lpl.so
run c4_default config4 A=1
run c4_lpl config4 AACFB_LIB=$PWD/aac.js_b200/libaacfb_lpl.so
done
tail -3 gpurun_out/bench.err
