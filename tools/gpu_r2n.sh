#!/bin/bash
# round-2 session-3 run 2: continuous prefetch across runs; config-5 launch list; config 2 against the previous build
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
run c2_new config2 A=1
run c2_head config2 AACFB_LIB=$PWD/aac.js_b200/libaacfb_head.so
run c5_head config5 AACFB_LIB=$PWD/aac.js_b200/libaacfb_head.so
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_config5.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-configs --workload config5 > gpurun_out/ncu_list5.log 2>&1
grep synth_kernel gpurun_out/launches_config5.csv | tail -6 | cut -d, -f5,12- 
tail -3 gpurun_out/bench.err
