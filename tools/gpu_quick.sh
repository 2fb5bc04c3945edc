#!/bin/bash
# Quick GPU iteration: parity tests, then kernel-only timing of the named workloads.
# usage: gpu_quick.sh [workloads...]   (default: config2 config3 config4 config5)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
for wl in ${@:-config2 config3 config4 config5}; do
  timeout 200 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --workload $wl > gpurun_out/bench_$wl.json 2>> gpurun_out/bench.err
  python -c "import json;d=json.load(open('gpurun_out/bench_$wl.json'));print('$wl %.4f ms  frac %.3f' % (d['ms_per_step'],d['roofline']['frac']))"
done
tail -3 gpurun_out/bench.err
