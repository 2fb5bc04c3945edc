#!/bin/bash
# Compare kernel variants (workers/stages) and chunk lengths (kernel-only timing).
mkdir -p gpurun_out
for lib in ${LIBS:-aac.js_b200/libaacfb.so}; do
  for chunk in ${CHUNKS:-0}; do
    for wl in ${WLS:-config2}; do
      ms=$(AACFB_LIB=$PWD/$lib AACFB_CHUNK_LEN=$chunk timeout 120 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --workload $wl 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4f ms  frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))")
      echo "$lib chunk=$chunk $wl: $ms"
    done
  done
done | tee gpurun_out/tune.log
