#!/bin/bash
# A/B on one box: kernel-only timing of the workloads with the in-tree library and with variant builds.
# usage: ab.sh "<lib1> <lib2> ..." [workloads...]     (libs relative to aac.js_b200/, "-" = the default build)
mkdir -p gpurun_out
libs=$1; shift
for rep in 1 2; do
for wl in ${@:-config2}; do
  for lib in $libs; do
    if [ "$lib" = "-" ]; then unset AACFB_LIB; else export AACFB_LIB=$PWD/aac.js_b200/$lib; fi
    rm -f gpurun_out/ab.json
    timeout 200 python bench.py --steps 100 --warmup 3 --no-e2e --no-cpu --workload $wl > gpurun_out/ab.json 2>> gpurun_out/bench.err
    [ -s gpurun_out/ab.json ] || { echo "$wl $lib: failed"; continue; }
    python -c "import json;d=json.load(open('gpurun_out/ab.json'));print('$wl %-24s %.4f ms  frac %.3f' % ('$lib', d['ms_per_step'],d['roofline']['frac']))"
  done
done
done
