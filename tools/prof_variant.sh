#!/bin/bash
# usage: prof_variant.sh <lib> <chunk> <tag> [workload]
mkdir -p gpurun_out
AACFB_LIB=$PWD/$1 AACFB_CHUNK_LEN=$2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:synth_kernel -s 3 -c 1 -f -o gpurun_out/prof_$3 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload ${4:-config2} > gpurun_out/ncu_$3.log 2>&1
tail -1 gpurun_out/ncu_$3.log | cut -c1-200
