#!/bin/bash
mkdir -p gpurun_out
for tp in 0 1 2 3; do
echo "AACFB_TAPER=$tp"
AACFB_TAPER=$tp bash tools/e2e_sweep.sh 8:2 16:2
done 2>&1 | tee gpurun_out/e2e_sweep_taper.log
