#!/bin/bash
# round-2 session-3 run 4: per-direction copy streams in the host pipeline: tests + e2e A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 180 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
for cs in 1 0; do
echo "AACFB_COPY_STREAMS=$cs"
AACFB_COPY_STREAMS=$cs bash tools/e2e_sweep.sh 8:2 16:2 32:2 16:4
done 2>&1 | tee gpurun_out/e2e_sweep_cs.log
