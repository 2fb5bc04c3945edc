#!/bin/bash
# round-2 session-2 run 1: tests, TNS chain micro-benchmark, W=5 A/B, generic-instantiation captures
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 60 tools/ubench/tns_chain > gpurun_out/tns_chain.log 2>&1; cat gpurun_out/tns_chain.log
bash tools/ab.sh "- libaacfb_w5.so" config2 config4 2>&1 | tee gpurun_out/ab_w5.log
for c in 3 5; do
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"synth_kernel<1" -s 3 -c 1 -f -o gpurun_out/r4_config$c python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --workload config$c > gpurun_out/ncu_c$c.log 2>&1; tail -1 gpurun_out/ncu_c$c.log | cut -c1-120
done
