#!/bin/bash
timeout 400 ncu --set full --clock-control none --import-source on -k regex:synth_tns -s 3 -c 1 -f -o gpurun_out/r4_c4fused python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --workload config4 > gpurun_out/ncu_c4f.log 2>&1; tail -1 gpurun_out/ncu_c4f.log | cut -c1-100
