#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
cat > /tmp/q16.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import aacjs_b200 as A
from tools import workloads as W
import bench
dev = torch.device("cuda:0")
b = bench.Batch(A, W, torch, "config2_q16_s16", 256, 256, 0, dev, 0, q16_s16=True)
st = torch.cuda.current_stream()
for _ in range(6):
    b.step(st)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:synth_kernel -s 8 -c 1 -f -o gpurun_out/r3_q16 python /tmp/q16.py > gpurun_out/ncu_q16.log 2>&1; tail -2 gpurun_out/ncu_q16.log | cut -c1-200
