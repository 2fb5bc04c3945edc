"""One warm device-resident pass of config 2 in a given input/output format, for ncu (run under gpurun):
   python tools/prof_io.py q16|f32 s16|f32"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aacjs_b200 as A
from tools import workloads as W

S, T, C = 256, 256, 2
dev = torch.device("cuda:0")
inq, out16 = sys.argv[1] == "q16", sys.argv[2] == "s16"
w = W.make_q(2, S, T, C, seed=0)
inp = torch.from_numpy(w["qframes"].view(np.uint8).reshape(S, T, C, 2304)).to(dev) if inq else torch.randn((S, T, C, 1024), device=dev) * 3e5
info = torch.from_numpy(w["info"].view(np.uint8).reshape(S, T, C, 8).copy()).to(dev)
out = torch.empty((S, T, 1024, C), device=dev, dtype=torch.int16 if out16 else torch.float32)
ctx = A.Context(S, C, 4, 0)
st = torch.cuda.current_stream()
for _ in range(6):
    ctx.process_device_io(inp.data_ptr(), A.IN_Q16 if inq else A.IN_F32, info.data_ptr(), out.data_ptr(), A.PCM_S16 if out16 else A.PCM_F32, T, st.cuda_stream)
torch.cuda.synchronize()
