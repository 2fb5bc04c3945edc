"""Device-resident timing of the four input/output format combinations of config 2 (run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aacjs_b200 as A
from oracle import oracle as O
from tools import workloads as W

S, T, C = 256, 256, 2
dev = torch.device("cuda:0")
sq = float(os.environ.get("SIGMA_Q", "40"))  # spread of the quantised integers
w = W.make_q(2, S, T, C, seed=0, sigma_q=sq)
qf = torch.from_numpy(w["qframes"].view(np.uint8).reshape(S, T, C, 2304)).to(dev)
info = torch.from_numpy(w["info"].view(np.uint8).reshape(S, T, C, 8).copy()).to(dev)
spec = torch.randn((S, T, C, 1024), device=dev) * 3e5
pf = torch.empty((S, T, 1024, C), device=dev)
p16 = torch.empty((S, T, 1024, C), device=dev, dtype=torch.int16)
st = torch.cuda.current_stream()
for name, inp, inf, out, outf in (("f32->f32", spec, A.IN_F32, pf, A.PCM_F32), ("f32->s16", spec, A.IN_F32, p16, A.PCM_S16),
                                  ("q16->f32", qf, A.IN_Q16, pf, A.PCM_F32), ("q16->s16", qf, A.IN_Q16, p16, A.PCM_S16)):
    ctx = A.Context(S, C, 4, 0)
    f = lambda: ctx.process_device_io(inp.data_ptr(), inf, info.data_ptr(), out.data_ptr(), outf, T, st.cuda_stream)
    for _ in range(5): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(50): f()
    b.record(st); torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b) / 50:.4f} ms  (sigma_q {sq})", flush=True)
    ctx.close()
