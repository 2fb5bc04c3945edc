#!/usr/bin/env python
"""Host<->device copy bandwidth of this box with pinned memory: H2D alone, D2H alone, and both
directions at once on two streams -- the ceiling of bench.py's `e2e` number (one step moves 538 MB
in and 537 MB out)."""
import json
import time

import torch

n = 512 << 20
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return n * reps / (time.perf_counter() - t0) / 1e9


run(True, True, 2)
res = {"h2d_alone_gbs": run(True, False), "d2h_alone_gbs": run(False, True), "both_each_way_gbs": run(True, True),
       "bytes": n}
res["e2e_floor_ms_for_538MB_each_way"] = 538e6 / (res["both_each_way_gbs"] * 1e9) * 1e3
print(json.dumps(res))
