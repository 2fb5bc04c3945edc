#!/usr/bin/env python
"""BASELINE config 5: a mixed long/short batch that starts on rank 0, sharded over N GPUs by an
NCCL scatter of frame batches and a gather of PCM (aac.js_b200/sharding.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/scatter_bench.py --streams 1024 --frames 256

Reports, as one JSON line from rank 0 (max over ranks, CUDA events):
  kernel_ms  : synthesis on the pre-sharded batch (what scales ~linearly: no communication)
  e2e_ms     : scatter + synthesis + gather from/to rank 0 -- bounded by rank 0's NVLink
               egress/ingress (~770 GB/s measured per direction), an order of magnitude below the
               kernel's HBM rate, so the two are reported separately.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aacjs_b200 as A  # noqa: E402
from aacjs_b200 import sharding  # noqa: E402
from tools import workloads as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=1024)
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    S, T, C = args.streams, args.frames, 2
    lo, hi = sharding.stream_range(S, world, rank)
    n = hi - lo
    full_spec = full_info = full_pcm = None
    if rank == 0:
        side = W.make(5, S, T, C, seed=0, side_only=True)
        full_spec = torch.randn((S, T, C, 1024), device=dev) * 1.0e5
        full_info = torch.from_numpy(side["info"].view(np.uint8).reshape(S, T, C, 8).copy()).to(dev)
        full_pcm = torch.empty((S, T, 1024, C), device=dev)
    spec = torch.empty((n, T, C, 1024), device=dev)
    info = torch.empty((n, T, C, 8), dtype=torch.uint8, device=dev)
    pcm = torch.empty((n, T, 1024, C), device=dev)
    ctx = A.Context(n, C, 4, 0, device=local)
    st = torch.cuda.current_stream()

    def kernel():
        ctx.process_device(spec.data_ptr(), info.data_ptr(), pcm.data_ptr(), T, st.cuda_stream)

    def e2e():
        sharding.scatter_streams(full_spec, spec, S)
        sharding.scatter_streams(full_info, info, S)
        kernel()
        sharding.gather_streams(pcm, full_pcm, S)

    def timed(fn):
        for _ in range(2):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(args.steps):
            fn()
        b.record(st)
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / args.steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    e2e_ms = timed(e2e)
    kernel_ms = timed(kernel)
    if rank == 0:
        moved = (S - (hi - lo)) * T * C * 4096
        print(json.dumps({"workload": f"config5 mixed long/short, {S} streams x {T} frames stereo, root scatter/gather",
                          "n_gpus": world, "kernel_ms": kernel_ms, "e2e_ms": e2e_ms,
                          "kernel_frames_per_s": S * T / kernel_ms * 1e3, "e2e_frames_per_s": S * T / e2e_ms * 1e3,
                          "nvlink_bytes_each_way": moved,
                          "nvlink_gbs_each_way": moved / ((e2e_ms - kernel_ms) / 2 * 1e-3) / 1e9 if e2e_ms > kernel_ms else None}))
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
