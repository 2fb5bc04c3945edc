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
  p2p_ms     : the fused path (sharding.share_from_root): every rank runs the synthesis kernel
               directly on rank 0's buffers mapped through CUDA IPC -- TMA row loads and PCM stores
               go over NVLink inside the kernel, both directions at once, no staging, no collective.
               Checked twice: every rank reads back over NVLink what its kernel left in rank 0's
               buffer and compares it bit for bit with its own local run; rank 0 compares the whole
               buffer with the PCM gathered through NCCL.
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
    ap.add_argument("--no-p2p", action="store_true", help="skip the fused peer-memory path")
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

    # ---- fused: the kernel reads from / writes to rank 0's buffers over NVLink -----------------
    p2p_ms = p2p_err = None
    if not args.no_p2p:
        # addresses of rank 0's buffers as seen from this rank's GPU (CUDA IPC, peer access over NVLink)
        a_spec = sharding.share_from_root(full_spec)
        a_info = sharding.share_from_root(full_info)
        a_pcm = sharding.share_from_root(full_pcm)
        row = T * C * 4096                      # bytes of spectra (= of PCM) per stream
        ctx2 = A.Context(n, C, 4, 0, device=local)

        def p2p():
            ctx2.process_device(a_spec + lo * row, a_info + lo * T * C * 8, a_pcm + lo * row, T, st.cuda_stream)

        # -- check, every rank: (1) local rows -> local PCM (what the scatter path computes),
        #    (2) rank 0's rows read over NVLink -> local PCM, (3) rank 0's rows -> rank 0's PCM buffer;
        #    rank 0 then gathers (1) and compares all S streams with what (3) left in full_pcm.
        def fresh(run_in_spec, run_in_info, run_out):
            c = A.Context(n, C, 4, 0, device=local)
            c.process_device(run_in_spec, run_in_info, run_out, T, st.cuda_stream)
            torch.cuda.synchronize()
            c.close()

        pcm_b = torch.empty_like(pcm)
        if rank == 0:
            full_pcm.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        fresh(spec.data_ptr(), info.data_ptr(), pcm.data_ptr())
        fresh(a_spec + lo * row, a_info + lo * T * C * 8, pcm_b.data_ptr())
        reads_ok = bool(torch.equal(pcm, pcm_b))
        p2p()                                   # first step of ctx2 (zero overlap) -> rank 0's buffer
        torch.cuda.synchronize()
        dist.barrier()
        # what this rank's kernel left in rank 0's buffer, read back over NVLink by this rank itself
        from cuda.bindings import runtime as rt
        back = torch.empty_like(pcm)
        rt.cudaMemcpy(back.data_ptr(), a_pcm + lo * row, back.numel() * 4, rt.cudaMemcpyKind.cudaMemcpyDefault)
        torch.cuda.synchronize()
        own_bad = (back != pcm).reshape(n, -1).any(dim=1)
        own = torch.tensor([int(own_bad.sum()), int(((back == 0) & (pcm != 0)).sum())], device=dev)
        own_all = [torch.zeros_like(own) for _ in range(world)]
        dist.all_gather(own_all, own)
        del back
        full_ref = torch.empty_like(full_pcm) if rank == 0 else None
        sharding.gather_streams(pcm, full_ref, S)
        flags = torch.tensor([int(reads_ok)], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            bad = (full_ref != full_pcm).reshape(S, -1).any(dim=1)
            p2p_err = {"peer_reads_bit_identical_on_every_rank": bool(flags.item()),
                       "streams_differing_in_root_buffer": int(bad.sum()),
                       "per_rank_readback_[streams_differing, samples_left_zero]": [t.tolist() for t in own_all],
                       "differing_streams_by_rank_region": [int(bad[slice(*sharding.stream_range(S, world, r))].sum()) for r in range(world)],
                       "max_abs_diff": float((full_ref - full_pcm).abs().max()), "max_abs_pcm": float(full_ref.abs().max())}
            del full_ref
        del pcm_b
        dist.barrier()   # nobody starts overwriting rank 0's buffer (the timing loop) while rank 0 still compares
        p2p_ms = timed(p2p)
        ctx2.close()
    if rank == 0:
        moved = (S - (hi - lo)) * T * C * 4096
        print(json.dumps({"workload": f"config5 mixed long/short, {S} streams x {T} frames stereo, root scatter/gather",
                          "n_gpus": world, "kernel_ms": kernel_ms, "e2e_ms": e2e_ms,
                          "kernel_frames_per_s": S * T / kernel_ms * 1e3, "e2e_frames_per_s": S * T / e2e_ms * 1e3,
                          "p2p_ms": p2p_ms, "p2p_frames_per_s": S * T / p2p_ms * 1e3 if p2p_ms else None,
                          "p2p_check": p2p_err,
                          "p2p_nvlink_gbs_each_way": moved / (p2p_ms * 1e-3) / 1e9 if p2p_ms else None,
                          "nvlink_bytes_each_way": moved,
                          "nvlink_gbs_each_way": moved / ((e2e_ms - kernel_ms) / 2 * 1e-3) / 1e9 if e2e_ms > kernel_ms else None}))
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
