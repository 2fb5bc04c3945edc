#!/usr/bin/env python
"""bench.py -- AAC-LC filterbank-synthesis throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A *step* is one pass of the hot path (TNS -> IMDCT -> window -> overlap-add -> interleave) over one
batch of synthetic spectra.  Default workload = BASELINE.json configs[1]: 65 536 stereo ONLY_LONG
frames (S=256 streams x T=256 frames), TNS off, per GPU.  Multi-GPU (torchrun, one rank per GPU)
shards by stream: every rank runs its own batch, no data-path collective -> weak scaling; `value`
is the whole-job frames/s over the max-over-ranks device time.

`value`   : inputs resident in HBM, timed with CUDA events on the launching stream.
`e2e`     : same metric through the C-ABI call a host makes (aacfb_process) with pinned HOST
            buffers; H2D of the spectra and D2H of the PCM are inside the timed region.
`roofline`: algorithmic bytes (8192 B per channel-frame + overlap state + side info)
            / mean launch duration of the synthesis kernel vs MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline`: the oracle (C restatement of the reference's JS) on the host cores, bounded sample.
`configs` : (N = 1) the other north_star configurations -- EIGHT_SHORT, LONG + TNS, mixed, config 2 with
            the stereo tools -- timed the same way in the same run: ms_per_step, frames/s, roofline.
`config5_scatter`: (N > 1) BASELINE.json configs[4] at full size: 1 048 576 stereo mixed frames that start
            on rank 0, through the NCCL scatter -> kernel -> gather path and through the fused
            peer-memory path (synth_kernel on rank 0's buffers over NVLink), each timed and bit-checked.
`pcie_concurrent`: host<->device copy rates per GPU with all ranks copying at once (256 MB pinned copies, each
            direction alone and both together): the link ceiling the `e2e` leg is to be read against.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "AAC-LC frames/sec (1024-pt IMDCT+OLA)"
UNIT = "frames/s"
WORKLOADS = {
    # name: (config id of tools/workloads.make, S, T, C, description)
    "config2": (2, 256, 256, 2, "batch=65536 stereo LC frames, ONLY_LONG_SEQUENCE, TNS off"),
    "config3": (3, 256, 256, 2, "batch=65536 stereo, EIGHT_SHORT_SEQUENCE"),
    "config4": (4, 256, 256, 2, "batch=65536 stereo, LONG + TNS (order 12, all bands), FIXED_AR"),
    "config5": (5, 256, 256, 2, "batch=65536 stereo per GPU, mixed long/short (t mod 16 pattern)"),
    # config2 with the stereo tools (processMS / processIS) applied on the device
    "config2_stereo": (2, 256, 256, 2, "config2 + M/S on bands 0-39 and intensity stereo on bands 40-45 of every frame, "
                                       "applied on the staged spectra"),
}


def algorithmic_bytes(S, T, C, tns_bytes=0, stereo_bytes=0):
    """4096 B spectrum in + 4096 B PCM out per channel-frame, plus the
    overlap state read+written once per (stream, channel), plus 8 B side info per channel-frame
    (+ the TNS blob / the 768-byte stereo records where the workload has them)."""
    n_cf = S * T * C
    return n_cf * 8192 + S * C * 8192 + n_cf * 8 + tns_bytes + stereo_bytes


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index`, BEFORE any pinned host
    memory is allocated, so that the e2e buffers are first-touched on the GPU's own NUMA node and
    the PCIe traffic of N ranks does not cross the socket interconnect.  Returns the CPU count of
    the mask (0: unavailable, nothing changed)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def oracle_rate(w, S_sample, threads, repeats=1, min_seconds=0.0):
    """frames/s of the CPU oracle on the first S_sample streams of workload w: the sample is run
    `repeats` times and then again until `min_seconds` of wall time have been spent; returns
    (frames/s over everything that was run, seconds, passes)."""
    from oracle import oracle as O

    sp, inf = w["spectra"][:S_sample], w["info"][:S_sample]
    T, C = sp.shape[1], sp.shape[2]
    blob, offs = w["tns_blob"], w["tns_offsets"]
    if offs is not None:
        offs = offs[: S_sample * T * C + 1]
    total, passes = 0.0, 0
    while passes < repeats or total < min_seconds:
        t0 = time.perf_counter()
        O.process(sp, inf, blob, offs, sample_index=w["sample_index"], flags=w["flags"], n_threads=threads)
        total += time.perf_counter() - t0
        passes += 1
    return S_sample * T * passes / total, total, passes


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is
    JavaScript and no JS runtime exists on this image, so it is the oracle's C restatement
    (kind "port"), all host threads, on a bounded sample of the same workload per step."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from tools import workloads as W

    cfg, S, T, C, desc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    S_sample = max(cores, min(S, 8 * cores))  # a few streams per thread: ~0.5-1 s per step
    w = W.make(cfg, S_sample, T, C, seed=0)
    from oracle import oracle as O

    O.lib()
    for _ in range(args.warmup):
        oracle_rate(w, S_sample, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_rate(w, S_sample, cores)  # one step = one pass over the bounded sample
    dt = time.perf_counter() - t0
    value = S_sample * T * args.steps / dt
    sample = f"{S_sample} streams x {T} frames x {C} ch of {args.workload} per step, {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 stores / f64 temporaries",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class Batch:
    """One synthetic batch of a workload, resident in HBM on this rank, with its context."""

    def __init__(self, A, W, torch, workload, S, T, rank, dev, local, q16_s16=False):
        self.A, self.q16_s16 = A, q16_s16
        if q16_s16:
            workload = workload[: -len("_q16_s16")]
        cfg, _, _, C, desc = WORKLOADS[workload]
        self.workload, self.cfg, self.S, self.T, self.C, self.desc = workload, cfg, S, T, C, desc
        stereo = workload.endswith("_stereo")
        sigma = {2: 3.0e5, 3: 1.0e5, 4: 0.75e5, 5: 1.0e5}[cfg] * (0.5 if stereo else 1.0)
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        self.spectra = torch.randn((S, T, C, 1024), device=dev, generator=g) * sigma
        self.side = side = W.make(cfg, S, T, C, seed=rank, side_only=True)  # info/TNS side data only
        self.info_np = side["info"]
        self.blob = self.offs = self.ops = self.ops_np = None
        self.tns_bytes = 0
        if side["tns_blob"] is not None:
            self.blob = torch.from_numpy(side["tns_blob"]).to(dev)
            self.offs = torch.from_numpy(side["tns_offsets"].view(np.int32).copy()).to(dev)
            self.tns_bytes = int(side["tns_blob"].size)
        if stereo:
            self.ops_np = W.joint_stereo_ops(S, T, seed=rank)
            self.info_np["stereo_present"][:, :, 0] = 1
            self.ops = torch.from_numpy(self.ops_np.view(np.uint8).reshape(S, T, 768).copy()).to(dev)
        self.info = torch.from_numpy(self.info_np.view(np.uint8).reshape(S, T, C, 8).copy()).to(dev)
        self.stereo_bytes = 0 if self.ops_np is None else int(self.ops_np.nbytes)
        self.pcm = torch.empty((S, T, 1024, C), device=dev, dtype=torch.int16 if q16_s16 else torch.float32)
        self.ctx = A.Context(S, C, side["sample_index"], side["flags"], device=local)
        self.alg_bytes = algorithmic_bytes(S, T, C, self.tns_bytes, self.stereo_bytes)
        if q16_s16:   # aacfb_qframe records in (2304 B), int16 PCM out (2048 B) per channel-frame
            wq = W.make_q(cfg, S, T, C, seed=rank)
            self.spectra = torch.from_numpy(wq["qframes"].view(np.uint8).reshape(S, T, C, 2304)).to(dev)
            self.alg_bytes -= S * T * C * (8192 - 2304 - 2048)

    def step(self, stream):
        if self.q16_s16:
            A = self.A
            self.ctx.process_device_io(self.spectra.data_ptr(), A.IN_Q16, self.info.data_ptr(), self.pcm.data_ptr(), A.PCM_S16,
                                       self.T, stream.cuda_stream)
            return
        self.ctx.process_device(self.spectra.data_ptr(), self.info.data_ptr(), self.pcm.data_ptr(), self.T,
                                stream.cuda_stream, self.blob.data_ptr() if self.blob is not None else 0,
                                self.offs.data_ptr() if self.offs is not None else 0, self.tns_bytes,
                                self.ops.data_ptr() if self.ops is not None else 0)

    def close(self):
        self.ctx.close()
        self.spectra = self.pcm = self.info = self.blob = self.offs = self.ops = None


def spin_up(torch, step, stream, ms=40.0, chunk=8, limit=2000):
    """Untimed steps until the GPU has been busy for `ms`: a configuration timed right after seconds of host-side
    data generation otherwise starts at idle clocks and its first steps run slow (the W warm-up steps of a 0.2 ms
    kernel last about a millisecond).  Extra warm-up only; nothing here is timed into a reported number."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    done = 0
    e0.record(stream)
    while done < limit:
        for _ in range(chunk):
            step()
        done += chunk
        e1.record(stream)
        e1.synchronize()
        if e0.elapsed_time(e1) >= ms:
            break


def time_steps(torch, step, steps, warmup, stream, barrier):
    """W untimed steps, then exactly `steps` steps between CUDA events on `stream`, barrier +
    synchronize on both sides.  Returns (total ms, [per-step ms])."""
    for _ in range(warmup):
        step()
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record(stream)
    for i in range(steps):
        step()
        evs[i + 1].record(stream)
    barrier()
    return evs[0].elapsed_time(evs[-1]), [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]


def load_traffic():
    """Measured DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum of one
    `ncu --set full` capture per kernel; profiles/traffic.json is written by tools/refresh_profiles.py
    from the captures of tools/gpu_capture.sh -- a profiler cannot run inside a timed bench)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def roofline_record(workload, alg, launch_ms, tns, traffic_tab):
    peak_gbs, peak_src = measured_peak()
    achieved = alg / (launch_ms * 1e-3) / 1e9
    t = traffic_tab.get(workload)
    return {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
            "traffic": t, "traffic_source": traffic_tab.get("_source") if t else None, "peak_source": peak_src,
            "kernel": "aacfb::tns_kernel + aacfb::synth_kernel" if tns else "aacfb::synth_kernel",
            "algorithmic_bytes_per_launch": alg, "launch_ms": launch_ms}


def pcie_probe(torch, dev, world, n_bytes=256 << 20, reps=6):
    """Host<->device copy rates of this rank while ALL ranks copy at once (pinned memory, one
    stream per direction): what the box gives N GPUs concurrently -- the ceiling of `e2e`."""
    h_in = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h):
        if world > 1:
            torch.distributed.barrier()
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
        return n_bytes * reps / (time.perf_counter() - t0) / 1e9

    run(True, True)
    mine = torch.tensor([run(True, False), run(False, True), run(True, True)], device=dev, dtype=torch.float64)
    allr = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        torch.distributed.all_gather(allr, mine)
    else:
        allr = [mine]
    m = torch.stack(allr).cpu().numpy()
    return {"concurrent_ranks": world, "bytes_per_copy": n_bytes,
            "h2d_alone_gbs_per_gpu": [round(float(v), 1) for v in m[:, 0]],
            "d2h_alone_gbs_per_gpu": [round(float(v), 1) for v in m[:, 1]],
            "duplex_each_way_gbs_per_gpu": [round(float(v), 1) for v in m[:, 2]],
            "duplex_each_way_gbs_sum": round(float(m[:, 2].sum()), 1)}


def bit_checksum(torch, t):
    """Order-independent 2 x 64-bit checksum of the raw bits of a float tensor (wrapping sums)."""
    v = t.reshape(-1).view(torch.int32).to(torch.int64)
    return torch.stack([v.sum(), (v * v + (v >> 3)).sum()])


def run_config5_scatter(A, W, torch, args, rank, world, local, dev):
    """BASELINE.json configs[4]: 1 048 576 stereo mixed long/short frames that START ON RANK 0, sharded
    by stream over the N GPUs.  Two exchange paths, both timed (CUDA events, max over ranks) and both
    checked bit for bit against each rank's own run on its shard:
      nccl : NCCL scatter of spectra + side info -> synth_kernel -> NCCL gather of PCM
      fused: every rank maps rank 0's buffers (CUDA IPC) and runs synth_kernel directly on the peer
             addresses: TMA row loads and PCM stores cross NVLink inside the kernel."""
    import torch.distributed as dist
    from aacjs_b200 import sharding

    S, T, C = args.scatter_streams, 256, 2
    lo, hi = sharding.stream_range(S, world, rank)
    n = hi - lo
    st = torch.cuda.current_stream()
    full_spec = full_info = full_pcm = None
    if rank == 0:
        side = W.make(5, S, T, C, seed=0, side_only=True)
        g = torch.Generator(device=dev).manual_seed(99)
        full_spec = torch.randn((S, T, C, 1024), device=dev, generator=g) * 1.0e5
        full_info = torch.from_numpy(side["info"].view(np.uint8).reshape(S, T, C, 8).copy()).to(dev)
        full_pcm = torch.zeros((S, T, 1024, C), device=dev)
    spec = torch.empty((n, T, C, 1024), device=dev)
    info = torch.empty((n, T, C, 8), dtype=torch.uint8, device=dev)
    pcm = torch.empty((n, T, 1024, C), device=dev)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def region_sums():   # rank 0: checksum of every rank's region of the root PCM buffer
        out = torch.zeros((world, 2), dtype=torch.int64, device=dev)
        if rank == 0:
            for r in range(world):
                a, b = sharding.stream_range(S, world, r)
                out[r] = bit_checksum(torch, full_pcm[a:b])
        dist.broadcast(out, src=0)
        return out

    def local_sums(t):
        mine = bit_checksum(torch, t)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        return torch.stack(allr)

    def timed(fn, steps):
        total, _ = time_steps(torch, fn, steps, 2, st, barrier)
        t = torch.tensor([total / steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    res = {"workload": f"config5: {S * T} stereo frames mixed long/short ({S} streams x {T}), resident on rank 0",
           "n_gpus": world}
    # ---- reference result of every shard: its own rows, its own GPU, fresh (zero) overlap ----------
    sharding.scatter_streams(full_spec, spec, S)
    sharding.scatter_streams(full_info, info, S)
    c0 = A.Context(n, C, 4, 0, device=local)
    c0.process_device(spec.data_ptr(), info.data_ptr(), pcm.data_ptr(), T, st.cuda_stream)
    torch.cuda.synchronize()
    c0.close()
    want = local_sums(pcm)

    # ---- NCCL scatter -> kernel -> NCCL gather -----------------------------------------------------
    ctx = A.Context(n, C, 4, 0, device=local)

    def kernel():
        ctx.process_device(spec.data_ptr(), info.data_ptr(), pcm.data_ptr(), T, st.cuda_stream)

    def nccl_path():
        sharding.scatter_streams(full_spec, spec, S)
        sharding.scatter_streams(full_info, info, S)
        kernel()
        sharding.gather_streams(pcm, full_pcm, S)

    nccl_path()          # first step of ctx: zero overlap, like the reference result
    barrier()
    res["nccl_bit_identical"] = bool(torch.equal(region_sums(), want))
    steps = max(2, min(args.steps, 5))
    res["nccl_ms"] = timed(nccl_path, steps)
    res["kernel_ms"] = timed(kernel, steps)
    ctx.close()
    moved = (S - n) * T * C * 4096 if rank == 0 else 0
    mv = torch.tensor([moved], device=dev, dtype=torch.float64)
    dist.broadcast(mv, src=0)
    moved = float(mv.item())
    res["nvlink_bytes_each_way_at_root"] = int(moved)
    res["nccl_frames_per_s"] = S * T / res["nccl_ms"] * 1e3
    res["kernel_frames_per_s"] = S * T / res["kernel_ms"] * 1e3
    if res["nccl_ms"] > res["kernel_ms"]:
        res["nccl_nvlink_gbs_each_way"] = moved / ((res["nccl_ms"] - res["kernel_ms"]) / 2 * 1e-3) / 1e9

    # ---- fused over peer memory --------------------------------------------------------------------
    try:
        a_spec = sharding.share_from_root(full_spec)
        a_info = sharding.share_from_root(full_info)
        a_pcm = sharding.share_from_root(full_pcm)
        row = T * C * 4096
        ctx2 = A.Context(n, C, 4, 0, device=local)

        def fused():
            ctx2.process_device(a_spec + lo * row, a_info + lo * T * C * 8, a_pcm + lo * row, T, st.cuda_stream)

        if rank == 0:
            full_pcm.zero_()
        barrier()
        fused()          # first step of ctx2 (zero overlap) straight into rank 0's buffer
        barrier()
        res["fused_bit_identical"] = bool(torch.equal(region_sums(), want))
        barrier()        # nobody overwrites rank 0's buffer while rank 0 still sums it
        res["fused_ms"] = timed(fused, steps)
        res["fused_frames_per_s"] = S * T / res["fused_ms"] * 1e3
        res["fused_nvlink_gbs_each_way"] = moved / (res["fused_ms"] * 1e-3) / 1e9
        ctx2.close()
    except Exception as e:  # the NCCL path above stays valid
        res["fused_error"] = f"{type(e).__name__}: {e}"[:300]
    barrier()
    return res


def run_ours(args):
    import torch

    import aacjs_b200 as A
    from tools import workloads as W

    rank, world, local = dist_env()
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    all_cpus = os.sched_getaffinity(0)
    numa_cpus = 0 if args.no_numa else bind_to_gpu_numa_node(local)
    cfg, S, T, C, desc = WORKLOADS[args.workload]
    if args.streams:
        S = args.streams
    if args.frames:
        T = args.frames
    stream = torch.cuda.current_stream()
    traffic_tab = load_traffic()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # --- headline: synthetic batch of this rank (seeded per rank), resident in HBM ------------
    b = Batch(A, W, torch, args.workload, S, T, rank, dev, local)
    warm = max(args.warmup, 3)
    spin_up(torch, lambda: b.step(stream), stream)
    for _ in range(warm):
        b.step(stream)
    barrier()
    launches0 = b.ctx.launches
    with ClockSampler(local) as clocks:
        total_ms, per_step = time_steps(torch, lambda: b.step(stream), args.steps, 0, stream, barrier)
    clock_kernel = clocks.summary()
    gpu_launches = b.ctx.launches - launches0
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * S * T * args.steps / (total_ms_max * 1e-3)

    # sanity: the timed kernels produced real PCM
    peak = float(b.pcm.abs().max())
    assert 0.05 < peak < 50 and bool(torch.isfinite(b.pcm).all()), peak

    # --- the other north_star configurations, same clock, same box (N = 1: BENCH and SCALE's first point) ---
    # (before the end-to-end legs: measured after them, with the device heap those legs leave behind, config 3 comes
    # out 3 % slower than in a process of its own -- 0.3075 against 0.2981 ms -- the other configurations do not move)
    configs = None
    if world == 1 and not args.no_configs and args.workload == "config2" and not args.streams and not args.frames:
        configs = {}
        for name in ("config3", "config4", "config5", "config2_stereo", "config2_q16_s16"):
            q16 = name.endswith("_q16_s16")
            _, S2, T2, C2, desc2 = WORKLOADS[name[: -len("_q16_s16")] if q16 else name]
            if q16:
                desc2 += "; input = aacfb_qframe records (inverse quantisation on the device), output = int16 PCM"
            bb = Batch(A, W, torch, name, S2, T2, rank, dev, local, q16_s16=q16)
            k = max(5, min(args.steps, 100))
            spin_up(torch, lambda: bb.step(stream), stream)
            for _ in range(3):
                bb.step(stream)
            l0 = bb.ctx.launches
            tot, per = time_steps(torch, lambda: bb.step(stream), k, 0, stream, barrier)
            ms = tot / k
            pk = float(bb.pcm.float().abs().max()) / (32768.0 if q16 else 1.0)
            assert 0.01 < pk < 100 and bool(torch.isfinite(bb.pcm.float()).all()), (name, pk)
            configs[name] = {"workload": desc2, "steps": k, "ms_per_step": ms, "value": S2 * T2 / ms * 1e3, "unit": UNIT,
                             "gpu_launches": int(bb.ctx.launches - l0),
                             "roofline": roofline_record(name, bb.alg_bytes, float(np.mean(per)), bb.tns_bytes > 0, traffic_tab)}
            bb.close()
            del bb
            torch.cuda.empty_cache()

    # --- end to end through the host-buffer C-ABI call (H2D + D2H inside the timed region) ---
    # Headline leg: the host hands over what the bit parse holds BEFORE inverse quantisation (aacfb_qframe:
    # int16 coefficients + scalefactor indices, 2304 B per channel-frame; the device runs ics.js:203-266) and
    # takes int16 PCM back (2048 B) -- 4.3 instead of 8.0 KiB per channel-frame over PCIe.  The float-in /
    # float-out call of round 1 (readChunk's own formats) is timed beside it, and so is the headline leg
    # from pageable memory (a host that neither allocates through aacfb_host_alloc nor registers its arrays).
    e2e = None
    side, info_np, ops_np, tns_bytes = b.side, b.info_np, b.ops_np, b.tns_bytes
    if not args.no_e2e:
        e2e_steps = args.e2e_steps or max(2, min(args.steps, 20))
        side_bytes = info_np.nbytes + (tns_bytes + side["tns_offsets"].nbytes if tns_bytes else 0) + b.stereo_bytes

        def leg(inp, in_format, out, pcm_format, steps):
            ctx2 = A.Context(S, C, side["sample_index"], side["flags"], device=local)
            call = lambda: ctx2.process_io(inp, info_np, side["tns_blob"], side["tns_offsets"], out=out, stereo_ops=ops_np,
                                           in_format=in_format, pcm_format=pcm_format)
            for _ in range(2):
                call()
            barrier()
            with ClockSampler(local) as cs:
                t0 = time.perf_counter()
                for _ in range(steps):
                    call()
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            ctx2.close()
            te = torch.tensor([dt], device=dev, dtype=torch.float64)
            if world > 1:
                torch.distributed.all_reduce(te, op=torch.distributed.ReduceOp.MAX)
            return float(te.item()) / steps * 1e3, cs.summary()

        def record(ms, inp, out, api):
            return {"value": world * S * T / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                    "h2d_bytes_per_step": int(inp.nbytes + side_bytes), "d2h_bytes_per_step": int(out.nbytes), "api": api}

        numa = f"process bound to the {numa_cpus} CPUs local to the GPU" if numa_cpus else "no binding"
        variants = {}
        # (1) float spectra in, float PCM out (pinned)
        h_spec = A.host_alloc((S, T, C, 1024), np.float32)
        h_spec[...] = b.spectra.cpu().numpy()
        h_pcm = A.host_alloc((S, T, 1024, C), np.float32)
        ms, ce = leg(h_spec, A.IN_F32, h_pcm, A.PCM_F32, e2e_steps)
        assert np.isfinite(h_pcm[0, 0]).all() and np.abs(h_pcm[-1, -1]).max() > 0
        variants["f32_in_f32_out"] = record(ms, h_spec, h_pcm, "aacfb_process (float spectra -> float PCM, page-locked buffers)")
        A.host_free(h_spec); A.host_free(h_pcm)
        del h_spec, h_pcm
        e2e = dict(variants["f32_in_f32_out"])
        if args.workload in ("config2", "config3", "config5") and ops_np is None:
            wq = W.make_q(cfg, S, T, C, seed=rank)
            h_q = A.host_alloc((S, T, C), A.QFRAME_DTYPE)
            h_q[...] = wq["qframes"]
            h_p16 = A.host_alloc((S, T, 1024, C), np.int16)
            ms, ce = leg(h_q, A.IN_Q16, h_p16, A.PCM_S16, e2e_steps)
            pk = int(np.abs(h_p16[0].astype(np.int32)).max())
            assert 500 < pk <= 32768, pk
            variants["q16_in_s16_out"] = record(ms, h_q, h_p16, "aacfb_process_io (aacfb_qframe -> int16 PCM, page-locked buffers "
                                                "from aacfb_host_alloc): inverse quantisation on the device")
            h_pf = A.host_alloc((S, T, 1024, C), np.float32)
            ms2, _ = leg(h_q, A.IN_Q16, h_pf, A.PCM_F32, max(2, e2e_steps // 2))
            variants["q16_in_f32_out"] = record(ms2, h_q, h_pf, "aacfb_process_io (aacfb_qframe -> float PCM)")
            A.host_free(h_pf)
            del h_pf
            # the same call from ordinary (pageable) numpy arrays
            p_q, p_p16 = wq["qframes"], np.empty((S, T, 1024, C), np.int16)
            ms3, _ = leg(p_q, A.IN_Q16, p_p16, A.PCM_S16, max(2, e2e_steps // 4))
            variants["q16_in_s16_out_pageable"] = record(ms3, p_q, p_p16, "the headline call from pageable memory")
            A.host_free(h_q); A.host_free(h_p16)
            del h_q, h_p16, wq, p_q, p_p16
            e2e = dict(variants["q16_in_s16_out"])
        e2e["steps"] = e2e_steps
        e2e["numa"] = numa
        e2e["variants"] = variants
        if ce["samples"]:  # the e2e loop is long enough for nvidia-smi's 100 ms sampling: merge
            clock_kernel = {"sm_mhz": ce["sm_mhz"] if not clock_kernel["samples"] else clock_kernel["sm_mhz"],
                            "sm_max_mhz": ce["sm_max_mhz"],
                            "reasons": sorted(set(ce["reasons"]) | set(clock_kernel["reasons"])),
                            "samples": clock_kernel["samples"], "samples_e2e": ce["samples"], "sm_mhz_e2e": ce["sm_mhz"]}
        # one call at the shape the drop-in decoder (js/decoder_b200.js) actually makes: 1 stream x K frames
        lat = {}
        for K in (64, 1):
            cl = A.Context(1, C, side["sample_index"], 0, device=local)
            wl = W.make_q(2, 1, K, C, seed=5)
            hq = A.host_alloc((1, K, C), A.QFRAME_DTYPE); hq[...] = wl["qframes"]
            hp = A.host_alloc((1, K, 1024, C), np.int16)
            hs = A.host_alloc((1, K, C, 1024), np.float32); hs[...] = 1000.0
            hf = A.host_alloc((1, K, 1024, C), np.float32)
            for name, fn in (("q16_s16", lambda: cl.process_io(hq, wl["info"], out=hp, in_format=A.IN_Q16, pcm_format=A.PCM_S16)),
                             ("f32_f32", lambda: cl.process_io(hs, wl["info"], out=hf))):
                for _ in range(10):
                    fn()
                t0 = time.perf_counter()
                for _ in range(200):
                    fn()
                lat[f"S1_T{K}_{name}"] = (time.perf_counter() - t0) / 200 * 1e6
            cl.close()
            for h in (hq, hp, hs, hf):
                A.host_free(h)
        e2e["latency_us_per_call"] = lat
    alg_head = b.alg_bytes
    b.close()
    del b
    torch.cuda.empty_cache()

    # --- N > 1: what the box's host<->device path gives all ranks at once, and config 5 from rank 0 ---
    pcie = scatter = None
    if not args.no_pcie_probe and not args.no_e2e:   # (N = 1 too: the link ceiling `e2e` is to be read against)
        pcie = pcie_probe(torch, dev, world)
    if world > 1 and not args.no_scatter:
        try:
            scatter = run_config5_scatter(A, W, torch, args, rank, world, local, dev)
        except Exception as e:
            scatter = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0:
        # synthesis-kernel launch time: each step is one memset + synth_kernel (x2 instantiations) on one
        # stream; event-to-event time of a step is the launch duration the roofline uses
        launch_ms = float(np.mean(per_step))
        cpu = None
        os.sched_setaffinity(0, all_cpus)  # the CPU baseline uses every host core again
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            S_cpu = max(cores, min(S, 4 * cores))
            wc = W.make(cfg, S_cpu, T, C, seed=0)
            oracle_rate(wc, S_cpu, cores)  # warm-up pass (page faults, thread start)
            # a bounded sample: the same S_cpu-stream slice over and over for >= 2 s of wall time on
            # all host cores (about 30 core-seconds on a 16-core box), then >= 1.5 s on one thread
            rate, secs, passes = oracle_rate(wc, S_cpu, cores, min_seconds=2.0)
            rate1, secs1, passes1 = oracle_rate(wc, max(1, S_cpu // cores), 1, min_seconds=1.5)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{S_cpu} streams x {T} frames x {C} ch of {args.workload}, {passes} passes in {secs:.2f} s "
                             f"({cores} threads = {secs * cores:.0f} core-seconds)",
                   "single_thread": rate1,
                   "note": "C restatement of aac.js under the JS rounding model (no JS engine on this image)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "streams_per_gpu": S, "frames_per_stream": T,
                       "channels": C, "parallelism": f"streams sharded over {world} GPU(s), no data-path collective",
                       "l2": "inputs+outputs 1 GiB per step >> 126 MB L2 (no flush needed)",
                       "spin_up": "40 ms of untimed steps before the W warm-up steps of every timed configuration (clock ramp)"},
            "clocks": clock_kernel,
            "e2e": e2e,
            "gpu_launches": int(gpu_launches),
            "roofline": roofline_record(args.workload, alg_head, launch_ms, tns_bytes > 0, traffic_tab),
            "cpu_baseline": cpu,
        }
        if configs is not None:
            line["configs"] = configs
        if pcie is not None:
            line["pcie_concurrent"] = pcie
        if scatter is not None:
            line["config5_scatter"] = scatter
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=0)
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed steps of the e2e leg (default min(steps, 20))")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process to the GPU's NUMA node")
    ap.add_argument("--no-configs", action="store_true", help="N = 1: skip the configs record (configs 3, 4, 5, config2_stereo)")
    ap.add_argument("--no-scatter", action="store_true", help="N > 1: skip BASELINE configs[4] (1 048 576 frames from rank 0)")
    ap.add_argument("--no-pcie-probe", action="store_true", help="skip the concurrent host<->device copy probe")
    ap.add_argument("--scatter-streams", type=int, default=4096, help="streams of the config-5 batch on rank 0 (x 256 frames)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
