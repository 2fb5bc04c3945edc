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

    # --- synthetic batch of this rank (seeded per rank), resident in HBM --------------------
    sigma = {2: 3.0e5, 3: 1.0e5, 4: 0.75e5, 5: 1.0e5}[cfg] * (0.5 if args.workload.endswith("_stereo") else 1.0)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    spectra = torch.randn((S, T, C, 1024), device=dev, generator=g) * sigma
    side = W.make(cfg, S, T, C, seed=rank, side_only=True)  # info/TNS side data only
    info_np = side["info"]
    info = torch.from_numpy(info_np.view(np.uint8).reshape(S, T, C, 8).copy()).to(dev)
    blob = offs = None
    tns_bytes = 0
    if side["tns_blob"] is not None:
        blob = torch.from_numpy(side["tns_blob"]).to(dev)
        offs = torch.from_numpy(side["tns_offsets"].view(np.int32).copy()).to(dev)
        tns_bytes = int(side["tns_blob"].size)
    ops_np = ops = None
    if args.workload.endswith("_stereo"):
        ops_np = W.joint_stereo_ops(S, T, seed=rank)
        info_np["stereo_present"][:, :, 0] = 1
        info = torch.from_numpy(info_np.view(np.uint8).reshape(S, T, C, 8).copy()).to(dev)
        ops = torch.from_numpy(ops_np.view(np.uint8).reshape(S, T, 768).copy()).to(dev)
    stereo_bytes = 0 if ops_np is None else int(ops_np.nbytes)
    pcm = torch.empty((S, T, 1024, C), device=dev)
    ctx = A.Context(S, C, side["sample_index"], side["flags"], device=local)
    stream = torch.cuda.current_stream()

    def step():
        ctx.process_device(spectra.data_ptr(), info.data_ptr(), pcm.data_ptr(), T, stream.cuda_stream,
                           blob.data_ptr() if blob is not None else 0, offs.data_ptr() if offs is not None else 0,
                           tns_bytes, ops.data_ptr() if ops is not None else 0)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = ctx.launches
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local) as clocks:
        barrier()
        evs[0].record(stream)
        for i in range(args.steps):
            step()
            evs[i + 1].record(stream)
        barrier()
    clock_kernel = clocks.summary()
    total_ms = evs[0].elapsed_time(evs[-1])
    per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    gpu_launches = ctx.launches - launches0
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * S * T * args.steps / (total_ms_max * 1e-3)

    # sanity: the timed kernels produced real PCM
    peak = float(pcm.abs().max())
    assert 0.05 < peak < 50 and bool(torch.isfinite(pcm).all()), peak

    # --- end to end through the host-buffer C-ABI call (H2D + D2H inside the timed region) ---
    e2e = None
    if not args.no_e2e:
        h_spec = torch.empty((S, T, C, 1024), dtype=torch.float32, pin_memory=True)
        h_spec.copy_(spectra)
        h_pcm = torch.empty((S, T, 1024, C), dtype=torch.float32, pin_memory=True)
        spec_np, pcm_np = h_spec.numpy(), h_pcm.numpy()
        ctx2 = A.Context(S, C, side["sample_index"], side["flags"], device=local)
        e2e_steps = args.e2e_steps or max(2, min(args.steps, 20))
        for _ in range(2):
            ctx2.process(spec_np, info_np, side["tns_blob"], side["tns_offsets"], out=pcm_np, stereo_ops=ops_np)
        barrier()
        with ClockSampler(local) as clocks_e2e:
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                ctx2.process(spec_np, info_np, side["tns_blob"], side["tns_offsets"], out=pcm_np, stereo_ops=ops_np)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        ce = clocks_e2e.summary()
        if ce["samples"]:  # the e2e loop is long enough for nvidia-smi's 100 ms sampling: merge
            clock_kernel = {"sm_mhz": ce["sm_mhz"] if not clock_kernel["samples"] else clock_kernel["sm_mhz"],
                            "sm_max_mhz": ce["sm_max_mhz"],
                            "reasons": sorted(set(ce["reasons"]) | set(clock_kernel["reasons"])),
                            "samples": clock_kernel["samples"], "samples_e2e": ce["samples"], "sm_mhz_e2e": ce["sm_mhz"]}
        te = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(te, op=torch.distributed.ReduceOp.MAX)
        e2e_val = world * S * T * e2e_steps / float(te.item())
        assert np.isfinite(pcm_np[0, 0]).all() and np.abs(pcm_np[-1, -1]).max() > 0
        side_bytes = info_np.nbytes + (tns_bytes + side["tns_offsets"].nbytes if tns_bytes else 0) + stereo_bytes
        e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(spec_np.nbytes + side_bytes),
               "d2h_bytes_per_step": int(pcm_np.nbytes), "steps": e2e_steps,
               "ms_per_step": float(te.item()) / e2e_steps * 1e3,
               "api": "aacfb_process (pinned host buffers, 2-lane copy/compute pipeline)",
               "numa": f"process bound to the {numa_cpus} CPUs local to the GPU" if numa_cpus else "no binding"}
        ctx2.close()

    if rank == 0:
        peak_gbs, peak_src = measured_peak()
        # synthesis-kernel launch time: each step is one memset + (tns_kernel) + synth_kernel on one
        # stream; event-to-event time of a step is the launch duration the roofline uses
        launch_ms = float(np.mean(per_step))
        alg = algorithmic_bytes(S, T, C, tns_bytes, stereo_bytes)
        achieved = alg / (launch_ms * 1e-3) / 1e9
        traffic = None
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get(args.workload)
            except Exception:
                traffic = None
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
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "streams_per_gpu": S, "frames_per_stream": T,
                       "channels": C, "parallelism": f"streams sharded over {world} GPU(s), no data-path collective",
                       "l2": "inputs+outputs 1 GiB per step >> 126 MB L2 (no flush needed)"},
            "clocks": clock_kernel,
            "e2e": e2e,
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                         "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "aacfb::synth_kernel" + (" (+ aacfb::tns_kernel)" if tns_bytes else ""),
                         "algorithmic_bytes_per_launch": alg,
                         "launch_ms": launch_ms},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=0)
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed steps of the e2e leg (default min(steps, 20))")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process to the GPU's NUMA node")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
