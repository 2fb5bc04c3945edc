"""Narrow down the synth_tns_kernel mismatch: sub-cases cut out of the big random TNS case."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aacjs_b200 as A  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tools import workloads as W  # noqa: E402

S, T, C, mode = 37, 33, 2, 1
w = W.random_case(S, T, C, np.random.default_rng(10 * mode + S), tns_mode=mode)


def cut(s0, s1, t0, t1, edit=None):
    spec = np.ascontiguousarray(w["spectra"][s0:s1, t0:t1])
    info = np.ascontiguousarray(w["info"][s0:s1, t0:t1])
    blobs, offs = [], [0]
    for s in range(s0, s1):
        for t in range(t0, t1):
            for c in range(C):
                cf = (s * T + t) * C + c
                b = bytearray(w["tns_blob"][w["tns_offsets"][cf]:w["tns_offsets"][cf + 1]].tobytes())
                if edit:
                    b = edit(s, t, c, b)
                blobs.append(bytes(b)); offs.append(offs[-1] + len(b))
    blob = np.frombuffer(b"".join(blobs) + b"\0" * 16, np.uint8)[: offs[-1]].copy()
    return spec, info, blob, np.array(offs, np.uint32)


def run(name, spec, info, blob, offs):
    s_, t_ = spec.shape[:2]
    ref, _ = O.process(spec, info, blob, offs, np.zeros((s_, C, 1024), np.float32), sample_index=w["sample_index"], flags=w["flags"], n_threads=4)
    ctx = A.Context(s_, C, w["sample_index"], w["flags"])
    got = ctx.process(spec, info, blob, offs)
    ctx.close()
    err = np.abs(np.nan_to_num(got.astype(np.float64) - ref)).max(axis=2)
    bad = np.argwhere(err > 1e-4)
    print(f"{name}: shape {spec.shape[:2]} max err {err.max():.3e} bad {[tuple(int(v) for v in b) for b in bad[:6]]}", flush=True)


def drop_second(s, t, c, b):
    if (s, t, c) == (15, 26, 1):
        b[0] = 1
        return b[: 8 + 4 + 4 * b[9]]
    return b


def drop_b(s, t, c, b):
    if (s, t, c) == (15, 26, 1):
        b[0] = 0
        return b[:8]
    return b


run("stream 15, frames 24..32", *cut(15, 16, 24, 33))
run("stream 15, frame 26 only", *cut(15, 16, 26, 27))
run("stream 15, frames 25..27", *cut(15, 16, 25, 28))
run("frame 26, B second filter dropped", *cut(15, 16, 26, 27, drop_second))
run("frame 26, B without filters", *cut(15, 16, 26, 27, drop_b))
run("streams 14..16 all frames", *cut(14, 17, 0, 33))
