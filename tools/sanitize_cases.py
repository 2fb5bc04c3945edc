"""Small invocations of every kernel instantiation, for `compute-sanitizer --tool racecheck|synccheck|memcheck`
(tools/sanitize.sh).  numpy + ctypes only (no torch: the sanitizer instruments every kernel of the process).
Each case is also checked against the CPU oracle, so a race that changes bits shows up twice."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aacjs_b200 as A  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tools import workloads as W  # noqa: E402

which = sys.argv[1:] or ["config1", "config2", "config3", "config5", "tns_ar", "tns_ma", "stereo", "stereo_tns", "surround",
                         "q16_f32", "q16_s16", "q16_s16_tns", "q16_mono"]
rng = np.random.default_rng(11)
worst = 0.0
def quantised(name):
    """aacfb_qframe input (dequant_stage in the IOV instantiations, dequant_kernel before a TNS pass) and int16 PCM."""
    global worst
    S, T, C = (3, 7, 1) if name == "q16_mono" else (4, 9, 2)
    case = W.random_q_case(S, T, C, rng, tns_mode=1 if name.endswith("_tns") else 0, p_noise=0.0)
    q = case["qframes"]
    q["q"][np.abs(q["q"]) == 8191] = 8190
    q["band"] = np.where((q["band"] & 0x1ff) > 427, (q["band"] & 0xc000) | 200, q["band"])
    fmt = A.PCM_F32 if name == "q16_f32" else A.PCM_S16
    kw = dict(sample_index=case["sample_index"], flags=case["flags"])
    ov = np.zeros((S, C, 1024), np.float32)
    ref, _ = O.process_io(q, O.IN_Q16, case["info"], case["tns_blob"], case["tns_offsets"], ov, pcm_format=fmt, **kw)
    ref2, _ = O.process_io(q, O.IN_Q16, case["info"], case["tns_blob"], case["tns_offsets"], ov, pcm_format=fmt, **kw)
    ctx = A.Context(S, C, case["sample_index"], case["flags"], device=0)
    got = ctx.process_io(q, case["info"], case["tns_blob"], case["tns_offsets"], in_format=A.IN_Q16, pcm_format=fmt)
    got2 = ctx.process_io(q, case["info"], case["tns_blob"], case["tns_offsets"], in_format=A.IN_Q16, pcm_format=fmt)
    f32, _ = O.process_io(q, O.IN_Q16, case["info"], case["tns_blob"], case["tns_offsets"], np.zeros_like(ov), **kw)
    peak = max(1.0, float(np.abs(f32).max()))
    if fmt == A.PCM_F32:
        e = max(float(np.abs(got - ref).max()), float(np.abs(got2 - ref2).max())) / peak
    else:   # 1 LSB (x the overshoot of full scale) is the bar for int16: report it on the 1e-5 scale
        lsb = max(int(np.abs(got.astype(np.int32) - ref).max()), int(np.abs(got2.astype(np.int32) - ref2).max()))
        e = 0.0 if lsb <= max(1, int(peak)) else 1.0
    worst = max(worst, e)
    print(f"{name}: S={S} T={T} C={C} launches={ctx.launches} error = {e:.3e}", flush=True)
    ctx.close()


for name in which:
    ops = None
    if name.startswith("q16"):
        quantised(name)
        continue
    if name in ("config1", "config2", "config3", "config5"):
        cfg = int(name[-1])
        S, T, C = (1, 1, 1) if cfg == 1 else (5, 19, 2)
        case = W.make(cfg, S, T, C, seed=3, shape_prev_mode="carried")
    elif name in ("tns_ar", "tns_ma"):
        case = W.random_case(4, 9, 2, rng, tns_mode=1 if name == "tns_ar" else 2)
    elif name == "surround":
        case = W.random_case(2, 6, 5, rng, tns_mode=0)
    else:
        # (TNS amplifies rounding differences: the amplitude the GPU tests use for this combination)
        case = W.random_stereo_case(3, 8, rng, tns_mode=1 if name == "stereo_tns" else 0,
                                    **(dict(sigma=2e4) if name == "stereo_tns" else {}))
        S, T = case["cpe"].shape
        ops = np.zeros((S, T, 1), A.STEREO_DTYPE)
        for s in range(S):
            for t in range(T):
                _, present = A.pack_stereo(case["cpe"][s, t], case["sample_index"], out=ops[s, t, 0])
                case["info"]["stereo_present"][s, t, 0] = int(present)
    S, T, C, _ = case["spectra"].shape
    ctx = A.Context(S, C, case["sample_index"], case["flags"], device=0)
    got = ctx.process(case["spectra"], case["info"], case["tns_blob"], case["tns_offsets"], stereo_ops=ops)
    got2 = ctx.process(case["spectra"], case["info"], case["tns_blob"], case["tns_offsets"], stereo_ops=ops)
    ov = np.zeros((S, C, 1024), np.float32)
    kw = dict(sample_index=case["sample_index"], flags=case["flags"])
    spec_ref = case["spectra"]
    if ops is not None:   # the oracle's processMS / processIS on every pair-frame, then the plain path
        spec_ref = case["spectra"].copy()
        for s in range(S):
            for t in range(T):
                spec_ref[s, t, 0], spec_ref[s, t, 1] = O.stereo(case["cpe"][s, t], case["sample_index"],
                                                                case["spectra"][s, t, 0], case["spectra"][s, t, 1])
    ref, _ = O.process(spec_ref, case["info"], case["tns_blob"], case["tns_offsets"], ov, **kw)
    ref2, _ = O.process(spec_ref, case["info"], case["tns_blob"], case["tns_offsets"], ov, **kw)
    peak = max(1.0, float(np.abs(ref).max()))   # stereo cases: random intensity scales push the PCM far above full scale
    e = max(float(np.abs(got - ref).max()), float(np.abs(got2 - ref2).max())) / peak
    worst = max(worst, e)
    print(f"{name}: S={S} T={T} C={C} launches={ctx.launches} max|pcm - oracle| = {e:.3e}", flush=True)
    ctx.close()
print("WORST", worst)
sys.exit(0 if worst <= 1e-5 else 1)
