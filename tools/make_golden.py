"""Generate the golden fixtures under tests/golden/ from the CPU oracle.

    python tools/make_golden.py

Inputs are regenerated from seeds by tools/workloads.py; only the expected PCM (and the
final overlap) are stored, as float32 .npz.  The reference itself cannot run here (no JS
engine), so these pin the *oracle* (a restatement of the reference) against drift; the
vectors produced by interpreting the reference's own source are made by
tools/js_reference.py (tests/golden/jsref_*.npz).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tools import workloads as W  # noqa: E402

CASES = {
    # name: (config, S, T, C, seed, shape_prev_mode)
    "config1_mono_long": (1, 1, 1, 1, 0, "as_shipped"),
    "config2_long": (2, 2, 6, 2, 1, "as_shipped"),
    "config3_short": (3, 2, 6, 2, 2, "as_shipped"),
    "config4_tns_ar": (4, 2, 6, 2, 3, "as_shipped"),
    "config5_mixed": (5, 2, 32, 2, 4, "carried"),
}


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, (cfg, S, T, C, seed, spm) in CASES.items():
        w = W.make(cfg, S, T, C, seed, spm)
        pcm, ov = O.process(w["spectra"], w["info"], w["tns_blob"], w["tns_offsets"],
                            sample_index=w["sample_index"], flags=w["flags"])
        np.savez_compressed(os.path.join(out_dir, f"oracle_{name}.npz"), pcm=pcm, overlap=ov,
                            meta=np.array([cfg, S, T, C, seed, int(spm == "carried")]))
        print(name, pcm.shape, float(np.abs(pcm).max()))
    rng = np.random.default_rng(1234)
    for mode, tag in ((1, "ar"), (2, "ma")):
        w = W.random_case(2, 10, 3, rng, tns_mode=mode)
        pcm, ov = O.process(w["spectra"], w["info"], w["tns_blob"], w["tns_offsets"], sample_index=4, flags=mode)
        np.savez_compressed(os.path.join(out_dir, f"oracle_random_{tag}.npz"), pcm=pcm, overlap=ov,
                            spectra=w["spectra"], info=w["info"].view(np.uint8), tns_blob=w["tns_blob"],
                            tns_offsets=w["tns_offsets"])
        print("random", tag, pcm.shape)


if __name__ == "__main__":
    main()
