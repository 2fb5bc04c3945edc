"""Oracle against the committed golden fixtures (tools/make_golden.py).  Pins the oracle
(and, through test_emulation / test_gpu_parity, the kernels) against drift."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as O
from tools import workloads as W

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load_case(path):
    z = np.load(path)
    if "spectra" in z:
        w = dict(spectra=z["spectra"], info=z["info"].view(W.INFO_DTYPE).reshape(z["spectra"].shape[:3]),
                 tns_blob=z["tns_blob"], tns_offsets=z["tns_offsets"], sample_index=4,
                 flags=1 if path.endswith("_ar.npz") else 2)
    else:
        cfg, S, T, C, seed, carried = (int(v) for v in z["meta"])
        w = W.make(cfg, S, T, C, seed, "carried" if carried else "as_shipped")
    return w, z["pcm"], z["overlap"]


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "oracle_*.npz"))), ids=os.path.basename)
def test_oracle_reproduces_golden(path):
    w, pcm, ov = load_case(path)
    got, gov = O.process(w["spectra"], w["info"], w["tns_blob"], w["tns_offsets"], sample_index=w["sample_index"],
                         flags=w["flags"])
    assert np.array_equal(got.view(np.uint32), pcm.view(np.uint32))
    assert np.array_equal(gov.view(np.uint32), ov.view(np.uint32))


def test_threaded_oracle_equals_single_thread():
    w = W.make(5, 8, 20, 2, seed=3)
    a, oa = O.process(w["spectra"], w["info"], sample_index=4, n_threads=1)
    b, ob = O.process(w["spectra"], w["info"], sample_index=4, n_threads=4)
    assert np.array_equal(a, b) and np.array_equal(oa, ob)
