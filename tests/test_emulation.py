"""The kernel's own phase code (aacfb_core.cuh / aacfb_worker.cuh), executed on the CPU by
64 host threads per worker, against the oracle.  This is how index maps, swizzles, window
switching, chunk/halo logic and the TNS chains are checked where no GPU exists; the GPU tier
(test_gpu_parity.py) repeats it through the C-ABI on the real kernels."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import emul
from tools import workloads as W

TOL = 1e-5  # BASELINE.json: PCM within 1e-5 max-abs of the reference


def run_both(w, S, T, C, chunk, seed=0):
    rng = np.random.default_rng(seed)
    ov0 = (rng.standard_normal((S, C, 1024)) * 0.25 * 32768).astype(np.float32)
    ovo = ov0.copy()
    ref, _ = O.process(w["spectra"], w["info"], w["tns_blob"], w["tns_offsets"], ovo,
                       sample_index=w["sample_index"], flags=w["flags"])
    ove = ov0.copy()
    got = emul.process(w["spectra"], w["info"], w["tns_blob"], w["tns_offsets"], ove, w["sample_index"],
                       w["flags"], chunk)
    return got, ref, ove, ovo


@pytest.mark.parametrize("cfg,S,T,C,chunk", [(1, 1, 1, 1, 8), (2, 2, 5, 2, 8), (2, 3, 7, 2, 3), (3, 2, 5, 2, 2),
                                              (4, 2, 5, 2, 4), (5, 1, 34, 2, 5), (5, 3, 17, 1, 4), (5, 1, 18, 3, 6)])
def test_baseline_configs(cfg, S, T, C, chunk):
    w = W.make(cfg, S, T, C, seed=cfg, shape_prev_mode="carried")
    got, ref, ove, ovo = run_both(w, S, T, C, chunk)
    assert np.abs(got.astype(np.float64) - ref).max() <= TOL
    assert np.abs(ove - ovo).max() / 32768 <= TOL


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("S,T,C,chunk", [(2, 6, 2, 2), (3, 4, 5, 3), (1, 9, 1, 32)])
def test_random_sequences_shapes_and_tns(mode, S, T, C, chunk):
    rng = np.random.default_rng(100 * mode + S + T)
    w = W.random_case(S, T, C, rng, tns_mode=mode)
    got, ref, ove, ovo = run_both(w, S, T, C, chunk)
    assert not np.isnan(ref).any()
    assert np.abs(got.astype(np.float64) - ref).max() <= TOL
    assert np.abs(ove - ovo).max() / 32768 <= TOL


def test_chunking_does_not_change_the_result():
    """Halo recomputation must make any cut of the time axis give identical bits."""
    w = W.make(5, 2, 20, 2, seed=9, shape_prev_mode="carried")
    outs = []
    for chunk in (1, 3, 7, 20):
        ov = np.zeros((2, 2, 1024), np.float32)
        outs.append((emul.process(w["spectra"], w["info"], None, None, ov, 4, 0, chunk), ov))
    for pcm, ov in outs[1:]:
        assert np.array_equal(pcm.view(np.uint32), outs[0][0].view(np.uint32))
        assert np.array_equal(ov.view(np.uint32), outs[0][1].view(np.uint32))


def test_ma_order_20_reproduces_the_reference_nan():
    """tns.js:43,169: tmp has 20 slots, so the MA branch with order 20 turns samples m >= 20 into NaN."""
    x = np.random.default_rng(0).standard_normal((1, 1, 1, 1024)).astype(np.float32)
    info = np.zeros((1, 1, 1), W.INFO_DTYPE)
    info["max_sfb"], info["tns_present"] = 49, 1
    coef = np.full(20, 0.18374951, np.float32)
    blob, offs = W.pack_tns([W.tns_block([1, 0, 0, 0, 0, 0, 0, 0], [(49, 20, 0, coef)])])
    ref, _ = O.process(x, info, blob, offs, sample_index=4, flags=2)
    got = emul.process(x, info, blob, offs, np.zeros((1, 1, 1024), np.float32), 4, 2, 8)
    assert np.isnan(ref).any() and np.array_equal(np.isnan(ref), np.isnan(got))
