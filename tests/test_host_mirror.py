"""The Python mirror of the reference's JS seam, driven the way decoder.js drives it:
FilterBank.process per channel (decoder.js:269,318-319), tns.process (decoder.js:264,310-313),
and the batched readChunk replacement."""
import numpy as np
import pytest

import aacjs_b200 as A
from oracle import oracle as O
from tools import workloads as W

pytestmark = pytest.mark.gpu


class Info:  # a JS-like ICSInfo (ics.js:270-332)
    def __init__(self, seq, shape_prev, shape_cur):
        self.windowSequence = seq
        self.windowShape = [shape_prev, shape_cur]


def test_filterbank_process_like_decoder_js():
    rng = np.random.default_rng(0)
    fb = A.FilterBank(False, 2)
    seqs = [0, 1, 2, 2, 3, 0]
    ov = np.zeros((2, 1024), np.float32)
    for t, sq in enumerate(seqs):
        for ch in range(2):
            x = (rng.standard_normal(1024) * 1e5).astype(np.float32)
            keep = x.copy()
            out = np.full(1024, np.nan, np.float32)
            info = Info(sq, t & 1, (t + ch) & 1)
            fb.process(info, x, out, ch)
            ref = O.filterbank(O.make_info(sq, t & 1, (t + ch) & 1), keep, ov[ch])
            assert np.array_equal(x, keep)                      # does not mutate input
            assert np.abs(out - ref).max() / 32768 <= 1e-5      # un-scaled samples, decoder.js:269
    assert np.abs(fb.overlaps - ov).max() / 32768 <= 1e-5


def test_filterbank_unknown_sequence_is_a_noop():
    """filter_bank.js:104-203 has no default case: output keeps its zeros, overlap untouched."""
    fb = A.FilterBank(False, 1)
    out = np.ones(1024, np.float32)
    fb.process({"windowSequence": 5, "windowShape": [0, 0]}, np.ones(1024, np.float32), out, 0)
    assert not out.any() and not fb.overlaps.any()


@pytest.mark.parametrize("decode", [True, False])
def test_tns_process_like_decoder_js(decode):
    rng = np.random.default_rng(1)
    t = A.TNS({"sampleIndex": 4})
    t.nFilt[0] = 2
    t.length[0][:2] = [20, 25]
    t.order[0][:2] = [7, 12]
    t.direction[0][:2] = [False, True]
    mild = np.asarray(W.TNS_COEF_0_4, np.float32)[W.MILD]
    t.coef[0][0][:7] = mild[rng.integers(0, 7, 7)]
    t.coef[0][1][:12] = mild[rng.integers(0, 7, 12)]
    data = (rng.standard_normal(1024) * 1e4).astype(np.float32)
    ics = {"info": Info(0, 0, 0), "maxSFB": 45}
    ref = O.tns(O.make_info(0, 0, 0, max_sfb=45, tns_present=1), t.block(), 4, O.TNS_FIXED_AR if decode else O.TNS_FIXED_MA, data)
    got = data.copy()
    t.process(ics, got, decode)
    assert not np.array_equal(got, data)
    assert np.abs(got - ref).max() <= 1e-5 * 32768 * 0.01
    # as shipped (tns.js:122) the reference's TNS is the identity
    t2 = A.TNS({"sampleIndex": 4}, mode=A.TNS_AS_SHIPPED)
    t2.nFilt[:], t2.length[:], t2.order[:], t2.coef[:] = t.nFilt, t.length, t.order, t.coef
    same = data.copy()
    t2.process(ics, same, decode)
    assert np.array_equal(same, data)


def test_decoder_readchunks_matches_per_frame_reference_flow():
    """K frames through the batched replacement == K times (tns.process, filter_bank.process,
    interleave /32768) of decoder.js, channel by channel."""
    rng = np.random.default_rng(2)
    Cn, K = 2, 9
    seqs = [0, 0, 1, 2, 3, 0, 1, 3, 0]
    dec = A.AACDecoder(Cn, 4, tns_mode=A.TNS_FIXED_AR)
    frames, ov, expect = [], np.zeros((Cn, 1024), np.float32), []
    mild = np.asarray(W.TNS_COEF_0_4, np.float32)[W.MILD]
    for t in range(K):
        fr, chans = [], []
        for ch in range(Cn):
            data = (rng.standard_normal(1024) * 5e4).astype(np.float32)
            tns = A.TNS({"sampleIndex": 4})
            present = bool((t + ch) & 1) and seqs[t] != 2
            if present:
                tns.nFilt[0] = 1
                tns.length[0][0], tns.order[0][0], tns.direction[0][0] = 30, 9, bool(t & 2)
                tns.coef[0][0][:9] = mild[rng.integers(0, 7, 9)]
            info = Info(seqs[t], 0, t & 1)
            fr.append({"info": info, "data": data, "tnsPresent": present, "tns": tns, "maxSFB": 40})
            rec = O.make_info(seqs[t], 0, t & 1, max_sfb=40, tns_present=int(present))
            d = O.tns(rec, tns.block(), 4, O.TNS_FIXED_AR, data) if present else data
            chans.append(O.filterbank(rec, d, ov[ch]))
        frames.append(fr)
        expect.append((np.stack(chans, 1).astype(np.float64) / 32768).astype(np.float32).reshape(-1))
    got = dec.readChunks(frames)
    assert got.shape == (K * 1024 * Cn,)
    assert np.abs(got - np.concatenate(expect)).max() <= 1e-5
