"""Stereo tools of a channel pair element: processMS / processIS,
reference src/decoder.js:337-404, moved onto the staged spectra.

CPU part: the oracle's restatement reproduces, bit for bit, what the reference's own two functions
produced when run by tools/jsmini.py (tests/golden/stereo/jsref_stereo.npz, made by
`python tools/js_reference.py stereo`; live as well where /root/reference exists), and the host's
op-table packer (aacjs_b200.pack_stereo, twin of js/stereo_pack.js) describes exactly that.
GPU part: the kernels applying the op table, through the C-ABI, against the oracle."""
import os

import numpy as np
import pytest

import aacjs_b200 as A
from oracle import oracle as O
from tools import workloads as W

GOLD = os.path.join(os.path.dirname(__file__), "golden", "stereo", "jsref_stereo.npz")
HAVE_REF = os.path.isdir("/root/reference/src")
TOL = 1e-5


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def fixture():
    from tools.js_reference import stereo_elements

    z = np.load(GOLD)
    n, seed, S, T, seed2 = (int(v) for v in z["meta"])
    cpe, left, right = stereo_elements(n, seed)
    return z, cpe, left, right, W.random_stereo_case(S, T, np.random.default_rng(seed2))


def apply_ops(rec, left, right):
    """numpy model of what the device does with one aacfb_stereo_ops record (include/aacfb.h)."""
    l, r = left.copy(), right.copy()
    op = np.repeat(rec["op"], 4)
    ms, it = op == A.STEREO_MS, op >= A.STEREO_IS
    l[ms] = left[ms] + right[ms]
    r[ms] = left[ms] - right[ms]
    r[it] = left[it] * rec["scale"][op[it] - A.STEREO_IS]
    return l, r


def oracle_stereo_spectra(case):
    sp = case["spectra"].copy()
    S, T = sp.shape[:2]
    for s in range(S):
        for t in range(T):
            sp[s, t, 0], sp[s, t, 1] = O.stereo(case["cpe"][s, t], case["sample_index"], sp[s, t, 0], sp[s, t, 1])
    return sp


def pack_case(case):
    """What the JS host hands over: info with stereo_present on the left channels + the op records."""
    S, T = case["spectra"].shape[:2]
    ops = np.zeros((S, T, 1), A.STEREO_DTYPE)
    info = case["info"].copy()
    for s in range(S):
        for t in range(T):
            _, present = A.pack_stereo(case["cpe"][s, t], case["sample_index"], out=ops[s, t, 0])
            info["stereo_present"][s, t, 0] = int(present)
    return info, ops


# ------------------------------------------------------------------------------ CPU
def test_oracle_equals_interpreted_reference_bit_for_bit():
    z, cpe, left, right, case = fixture()
    touched = 0
    for k in range(len(cpe)):
        l, r = O.stereo(cpe[k], 4, left[k], right[k])
        assert np.array_equal(bits(l), bits(z["out_left"][k])) and np.array_equal(bits(r), bits(z["out_right"][k]))
        touched += int((l != left[k]).any()) + int((r != right[k]).any())
    assert touched > len(cpe) // 2   # the vectors do exercise both tools
    pcm, ovl = O.process(oracle_stereo_spectra(case), case["info"], sample_index=4)
    assert np.array_equal(bits(pcm), bits(z["pcm"])) and np.array_equal(bits(ovl), bits(z["overlap"]))


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference is not on this machine")
def test_oracle_equals_reference_live():
    from tools.js_reference import STEREO_SEQS, Reference

    ref, rng = Reference(), np.random.default_rng(2024)
    for k in range(40):
        a, b = STEREO_SEQS[k % len(STEREO_SEQS)]
        si = int(rng.integers(0, 12))
        e = W.random_cpe(rng, a, b, sample_index=si)
        l = (rng.standard_normal(1024) * 1e4).astype(np.float32)
        r = (rng.standard_normal(1024) * 1e4).astype(np.float32)
        jl, jr = ref.stereo(e, si, l, r)
        ol, orr = O.stereo(e, si, l, r)
        assert np.array_equal(bits(jl), bits(ol)) and np.array_equal(bits(jr), bits(orr))


def test_packed_ops_describe_the_reference_walk():
    """pack_stereo + the op semantics of aacfb.h == processMS / processIS, for every sample rate."""
    rng = np.random.default_rng(7)
    seqs = [(0, 0), (2, 2), (0, 2), (2, 0), (1, 1), (3, 3)]
    n_ms = n_is = 0
    for k in range(120):
        si = k % 12
        e = W.random_cpe(rng, *seqs[k % 6], sample_index=si, p_intensity=0.4)
        l = (rng.standard_normal(1024) * 1e4).astype(np.float32)
        r = (rng.standard_normal(1024) * 1e4).astype(np.float32)
        rec, present = A.pack_stereo(e, si)
        pl, pr = apply_ops(rec, l, r)
        ol, orr = O.stereo(e, si, l, r)
        assert np.array_equal(bits(pl), bits(ol)) and np.array_equal(bits(pr), bits(orr))
        assert present == bool((rec["op"] != 0).any())
        n_ms += int((rec["op"] == 1).sum())
        n_is += int((rec["op"] >= 2).sum())
    assert n_ms > 500 and n_is > 500


def test_js_packer_equals_python_twin():
    """js/stereo_pack.js (what the Node host runs) executed by the interpreter writes the same record."""
    from tools import jsmini as J

    js_dir = os.path.join(os.path.dirname(os.path.dirname(__file__)), "aac.js_b200", "js")
    rt = J.Runtime(js_dir, stubs={"aac/src/ics": J.obj(NOISE_BT=13, INTENSITY_BT2=14, INTENSITY_BT=15)})
    pack = rt.require("./stereo_pack").get("pack")
    rng = np.random.default_rng(11)
    for k in range(30):
        a, b = [(0, 0), (2, 2), (0, 2), (2, 0), (1, 1)][k % 5]
        e = W.random_cpe(rng, a, b, sample_index=4, p_intensity=0.4)

        def ics(c):
            info = J.obj(groupCount=int(e["group_count"][c]), maxSFB=int(e["max_sfb"][c]),
                         groupLength=J.int32array(e["group_length"][c]),
                         swbOffsets=J.JSTyped("Uint16Array", A.swb_offsets(4, e["window_sequence"][c] == 2)))
            return J.obj(info=info, bandTypes=J.int32array(e["band_types"][c]), sectEnd=J.int32array(e["sect_end"][c]),
                         scaleFactors=J.float32array(e["scale_factors"][c]))
        element = J.obj(commonWindow=bool(e["common_window"]), maskPresent=bool(e["mask_present"]),
                        ms_used=J.JSArray([bool(v) for v in e["ms_used"]]), left=ics(0), right=ics(1))
        ops, scales = J.JSTyped("Uint8Array", np.full(256, 9, np.uint8)), J.float32array(np.zeros(128))
        present = pack.call(J.UNDEF, [element, ops, scales])
        rec, want = A.pack_stereo(e, 4)
        assert present is want
        assert np.array_equal(ops.a, rec["op"])
        n = int(rec["op"].max()) - 1
        assert n <= 0 or np.array_equal(bits(scales.a[:n]), bits(rec["scale"][:n]))


def test_swb_offsets_are_multiples_of_four_and_match_the_counts():
    """The op table works on groups of 4 coefficients: every band edge must be one (tables.js:34-124)."""
    for si in range(12):
        for short, counts in ((False, W.SWB_LONG_COUNT), (True, W.SWB_SHORT_COUNT)):
            off = A.swb_offsets(si, short)
            assert len(off) == counts[si] + 1 and off[0] == 0 and off[-1] == (128 if short else 1024)
            assert (off % 4 == 0).all() and (np.diff(off.astype(int)) > 0).all()


# ------------------------------------------------------------------------------ GPU
def gpu_vs_oracle(case, S, T, flags):
    info, ops = pack_case(case)
    ref, rov = O.process(oracle_stereo_spectra(case), case["info"], case["tns_blob"], case["tns_offsets"],
                         sample_index=case["sample_index"], flags=flags, n_threads=8)
    ctx = A.Context(S, 2, case["sample_index"], flags)
    got = ctx.process(case["spectra"], info, case["tns_blob"], case["tns_offsets"], stereo_ops=ops)
    gov = ctx.get_overlap()
    launches = ctx.launches
    ctx.close()
    m = ~np.isnan(ref)
    assert np.array_equal(np.isnan(got), ~m)
    tol = TOL * max(1.0, float(np.abs(ref[m]).max()))
    assert np.abs(got[m].astype(np.float64) - ref[m]).max() <= tol
    assert np.abs(gov - rov)[~np.isnan(rov)].max() / 32768 <= tol
    return launches


@pytest.mark.gpu
@pytest.mark.parametrize("S,T", [(1, 1), (3, 9), (37, 33)])
def test_gpu_stereo_tools_fused_into_synthesis(S, T):
    """No TNS pass: synth_kernel applies the ops to the staged rows (no pre-pass: the long-only
    instantiation, plus the generic one unless the host-side check saw no EIGHT_SHORT frame)."""
    case = W.random_stereo_case(S, T, np.random.default_rng(100 + S), sigma=3e4)
    launches = gpu_vs_oracle(case, S, T, A.TNS_AS_SHIPPED)
    if S * T * 2 * 4096 < (8 << 20):   # one sub-batch (aacfb_process splits larger ones)
        assert launches == (2 if (case["info"]["window_sequence"] == 2).any() else 1)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [A.TNS_FIXED_AR, A.TNS_FIXED_MA])
def test_gpu_stereo_tools_before_tns(mode):
    """TNS runs between the stereo tools and the IMDCT (decoder.js:300-319): stereo pre-pass, TNS pre-pass,
    long-only and generic instantiation.  AACFB_TNS_FUSED=1 (opt-in): stereo pre-pass kernel, then the synthesis
    kernel that filters its own rows (synth_tns_kernel; with an EIGHT_SHORT frame in the batch also the TNS
    pre-pass and the generic instantiation for the items that have one)."""
    case = W.random_stereo_case(5, 11, np.random.default_rng(200 + mode), tns_mode=mode, sigma=2e4)
    any_short = (case["info"]["window_sequence"] == 2).any()
    fused = os.environ.get("AACFB_TNS_FUSED", "0") != "0"
    assert gpu_vs_oracle(case, 5, 11, mode) == ((4 if any_short else 2) if fused else (4 if any_short else 3))


@pytest.mark.gpu
def test_gpu_stereo_golden_from_the_reference():
    z, cpe, left, right, case = fixture()
    S, T = case["spectra"].shape[:2]
    info, ops = pack_case(case)
    ctx = A.Context(S, 2, 4, 0)
    got = ctx.process(case["spectra"], info, stereo_ops=ops)
    ctx.close()
    assert np.abs(got.astype(np.float64) - z["pcm"]).max() <= TOL * max(1.0, float(np.abs(z["pcm"]).max()))


@pytest.mark.gpu
def test_gpu_stereo_records_without_flag_are_ignored_and_bad_ops_rejected():
    case = W.random_stereo_case(2, 5, np.random.default_rng(5), sigma=3e4)
    info, ops = pack_case(case)
    plain = case["info"].copy()            # stereo_present = 0 everywhere: the records must not be read
    ctx = A.Context(2, 2, 4, 0)
    a = ctx.process(case["spectra"], plain, stereo_ops=ops)
    ctx.reset()
    b = ctx.process(case["spectra"], plain)
    assert np.array_equal(a, b)
    bad = ops.copy()
    s, t = np.argwhere(info["stereo_present"][:, :, 0] != 0)[0]
    bad["op"][s, t, 0, 3] = 200
    with pytest.raises(A.AacfbError):
        ctx.process(case["spectra"], info, stereo_ops=bad)
    wrong = info.copy()
    wrong["stereo_present"][0, 0, 1] = 1   # the flag belongs to the left channel
    with pytest.raises(A.AacfbError):
        ctx.process(case["spectra"], wrong, stereo_ops=ops)
    ctx.close()


# ------------------------------------------------- the reference's own decoder.process as the driver
DEC_GOLD = os.path.join(os.path.dirname(__file__), "golden", "stereo", "jsref_decoder_process.npz")


def decoder_fixture():
    from tools.js_reference import decoder_case

    z = np.load(DEC_GOLD)
    seed, T, mseed, mT = (int(v) for v in z["meta"])
    return z, decoder_case(seed, T), W.make(5, 1, mT, 1, seed=mseed, shape_prev_mode="carried")


def test_oracle_equals_unmodified_decoder_process_bit_for_bit():
    """tests/golden/stereo/jsref_decoder_process.npz was produced by the reference's UNMODIFIED decoder.js
    (AACDecoder.prototype.process: processPair / processSingle -> processMS, processIS, tns.process,
    filter_bank.process, then the interleave of readChunk) on elements built with the reference's own
    ICStream / CPEElement constructors (tools/js_reference.DecoderReference).  The oracle's restatement
    composed the same way reproduces every bit of PCM and overlap, stereo and mono."""
    z, case, mono = decoder_fixture()
    pcm, ovl = O.process(oracle_stereo_spectra(case), case["info"], sample_index=4)
    assert np.array_equal(bits(pcm[0]), bits(z["pcm"])) and np.array_equal(bits(ovl[0]), bits(z["overlap"]))
    mp, mo = O.process(mono["spectra"], mono["info"], sample_index=4)
    assert np.array_equal(bits(mp[0]), bits(z["mono_pcm"])) and np.array_equal(bits(mo[0]), bits(z["mono_overlap"]))
    assert float(np.abs(z["pcm"]).max()) > 0.1


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference is not on this machine")
def test_oracle_equals_decoder_process_live():
    from tools.js_reference import DecoderReference, decoder_case

    case = decoder_case(seed=4242, T=6)
    jp, jo = DecoderReference(2).run_stereo_stream(case["spectra"][0], case["info"][0], case["cpe"][0])
    pcm, ovl = O.process(oracle_stereo_spectra(case), case["info"], sample_index=4)
    assert np.array_equal(bits(pcm[0]), bits(jp)) and np.array_equal(bits(ovl[0]), bits(jo))


@pytest.mark.gpu
def test_gpu_equals_unmodified_decoder_process():
    """The CUDA path (stereo tools on the device) against the reference's own decoder.process output."""
    z, case, mono = decoder_fixture()
    info, ops = pack_case(case)
    ctx = A.Context(1, 2, 4, 0)
    got = ctx.process(case["spectra"], info, stereo_ops=ops)
    gov = ctx.get_overlap()
    ctx.close()
    scale = max(1.0, float(np.abs(z["pcm"]).max()))
    assert np.abs(got[0].astype(np.float64) - z["pcm"]).max() <= TOL * scale
    assert np.abs(gov[0] - z["overlap"]).max() / 32768 <= TOL * scale
    ctx = A.Context(1, 1, 4, 0)
    got = ctx.process(mono["spectra"], mono["info"])
    ctx.close()
    assert np.abs(got[0].astype(np.float64) - z["mono_pcm"]).max() <= TOL * max(1.0, float(np.abs(z["mono_pcm"]).max()))
