"""inverse quantisation + scalefactors + PNS on the device (aacfb_qframe input)
and the int16 PCM sink.

Oracle pin: oracle/aacfb_oracle.c:decode_spectral_data restates ICStream.decodeSpectralData
(ics.js:203-266); tests/golden/dequant/jsref_dequant.npz holds 40 records run through the reference's
own, unmodified function by tools/js_reference.DequantReference (Huffman.decodeSpectralData replaced by
a feed of the record's integers) -- every band kind, window groups, out-of-table reads, the PNS
generator as shipped.  The oracle reproduces them BIT FOR BIT (NaN positions and the sign of zeros
included), live too when /root/reference exists.  The kernel's phase code (dequant4 / dequant_stage /
pcm_s16, run on the CPU by the emulation and on the GPU through the C-ABI) is then held to the oracle:
bit-exact for the dequantised spectrum and the lookup tables, 1e-5 on float PCM, 1 LSB on int16 PCM
(the float sample it rounds is itself within 1e-5 * 32768 = 0.33 of the oracle's)."""
import os

import numpy as np
import pytest

import aacjs_b200 as A
from oracle import oracle as O
from tests import emul
from tools import workloads as W

GOLD = os.path.join(os.path.dirname(__file__), "golden", "dequant", "jsref_dequant.npz")


def same_bits(a, b):
    """equal as floats bit for bit, any NaN == any NaN"""
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a[~na].view(np.int32), b[~nb].view(np.int32))


def test_oracle_reproduces_the_reference_decodeSpectralData_bit_for_bit():
    d = np.load(GOLD)
    assert np.isnan(d["data"]).any() and (d["qframes"]["band"] & 0x8000).any()   # PNS and NaN cases are in there
    for q, fi, want in zip(d["qframes"], d["info"], d["data"]):
        assert same_bits(O.dequant(q, fi, int(d["sample_index"])), want)
    assert np.array_equal(O.dequant_table(0), d["iq_table"])      # tables.js:181-191 as the interpreter built it
    assert np.array_equal(O.dequant_table(1), d["sf_table"])      # tables.js:168-176


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference sources not on this box")
def test_oracle_equals_the_reference_live_on_fresh_records():
    from tools import js_reference as R

    rng = np.random.default_rng(2024)
    for si in (4, 8, 11):
        ref = R.DequantReference(sample_index=si)
        case = W.random_q_case(6, 1, 1, rng, sample_index=si, p_noise=0.15)
        for q, fi in zip(case["qframes"].reshape(-1), case["info"].reshape(-1)):
            assert same_bits(O.dequant(q, fi, si), ref.run(q, fi, rng))


def test_pns_generator_as_shipped_collapses_after_11_values():
    """ics.js:234: randomState = (randomState * (1664525 + 1013904223))|0 -- an even multiplier and no
    increment: 11 non-zero outputs, then 0 for ever (reference defects)."""
    noise = np.empty(8192, np.float32)
    assert A.lib().aacfb_get_table(10, noise.ctypes.data, noise.size) == 32
    assert np.count_nonzero(noise[:32]) == 11 and not noise[11:32].any()
    assert noise[0] == np.float32(-1691269120) and noise[10] == np.float32(-2147483648)
    # four noise bands of 4 coefficients at the start of a frame: 11 scaled values, then one signed zero (its
    # band still has energy), then NaN throughout -- 0 * (sf / sqrt(0)) -- also in every later noise band
    q = np.zeros((), A.QFRAME_DTYPE)
    q["group_len"][0] = 1
    q["band"][:8] = [A.BAND_NOISE | 200] * 4 + [A.BAND_SPECTRAL | 200] * 2 + [A.BAND_NOISE | 200] * 2
    fi = O.make_info(max_sfb=8)
    x = O.dequant(q, fi, 4)                       # 44.1 kHz: bands of 4 coefficients
    assert np.isfinite(x[:12]).all() and (x[:11] != 0).all() and x[11] == 0 and np.signbit(x[11])
    assert np.isnan(x[12:16]).all() and np.isnan(x[24:32]).all() and not np.isnan(x[16:24]).any()


def test_library_tables_equal_the_oracles():
    iq = np.empty(8192, np.float32)
    assert A.lib().aacfb_get_table(8, iq.ctypes.data, iq.size) == 8192
    assert np.array_equal(iq[:8191], O.dequant_table(0)) and np.isnan(iq[8191])
    sf = np.empty(8192, np.float32)
    assert A.lib().aacfb_get_table(9, sf.ctypes.data, sf.size) == 428
    assert np.array_equal(sf[:428], O.dequant_table(1))


def test_kernel_dequant_code_equals_the_oracle_bit_for_bit_on_the_cpu():
    """dequant4 / dequant_row of aacfb_core.cuh, compiled for the host."""
    d = np.load(GOLD)
    for q, fi, want in zip(d["qframes"], d["info"], d["data"]):
        assert same_bits(emul.dequant(q, fi, 4), want)
    rng = np.random.default_rng(5)
    for si in (0, 3, 6, 9, 11):
        case = W.random_q_case(8, 1, 1, rng, sample_index=si, p_noise=0.2)
        for q, fi in zip(case["qframes"].reshape(-1), case["info"].reshape(-1)):
            assert same_bits(emul.dequant(q, fi, si), O.dequant(q, fi, si))


S16_KAT = [(0.5, 1), (-0.5, 0), (1.5, 2), (-1.5, -1), (2.5, 3), (-2.5, -2), (0.49999997, 0), (-0.50000006, -1),
           (32766.5, 32767), (32767.4, 32767), (32767.5, 32767), (40000.0, 32767), (-32768.4, -32768), (-32768.5, -32768),
           (-32769.0, -32768), (float("nan"), 0), (float("inf"), 32767), (float("-inf"), -32768), (-0.0, 0), (1e-30, 0),
           (8388607.5, 32767), (123.0, 123), (-123.0, -123)]


def test_pcm_s16_known_answers():
    """AACFB_PCM_S16 = max(-32768, min(32767, Math.round(x))) stored into an Int16Array (include/aacfb.h):
    round half UP, saturate, NaN -> 0.  No such conversion exists inside the reference tree (Aurora's
    sinks do it): parity is pinned by these known answers, for the oracle and for the kernel's code."""
    x = np.array([v for v, _ in S16_KAT], np.float32)
    want = np.array([w for _, w in S16_KAT], np.int16)
    assert np.array_equal(O.pcm_s16(x), want)
    assert np.array_equal(emul.pcm_s16(x), want)
    rng = np.random.default_rng(0)
    y = (rng.standard_normal(300000) * 20000).astype(np.float32)
    y[:100000] = np.round(y[:100000] * 2) / 2          # lots of exact ties
    ref = np.clip(np.floor(y.astype(np.float64) + 0.5), -32768, 32767).astype(np.int16)
    assert np.array_equal(O.pcm_s16(y), ref) and np.array_equal(emul.pcm_s16(y), ref)


def _case(rng, S, T, C, tns_mode=0):
    return W.random_q_case(S, T, C, rng, tns_mode=tns_mode, p_noise=0.0)   # PCM-level checks: no NaN rows


def test_emulated_kernel_schedule_with_quantised_input_and_s16_output():
    rng = np.random.default_rng(17)
    for (S, T, C, tns) in [(2, 5, 2, 0), (1, 4, 1, 0), (2, 4, 3, 0), (2, 5, 2, 1)]:
        case = _case(rng, S, T, C, tns)
        q = case["qframes"]
        q["q"][np.abs(q["q"]) == 8191] = 8190       # NaN spreads over the whole frame: kept for the spectrum-level tests
        q["band"] = np.where((q["band"] & 0x1ff) > 427, (q["band"] & 0xc000) | 200, q["band"])
        kw = dict(sample_index=case["sample_index"], flags=case["flags"])
        ref, _ = O.process_io(q, O.IN_Q16, case["info"], case["tns_blob"], case["tns_offsets"], **kw)
        scale = max(1.0, float(np.abs(ref).max()))
        ov = np.zeros((S, C, 1024), np.float32)
        got = emul.process_io(q, 1, case["info"], case["tns_blob"], case["tns_offsets"], ov, case["sample_index"], case["flags"], 3)
        assert np.abs(got - ref).max() <= 1e-5 * scale
        ref16, _ = O.process_io(q, O.IN_Q16, case["info"], case["tns_blob"], case["tns_offsets"], pcm_format=O.PCM_S16, **kw)
        ov = np.zeros((S, C, 1024), np.float32)
        got16 = emul.process_io(q, 1, case["info"], case["tns_blob"], case["tns_offsets"], ov, case["sample_index"], case["flags"], 3,
                                pcm_format=1)
        # |float sample - oracle's| <= 1e-5 * 32768 * scale < 0.5 * scale: the rounded values differ by at most 1 (x scale)
        assert np.abs(got16.astype(np.int32) - ref16).max() <= max(1, int(scale))
        # float spectra in, int16 out
        spec = O.dequant_batch(q, case["info"], case["sample_index"])
        ov = np.zeros((S, C, 1024), np.float32)
        got16b = emul.process_io(spec, 0, case["info"], case["tns_blob"], case["tns_offsets"], ov, case["sample_index"], case["flags"], 4,
                                 pcm_format=1)
        assert np.array_equal(got16b, got16)       # same arithmetic after the (bit-exact) dequantisation


def test_pack_qframe_mirrors_the_band_walk():
    """pack_qframe (Python twin of js/quant_pack.js): band types + Float32 scalefactors -> band codes."""
    sf_tab = O.dequant_table(1)
    ics = {"info": {"windowSequence": 2, "groupCount": 3, "groupLength": [2, 3, 3, 0, 0, 0, 0, 0], "maxSFB": 4},
           "bandTypes": [1, 0, 13, 15, 11, 14, 5, 13, 0, 0, 3, 3] + [0] * 108,
           "scaleFactors": [sf_tab[210], 0.0, -sf_tab[150], sf_tab[220], sf_tab[0], sf_tab[1], sf_tab[427], -sf_tab[355],
                            0.0, 0.0, np.float32("nan"), sf_tab[200]] + [0.0] * 108}
    quant = np.arange(1024) % 50 - 25
    rec = A.pack_qframe(ics, quant)
    assert list(rec["group_len"]) == [2, 3, 3, 0, 0, 0, 0, 0]
    assert list(rec["band"][:12]) == [0x4000 | 210, 0, 0x8000 | 150, 0, 0x4000, 0, 0x4000 | 427, 0x8000 | 355, 0, 0,
                                      0x4000 | 0x1ff, 0x4000 | 200]
    assert not rec["band"][12:].any() and np.array_equal(rec["q"], quant)


# ------------------------------------------------------------------------------------------ GPU
gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("S,T,C,tns", [(3, 7, 2, 0), (2, 9, 1, 0), (2, 5, 5, 0), (3, 6, 2, 1), (2, 6, 2, 2), (40, 21, 2, 0)])
def test_gpu_quantised_input_matches_the_oracle(S, T, C, tns):
    rng = np.random.default_rng(100 + S + T + C + tns)
    case = _case(rng, S, T, C, tns)
    q = case["qframes"]
    q["q"][np.abs(q["q"]) == 8191] = 8190
    q["band"] = np.where((q["band"] & 0x1ff) > 427, (q["band"] & 0xc000) | 200, q["band"])
    kw = dict(sample_index=case["sample_index"], flags=case["flags"])
    ov0 = (rng.standard_normal((S, C, 1024)) * 3000).astype(np.float32)
    for pcm_format in (A.PCM_F32, A.PCM_S16):
        ovr = ov0.copy()
        ref, _ = O.process_io(q, O.IN_Q16, case["info"], case["tns_blob"], case["tns_offsets"], ovr, pcm_format=pcm_format, **kw)
        ref2, _ = O.process_io(q, O.IN_Q16, case["info"], case["tns_blob"], case["tns_offsets"], ovr, pcm_format=pcm_format, **kw)
        ctx = A.Context(S, C, case["sample_index"], case["flags"])
        ctx.set_overlap(ov0)
        got = ctx.process_io(q, case["info"], case["tns_blob"], case["tns_offsets"], in_format=A.IN_Q16, pcm_format=pcm_format)
        got2 = ctx.process_io(q, case["info"], case["tns_blob"], case["tns_offsets"], in_format=A.IN_Q16, pcm_format=pcm_format)
        assert ctx.launches >= 2   # two calls
        ovg = ctx.get_overlap()
        ctx.close()
        if pcm_format == A.PCM_F32:
            scale = max(1.0, float(np.abs(ref).max()))
            assert np.abs(got - ref).max() <= 1e-5 * scale and np.abs(got2 - ref2).max() <= 1e-5 * scale
        else:
            scale = max(1, int(np.abs(O.process_io(q, O.IN_Q16, case["info"], case["tns_blob"], case["tns_offsets"], ov0.copy(), **kw)[0]).max()))
            assert np.abs(got.astype(np.int32) - ref).max() <= scale and np.abs(got2.astype(np.int32) - ref2).max() <= scale
        assert np.abs(ovg - ovr).max() <= 1e-5 * 32768 * max(1.0, float(np.abs(ovr).max()) / 32768)


@gpu
def test_gpu_dequantised_spectrum_is_bit_exact_through_the_filterbank_identity():
    """The device's inverse quantisation alone: run the SAME batch as aacfb_qframe records and as the
    oracle's dequantised float spectra through the library -- the PCM must be identical bit for bit
    (everything after the dequantisation is the same code on the same floats), for every band kind,
    including the NaN rows of PNS as shipped and of out-of-table reads."""
    rng = np.random.default_rng(9)
    for si in (4, 11):
        case = W.random_q_case(6, 8, 2, rng, sample_index=si, p_noise=0.15)
        spec = O.dequant_batch(case["qframes"], case["info"], si)
        a = A.Context(6, 2, si, 0)
        b = A.Context(6, 2, si, 0)
        pa = a.process_io(case["qframes"], case["info"], in_format=A.IN_Q16)
        pb = b.process_io(spec, case["info"])
        assert np.isnan(pb).any() and same_bits(pa, pb)
        pa16 = a.process_io(case["qframes"], case["info"], in_format=A.IN_Q16, pcm_format=A.PCM_S16)
        pb16 = b.process_io(spec, case["info"], pcm_format=A.PCM_S16)
        assert np.array_equal(pa16, pb16)
        a.close(); b.close()


@gpu
def test_gpu_stereo_tools_after_device_dequantisation():
    """Quantised input + stereo records: inverse quantisation, then processMS / processIS, then the IMDCT
    -- all on the staged rows (decoder.js:294-301 order)."""
    rng = np.random.default_rng(23)
    for tns in (0, 1):
        st = W.random_stereo_case(3, 6, rng, tns_mode=tns)
        S, T = st["cpe"].shape
        qc = W.random_q_case(S, T, 2, rng, p_noise=0.0)
        q = qc["qframes"]
        q["q"][np.abs(q["q"]) == 8191] = 8190
        q["band"] = np.where((q["band"] & 0x1ff) > 427, (q["band"] & 0xc000) | 200, q["band"])
        info = st["info"].copy()
        info["max_sfb"] = qc["info"]["max_sfb"]
        for f in ("window_sequence",):
            assert info[f].shape == qc["info"][f].shape
        # the q-case's groups belong to its own window sequences: rebuild the group lengths for st's sequences
        short = info["window_sequence"] == 2
        q["group_len"][...] = 0
        q["group_len"][..., 0] = np.where(short, 8, 1)
        info["max_sfb"] = np.where(short, np.minimum(info["max_sfb"], 14), info["max_sfb"])
        ops = np.zeros((S, T, 1), A.STEREO_DTYPE)
        for s in range(S):
            for t in range(T):
                _, present = A.pack_stereo(st["cpe"][s, t], 4, out=ops[s, t, 0])
                info["stereo_present"][s, t, 0] = int(present)
        spec = O.dequant_batch(q, info, 4)
        ref_in = spec.copy()
        for s in range(S):
            for t in range(T):
                ref_in[s, t, 0], ref_in[s, t, 1] = O.stereo(st["cpe"][s, t], 4, spec[s, t, 0], spec[s, t, 1])
        ref, _ = O.process(ref_in, info, st["tns_blob"], st["tns_offsets"], sample_index=4, flags=st["flags"])
        ctx = A.Context(S, 2, 4, st["flags"])
        got = ctx.process_io(q, info, st["tns_blob"], st["tns_offsets"], stereo_ops=ops, in_format=A.IN_Q16)
        ctx.close()
        assert np.abs(got - ref).max() <= 1e-5 * max(1.0, float(np.abs(ref).max()))


@gpu
def test_gpu_s16_output_from_float_spectra_and_saturation():
    w = W.make(5, 8, 40, 2, seed=3, shape_prev_mode="carried")
    w["spectra"] *= 4.0                      # peaks well above full scale: saturation on both sides
    ref, _ = O.process_io(w["spectra"], O.IN_F32, w["info"], pcm_format=O.PCM_S16)
    ctx = A.Context(8, 2, 4, 0)
    got = ctx.process_io(w["spectra"], w["info"], pcm_format=A.PCM_S16)
    ctx.close()
    assert (got == 32767).any() and (got == -32768).any()
    assert np.abs(got.astype(np.int32) - ref).max() <= 4


@gpu
def test_gpu_full_size_quantised_config2_device_path_roundtrip():
    """BASELINE config 2 at full size (65 536 stereo frames) as aacfb_qframe records through the
    device-pointer entry point with int16 PCM: spot-check streams against the oracle and the whole
    batch through a size-independent property -- a second context fed the oracle-dequantised floats
    of the same records gives the same int16 PCM bit for bit."""
    import torch

    S, T, C = 256, 256, 2
    w = W.make_q(2, S, T, C, seed=1)
    dev = torch.device("cuda:0")
    qf = torch.from_numpy(w["qframes"].view(np.uint8).reshape(S, T, C, 2304)).to(dev)
    info = torch.from_numpy(w["info"].view(np.uint8).reshape(S, T, C, 8).copy()).to(dev)
    pcm = torch.empty((S, T, 1024, C), dtype=torch.int16, device=dev)
    ctx = A.Context(S, C, 4, 0)
    st = torch.cuda.current_stream().cuda_stream
    ctx.process_device_io(qf.data_ptr(), A.IN_Q16, info.data_ptr(), pcm.data_ptr(), A.PCM_S16, T, st)
    torch.cuda.synchronize()
    got = pcm.cpu().numpy()
    ctx.close()
    for s in (0, 100, 255):
        ref, _ = O.process_io(w["qframes"][s:s + 1], O.IN_Q16, w["info"][s:s + 1], pcm_format=O.PCM_S16)
        assert np.abs(got[s:s + 1].astype(np.int32) - ref).max() <= 1
    sub = slice(0, 64)
    spec = torch.from_numpy(O.dequant_batch(w["qframes"][sub], w["info"][sub], 4)).to(dev)
    pcm2 = torch.empty((64, T, 1024, C), dtype=torch.int16, device=dev)
    ctx2 = A.Context(64, C, 4, 0)
    ctx2.process_device_io(spec.data_ptr(), A.IN_F32, info[sub].contiguous().data_ptr(), pcm2.data_ptr(), A.PCM_S16, T, st)
    torch.cuda.synchronize()
    ctx2.close()
    assert np.array_equal(pcm2.cpu().numpy(), got[sub])


@gpu
def test_gpu_host_memory_helpers_and_validation():
    a = A.host_alloc((4, 3, 2, 1024), np.float32)
    a[...] = 1000.0
    info = np.zeros((4, 3, 2), A.INFO_DTYPE)
    out = A.host_alloc((4, 3, 1024, 2), np.float32)
    ctx = A.Context(4, 2, 4, 0)
    ctx.process_io(a, info, out=out)
    assert np.isfinite(out).all() and np.abs(out).max() > 0
    b = np.ascontiguousarray(a.copy())
    A.host_register(b)
    got = ctx.process_io(b, info)
    A.host_unregister(b)
    assert got.shape == out.shape
    # quantised input: maxSFB beyond the band table and groups that do not cover the frame are rejected
    q = np.zeros((4, 3, 2), A.QFRAME_DTYPE)
    q["group_len"][..., 0] = 1
    bad = info.copy(); bad["max_sfb"] = 50
    with pytest.raises(A.AacfbError):
        ctx.process_io(q, bad, in_format=A.IN_Q16)
    q["group_len"][0, 0, 0, 0] = 0
    with pytest.raises(A.AacfbError):
        ctx.process_io(q, info, in_format=A.IN_Q16)
    ctx.close()
    A.host_free(a); A.host_free(out)
