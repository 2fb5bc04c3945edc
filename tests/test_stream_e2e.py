"""End to end at the plugin surface: an AAC-LC ADTS byte stream -> interleaved Float32 PCM.

tests/golden/stream/jsref_stream_*.npz (made by `python tools/js_reference.py stream`, where
/root/reference exists) hold, for a stereo and a mono synthetic stream (tools/aac_bitstream.py: valid
raw_data_block syntax -- channel pair / single channel elements, M/S masks, intensity stereo, section
and scalefactor data, TNS data, all eleven spectral Huffman codebooks with escapes, window switching --
written with the reference's own code tables):

  * `pcm`: what the reference's UNMODIFIED decoder produced -- setCookie, then readChunk per access unit:
    ADTS header, the whole bit parse, process, interleave -- run by tools/jsmini.py on a stand-in for
    the `av` peer package;
  * per addon call, the typed arrays that aac.js_b200/js/decoder_b200.js -- running on top of that same
    unmodified decoder, its parse intercepted -- staged for aacfb_process / aacfb_process_stereo.

(A third stream is 5.1: centre SCE, two channel pair elements, LFE per access unit; the stereo tools of
its pairs stay on the host, the device path takes them for 2-channel streams only.)

Replaying the staged calls through the library must give the reference's PCM: bit for bit with the
oracle standing in for the library (CPU), within 1e-5 with the CUDA library (GPU).  Where the
reference tree is present the whole thing is also run live on a fresh stream."""
import glob
import os

import numpy as np
import pytest

import aacjs_b200 as A
from tools import workloads as W

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "stream", "jsref_stream_*.npz")))
HAVE_REF = os.path.isdir("/root/reference/src")
TOL = 1e-5


def tns_flags(path):
    """`*_tnsfixed`: the reference ran with its two TNS tokens fixed (tools/js_reference.TNS_FIXES: `tmp` -> `top`
    at tns.js:122 and `ics.maxSFB` -> `ics.info.maxSFB` at :106), so its decoder -- which passes decode = false
    -- really runs the MA filter: the library mode is TNS_FIXED_MA.  Otherwise TNS is the identity."""
    return A.TNS_FIXED_MA if path.endswith("_tnsfixed.npz") else A.TNS_AS_SHIPPED


def load(path):
    z = np.load(path)
    C, n, seed, K, n_calls = (int(v) for v in z["meta"])
    calls = []
    for i in range(n_calls):
        sp = z[f"c{i}_spectra"]
        T = sp.shape[0]
        calls.append({"spectra": sp, "info": z[f"c{i}_info"].view(W.INFO_DTYPE).reshape(T, C),
                      "stereo_ops": z[f"c{i}_stereo"] if z[f"c{i}_stereo"].size else None,
                      "tns_blob": z[f"c{i}_tns_blob"] if z[f"c{i}_tns_blob"].size else None,
                      "tns_offsets": z[f"c{i}_tns_offsets"] if z[f"c{i}_tns_offsets"].size else None})
    return z, C, n, calls


def test_fixtures_cover_both_layouts():
    assert [os.path.basename(p) for p in GOLD] == ["jsref_stream_mono.npz", "jsref_stream_stereo.npz",
                                                   "jsref_stream_stereo_tnsfixed.npz",
                                                   "jsref_stream_surround.npz"]   # 1, 2 and 5.1 channels


@pytest.mark.parametrize("path", GOLD, ids=os.path.basename)
def test_adts_index_finds_every_access_unit_of_the_stream(path):
    z, C, n, calls = load(path)
    frames, consumed = A.adts_index(z["adts"])
    assert len(frames) == n and consumed == z["adts"].size
    assert sum(c["spectra"].shape[0] for c in calls) == n          # the JS batched all of them
    assert set(int(v) for v in frames["chan_config"]) == {C} and set(int(v) for v in frames["profile"]) == {2}


@pytest.mark.parametrize("path", GOLD, ids=os.path.basename)
def test_oracle_replay_of_the_staged_calls_equals_the_reference_decoder(path):
    from tools.js_reference import OracleLibrary

    z, C, n, calls = load(path)
    lib = OracleLibrary(C, flags=tns_flags(path))
    pcm = np.concatenate([lib(c).reshape(-1) for c in calls])
    assert pcm.size == n * 1024 * C
    assert np.array_equal(pcm.view(np.uint32), z["pcm"].view(np.uint32))
    if C == 2:   # the stream does exercise the stereo tools and TNS side info
        assert any(c["stereo_ops"] is not None and c["info"]["stereo_present"].any() for c in calls)
        assert any(c["tns_blob"] is not None for c in calls)


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference is not on this machine")
@pytest.mark.parametrize("channels,stereo_on_device", [(2, True), (2, False), (1, True), (6, True)])
def test_batching_decoder_equals_stock_decoder_live(channels, stereo_on_device):
    """decoder_b200.js on top of the unmodified reference == the unmodified reference, sample for sample,
    for any chunking (frames per chunk 1, 3, 64) and with the stereo tools on either side."""
    from tools import aac_bitstream as B
    from tools.js_reference import B200DecoderHarness, OracleLibrary, StreamReference

    data = B.write_adts_stream(B.random_frames(np.random.default_rng(500 + channels), 7, channels=channels),
                               B.codebooks(), channels=channels)
    ref = StreamReference(data, channels=channels).decode_all()
    assert ref.size == 7 * 1024 * channels and np.isfinite(ref).all() and np.abs(ref).max() > 0.01
    for K in (1, 3, 64):
        h = B200DecoderHarness(data, OracleLibrary(channels), channels=channels, frames_per_chunk=K,
                               stereo_on_device=stereo_on_device)
        got = h.decode_all()
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
        assert len(h.calls) == -(-7 // K)
        assert all((c["entry"] == "aacfb_process_stereo") <= (stereo_on_device and channels == 2) for c in h.calls)


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference is not on this machine")
@pytest.mark.parametrize("channels", [2, 1])
def test_batching_decoder_quantised_staging_s16_adts_index_and_cpu_frames_live(channels):
    """The round-2 additions of decoder_b200.js, each against the stock decoder's PCM, sample for sample:
    quantOnDevice (ICStream.decodeSpectralData swapped for quant_pack.js's walk while the reference parses;
    the oracle's restatement of ics.js:203-266 stands in for the device), int16 PCM, the ADTS frame index
    bounding each batch (no underflow probing), and frames that have to take the reference's own CPU path
    (a coupling element in this.cces, simulated) with the overlap state handed over in both directions."""
    from oracle import oracle as O
    from tools import aac_bitstream as B
    from tools.js_reference import B200DecoderHarness, OracleLibrary, StreamReference

    data = B.write_adts_stream(B.random_frames(np.random.default_rng(600 + channels), 7, channels=channels),
                               B.codebooks(), channels=channels)
    ref = StreamReference(data, channels=channels).decode_all()
    for kw in (dict(quant_on_device=True), dict(quant_on_device=True, pcm_format="s16"), dict(force_cpu_frames=(2, 5)),
               dict(quant_on_device=True, force_cpu_frames=(0, 3, 6)), dict(quant_on_device=True, adts_index=False)):
        for K in (3, 64):
            h = B200DecoderHarness(data, OracleLibrary(channels), channels=channels, frames_per_chunk=K, **kw)
            got = h.decode_all()
            if kw.get("pcm_format") == "s16":
                assert got.dtype == np.int16 and np.array_equal(got, O.pcm_s16(ref * 32768))
            else:
                assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
            if kw.get("quant_on_device"):
                assert all(c["entry"] == "aacfb_process_io" and "qframes" in c for c in h.calls)
            if kw.get("adts_index", True):   # the index told readChunk how many complete frames each chunk holds
                assert h.index_calls[0] == min(K, 7) and sum(c["info"].shape[0] for c in h.calls) + h.cpu_frames == 7
            assert h.cpu_frames == len(kw.get("force_cpu_frames", ()))


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference is not on this machine")
@pytest.mark.parametrize("channels,kw", [(2, dict()), (1, dict(pcm_format="s16")),
                                         (2, dict(quant_on_device=False, stereo_on_device=False, adts_index=False))])
def test_decoder_pool_equals_the_stock_decoder_on_every_stream_live(channels, kw):
    """decoder_pool.js: S decoders in lock step behind ONE library context (the operating point bench.py
    measures).  Three streams of different lengths, each parsed by the reference's own code; every round is one
    addon call over [S][T][C] with T what every stream can deliver; each stream's PCM equals the stock
    decoder's, sample for sample, for as many frames as the shortest stream has."""
    from oracle import oracle as O
    from tools import aac_bitstream as B
    from tools.js_reference import B200PoolHarness, OracleLibrary, StreamReference

    lengths = (4, 3, 5)
    streams = [B.write_adts_stream(B.random_frames(np.random.default_rng(700 + 10 * channels + i), n, channels=channels),
                                   B.codebooks(), channels=channels) for i, n in enumerate(lengths)]
    refs = [StreamReference(d, channels=channels).decode_all() for d in streams]
    for K in (2, 64):
        h = B200PoolHarness(streams, OracleLibrary(channels, n_streams=3), channels=channels, frames_per_chunk=K, **kw)
        got = h.decode_all()
        assert h.created == [[0, 3, channels, 4, 0, 0]]           # one context, n_streams = 3
        n = min(lengths) * 1024 * channels                         # the pool stops when a stream runs dry
        for s in range(3):
            if kw.get("pcm_format") == "s16":
                assert got[s].dtype == np.int16 and np.array_equal(got[s], O.pcm_s16(refs[s][:n] * 32768))
            else:
                assert got[s].size == n and np.array_equal(got[s].view(np.uint32), refs[s][:n].view(np.uint32))
        assert [c["info"].shape[:2] for c in h.calls] == [(3, t) for t in ([2, 1] if K == 2 else [3])]
        assert all(("qframes" in c) == kw.get("quant_on_device", True) for c in h.calls)
        # nothing beyond the delivered frames was consumed: the longer streams continue where the pool left them
        assert h.read_chunks() is None


QGOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "stream", "jsref_streamq_*.npz")))


def load_q(path):
    z = np.load(path)
    C, n, seed, K, n_calls, s16 = (int(v) for v in z["meta"])
    calls = []
    for i in range(n_calls):
        q = z[f"c{i}_qframes"].view(A.QFRAME_DTYPE)
        T = q.size // C
        calls.append({"qframes": q.reshape(T, C), "info": z[f"c{i}_info"].view(W.INFO_DTYPE).reshape(T, C),
                      "stereo_ops": z[f"c{i}_stereo"] if z[f"c{i}_stereo"].size else None, "tns_blob": None, "tns_offsets": None,
                      "pcm_format": s16})
    return z, C, n, s16, calls


@pytest.mark.parametrize("path", QGOLD, ids=os.path.basename)
def test_oracle_replay_of_the_quantised_staged_calls_equals_the_reference_decoder(path):
    from oracle import oracle as O
    from tools.js_reference import OracleLibrary

    z, C, n, s16, calls = load_q(path)
    assert len(QGOLD) == 2
    lib = OracleLibrary(C)
    pcm = np.concatenate([lib(c).reshape(-1) for c in calls])
    want = O.pcm_s16(z["pcm"] * 32768) if s16 else z["pcm"]
    assert pcm.size == n * 1024 * C and np.array_equal(pcm, want)


@pytest.mark.gpu
@pytest.mark.parametrize("path", QGOLD, ids=os.path.basename)
def test_gpu_replay_of_the_quantised_staged_calls_equals_the_reference_decoder(path):
    """What decoder_b200.js staged with quantOnDevice -- aacfb_qframe records straight from the reference's
    bit parse -- through aacfb_process_io on the GPU, against the stock decoder's PCM."""
    z, C, n, s16, calls = load_q(path)
    ctx = A.Context(1, C, 4, 0)
    out = []
    for c in calls:
        T = c["qframes"].shape[0]
        ops = c["stereo_ops"].view(A.STEREO_DTYPE).reshape(1, T, 1) if c["stereo_ops"] is not None else None
        out.append(ctx.process_io(c["qframes"][None], c["info"][None], stereo_ops=ops, in_format=A.IN_Q16,
                                  pcm_format=A.PCM_S16 if s16 else A.PCM_F32).reshape(-1))
    ctx.close()
    pcm, ref = np.concatenate(out), z["pcm"]
    if s16:
        want = np.clip(np.floor(ref.astype(np.float64) * 32768 + 0.5), -32768, 32767)
        assert np.abs(pcm.astype(np.int32) - want).max() <= max(1, int(np.ceil(np.abs(ref).max())))   # 1e-5 * 32768 * peak < 0.5 * peak
    else:
        assert np.abs(pcm.astype(np.float64) - ref).max() <= TOL * max(1.0, float(np.abs(ref).max()))


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=os.path.basename)
def test_gpu_replay_of_the_staged_calls_equals_the_reference_decoder(path):
    z, C, n, calls = load(path)
    ctx = A.Context(1, C, 4, tns_flags(path))
    out = []
    for c in calls:
        T = c["spectra"].shape[0]
        ops = c["stereo_ops"].view(A.STEREO_DTYPE).reshape(1, T, 1) if c["stereo_ops"] is not None else None
        out.append(ctx.process(c["spectra"][None], c["info"][None], c["tns_blob"], c["tns_offsets"], stereo_ops=ops).reshape(-1))
    assert ctx.launches >= len(calls)
    ctx.close()
    pcm = np.concatenate(out)
    ref = z["pcm"]
    assert np.abs(pcm.astype(np.float64) - ref).max() <= TOL * max(1.0, float(np.abs(ref).max()))


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference is not on this machine")
def test_tns_end_to_end_with_the_reference_tokens_fixed_live():
    """With both defects fixed the reference's real decoder filters (MA branch); the batching decoder with
    the library in TNS_FIXED_MA mode gives the same bits -- TNS data parsed from the bitstream (coefficient
    tables, compression, short-window filters) -> tns_pack.js blocks -> filter, end to end."""
    from tools import aac_bitstream as B
    from tools.js_reference import TNS_FIXES, B200DecoderHarness, OracleLibrary, StreamReference

    data = B.write_adts_stream(B.random_frames(np.random.default_rng(77), 6, channels=2), B.codebooks(), channels=2)
    shipped = StreamReference(data).decode_all()
    fixed = StreamReference(data, patches=TNS_FIXES).decode_all()
    assert not np.array_equal(shipped, fixed) and np.isfinite(fixed).all()
    got = B200DecoderHarness(data, OracleLibrary(2, flags=A.TNS_FIXED_MA), channels=2, frames_per_chunk=4).decode_all()
    assert np.array_equal(got.view(np.uint32), fixed.view(np.uint32))


@pytest.mark.parametrize("path", GOLD, ids=os.path.basename)
def test_emulated_kernel_replay_of_the_staged_calls(path):
    """The kernels' own arithmetic (the same source compiled for the host, tests/emul.py) on the staged
    calls, against the reference decoder's PCM: what the GPU replay below sees, minus the GPU."""
    from tests import emul

    z, C, n, calls = load(path)
    ov = np.zeros((1, C, 1024), np.float32)
    out = []
    for c in calls:
        sp = c["spectra"].copy()
        if c["stereo_ops"] is not None:      # the op semantics of include/aacfb.h (stereo_apply on the device)
            recs = c["stereo_ops"].view(A.STEREO_DTYPE)
            for t in range(sp.shape[0]):
                if c["info"][t, 0]["stereo_present"]:
                    op = np.repeat(recs[t]["op"], 4)
                    l, r = sp[t, 0].copy(), sp[t, 1].copy()
                    ms, it = op == A.STEREO_MS, op >= A.STEREO_IS
                    sp[t, 0][ms], sp[t, 1][ms] = l[ms] + r[ms], l[ms] - r[ms]
                    sp[t, 1][it] = l[it] * recs[t]["scale"][op[it] - A.STEREO_IS]
        out.append(np.asarray(emul.process(sp[None], c["info"][None], c["tns_blob"], c["tns_offsets"], ov, 4,
                                           tns_flags(path), 3)).reshape(-1))
    ref = z["pcm"]
    err = np.abs(np.concatenate(out).astype(np.float64) - ref).max()
    assert err <= TOL * max(1.0, float(np.abs(ref).max())) / 5, err      # 5x inside the bar
