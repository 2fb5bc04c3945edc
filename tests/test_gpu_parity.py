"""GPU parity tests proper: the CUDA path, called through the C-ABI, against the CPU oracle
on the same seeded inputs, against the committed golden fixtures, and -- at BASELINE.json's
full sizes -- through size-independent properties (TDAC round trip, linearity, invariance
to how the batch is cut).  Tolerance: 1e-5 max-abs on full-scale PCM (BASELINE.json
north_star, floating-point IMDCT); measured ~2e-7."""
import glob
import os

import numpy as np
import pytest

import aacjs_b200 as A
from oracle import oracle as O
from tools import workloads as W

pytestmark = pytest.mark.gpu
TOL = 1e-5
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gpu_process(w, S, C, ov0=None, T=None):
    ctx = A.Context(S, C, w["sample_index"], w["flags"])
    if ov0 is not None:
        ctx.set_overlap(ov0)
    pcm = ctx.process(w["spectra"], w["info"], w["tns_blob"], w["tns_offsets"])
    ov = ctx.get_overlap()
    n = ctx.launches
    ctx.close()
    assert n >= 1
    return pcm, ov


def check(w, S, T, C, seed=0):
    rng = np.random.default_rng(seed)
    ov0 = (rng.standard_normal((S, C, 1024)) * 0.25 * 32768).astype(np.float32)
    ovo = ov0.copy()
    ref, _ = O.process(w["spectra"], w["info"], w["tns_blob"], w["tns_offsets"], ovo, sample_index=w["sample_index"],
                       flags=w["flags"], n_threads=8)
    got, ovg = gpu_process(w, S, C, ov0)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    m = ~np.isnan(ref)
    # the 1e-5 bar is on full-scale PCM ([-1, 1)); scale it for inputs that overshoot full scale
    tol = TOL * max(1.0, float(np.abs(ref[m]).max()))
    assert np.abs(got[m].astype(np.float64) - ref[m]).max() <= tol
    assert np.abs(ovg - ovo)[~np.isnan(ovo)].max() / 32768 <= tol


@pytest.mark.parametrize("cfg,S,T,C", [(1, 1, 1, 1), (2, 4, 9, 2), (2, 64, 40, 2), (3, 8, 7, 2), (4, 8, 9, 2),
                                        (5, 5, 37, 2), (5, 3, 33, 1), (5, 2, 19, 3), (5, 1, 17, 8)])
def test_baseline_configs_small(cfg, S, T, C):
    check(W.make(cfg, S, T, C, seed=cfg, shape_prev_mode="carried"), S, T, C)


@pytest.mark.parametrize("slice_len", [7, 40, 97, 1000])
def test_runs_of_long_and_short_frames_inside_an_item(slice_len, monkeypatch):
    """Work items of different lengths (AACFB_SLICE_LEN: shorter than a block of EIGHT_SHORT frames, longer than
    32 frames, longer than a stream, the whole batch as one item) over streams that mix the two kinds of
    frames in every way: EIGHT_SHORT blocks at the first / last frame of a stream, cut by items and by pair
    boundaries, one chain of a pair-frame short and the other not, streams without any, one all EIGHT_SHORT."""
    monkeypatch.setenv("AACFB_SLICE_LEN", str(slice_len))
    S, T, C = 6, 150, 2
    rng = np.random.default_rng(slice_len)
    w = W.random_case(S, T, C, rng)
    seq = w["info"]["window_sequence"]
    seq[0] = 0                                    # stream 0: ONLY_LONG throughout
    seq[1] = W.config5_sequence(T)[:, None]       # stream 1: the config-5 pattern
    seq[2] = 0
    seq[2, :2] = 2; seq[2, 2] = 3; seq[2, -3] = 1; seq[2, -2:] = 2   # blocks at both ends of the stream
    seq[3] = 2                                    # all EIGHT_SHORT
    seq[4] = 0
    seq[4, 70] = 1; seq[4, 71, 0] = 2; seq[4, 71, 1] = 3; seq[4, 72] = 3   # one chain of a pair-frame short, the other not
    w["info"]["max_sfb"] = np.where(seq == 2, 14, 49)
    check(w, S, T, C, seed=slice_len)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("S,T,C", [(2, 6, 2), (3, 4, 5), (1, 9, 1), (37, 33, 2)])
def test_random_sequences_shapes_and_tns(mode, S, T, C):
    check(W.random_case(S, T, C, np.random.default_rng(10 * mode + S), tns_mode=mode), S, T, C)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "*.npz"))), ids=os.path.basename)
def test_golden_fixtures(path):
    from tests.test_golden import load_case

    if os.path.basename(path).startswith("jsref_"):
        from tests.test_oracle_pin import load_jsref

        w, pcm, ov = load_jsref(path)  # vectors produced by the reference's own JS (tools/js_reference.py)
        w["sample_index"] = 4
    else:
        w, pcm, ov = load_case(path)
    S, T, C = w["spectra"].shape[:3]
    got, ovg = gpu_process(w, S, C)
    assert np.abs(got.astype(np.float64) - pcm).max() <= TOL
    assert np.abs(ovg - ov).max() / 32768 <= TOL


def test_streams_continue_across_calls_and_reset():
    """overlaps[] persists between readChunk calls (filter_bank.js:38-41); reset() zeroes it."""
    w = W.make(5, 3, 40, 2, seed=1, shape_prev_mode="carried")
    ref, _ = O.process(w["spectra"], w["info"], sample_index=4)
    ctx = A.Context(3, 2)
    parts = [ctx.process(w["spectra"][:, a:b], w["info"][:, a:b]) for a, b in ((0, 1), (1, 14), (14, 40))]
    got = np.concatenate(parts, axis=1)
    assert np.abs(got.astype(np.float64) - ref).max() <= TOL
    ctx.reset()
    again = ctx.process(w["spectra"], w["info"])
    assert np.array_equal(again.view(np.uint32), got.view(np.uint32))
    ctx.close()


def test_edge_cases():
    ctx = A.Context(2, 2)
    # empty batch
    out = ctx.process(np.zeros((2, 0, 2, 1024), np.float32), np.zeros((2, 0, 2), W.INFO_DTYPE))
    assert out.shape == (2, 0, 1024, 2)
    # all-zero spectra -> exact zeros
    z = ctx.process(np.zeros((2, 3, 2, 1024), np.float32), np.zeros((2, 3, 2), W.INFO_DTYPE))
    assert not z.any()
    # invalid window sequence is rejected on the host path
    bad = np.zeros((2, 1, 2), W.INFO_DTYPE)
    bad["window_sequence"][1, 0, 1] = 7
    with pytest.raises(A.AacfbError, match="AACFB_ERR_SEQUENCE"):
        ctx.process(np.zeros((2, 1, 2, 1024), np.float32), bad)
    # TNS order > 20 is rejected like tns.js:85
    inf = np.zeros((2, 1, 2), W.INFO_DTYPE)
    inf["tns_present"][0, 0, 0] = 1
    blk = bytes([1, 0, 0, 0, 0, 0, 0, 0, 10, 21, 0, 0]) + bytes(84)
    blob, offs = W.pack_tns([blk, None, None, None])
    c2 = A.Context(2, 2, flags=A.TNS_FIXED_AR)
    with pytest.raises(A.AacfbError, match="TNS filter out of range"):
        c2.process(np.zeros((2, 1, 2, 1024), np.float32), inf, blob, offs)
    c2.close()
    ctx.close()


def test_error_in_a_late_sub_batch_drains_the_pipeline_and_leaves_the_state():
    """The host path pipelines sub-batches of streams (copies on one stream per PCIe direction, kernels on the
    buffer sets' streams) and validates each one's side info right before queueing it.  A bad record in the LAST
    stream is found when earlier sub-batches are already in flight: the call has to drain them, report the
    reference-style error, leave the overlap state as it was, and the context has to keep working."""
    S, T, C = 48, 48, 2                      # 36 MiB in + out: several sub-batches
    w = W.make(5, S, T, C, seed=9)
    ctx = A.Context(S, C)
    first = ctx.process(w["spectra"], w["info"])
    state = ctx.get_overlap()
    bad = w["info"].copy()
    bad["window_sequence"][S - 1, T - 1, 1] = 9
    with pytest.raises(A.AacfbError, match="AACFB_ERR_SEQUENCE"):
        ctx.process(w["spectra"], bad)
    assert np.array_equal(ctx.get_overlap(), state)
    # the same frames again, from the state the first call left: equals a fresh context fed both calls
    again = ctx.process(w["spectra"], w["info"])
    ref = A.Context(S, C)
    assert np.array_equal(ref.process(w["spectra"], w["info"]), first)
    assert np.array_equal(ref.process(w["spectra"], w["info"]), again)
    ref.close()
    ctx.close()


def test_max_magnitude_inputs_stay_finite():
    w = W.make(2, 2, 4, 2, seed=5)
    w["spectra"] = np.sign(w["spectra"]) * np.float32(3.0e7)
    check(w, 2, 4, 2)


# ----- full-size properties (BASELINE.json configs: 65536 stereo frames) ---------------------
FULL_S, FULL_T = 256, 256


def _device_run(ctx, spectra_t, info_t, T):
    import torch

    pcm = torch.empty((spectra_t.shape[0], T, 1024, spectra_t.shape[2]), device=spectra_t.device)
    ctx.process_device(spectra_t.data_ptr(), info_t.data_ptr(), pcm.data_ptr(), T, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return pcm


@pytest.mark.parametrize("cfg", [2, 3, 5])
def test_full_size_tdac_round_trip(cfg):
    """encode -> decode at full batch: forward-MDCT a known signal on the GPU with torch (float64),
    run the kernels, recover the signal (Princen-Bradley), for every stream of the 65536-frame batch."""
    import torch

    dev = torch.device("cuda:0")
    S, T, C = FULL_S, FULL_T, 2
    g = torch.Generator(device=dev).manual_seed(cfg)
    sig = torch.randn((S, C, (T + 1) * 1024), device=dev, generator=g, dtype=torch.float32) * 8000.0
    seq = np.zeros(T, np.uint8) if cfg == 2 else (np.full(T, 2, np.uint8) if cfg == 3 else W.config5_sequence(T))
    if cfg == 5:
        seq[0] = 0
    wl = torch.tensor(O.table(4), device=dev, dtype=torch.float64)
    ws = torch.tensor(O.table(6), device=dev, dtype=torch.float64)

    def basis(N2):
        n = torch.arange(N2, device=dev, dtype=torch.float64)[None, :]
        k = torch.arange(N2 // 2, device=dev, dtype=torch.float64)[:, None]
        return 2.0 * torch.cos(2 * np.pi / N2 * (n + 0.5 + N2 / 4) * (k + 0.5))

    B2048, B256 = basis(2048).T.contiguous(), basis(256).T.contiguous()
    one, zero = torch.ones(448, device=dev, dtype=torch.float64), torch.zeros(448, device=dev, dtype=torch.float64)
    spectra = torch.empty((S, T, C, 1024), device=dev, dtype=torch.float32)
    for t in range(T):
        blk = sig[:, :, t * 1024:(t + 2) * 1024].double()
        sq = int(seq[t])
        if sq == 2:
            win = torch.cat([ws, ws.flip(0)])
            segs = torch.stack([blk[:, :, 448 + 128 * w: 448 + 128 * w + 256] for w in range(8)], 2) * win
            spectra[:, t] = (segs @ B256).reshape(S, C, 1024).float()
        else:
            first = wl if sq in (0, 1) else torch.cat([zero, ws, one])
            second = wl.flip(0) if sq in (0, 3) else torch.cat([one, ws.flip(0), zero])
            spectra[:, t] = ((blk * torch.cat([first, second])) @ B2048).float()
    info = torch.zeros((S, T, C, 8), dtype=torch.uint8, device=dev)
    info[..., 0] = torch.tensor(seq, device=dev)[None, :, None]
    ctx = A.Context(S, C)
    pcm = _device_run(ctx, spectra, info, T)
    rec = pcm[:, 1:].permute(0, 3, 1, 2).reshape(S, C, (T - 1) * 1024) * 32768.0
    err = (rec - sig[:, :, 1024:T * 1024]).abs().max().item()
    assert err < 0.25, err  # |sig| peaks ~4e4 and spectra are rounded to f32: relative ~6e-6
    ctx.close()


def test_full_size_linearity_and_batch_cut_invariance():
    """f(a*x) = a*f(x) exactly for a power of two; and the PCM does not depend on how the batch
    is cut into device calls (chunk/halo logic at full size)."""
    import torch

    dev = torch.device("cuda:0")
    S, T, C = FULL_S, FULL_T, 2
    w = W.make(5, 1, T, C, seed=2)
    info = torch.tensor(np.broadcast_to(w["info"].view(np.uint8).reshape(1, T, C, 8), (S, T, C, 8)).copy(), device=dev)
    x = torch.randn((S, T, C, 1024), device=dev) * 1.0e5
    ctx = A.Context(S, C)
    y = _device_run(ctx, x, info, T)
    ctx.reset()
    y4 = _device_run(ctx, x * 4.0, info, T)
    assert torch.equal(y4, y * 4.0)
    ctx.reset()
    parts = []
    for a, b in ((0, 100), (100, 101), (101, 256)):
        parts.append(_device_run(ctx, x[:, a:b].contiguous(), info[:, a:b].contiguous(), b - a))
    assert torch.equal(torch.cat(parts, 1), y)
    # a sample of streams against the oracle at full T
    sel = [0, 1, S // 2, S - 1]
    ref, _ = O.process(x[sel].cpu().numpy(), info[sel].cpu().numpy().view(W.INFO_DTYPE).reshape(len(sel), T, C),
                       sample_index=4, n_threads=4)
    assert np.abs(y[sel].cpu().numpy().astype(np.float64) - ref).max() <= TOL
    ctx.close()


@pytest.mark.parametrize("env", [dict(AACFB_SUB_BATCHES="1"), dict(AACFB_SUB_BATCHES="5", AACFB_LANES="1"),
                                 dict(AACFB_SUB_BATCHES="7", AACFB_LANES="4"), dict(AACFB_SUB_BATCHES="8", AACFB_TAPER="2"),
                                 dict(AACFB_SUB_BATCHES="48", AACFB_TAPER="1")])
def test_host_pipeline_shape_does_not_change_the_pcm(env, monkeypatch):
    """aacfb_process_io pipelines sub-batches of streams over buffer sets, with one copy stream per PCIe direction
    tied to the kernels by events.  However the call is cut (sub-batch count, buffer sets, tapered ends), the PCM
    and the overlap state are those of the default shape, bit for bit, over two consecutive calls (buffer reuse)."""
    S, T, C = 48, 40, 2                       # 30 MiB in + out: pipelined by default
    w = W.make(5, S, T, C, seed=21, shape_prev_mode="carried")
    ctx = A.Context(S, C)
    want = [ctx.process(w["spectra"], w["info"]), ctx.process(w["spectra"], w["info"])]
    want_ov = ctx.get_overlap()
    ctx.close()
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    ctx = A.Context(S, C)
    got = [ctx.process(w["spectra"], w["info"]), ctx.process(w["spectra"], w["info"])]
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert np.array_equal(ctx.get_overlap(), want_ov)
    ctx.close()


def test_full_size_tns_sample_against_oracle():
    """config 4 at full batch (65536 stereo frames, TNS order 12): spot-check streams against the oracle."""
    import torch

    S, T, C = FULL_S, FULL_T, 2
    w = W.make(4, S, T, C, seed=4)
    ctx = A.Context(S, C, 4, A.TNS_FIXED_AR)
    got = ctx.process(w["spectra"], w["info"], w["tns_blob"], w["tns_offsets"])
    sel = [0, 77, S - 1]
    for s in sel:
        lo, hi = s * T * C, (s + 1) * T * C
        offs = w["tns_offsets"][lo:hi + 1]
        ref, _ = O.process(w["spectra"][s:s + 1], w["info"][s:s + 1], w["tns_blob"], offs, sample_index=4, flags=1)
        assert np.abs(got[s:s + 1].astype(np.float64) - ref).max() <= TOL
    ctx.close()
