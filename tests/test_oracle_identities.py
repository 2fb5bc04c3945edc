"""Reference-independent known-answer checks that keep the oracle honest
(SURVEY.md section 4 iii): direct O(N^2) IMDCT, numpy IFFT, TDAC perfect
reconstruction through forward MDCT -> IMDCT+window+OLA, window identities."""
import numpy as np
import pytest

from oracle import oracle as O


def direct_imdct(N, x):
    n = np.arange(N)[:, None]
    k = np.arange(N // 2)[None, :]
    return (2.0 / N) * (np.cos(2 * np.pi / N * (n + 0.5 + N / 4) * (k + 0.5)) @ x.astype(np.float64))


@pytest.mark.parametrize("N,sigma", [(2048, 3e5), (256, 1e5)])
def test_imdct_matches_direct_formula(N, sigma):
    rng = np.random.default_rng(N)
    for _ in range(3):
        x = (rng.standard_normal(N // 2) * sigma).astype(np.float32)
        y = O.imdct(N, x).astype(np.float64)
        yd = direct_imdct(N, x)
        # full-scale PCM after /32768; float64-exact vs JS rounding model is ~1.2e-6
        assert np.abs(y - yd).max() / 32768 < 4e-6
        # IMDCT symmetries hold bit-exactly (each bin is written twice with sign, mdct.js:90-114)
        assert np.array_equal(y[: N // 2][::-1], -y[: N // 2])
        assert np.array_equal(y[N // 2:][::-1], y[N // 2:])


@pytest.mark.parametrize("L", [512, 64])
def test_fft_is_unscaled_inverse_dft(L):
    rng = np.random.default_rng(L)
    z = rng.standard_normal((L, 2)).astype(np.float32)
    got = O.fft_inverse(z)
    ref = np.fft.ifft(z[:, 0].astype(np.float64) + 1j * z[:, 1]) * L  # e^{+2 pi i nk/L}, no 1/L (fft.js:108)
    err = np.abs((got[:, 0] + 1j * got[:, 1]) - ref).max()
    assert err < 2e-4 * np.sqrt(L)


def test_window_power_complementarity():
    for which, n in ((4, 1024), (5, 1024), (6, 128), (7, 128)):
        w = O.table(which).astype(np.float64)
        assert w.size == n
        assert np.abs(w ** 2 + w[::-1] ** 2 - 1).max() < 2e-7
    assert O.table(5)[-1] == np.float32(1.0)  # KBD[len-1] rounds to exactly 1.0f


def forward_mdct(frame2N):
    """Forward MDCT of N2 windowed samples -> N2/2 coefficients with the AAC encoder's factor 2,
    so that the reference's IMDCT y = (2/N2) sum X cos (mdct.js) inverts it under TDAC."""
    N2 = frame2N.size  # 2048 or 256
    n = np.arange(N2)[None, :]
    k = np.arange(N2 // 2)[:, None]
    return (np.cos(2 * np.pi / N2 * (n + 0.5 + N2 / 4) * (k + 0.5)) @ frame2N) * 2.0


@pytest.mark.parametrize("shape", [0, 1])
def test_tdac_long_only(shape):
    """Princen-Bradley: window -> MDCT -> (oracle) IMDCT+window+OLA recovers the signal."""
    rng = np.random.default_rng(5 + shape)
    T = 6
    sig = rng.standard_normal((T + 1) * 1024) * 8000.0
    w = O.table(4 + shape).astype(np.float64)
    win = np.concatenate([w, w[::-1]])
    ov = np.zeros(1024, np.float32)
    info = O.make_info(0, shape, shape)
    outs = []
    for t in range(T):
        X = forward_mdct(sig[t * 1024:(t + 2) * 1024] * win)
        outs.append(O.filterbank(info, X.astype(np.float32), ov))
    rec = np.concatenate(outs[1:])  # first frame lacks its left neighbour
    assert np.abs(rec - sig[1024:T * 1024]).max() < 0.25  # |sig| peaks ~3e4 and X is rounded to f32: relative ~1e-5


def test_tdac_window_switching():
    """ONLY_LONG -> LONG_START -> EIGHT_SHORT x2 -> LONG_STOP -> ONLY_LONG reconstructs too."""
    rng = np.random.default_rng(11)
    seqs = [0, 0, 1, 2, 2, 3, 0, 0]
    T = len(seqs)
    sig = rng.standard_normal((T + 1) * 1024) * 8000.0
    wl = O.table(4).astype(np.float64)
    ws = O.table(6).astype(np.float64)
    ov = np.zeros(1024, np.float32)
    outs = []
    for t, sq in enumerate(seqs):
        blk = sig[t * 1024:(t + 2) * 1024]
        if sq == 2:
            X = np.empty(1024)
            for wdw in range(8):
                seg = blk[448 + 128 * wdw: 448 + 128 * wdw + 256]
                X[128 * wdw:128 * (wdw + 1)] = forward_mdct(seg * np.concatenate([ws, ws[::-1]]))
        else:
            first = wl if sq in (0, 1) else np.concatenate([np.zeros(448), ws, np.ones(448)])
            second = wl[::-1] if sq in (0, 3) else np.concatenate([np.ones(448), ws[::-1], np.zeros(448)])
            X = forward_mdct(blk * np.concatenate([first, second]))
        outs.append(O.filterbank(O.make_info(sq, 0, 0), X.astype(np.float32), ov))
    rec = np.concatenate(outs[1:])
    assert np.abs(rec - sig[1024:T * 1024]).max() < 0.25


def test_tns_as_shipped_is_identity_and_ar_inverts_ma():
    """AS_SHIPPED leaves the spectrum alone (tns.js:122); with the one-token fix the all-pole
    branch undoes the MA branch built from the same reflection coefficients."""
    from tools import workloads as W

    rng = np.random.default_rng(3)
    x = (rng.standard_normal(1024) * 1e4).astype(np.float32)
    coef = np.asarray(W.TNS_COEF_0_4, np.float32)[rng.choice(W.MILD, 12)]
    blk = W.tns_block([1, 0, 0, 0, 0, 0, 0, 0], [(49, 12, 0, coef)])
    info = O.make_info(0, 0, 0, max_sfb=49, tns_present=1)
    assert np.array_equal(O.tns(info, blk, 4, O.TNS_AS_SHIPPED, x), x)
    ma = O.tns(info, blk, 4, O.TNS_FIXED_MA, x)
    assert not np.array_equal(ma, x) and np.array_equal(ma[736:], x[736:])
    back = O.tns(info, blk, 4, O.TNS_FIXED_AR, ma)
    assert np.abs(back - x).max() < 0.1  # |x| ~ 3e4 peak: relative ~3e-6
