"""Synthetic workloads of BASELINE.json's configs, shared by
tests/, bench.py and __graft_entry__.smoke().  numpy only.

Layout: spectra [S][T][C][1024] f32, info [S][T][C] (aacfb_frame_info), TNS blob +
offsets in the include/aacfb.h format.
"""
from __future__ import annotations

import numpy as np

INFO_DTYPE = np.dtype(
    [("window_sequence", "u1"), ("shape_prev", "u1"), ("shape_cur", "u1"),
     ("max_sfb", "u1"), ("tns_present", "u1"), ("reserved", "u1", (3,))])

ONLY_LONG, LONG_START, EIGHT_SHORT, LONG_STOP = 0, 1, 2, 3

# TNS_COEF_0_4 (reference src/tns.js:58-61); the "mild" subset keeps |k| <= 0.588 so the
# all-pole filter stays well conditioned
TNS_COEF_0_4 = [0.0, -0.20791170, -0.40673664, -0.58778524, -0.74314481, -0.86602539, -0.95105654, -0.99452192,
                0.99573416, 0.96182561, 0.89516330, 0.79801720, 0.67369562, 0.52643216, 0.36124167, 0.18374951]
MILD = [0, 1, 2, 3, 13, 14, 15]


def pack_tns(blocks):
    offs, blob = [0], bytearray()
    for b in blocks:
        if b:
            blob += b
            blob += b"\0" * (-len(blob) % 4)
        offs.append(len(blob))
    blob_a = np.frombuffer(bytes(blob), np.uint8).copy() if blob else np.zeros(4, np.uint8)
    return blob_a, np.asarray(offs, np.uint32)


def tns_block(n_filt, filters) -> bytes:
    """n_filt: 8 counts; filters: list of (length, order, direction, coef[order]) in (w, filt) order."""
    b = bytearray(int(v) for v in n_filt)
    for length, order, direction, coef in filters:
        b += bytes([length, order, int(bool(direction)), 0])
        b += np.asarray(coef[:order], np.float32).tobytes()
    return bytes(b)


def config5_sequence(T):
    """per stream: t mod 16 == 11 -> LONG_START, 12,13 -> EIGHT_SHORT, 14 -> LONG_STOP, else ONLY_LONG."""
    m = np.arange(T) % 16
    seq = np.zeros(T, np.uint8)
    seq[m == 11] = LONG_START
    seq[(m == 12) | (m == 13)] = EIGHT_SHORT
    seq[m == 14] = LONG_STOP
    return seq


def legal_random_sequence(T, rng):
    """A random walk through the legal window-sequence transitions."""
    seq, cur = np.zeros(T, np.uint8), 0
    for t in range(T):
        if cur in (ONLY_LONG, LONG_STOP):
            nxt = rng.choice([ONLY_LONG, ONLY_LONG, LONG_START])
        elif cur == LONG_START:
            nxt = EIGHT_SHORT if rng.random() < 0.7 else LONG_STOP
        else:
            nxt = rng.choice([EIGHT_SHORT, LONG_STOP])
        seq[t] = cur = nxt
    return seq


def make(config: int, S: int, T: int, C: int = 2, seed: int = 0, shape_prev_mode: str = "as_shipped",
         side_only: bool = False):
    """Returns dict(spectra, info, tns_blob, tns_offsets, flags, sample_index).

    config 1/2: ONLY_LONG, TNS off.  3: EIGHT_SHORT.  4: ONLY_LONG + TNS (order 12, one filter of
    49 bands, direction alternating by frame, mild coefficients), mode FIXED_AR, 44.1 kHz.
    5: mixed long/short per config5_sequence.  shape_cur = s & 1; shape_prev = 0 ("as_shipped",
    reference defect C5) or the previous frame's shape_cur ("carried")."""
    rng = np.random.default_rng(seed)
    sigma = {1: 3.0e5, 2: 3.0e5, 3: 1.0e5, 4: 0.75e5, 5: 1.0e5}[config]
    spectra = None if side_only else rng.standard_normal((S, T, C, 1024), dtype=np.float32) * np.float32(sigma)
    info = np.zeros((S, T, C), INFO_DTYPE)
    info["shape_cur"] = (np.arange(S) & 1).astype(np.uint8)[:, None, None]
    if shape_prev_mode == "carried":
        info["shape_prev"] = info["shape_cur"]
    info["max_sfb"] = 49
    if config == 3:
        info["window_sequence"] = EIGHT_SHORT
        info["max_sfb"] = 14
    elif config == 5:
        info["window_sequence"] = config5_sequence(T)[None, :, None]
        info["max_sfb"] = np.where(info["window_sequence"] == EIGHT_SHORT, 14, 49)
    blob = offs = None
    flags = 0
    if config == 4:
        flags = 1  # AACFB_TNS_FIXED_AR
        info["tns_present"] = 1
        idx = rng.choice(MILD, size=(S, T, C, 12))
        coef = np.asarray(TNS_COEF_0_4, np.float32)[idx]
        hdr = np.zeros((S, T, C, 12), np.uint8)  # n_filt[8] + filter header(4)
        hdr[..., 0] = 1
        hdr[..., 8] = 49
        hdr[..., 9] = 12
        hdr[..., 10] = (np.arange(T) & 1).astype(np.uint8)[None, :, None]
        blk = np.concatenate([hdr, coef.view(np.uint8).reshape(S, T, C, 48)], axis=-1)  # 60 bytes each
        blob = np.ascontiguousarray(blk).reshape(-1)
        offs = (np.arange(S * T * C + 1, dtype=np.uint64) * 60).astype(np.uint32)
    return dict(spectra=spectra, info=info, tns_blob=blob, tns_offsets=offs, flags=flags, sample_index=4)


def random_case(S, T, C, rng, tns_mode=0, sigma=1e5, short_tns_orders=True):
    """Irregular test input: legal random sequences, random shapes, random TNS filters."""
    spectra = (rng.standard_normal((S, T, C, 1024)) * sigma).astype(np.float32)
    info = np.zeros((S, T, C), INFO_DTYPE)
    for s in range(S):
        for c in range(C):
            info["window_sequence"][s, :, c] = legal_random_sequence(T, rng)
    info["shape_cur"] = rng.integers(0, 2, (S, T, C))
    info["shape_prev"] = rng.integers(0, 2, (S, T, C))
    info["max_sfb"] = rng.integers(0, 52, (S, T, C))
    blob = offs = None
    if tns_mode:
        info["tns_present"] = rng.integers(0, 2, (S, T, C))
        mild = np.asarray(TNS_COEF_0_4, np.float32)[MILD]
        blocks = []
        for fi in info.reshape(-1):
            if not fi["tns_present"]:
                blocks.append(None)
                continue
            short = fi["window_sequence"] == EIGHT_SHORT
            nw = 8 if short else 1
            nf = [int(rng.integers(0, 2 if short else 4)) if w < nw else 0 for w in range(8)]
            filters = []
            for w in range(8):
                for _ in range(nf[w]):
                    # MA branch with order 20 is NaN by construction (tns.js:43,169): keep below 20 there
                    hi = 8 if short else (20 if tns_mode == 2 else 21)
                    order = int(rng.integers(0, hi))
                    filters.append((int(rng.integers(1, 15 if short else 40)), order, int(rng.integers(0, 2)),
                                    mild[rng.integers(0, len(mild), order)]))
            blocks.append(tns_block(nf, filters))
        blob, offs = pack_tns(blocks)
    return dict(spectra=spectra, info=info, tns_blob=blob, tns_offsets=offs, flags=tns_mode, sample_index=4)
