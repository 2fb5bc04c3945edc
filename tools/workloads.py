"""Synthetic workloads of BASELINE.json's configs, shared by
tests/, bench.py and __graft_entry__.smoke().  numpy only.

Layout: spectra [S][T][C][1024] f32, info [S][T][C] (aacfb_frame_info), TNS blob +
offsets in the include/aacfb.h format.
"""
from __future__ import annotations

import numpy as np

INFO_DTYPE = np.dtype(
    [("window_sequence", "u1"), ("shape_prev", "u1"), ("shape_cur", "u1"),
     ("max_sfb", "u1"), ("tns_present", "u1"), ("stereo_present", "u1"), ("reserved", "u1", (2,))])

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


# ------------------------------------------------------------------ stereo tools (CPE side info)
# What processMS / processIS read of a CPEElement and its two ICStreams (reference src/cpe.js:24-75,
# src/ics.js:25-34,270-310); index 0 = left, 1 = right.  Same layout as struct oracle_cpe.
CPE_DTYPE = np.dtype([("common_window", "i4"), ("mask_present", "i4"), ("ms_used", "u1", (128,)),
                      ("window_sequence", "i4", (2,)), ("group_count", "i4", (2,)), ("group_length", "i4", (2, 8)),
                      ("max_sfb", "i4", (2,)), ("band_types", "i4", (2, 120)), ("sect_end", "i4", (2, 120)),
                      ("scale_factors", "f4", (2, 120))])
SWB_LONG_COUNT = [41, 41, 47, 49, 49, 51, 47, 47, 43, 43, 43, 40]    # tables.js:161-163
SWB_SHORT_COUNT = [12, 12, 12, 14, 14, 14, 15, 15, 15, 15, 15, 15]   # tables.js:157-159
NOISE_BT, INTENSITY_BT2, INTENSITY_BT = 13, 14, 15                   # ics.js:39-41


def _sections(rng, n_groups, max_sfb, p_intensity):
    """band_types / sect_end / scale_factors: per group, sections of random length with one band
    type each, the way ics.js:130-170 leaves them (sect_end = end band of the section, per band)."""
    bt, se = np.zeros(120, np.int32), np.zeros(120, np.int32)
    idx = 0
    for _ in range(n_groups):
        k = 0
        while k < max_sfb:
            end = min(max_sfb, k + int(rng.integers(1, 9)))
            r = rng.random()
            t = INTENSITY_BT if r < p_intensity / 2 else INTENSITY_BT2 if r < p_intensity else \
                NOISE_BT if r < p_intensity + 0.1 else int(rng.integers(0, 12))
            bt[idx:idx + end - k] = t
            se[idx:idx + end - k] = end
            idx += end - k
            k = end
    sf = (0.5 ** (rng.integers(-40, 40, 120) / 4.0)).astype(np.float32)
    return bt, se, sf


def _random_ics(rng, seq, sample_index, p_intensity, like=None):
    """(group_count, group_length[8], max_sfb, band_types, sect_end, scale_factors) of one ICStream;
    `like`: take the groups and maxSFB of that stream (common window)."""
    if like is not None:
        n_groups, gl, max_sfb = like[0], like[1], like[2]
    else:
        if seq == EIGHT_SHORT:
            cuts = np.sort(rng.choice(np.arange(1, 8), size=int(rng.integers(0, 8)), replace=False))
            glen = np.diff(np.concatenate([[0], cuts, [8]])).astype(np.int32)
            max_sfb = int(rng.integers(0, SWB_SHORT_COUNT[sample_index] + 1))
        else:
            glen = np.array([1], np.int32)
            max_sfb = int(rng.integers(0, SWB_LONG_COUNT[sample_index] + 1))
        n_groups, gl = len(glen), np.zeros(8, np.int32)
        gl[:n_groups] = glen
    return (n_groups, gl, max_sfb) + _sections(rng, n_groups, max_sfb, p_intensity)


def random_cpe(rng, seq_left, seq_right, sample_index=4, common_window=None, p_intensity=0.3):
    """One random channel pair element for the window sequences given (common window forces the
    left one on both channels, cpe.js:40-42)."""
    e = np.zeros((), CPE_DTYPE)
    cw = bool(rng.integers(0, 2)) if common_window is None else bool(common_window)
    if seq_left != seq_right:
        cw = False
    e["common_window"] = cw
    mask = int(rng.integers(0, 3)) if cw else 0          # cpe.js:44-66
    e["mask_present"] = int(mask != 0)
    e["ms_used"] = rng.integers(0, 2, 128) if mask == 1 else (1 if mask == 2 else 0)
    left = _random_ics(rng, seq_left, sample_index, 0.0)  # intensity codebooks occur in the right channel
    right = _random_ics(rng, seq_right, sample_index, p_intensity)
    if cw:   # right.info = left.info: groups and maxSFB are shared, the sections are not
        right = _random_ics(rng, seq_left, sample_index, p_intensity, like=left)
    for c, ics in enumerate((left, right)):
        e["window_sequence"][c] = seq_left if (cw or c == 0) else seq_right
        e["group_count"][c], e["group_length"][c], e["max_sfb"][c] = ics[0], ics[1], ics[2]
        e["band_types"][c], e["sect_end"][c], e["scale_factors"][c] = ics[3], ics[4], ics[5]
    return e


def random_stereo_case(S, T, rng, tns_mode=0, sigma=1e5, sample_index=4):
    """random_case for stereo streams plus one CPE per pair-frame: spectra are ics.data BEFORE
    processMS / processIS.  Returns the case dict with `cpe` [S][T] (CPE_DTYPE)."""
    case = random_case(S, T, 2, rng, tns_mode=tns_mode, sigma=sigma)
    cpe = np.zeros((S, T), CPE_DTYPE)
    for s in range(S):
        for t in range(T):
            ws = case["info"]["window_sequence"][s, t]
            cpe[s, t] = random_cpe(rng, int(ws[0]), int(ws[1]), sample_index)
    case["cpe"] = cpe
    case["sample_index"] = sample_index
    return case


STEREO_DTYPE = np.dtype([("op", "u1", (256,)), ("scale", "f4", (128,))])   # aacfb_stereo_ops


def joint_stereo_ops(S, T, seed=0):
    """Stereo side info of a "joint stereo" batch for bench.py: every frame of every stream is a
    common-window pair with M/S on scalefactor bands 0..39 (coefficients 0..511, ms_used all set)
    and intensity stereo on bands 40..45 (512..767, six scales), the rest untouched -- long-window
    band edges of 44.1 kHz (tables.js:64-68).  Returns ops [S][T][1] (STEREO_DTYPE)."""
    rng = np.random.default_rng(seed)
    ops = np.zeros((S, T, 1), STEREO_DTYPE)
    edges = [512, 544, 576, 608, 640, 672, 704]   # swbOffsets[40..46] at 44.1/48 kHz
    ops["op"][..., : 512 // 4] = 1
    for k in range(6):
        ops["op"][..., edges[k] // 4: edges[k + 1] // 4] = 2 + k
    ops["scale"][..., :6] = (0.5 ** (rng.integers(-8, 8, (S, T, 1, 6)) / 4.0)).astype(np.float32)
    return ops


def adts_stream(rng, n_frames, crc_every=3, sampling_index=4, chan_config=2):
    """A synthetic ADTS byte stream: n_frames access units with random payload sizes; every
    `crc_every`-th header carries the 16-bit CRC field (protection_absent = 0).  Returns
    (bytes, [(offset, frame_length, header_bytes, profile, num_frames)])."""
    out, meta = bytearray(), []
    for i in range(n_frames):
        prot_absent = 0 if (crc_every and i % crc_every == crc_every - 1) else 1
        hdr = 7 if prot_absent else 9
        length = hdr + int(rng.integers(0, 1500))
        profile_field, nf_field = int(rng.integers(0, 4)), int(rng.integers(0, 4))
        bits = 0
        def put(v, n):
            nonlocal bits
            bits = (bits << n) | (v & ((1 << n) - 1))
        put(0xfff, 12); put(int(rng.integers(0, 8)), 3); put(prot_absent, 1)
        put(profile_field, 2); put(sampling_index, 4); put(int(rng.integers(0, 2)), 1); put(chan_config, 3)
        put(int(rng.integers(0, 16)), 4); put(length, 13); put(int(rng.integers(0, 2048)), 11); put(nf_field, 2)
        h = bits.to_bytes(7, "big")
        meta.append((len(out), length, hdr, profile_field + 1, nf_field + 1))
        out += h + bytes(rng.integers(0, 256, length - 7, dtype=np.uint8))
    return bytes(out), meta


# ------------------------------------------------------------------ quantised input
QFRAME_DTYPE = np.dtype([("group_len", "u1", (8,)), ("band", "u2", (120,)), ("reserved", "u1", (8,)), ("q", "i2", (1024,))])
BAND_ZERO, BAND_SPECTRAL, BAND_NOISE = 0x0000, 0x4000, 0x8000


def make_q(config: int, S: int, T: int, C: int = 2, seed: int = 0, sigma_q: float = 40.0):
    """The quantised twin of make(): the same window sequences / shapes / TNS side info, but the input is
    what the bit parse holds BEFORE inverse quantisation -- aacfb_qframe records: Huffman-decoded integers
    q ~ round(N(0, sigma_q^2)) and one scalefactor per band, drawn within +-2 octaves of the value that gives
    the float workload's spectral level (rms(|q|^(4/3)) * 2^((i - 200) / 4) ~ sigma).  Every band below
    maxSFB is spectral; everything above it is zero (as in a real stream).  Returns make()'s dict with
    `qframes` [S][T][C] instead of `spectra`."""
    w = make(config, S, T, C, seed=seed, side_only=True)
    rng = np.random.default_rng(seed + 7919)
    sigma = {1: 3.0e5, 2: 3.0e5, 3: 1.0e5, 4: 0.75e5, 5: 1.0e5}[config]
    qf = np.zeros((S, T, C), QFRAME_DTYPE)
    q = np.rint(rng.standard_normal((S, T, C, 1024), dtype=np.float32) * np.float32(sigma_q))
    qf["q"] = np.clip(q, -8190, 8190).astype(np.int16)
    level = float(np.sqrt(np.mean(np.abs(q[:1]).astype(np.float64) ** (8.0 / 3.0))))
    i0 = int(round(200 + 4 * np.log2(sigma / level))) - 4   # the +-2 octave spread below raises the rms by ~1.9
    is_short = w["info"]["window_sequence"] == EIGHT_SHORT
    qf["group_len"][..., 0] = np.where(is_short, 2, 1)          # EIGHT_SHORT: groups of 2, 3, 3 windows
    qf["group_len"][..., 1] = np.where(is_short, 3, 0)
    qf["group_len"][..., 2] = np.where(is_short, 3, 0)
    qf["band"] = (BAND_SPECTRAL | np.clip(i0 + rng.integers(-8, 9, (S, T, C, 120)), 0, 427)).astype(np.uint16)
    w["qframes"] = qf
    return w


def random_q_case(S, T, C, rng, tns_mode=0, sample_index=4, p_noise=0.08, p_zero=0.1, sigma_q=60.0):
    """Irregular quantised input: random_case's window sequences / shapes / TNS, random window groups,
    random maxSFB, every band kind (zero, spectral, noise), scalefactor indices over the whole table and a
    sprinkling of extreme integers (+-8190, +-8191 = the reference's out-of-table read)."""
    case = random_case(S, T, C, rng, tns_mode=tns_mode)
    del case["spectra"]
    info = case["info"]
    qf = np.zeros((S, T, C), QFRAME_DTYPE)
    q = np.rint(rng.standard_normal((S, T, C, 1024)) * sigma_q)
    big = rng.random((S, T, C, 1024)) < 0.002
    q[big] = rng.choice([-8191, -8190, -4000, 4000, 8190, 8191], size=int(big.sum()))
    qf["q"] = q.astype(np.int16)
    for s in range(S):
        for t in range(T):
            for c in range(C):
                short = info["window_sequence"][s, t, c] == EIGHT_SHORT
                if short:
                    cuts = np.sort(rng.choice(np.arange(1, 8), size=int(rng.integers(0, 8)), replace=False))
                    glen = np.diff(np.concatenate([[0], cuts, [8]]))
                    info["max_sfb"][s, t, c] = rng.integers(0, SWB_SHORT_COUNT[sample_index] + 1)
                else:
                    glen = np.array([1])
                    info["max_sfb"][s, t, c] = rng.integers(0, SWB_LONG_COUNT[sample_index] + 1)
                qf["group_len"][s, t, c, :len(glen)] = glen
                r = rng.random(120)
                kind = np.where(r < p_zero, BAND_ZERO, np.where(r < p_zero + p_noise, BAND_NOISE, BAND_SPECTRAL))
                idx = rng.integers(150, 300, 120)
                odd = rng.random(120) < 0.02
                idx[odd] = rng.choice([0, 427, 428, 511], size=int(odd.sum()))
                qf["band"][s, t, c] = (kind | idx).astype(np.uint16)
    case["qframes"] = qf
    case["sample_index"] = sample_index
    return case
