"""Synthetic AAC-LC bitstreams for end-to-end tests (test infrastructure, CPU only).

There is no encoded AAC file anywhere in the reference tree or in this image, so the tests that
drive the reference's whole `readChunk` (bit parse included) write their own access units: valid
raw_data_block syntax (ISO 14496-3 4.4.2: SCE / CPE elements, ics_info, section data, scalefactor
data, TNS data, Huffman-coded spectral data, END, byte alignment) behind ADTS headers, with random
but legal content.  The Huffman code tables are not restated here: they are read as DATA from the
reference's own src/huffman.js at run time through tools/jsmini.py, so this module only works
where /root/reference exists; what it produces is committed as fixtures (tests/golden/stream/).

    frames = random_frames(rng, n_frames, channels=2)       # descriptions (dicts)
    data   = write_adts_stream(frames, codebooks(), sample_index=4, channels=2)
"""
from __future__ import annotations

import os

import numpy as np

ONLY_LONG, LONG_START, EIGHT_SHORT, LONG_STOP = 0, 1, 2, 3
ZERO_BT, ESC_BT, NOISE_BT, INTENSITY_BT2, INTENSITY_BT = 0, 11, 13, 14, 15
SWB_LONG_COUNT = [41, 41, 47, 49, 49, 51, 47, 47, 43, 43, 43, 40]
SWB_SHORT_COUNT = [12, 12, 12, 14, 14, 14, 15, 15, 15, 15, 15, 15]
# largest absolute value each spectral codebook can carry (11: 16 = escape)
LAV = {1: 1, 2: 1, 3: 2, 4: 2, 5: 4, 6: 4, 7: 7, 8: 7, 9: 12, 10: 12, 11: 16}


class BitWriter:
    def __init__(self):
        self.bits = []

    def put(self, value, n):
        assert 0 <= value < (1 << n), (value, n)
        self.bits.extend((value >> (n - 1 - i)) & 1 for i in range(n))

    def align(self):
        self.bits.extend([0] * (-len(self.bits) % 8))

    def tobytes(self):
        self.align()
        return np.packbits(np.asarray(self.bits, np.uint8)).tobytes()


_CODEBOOKS = None


def codebooks(src_dir="/root/reference/src"):
    """{cb: {values tuple: (length, codeword)}} for cb 1..11 and {'sf': {value: (length, codeword)}},
    from the tables of the reference's huffman.js ([bit length, codeword, values...], :21-1419)."""
    global _CODEBOOKS
    if _CODEBOOKS is None:
        from tools import jsmini as J

        scope = J.Runtime(src_dir).run(open(os.path.join(src_dir, "huffman.js")).read(),
                                       {"module": J.obj(exports=J.obj())})
        books = {}
        for cb in range(1, 12):
            rows = scope[f"HCB{cb}"].items
            books[cb] = {tuple(int(v) for v in r.items[2:]): (int(r.items[0]), int(r.items[1])) for r in rows}
        books["sf"] = {int(r.items[2]): (int(r.items[0]), int(r.items[1])) for r in scope["HCB_SF"].items}
        _CODEBOOKS = books
    return _CODEBOOKS


# ---------------------------------------------------------------------------------- writing
def _write_ics_info(w, ics):
    w.put(0, 1)                                   # ics_reserved_bit
    w.put(ics["window_sequence"], 2)
    w.put(ics["window_shape"], 1)
    if ics["window_sequence"] == EIGHT_SHORT:
        w.put(ics["max_sfb"], 4)
        glen = ics["group_length"]
        bits = []
        for n in glen:                            # scale_factor_grouping: 1 = same group as the previous window
            bits.extend([1] * (n - 1) + [0])
        for b in bits[:7]:
            w.put(b, 1)
    else:
        w.put(ics["max_sfb"], 6)
        w.put(0, 1)                               # predictor_data_present


def _write_section_data(w, ics):
    short = ics["window_sequence"] == EIGHT_SHORT
    bits = 3 if short else 5
    esc = (1 << bits) - 1
    for sections in ics["sections"]:              # per group: [(codebook, n_bands), ...] covering max_sfb
        for cb, n in sections:
            w.put(cb, 4)
            while n >= esc:
                w.put(esc, bits)
                n -= esc
            w.put(n, bits)


def _write_scale_factors(w, ics, books):
    sf = books["sf"]
    offset = [ics["global_gain"], ics["global_gain"] - 90, 0]
    noise_first = True
    for g, sections in enumerate(ics["sections"]):
        k = 0
        for cb, n in sections:
            for b in range(k, k + n):
                v = ics["sf"][g][b]
                if cb == ZERO_BT:
                    continue
                if cb in (INTENSITY_BT, INTENSITY_BT2):
                    d = v - offset[2]; offset[2] = v
                elif cb == NOISE_BT:
                    if noise_first:
                        w.put(v - offset[1] + 256, 9); offset[1] = v; noise_first = False
                        continue
                    d = v - offset[1]; offset[1] = v
                else:
                    d = v - offset[0]; offset[0] = v
                length, code = sf[d + 60]
                w.put(code, length)
            k += n


def _write_tns(w, ics):
    short = ics["window_sequence"] == EIGHT_SHORT
    nb, lb, ob = (1, 4, 3) if short else (2, 6, 5)
    for filters in ics["tns"]:                    # per window: [(length, order, direction, coef_compress, [idx...]), ...]
        w.put(len(filters), nb)
        if not filters:
            continue
        coef_res = ics["tns_coef_res"]
        w.put(coef_res, 1)
        for length, order, direction, compress, idx in filters:
            w.put(length, lb)
            w.put(order, ob)
            if order:
                w.put(direction, 1)
                w.put(compress, 1)
                for i in idx:
                    w.put(i, coef_res + 3 - compress)


def _write_spectral(w, ics, books):
    short = ics["window_sequence"] == EIGHT_SHORT
    offsets = ics["swb_offsets"]
    q = ics["quant"]                              # [1024] ints, window-major for short frames
    group_off = 0
    for g, sections in enumerate(ics["sections"]):
        glen = ics["group_length"][g] if short else 1
        k = 0
        for cb, n in sections:
            for b in range(k, k + n):
                if cb in (ZERO_BT, NOISE_BT, INTENSITY_BT, INTENSITY_BT2):
                    continue
                width = offsets[b + 1] - offsets[b]
                step = 4 if cb < 5 else 2
                unsigned = cb in (3, 4, 7, 8, 9, 10, 11)
                for win in range(glen):
                    base = group_off + win * 128 + offsets[b]
                    for i in range(0, width, step):
                        vals = [int(v) for v in q[base + i: base + i + step]]
                        key = tuple(min(abs(v), 16) if cb == 11 else (abs(v) if unsigned else v) for v in vals)
                        length, code = books[cb][key]
                        w.put(code, length)
                        if unsigned:
                            for v in vals:
                                if v != 0:
                                    w.put(1 if v < 0 else 0, 1)
                        if cb == 11:
                            for v in vals:
                                if abs(v) >= 16:   # escape: N ones, a zero, then N+4 bits of |v| - 2^(N+4)
                                    nbits = abs(v).bit_length() - 1
                                    for _ in range(nbits - 4):
                                        w.put(1, 1)
                                    w.put(0, 1)
                                    w.put(abs(v) - (1 << nbits), nbits)
            k += n
        group_off += glen * 128


def _write_ics(w, ics, books, common_window):
    w.put(ics["global_gain"], 8)
    if not common_window:
        _write_ics_info(w, ics)
    _write_section_data(w, ics)
    _write_scale_factors(w, ics, books)
    w.put(0, 1)                                   # pulse_data_present
    w.put(1 if ics.get("tns") else 0, 1)
    if ics.get("tns"):
        _write_tns(w, ics)
    w.put(0, 1)                                   # gain_control_data_present
    _write_spectral(w, ics, books)


def write_raw_data_block(frame, books):
    """frame: list of elements, ("sce", ics) or ("cpe", common_window, ms_mask, ms_used, left, right)."""
    w = BitWriter()
    for tag, el in enumerate(frame):
        if el[0] in ("sce", "lfe"):               # single_channel_element / lfe_channel_element: same syntax
            w.put(0 if el[0] == "sce" else 3, 3); w.put(tag, 4)
            _write_ics(w, el[1], books, False)
        else:
            _, common, mask, ms_used, left, right = el
            w.put(1, 3); w.put(tag, 4)
            w.put(int(common), 1)
            if common:
                _write_ics_info(w, left)
                w.put(mask, 2)
                if mask == 1:
                    n = len(left["sections"]) * left["max_sfb"]
                    for i in range(n):
                        w.put(int(ms_used[i]), 1)
            _write_ics(w, left, books, common)
            _write_ics(w, right, books, common)
    w.put(7, 3)                                   # END
    return w.tobytes()


def write_adts_stream(frames, books, sample_index=4, channels=2):
    out = bytearray()
    for frame in frames:
        payload = write_raw_data_block(frame, books)
        w = BitWriter()
        w.put(0xfff, 12); w.put(0, 1); w.put(0, 2); w.put(1, 1)       # sync, MPEG-4, layer, protection_absent
        w.put(1, 2); w.put(sample_index, 4); w.put(0, 1); w.put(channels, 3)   # AAC LC (profile - 1 = 1)
        w.put(0, 4); w.put(7 + len(payload), 13); w.put(0x7ff, 11); w.put(0, 2)
        out += w.tobytes() + payload
    return bytes(out)


# ------------------------------------------------------------------------- random content
def _random_ics(rng, seq, shape, sample_index, stereo_right=False, allow_tns=True, like=None, p_noise=0.0):
    import aacjs_b200 as A   # scalefactor-band tables (host-side call, no GPU involved)

    short = seq == EIGHT_SHORT
    ics = {"window_sequence": seq, "window_shape": shape}
    if like is not None:                          # common window: groups and max_sfb are the left channel's
        ics["group_length"], ics["max_sfb"] = like["group_length"], like["max_sfb"]
    elif short:
        cuts = np.sort(rng.choice(np.arange(1, 8), size=int(rng.integers(0, 4)), replace=False))
        ics["group_length"] = [int(v) for v in np.diff(np.concatenate([[0], cuts, [8]]))]
        ics["max_sfb"] = int(rng.integers(4, SWB_SHORT_COUNT[sample_index] + 1))
    else:
        ics["group_length"] = [1]
        ics["max_sfb"] = int(rng.integers(20, SWB_LONG_COUNT[sample_index] + 1))
    n_groups, max_sfb = len(ics["group_length"]), ics["max_sfb"]
    offs_t = [int(v) for v in A.swb_offsets(sample_index, short)]
    ics["swb_offsets"] = offs_t
    ics["global_gain"] = int(rng.integers(140, 156))
    sections, sfs = [], []
    gain = ics["global_gain"]
    noise, inten = gain - 90, 0
    for _ in range(n_groups):
        secs, k, row = [], 0, []
        while k < max_sfb:
            n = min(max_sfb - k, int(rng.integers(1, 9)))
            r = rng.random()
            if r < 0.08:
                cb = ZERO_BT
            elif r < 0.08 + p_noise:   # perceptual noise substitution: off by default -- the reference's generator
                cb = NOISE_BT          # (ics.js:232-234) degenerates to zeros after ~15 values and the band becomes
                                       # 0 * (sf / sqrt(0)) = NaN (reference defects)
            elif stereo_right and r < 0.30:
                cb = INTENSITY_BT if rng.random() < 0.5 else INTENSITY_BT2
            else:
                cb = int(rng.integers(1, 12))
            secs.append((cb, n))
            for _b in range(n):
                if cb == ZERO_BT:
                    row.append(0)
                elif cb in (INTENSITY_BT, INTENSITY_BT2):
                    inten = int(np.clip(inten + rng.integers(-4, 5), -40, 40)); row.append(inten)
                elif cb == NOISE_BT:
                    noise = int(np.clip(noise + rng.integers(-3, 4), gain - 120, gain - 70)); row.append(noise)
                else:
                    gain = int(np.clip(gain + rng.integers(-5, 6), 130, 165)); row.append(gain)
            k += n
        sections.append(secs)
        sfs.append(row)
    ics["sections"], ics["sf"] = sections, sfs
    q = np.zeros(1024, np.int64)
    group_off = 0
    for g, secs in enumerate(sections):
        glen = ics["group_length"][g] if short else 1
        k = 0
        for cb, n in secs:
            if 1 <= cb <= 11:
                lav = LAV[cb]
                for b in range(k, k + n):
                    for win in range(glen):
                        lo, hi = group_off + win * 128 + offs_t[b], group_off + win * 128 + offs_t[b + 1]
                        v = rng.integers(-lav, lav + 1, hi - lo)
                        if cb == 11 and rng.random() < 0.5:     # a few escapes
                            v[rng.integers(0, hi - lo)] = int(rng.integers(17, 200)) * (1 if rng.random() < 0.5 else -1)
                        q[lo:hi] = v
            k += n
        group_off += glen * 128
    ics["quant"] = q
    if allow_tns and rng.random() < 0.4:
        ics["tns_coef_res"] = int(rng.integers(0, 2))
        tns = []
        for _w in range(8 if short else 1):
            filters = []
            for _f in range(int(rng.integers(0, 2 if short else 3))):
                order = int(rng.integers(0, 8 if short else 13))
                compress = int(rng.integers(0, 2))
                nb = ics["tns_coef_res"] + 3 - compress
                filters.append((int(rng.integers(1, 10 if short else 40)), order, int(rng.integers(0, 2)), compress,
                                [int(v) for v in rng.integers(0, 1 << nb, order)]))
            tns.append(filters)
        ics["tns"] = tns
    return ics


def random_frames(rng, n_frames, channels=2, sample_index=4):
    """Frame descriptions for a stream of `channels` (1: one SCE per frame, 2: one CPE per frame) whose
    window sequences follow the legal transitions."""
    from tools import workloads as W

    seqs = [W.legal_random_sequence(n_frames, rng) for _ in range(channels)]
    frames = []

    def pair(t, ca, cb):
        common = bool(rng.random() < 0.6)
        sl = int(seqs[ca][t])
        sr = sl if common else int(seqs[cb][t])
        shape = int(rng.integers(0, 2))
        left = _random_ics(rng, sl, shape, sample_index)
        right = _random_ics(rng, sr, shape if common else int(rng.integers(0, 2)), sample_index, stereo_right=True,
                            like=left if common else None)
        return ("cpe", common, int(rng.integers(0, 3)) if common else 0, rng.integers(0, 2, 128), left, right)

    for t in range(n_frames):
        if channels == 1:
            frames.append([("sce", _random_ics(rng, int(seqs[0][t]), int(rng.integers(0, 2)), sample_index))])
            continue
        if channels == 6:   # 5.1: centre, front pair, rear pair, LFE (channel configuration 6)
            frames.append([("sce", _random_ics(rng, int(seqs[0][t]), int(rng.integers(0, 2)), sample_index)),
                           pair(t, 1, 2), pair(t, 3, 4),
                           ("lfe", _random_ics(rng, int(seqs[5][t]), int(rng.integers(0, 2)), sample_index, allow_tns=False))])
            continue
        common = bool(rng.random() < 0.6)
        sl = int(seqs[0][t])
        sr = sl if common else int(seqs[1][t])
        shape = int(rng.integers(0, 2))
        left = _random_ics(rng, sl, shape, sample_index)
        right = _random_ics(rng, sr, shape if common else int(rng.integers(0, 2)), sample_index, stereo_right=True,
                            like=left if common else None)
        mask = int(rng.integers(0, 3)) if common else 0
        ms_used = rng.integers(0, 2, 128)
        frames.append([("cpe", common, mask, ms_used, left, right)])
    return frames
