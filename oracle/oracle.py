"""ctypes binding of the CPU oracle (oracle/aacfb_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of aacfb_oracle.c.  Importable
from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs; never from the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libaacfb_oracle.so")

INFO_DTYPE = np.dtype(
    [("window_sequence", "u1"), ("shape_prev", "u1"), ("shape_cur", "u1"),
     ("max_sfb", "u1"), ("tns_present", "u1"), ("stereo_present", "u1"), ("reserved", "u1", (2,))])
assert INFO_DTYPE.itemsize == 8

TNS_AS_SHIPPED, TNS_FIXED_AR, TNS_FIXED_MA = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    src = os.path.join(_HERE, "aacfb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.aacfb_oracle_init.restype = None
        L.aacfb_oracle_table.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.aacfb_oracle_table_f64.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.aacfb_oracle_fft.argtypes = [C.c_int, C.c_void_p]
        L.aacfb_oracle_fft.restype = None
        L.aacfb_oracle_mdct.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.aacfb_oracle_mdct.restype = None
        L.aacfb_oracle_filterbank.argtypes = [C.c_void_p] * 4
        L.aacfb_oracle_filterbank.restype = None
        L.aacfb_oracle_tns.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_void_p]
        L.aacfb_oracle_tns.restype = None
        L.aacfb_oracle_process.argtypes = [C.c_void_p] * 6 + [C.c_int] * 4 + [C.c_uint32, C.c_int]
        L.aacfb_oracle_process.restype = C.c_int
        L.aacfb_oracle_stereo.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.aacfb_oracle_stereo.restype = None
        L.aacfb_oracle_adts_header.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.aacfb_oracle_adts_header.restype = C.c_int
        L.aacfb_oracle_dequant.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.aacfb_oracle_dequant.restype = None
        L.aacfb_oracle_dequant_table.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.aacfb_oracle_pcm_s16.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.aacfb_oracle_pcm_s16.restype = None
        L.aacfb_oracle_process_io.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 5 + [C.c_uint32] + [C.c_int] * 4 + [C.c_uint32, C.c_int]
        L.aacfb_oracle_process_io.restype = C.c_int
        L.aacfb_oracle_init()
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def table(which: int) -> np.ndarray:
    out = np.empty(1024, np.float32)
    n = lib().aacfb_oracle_table(which, _p(out), out.size)
    assert n > 0
    return out[:n].copy()


def table_f64(which: int) -> np.ndarray:
    out = np.empty(1024, np.float64)
    n = lib().aacfb_oracle_table_f64(which, _p(out), out.size)
    assert n > 0
    return out[:n].copy()


def fft_inverse(z: np.ndarray) -> np.ndarray:
    """fft.js:105-192 with forward=false on an [L][2] float32 AoS array."""
    a = np.ascontiguousarray(z, np.float32).copy()
    lib().aacfb_oracle_fft(a.shape[0], _p(a))
    return a


def imdct(N: int, x: np.ndarray) -> np.ndarray:
    """mdct.js:62-115: N/2 coefficients -> N samples."""
    x = np.ascontiguousarray(x, np.float32)
    assert x.size == N // 2
    y = np.empty(N, np.float32)
    lib().aacfb_oracle_mdct(N, _p(x), _p(y))
    return y


def make_info(window_sequence=0, shape_prev=0, shape_cur=0, max_sfb=0, tns_present=0) -> np.ndarray:
    r = np.zeros((), INFO_DTYPE)
    r["window_sequence"], r["shape_prev"], r["shape_cur"] = window_sequence, shape_prev, shape_cur
    r["max_sfb"], r["tns_present"] = max_sfb, tns_present
    return r


def filterbank(info: np.ndarray, x: np.ndarray, overlap: np.ndarray) -> np.ndarray:
    """filter_bank.js:88-204 for one channel-frame; `overlap` is updated in place."""
    x = np.ascontiguousarray(x, np.float32)
    assert overlap.dtype == np.float32 and overlap.flags.c_contiguous and overlap.size == 1024
    info = np.ascontiguousarray(info)
    out = np.empty(1024, np.float32)
    lib().aacfb_oracle_filterbank(_p(info), _p(x), _p(out), _p(overlap))
    return out


def tns(info: np.ndarray, block: bytes, sample_index: int, mode: int, data: np.ndarray) -> np.ndarray:
    """tns.js:105-177 on a copy of `data`."""
    d = np.ascontiguousarray(data, np.float32).copy()
    info = np.ascontiguousarray(info)
    b = np.frombuffer(block, np.uint8).copy() if len(block) else np.zeros(8, np.uint8)
    lib().aacfb_oracle_tns(_p(info), _p(b), len(block), sample_index, mode, _p(d))
    return d


def process(spectra, info, tns_blob=None, tns_offsets=None, overlap=None, *, sample_index=4,
            flags=TNS_AS_SHIPPED, n_threads=1):
    """Whole path for a batch (same contract as aacfb_process).

    spectra [S][T][C][1024] f32, info [S][T][C] INFO_DTYPE, overlap [S][C][1024]
    (updated in place; zeros if None).  Returns (pcm [S][T][1024][C], overlap)."""
    spectra = np.ascontiguousarray(spectra, np.float32)
    S, T, Cn, n = spectra.shape
    assert n == 1024
    info = np.ascontiguousarray(info, INFO_DTYPE).reshape(S, T, Cn)
    if overlap is None:
        overlap = np.zeros((S, Cn, 1024), np.float32)
    assert overlap.shape == (S, Cn, 1024) and overlap.dtype == np.float32 and overlap.flags.c_contiguous
    pcm = np.empty((S, T, 1024, Cn), np.float32)
    if tns_blob is not None:
        tns_blob = np.ascontiguousarray(tns_blob, np.uint8)
        tns_offsets = np.ascontiguousarray(tns_offsets, np.uint32)
        assert tns_offsets.size == S * T * Cn + 1
    rc = lib().aacfb_oracle_process(_p(spectra), _p(info), _p(tns_blob), _p(tns_offsets), _p(overlap), _p(pcm),
                                    S, T, Cn, sample_index, flags, n_threads)
    assert rc == 0
    return pcm, overlap


# What processMS / processIS read of a CPEElement and its two ICStreams (struct oracle_cpe).
CPE_DTYPE = np.dtype([("common_window", "i4"), ("mask_present", "i4"), ("ms_used", "u1", (128,)),
                      ("window_sequence", "i4", (2,)), ("group_count", "i4", (2,)), ("group_length", "i4", (2, 8)),
                      ("max_sfb", "i4", (2,)), ("band_types", "i4", (2, 120)), ("sect_end", "i4", (2, 120)),
                      ("scale_factors", "f4", (2, 120))])


def stereo(cpe: np.ndarray, sample_index: int, left: np.ndarray, right: np.ndarray):
    """processPair's stereo part (decoder.js:300-307): processMS (:379-404) if commonWindow and
    maskPresent, then processIS (:337-376), on copies of `left` / `right` (1024 f32 each)."""
    cpe = np.ascontiguousarray(cpe, CPE_DTYPE)
    l = np.ascontiguousarray(left, np.float32).copy()
    r = np.ascontiguousarray(right, np.float32).copy()
    lib().aacfb_oracle_stereo(_p(cpe), sample_index, _p(l), _p(r))
    return l, r


def adts_header(data: bytes):
    """ADTSDemuxer.readHeader (adts_demuxer.js:28-52) on the bytes at the start of `data`:
    dict(profile, samplingIndex, chanConfig, frameLength, numFrames, bits) or None for
    'Invalid ADTS header.'"""
    buf = np.frombuffer(bytes(data), np.uint8)
    out = np.zeros(6, np.uint32)
    if lib().aacfb_oracle_adts_header(_p(buf), buf.size, _p(out)) != 0:
        return None
    return dict(zip(("profile", "samplingIndex", "chanConfig", "frameLength", "numFrames", "bits"), (int(v) for v in out)))


# ---- inverse quantisation (ics.js:203-266) and the int16 PCM sink -----------------------------
QFRAME_DTYPE = np.dtype([("group_len", "u1", (8,)), ("band", "u2", (120,)), ("reserved", "u1", (8,)), ("q", "i2", (1024,))])
assert QFRAME_DTYPE.itemsize == 2304
IN_F32, IN_Q16, PCM_F32, PCM_S16 = 0, 1, 0, 1


def dequant(qframe: np.ndarray, info: np.ndarray, sample_index: int = 4) -> np.ndarray:
    """ICStream.decodeSpectralData's arithmetic on one aacfb_qframe record -> ics.data (1024 f32)."""
    qf = np.ascontiguousarray(qframe, QFRAME_DTYPE)
    info = np.ascontiguousarray(info, INFO_DTYPE)
    out = np.empty(1024, np.float32)
    lib().aacfb_oracle_dequant(_p(qf), _p(info), sample_index, _p(out))
    return out


def dequant_batch(qframes: np.ndarray, info: np.ndarray, sample_index: int = 4) -> np.ndarray:
    """dequant over an [S][T][C] array of records -> spectra [S][T][C][1024]."""
    qf = np.ascontiguousarray(qframes, QFRAME_DTYPE)
    info = np.ascontiguousarray(info, INFO_DTYPE).reshape(qf.shape)
    out = np.empty(qf.shape + (1024,), np.float32)
    flat_q, flat_i, flat_o = qf.reshape(-1), info.reshape(-1), out.reshape(-1, 1024)
    L = lib()
    for i in range(flat_q.size):
        L.aacfb_oracle_dequant(C.c_void_p(flat_q[i:i + 1].ctypes.data), C.c_void_p(flat_i[i:i + 1].ctypes.data),
                               sample_index, C.c_void_p(flat_o[i].ctypes.data))
    return out


def dequant_table(which: int) -> np.ndarray:
    """0: IQ_TABLE (8191), 1: SCALEFACTOR_TABLE (428) as the oracle builds them (tables.js:168-191)."""
    out = np.empty(8192, np.float32)
    n = lib().aacfb_oracle_dequant_table(which, _p(out), out.size)
    assert n > 0
    return out[:n].copy()


def pcm_s16(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty(x.shape, np.int16)
    lib().aacfb_oracle_pcm_s16(_p(x), _p(out), x.size)
    return out


def process_io(inp, in_format, info, tns_blob=None, tns_offsets=None, overlap=None, *, pcm_format=PCM_F32,
               sample_index=4, flags=TNS_AS_SHIPPED, n_threads=1):
    """aacfb_process_io's contract on the CPU: inp = spectra [S][T][C][1024] f32 (IN_F32) or records
    [S][T][C] (IN_Q16); returns (pcm [S][T][1024][C] f32 or i16, overlap)."""
    if in_format == IN_Q16:
        inp = np.ascontiguousarray(inp, QFRAME_DTYPE)
        S, T, Cn = inp.shape
    else:
        inp = np.ascontiguousarray(inp, np.float32)
        S, T, Cn, _ = inp.shape
    info = np.ascontiguousarray(info, INFO_DTYPE).reshape(S, T, Cn)
    if overlap is None:
        overlap = np.zeros((S, Cn, 1024), np.float32)
    pcm = np.empty((S, T, 1024, Cn), np.int16 if pcm_format == PCM_S16 else np.float32)
    if tns_blob is not None:
        tns_blob = np.ascontiguousarray(tns_blob, np.uint8)
        tns_offsets = np.ascontiguousarray(tns_offsets, np.uint32)
    rc = lib().aacfb_oracle_process_io(_p(inp), in_format, _p(info), _p(tns_blob), _p(tns_offsets), _p(overlap), _p(pcm),
                                       pcm_format, S, T, Cn, sample_index, flags, n_threads)
    assert rc == 0
    return pcm, overlap
