"""ctypes binding of the test-only CPU emulation of the kernel schedule
(aac.js_b200/csrc/aacfb_emul.cpp): the same __host__ __device__ phase code the
CUDA kernel runs, driven by 64 host threads per worker."""
import ctypes as C
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "aac.js_b200", "libaacfb_emul.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_SO)
        L.aacfb_emul_process.argtypes = [C.c_void_p] * 6 + [C.c_int] * 4 + [C.c_uint32, C.c_int]
        L.aacfb_emul_table.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.aacfb_emul_process_io.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 5 + [C.c_uint32] + [C.c_int] * 4 + [C.c_uint32, C.c_int]
        L.aacfb_emul_dequant.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.aacfb_emul_dequant.restype = None
        L.aacfb_emul_pcm_s16.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.aacfb_emul_pcm_s16.restype = None
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def process(spectra, info, tns_blob, tns_offsets, overlap, sample_index, flags, chunk_len):
    S, T, Cn, _ = spectra.shape
    pcm = np.empty((S, T, 1024, Cn), np.float32)
    rc = lib().aacfb_emul_process(_p(spectra), _p(np.ascontiguousarray(info)), _p(tns_blob), _p(tns_offsets),
                                  _p(overlap), _p(pcm), S, T, Cn, sample_index, flags, chunk_len)
    assert rc == 0
    return pcm


def table(which):
    out = np.empty(1024, np.float32)
    n = lib().aacfb_emul_table(which, _p(out), out.size)
    assert n > 0
    return out[:n].copy()


def process_io(inp, in_format, info, tns_blob, tns_offsets, overlap, sample_index, flags, chunk_len, pcm_format=0):
    """The emulated kernel schedule with quantised input (in_format 1) and / or int16 PCM (pcm_format 1)."""
    if in_format == 1:
        S, T, Cn = inp.shape
    else:
        S, T, Cn, _ = inp.shape
    pcm = np.empty((S, T, 1024, Cn), np.int16 if pcm_format == 1 else np.float32)
    rc = lib().aacfb_emul_process_io(_p(np.ascontiguousarray(inp)), in_format, _p(np.ascontiguousarray(info)), _p(tns_blob),
                                     _p(tns_offsets), _p(overlap), _p(pcm), pcm_format, S, T, Cn, sample_index, flags, chunk_len)
    assert rc == 0
    return pcm


def dequant(qframe, info, sample_index=4):
    out = np.empty(1024, np.float32)
    lib().aacfb_emul_dequant(_p(np.ascontiguousarray(qframe)), _p(np.ascontiguousarray(info)), sample_index, _p(out))
    return out


def pcm_s16(x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty(x.shape, np.int16)
    lib().aacfb_emul_pcm_s16(_p(x), _p(out), x.size)
    return out
