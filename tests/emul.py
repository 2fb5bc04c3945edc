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
