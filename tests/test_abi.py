"""The C-ABI library: loads, exports exactly what include/aacfb.h declares, validates
arguments like the reference throws, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import aacjs_b200 as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    h = open(os.path.join(ROOT, "include", "aacfb.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(aacfb_[a-z_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert sorted(A.ABI_SYMBOLS) == syms
    out = subprocess.run(["nm", "-D", "--defined-only", A.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (aacfb_\w+)", out))
    assert exported == set(syms)
    L = A.lib()
    for s in syms:
        assert hasattr(L, s)
    assert L.aacfb_version() >= 100


def test_library_carries_sm100a_code_and_tma():
    """The shipped image is sm_100a SASS with 1-D TMA bulk copies (UBLKCP) and mbarriers."""
    r = subprocess.run(["cuobjdump", "-sass", A.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in r.stdout
    assert "UBLKCP" in r.stdout and "SYNCS" in r.stdout
    assert "synth_kernel" in r.stdout and "tns_kernel" in r.stdout
    # the two-chain arithmetic is packed FP32 (fma/mul/add.rn.f32x2), DESIGN.md section 4.1
    assert "FFMA2" in r.stdout and "FMUL2" in r.stdout and "FADD2" in r.stdout


def test_oracle_is_not_linked_into_the_product():
    out = subprocess.run(["ldd", A.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    strings = subprocess.run(["nm", "-D", A.LIB_PATH], capture_output=True, text=True).stdout
    assert "aacfb_oracle" not in strings and "aacfb_emul" not in strings


def test_small_frames_throw_like_the_reference():
    with pytest.raises(A.AacfbError, match="No small frames allowed"):  # filter_bank.js:26
        A.Context(1, 2, small_frames=True)
    with pytest.raises(A.AacfbError, match="No small frames allowed"):
        A.FilterBank(True, 2)


@pytest.mark.parametrize("kw", [dict(n_streams=0), dict(channels=0), dict(channels=9), dict(sample_index=12),
                                dict(flags=3), dict(flags=8)])
def test_create_rejects_bad_arguments(kw):
    with pytest.raises(A.AacfbError, match="AACFB_ERR_ARG"):
        A.Context(**{**dict(n_streams=1, channels=2), **kw})


def test_no_gpu_means_error_not_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(A.AacfbError, match="AACFB_ERR_CUDA"):
        A.Context(1, 2)


def test_tns_block_serialisation_and_order_check():
    t = A.TNS({"sampleIndex": 4})
    t.nFilt[0] = 1
    t.length[0][0], t.order[0][0], t.direction[0][0] = 49, 3, True
    t.coef[0][0][:3] = [0.1, 0.2, 0.3]
    b = t.block()
    assert b[:8] == bytes([1, 0, 0, 0, 0, 0, 0, 0]) and b[8:12] == bytes([49, 3, 1, 0]) and len(b) == 24
    assert np.allclose(np.frombuffer(b[12:], np.float32), [0.1, 0.2, 0.3])
    t.order[0][0] = 21
    with pytest.raises(A.AacfbError, match="TNS filter out of range"):  # tns.js:85
        t.block()


def test_napi_addon_source_compiles_against_the_header():
    """No Node on this image: syntax-check the addon against a stub node_api.h so that the
    JS binding cannot drift from include/aacfb.h unnoticed."""
    src = os.path.join(ROOT, "aac.js_b200", "js", "napi", "aacfb_napi.c")
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "stubs"), src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(src).read()
    for sym in ("aacfb_create", "aacfb_process", "aacfb_filterbank_process", "aacfb_tns_process", "aacfb_reset",
                "aacfb_get_overlap", "aacfb_set_overlap", "aacfb_destroy", "aacfb_last_error", "aacfb_process_io",
                "aacfb_host_alloc", "aacfb_host_free", "aacfb_get_swb_offsets", "aacfb_adts_index"):
        assert sym in text
