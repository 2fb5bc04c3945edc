"""aac.js_b200 -- host-side mirror of the aac.js filterbank seam over the B200 C-ABI.

The reference is JavaScript; no JS runtime exists in this image, so the host
side that tests and benchmarks drive is this Python mirror of the reference's
own interface for the path (same class names, argument meaning and error
behaviour):

    FilterBank(smallFrames, channels).process(info, input, output, channel)
        reference src/filter_bank.js:24-44, 88-204
    TNS(config).process(ics, data, decode)
        reference src/tns.js:22-44, 105-177
    AACDecoder.process(elements) + the interleave of readChunk
        reference src/decoder.js:204-213, 218-334

Everything is a thin ctypes call into ``libaacfb.so`` (include/aacfb.h) -- the
same C entry points the N-API addon in ``js/`` binds.  There is NO CPU
implementation here: if the CUDA library is missing or no sm_100 device is
usable, construction raises.  (Because the directory name contains a dot the
package is imported through the ``aacjs_b200`` shim at the repo root.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AACFB_LIB") or os.path.join(_HERE, "libaacfb.so")  # AACFB_LIB: tuning variants

ONLY_LONG_SEQUENCE, LONG_START_SEQUENCE, EIGHT_SHORT_SEQUENCE, LONG_STOP_SEQUENCE = 0, 1, 2, 3  # ics.js:44-47
TNS_AS_SHIPPED, TNS_FIXED_AR, TNS_FIXED_MA = 0, 1, 2
TNS_MAX_ORDER = 20  # tns.js:46

INFO_DTYPE = np.dtype(
    [("window_sequence", "u1"), ("shape_prev", "u1"), ("shape_cur", "u1"),
     ("max_sfb", "u1"), ("tns_present", "u1"), ("stereo_present", "u1"), ("reserved", "u1", (2,))])

ERRORS = {-1: "AACFB_ERR_ARG", -2: "AACFB_ERR_SMALL", -3: "AACFB_ERR_SEQUENCE", -4: "AACFB_ERR_TNS",
          -5: "AACFB_ERR_CUDA", -6: "AACFB_ERR_NOMEM", -7: "AACFB_ERR_ADTS"}

# every symbol include/aacfb.h declares
ABI_SYMBOLS = ["aacfb_create", "aacfb_destroy", "aacfb_reset", "aacfb_process", "aacfb_process_device",
               "aacfb_filterbank_process", "aacfb_tns_process", "aacfb_get_overlap", "aacfb_set_overlap",
               "aacfb_last_error", "aacfb_version", "aacfb_launch_count", "aacfb_get_table",
               "aacfb_process_stereo", "aacfb_process_device_stereo", "aacfb_get_swb_offsets", "aacfb_adts_index",
               "aacfb_process_io", "aacfb_process_device_io", "aacfb_host_alloc", "aacfb_host_free",
               "aacfb_host_register", "aacfb_host_unregister"]

# aacfb_qframe: one channel-frame BEFORE inverse quantisation (include/aacfb.h)
QFRAME_DTYPE = np.dtype([("group_len", "u1", (8,)), ("band", "u2", (120,)), ("reserved", "u1", (8,)), ("q", "i2", (1024,))])
assert QFRAME_DTYPE.itemsize == 2304
BAND_ZERO, BAND_SPECTRAL, BAND_NOISE, BAND_UNDEFINED = 0x0000, 0x4000, 0x8000, 0x01ff
IN_F32, IN_Q16 = 0, 1
PCM_F32, PCM_S16 = 0, 1
ZERO_BT = 0  # ics.js:35

ADTS_FRAME_DTYPE = np.dtype([("offset", "u8"), ("frame_length", "u4"), ("header_bytes", "u1"), ("profile", "u1"),
                             ("sampling_index", "u1"), ("chan_config", "u1"), ("num_frames", "u1"), ("reserved", "u1", (7,))])
assert ADTS_FRAME_DTYPE.itemsize == 24

# aacfb_stereo_ops: what to do to each group of 4 coefficients of a channel pair (include/aacfb.h)
STEREO_DTYPE = np.dtype([("op", "u1", (256,)), ("scale", "f4", (128,))])
STEREO_NONE, STEREO_MS, STEREO_IS = 0, 1, 2
NOISE_BT, INTENSITY_BT2, INTENSITY_BT = 13, 14, 15  # ics.js:39-41


class AacfbError(RuntimeError):
    """The JS shim rethrows library errors as ``Error(msg)``; this is its Python twin."""

    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load libaacfb.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU implementation of this path)")
        L = C.CDLL(LIB_PATH)
        vp, ci, u32 = C.c_void_p, C.c_int, C.c_uint32
        L.aacfb_create.argtypes = [C.POINTER(vp), ci, ci, ci, ci, ci, u32]
        L.aacfb_destroy.argtypes = [vp]
        L.aacfb_reset.argtypes = [vp]
        L.aacfb_process.argtypes = [vp, vp, vp, vp, vp, vp, ci]
        L.aacfb_process_device.argtypes = [vp, vp, vp, vp, vp, C.c_size_t, vp, ci, vp]
        L.aacfb_filterbank_process.argtypes = [vp, ci, ci, vp, vp, vp]
        L.aacfb_tns_process.argtypes = [vp, vp, vp, C.c_size_t, vp, u32]
        L.aacfb_get_overlap.argtypes = [vp, vp]
        L.aacfb_set_overlap.argtypes = [vp, vp]
        L.aacfb_last_error.argtypes = [vp]
        L.aacfb_last_error.restype = C.c_char_p
        L.aacfb_launch_count.argtypes = [vp]
        L.aacfb_launch_count.restype = C.c_uint64
        L.aacfb_get_table.argtypes = [ci, vp, ci]
        if hasattr(L, "aacfb_process_stereo"):  # (older tuning variants loaded through AACFB_LIB lack them)
            L.aacfb_process_stereo.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci]
            L.aacfb_process_device_stereo.argtypes = [vp, vp, vp, vp, vp, vp, C.c_size_t, vp, ci, vp]
            L.aacfb_get_swb_offsets.argtypes = [ci, ci, vp, ci]
        if hasattr(L, "aacfb_adts_index"):
            L.aacfb_adts_index.argtypes = [vp, C.c_size_t, vp, ci, C.POINTER(C.c_size_t)]
        if hasattr(L, "aacfb_process_io"):
            L.aacfb_process_io.argtypes = [vp, vp, u32, vp, vp, vp, vp, vp, u32, ci]
            L.aacfb_process_device_io.argtypes = [vp, vp, u32, vp, vp, vp, vp, C.c_size_t, vp, u32, ci, vp]
            L.aacfb_host_alloc.argtypes = [C.c_size_t]
            L.aacfb_host_alloc.restype = vp
            L.aacfb_host_free.argtypes = [vp]
            L.aacfb_host_register.argtypes = [vp, C.c_size_t]
            L.aacfb_host_unregister.argtypes = [vp]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def get_table(which: int) -> np.ndarray:
    out = np.empty(1024, np.float32)
    n = lib().aacfb_get_table(which, _ptr(out), out.size)
    if n < 0:
        raise AacfbError(n, "bad table index")
    return out[:n].copy()


def swb_offsets(sample_index: int, is_short: bool) -> np.ndarray:
    """info.swbOffsets of one sample rate (tables.js:126-154): swbCount + 1 entries."""
    out = np.empty(64, np.uint16)
    n = lib().aacfb_get_swb_offsets(sample_index, int(bool(is_short)), _ptr(out), out.size)
    if n < 0:
        raise AacfbError(n, "bad sample index")
    return out[:n + 1].copy()


def adts_index(data, capacity: int | None = None):
    """Locate the access units of an ADTS byte buffer from their headers alone (the reference's
    ADTSDemuxer.readHeader, adts_demuxer.js:28-52, hopping by frameLength): returns
    (frames [n] ADTS_FRAME_DTYPE, consumed) -- `consumed` is where the first incomplete frame
    starts.  Raises like the reference on a bad syncword ('Invalid ADTS header.')."""
    buf = np.frombuffer(bytes(data), np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
    consumed = C.c_size_t(0)
    if capacity is None:
        capacity = lib().aacfb_adts_index(_ptr(buf) if buf.size else None, buf.size, None, 0, C.byref(consumed))
        if capacity < 0:
            raise AacfbError(capacity, lib().aacfb_last_error(None).decode())
    frames = np.zeros(max(capacity, 1), ADTS_FRAME_DTYPE)
    n = lib().aacfb_adts_index(_ptr(buf) if buf.size else None, buf.size, _ptr(frames), capacity, C.byref(consumed))
    if n < 0:
        raise AacfbError(n, lib().aacfb_last_error(None).decode())
    return frames[:n], int(consumed.value)


def host_alloc(shape, dtype) -> np.ndarray:
    """A numpy array on page-locked memory owned by the library (aacfb_host_alloc) -- what the N-API
    addon wraps in an external ArrayBuffer for the decoder's staging typed arrays.  Free with host_free."""
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    p = lib().aacfb_host_alloc(n)
    if not p:
        raise AacfbError(-5, lib().aacfb_last_error(None).decode())
    buf = (C.c_uint8 * n).from_address(p)
    a = np.frombuffer(buf, dtype=dt).reshape(shape)
    _host_allocs[a.ctypes.data] = p
    return a


_host_allocs: dict = {}


def host_free(a: np.ndarray):
    p = _host_allocs.pop(a.ctypes.data, None)
    if p:
        lib().aacfb_host_free(C.c_void_p(p))


def host_register(a: np.ndarray):
    """Page-lock an existing contiguous array for the life of the decoder (aacfb_host_register)."""
    rc = lib().aacfb_host_register(C.c_void_p(a.ctypes.data), a.nbytes)
    if rc != 0:
        raise AacfbError(rc, lib().aacfb_last_error(None).decode())


def host_unregister(a: np.ndarray):
    lib().aacfb_host_unregister(C.c_void_p(a.ctypes.data))


class Context:
    """One aacfb_ctx: S decoder instances of `channels` channels on one GPU."""

    def __init__(self, n_streams=1, channels=2, sample_index=4, flags=TNS_AS_SHIPPED, device=0, small_frames=False):
        self._h = C.c_void_p()
        self.n_streams, self.channels, self.sample_index, self.flags = n_streams, channels, sample_index, flags
        rc = lib().aacfb_create(C.byref(self._h), device, n_streams, channels, sample_index, int(bool(small_frames)),
                                flags)
        if rc != 0:
            raise AacfbError(rc, lib().aacfb_last_error(None).decode())

    def _check(self, rc):
        if rc != 0:
            raise AacfbError(rc, lib().aacfb_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().aacfb_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def reset(self):
        self._check(lib().aacfb_reset(self._h))

    @property
    def launches(self) -> int:
        return int(lib().aacfb_launch_count(self._h))

    def get_overlap(self) -> np.ndarray:
        out = np.empty((self.n_streams, self.channels, 1024), np.float32)
        self._check(lib().aacfb_get_overlap(self._h, _ptr(out)))
        return out

    def set_overlap(self, ov):
        ov = np.ascontiguousarray(ov, np.float32)
        assert ov.shape == (self.n_streams, self.channels, 1024)
        self._check(lib().aacfb_set_overlap(self._h, _ptr(ov)))

    def process(self, spectra, info, tns_blob=None, tns_offsets=None, out=None, stereo_ops=None) -> np.ndarray:
        """Batched hot path with host arrays: spectra [S][T][C][1024] -> pcm [S][T][1024][C].

        stereo_ops [S][T][C/2] (STEREO_DTYPE): the spectra are ics.data BEFORE processMS / processIS
        (decoder.js:300-307) and the device applies the ops of the records flagged by stereo_present."""
        spectra = np.ascontiguousarray(spectra, np.float32)
        S, T, Cn, n = spectra.shape
        assert (S, Cn, n) == (self.n_streams, self.channels, 1024), spectra.shape
        info = np.ascontiguousarray(info, INFO_DTYPE)
        assert info.size == S * T * Cn
        if tns_blob is not None:
            tns_blob = np.ascontiguousarray(tns_blob, np.uint8)
            tns_offsets = np.ascontiguousarray(tns_offsets, np.uint32)
            assert tns_offsets.size == S * T * Cn + 1
        pcm = out if out is not None else np.empty((S, T, 1024, Cn), np.float32)
        assert pcm.dtype == np.float32 and pcm.flags.c_contiguous and pcm.size == spectra.size
        if stereo_ops is not None:
            stereo_ops = np.ascontiguousarray(stereo_ops, STEREO_DTYPE)
            assert stereo_ops.size * 2 == S * T * Cn, stereo_ops.shape
            self._check(lib().aacfb_process_stereo(self._h, _ptr(spectra), _ptr(info), _ptr(stereo_ops), _ptr(tns_blob),
                                                   _ptr(tns_offsets), _ptr(pcm), T))
            return pcm
        self._check(lib().aacfb_process(self._h, _ptr(spectra), _ptr(info), _ptr(tns_blob), _ptr(tns_offsets),
                                        _ptr(pcm), T))
        return pcm

    def process_io(self, inp, info, tns_blob=None, tns_offsets=None, out=None, stereo_ops=None, *, in_format=IN_F32,
                   pcm_format=PCM_F32) -> np.ndarray:
        """aacfb_process_io: `inp` is spectra [S][T][C][1024] f32 (IN_F32) or aacfb_qframe records [S][T][C]
        (IN_Q16: the device dequantises, ics.js:203-266); returns pcm [S][T][1024][C] as f32 (/32768) or int16."""
        if in_format == IN_Q16:
            inp = np.ascontiguousarray(inp, QFRAME_DTYPE)
            S, T, Cn = inp.shape
        else:
            inp = np.ascontiguousarray(inp, np.float32)
            S, T, Cn, n = inp.shape
            assert n == 1024
        assert (S, Cn) == (self.n_streams, self.channels), inp.shape
        info = np.ascontiguousarray(info, INFO_DTYPE)
        assert info.size == S * T * Cn
        if tns_blob is not None:
            tns_blob = np.ascontiguousarray(tns_blob, np.uint8)
            tns_offsets = np.ascontiguousarray(tns_offsets, np.uint32)
            assert tns_offsets.size == S * T * Cn + 1
        dt = np.int16 if pcm_format == PCM_S16 else np.float32
        pcm = out if out is not None else np.empty((S, T, 1024, Cn), dt)
        assert pcm.dtype == dt and pcm.flags.c_contiguous and pcm.size == S * T * Cn * 1024
        if stereo_ops is not None:
            stereo_ops = np.ascontiguousarray(stereo_ops, STEREO_DTYPE)
            assert stereo_ops.size * 2 == S * T * Cn, stereo_ops.shape
        self._check(lib().aacfb_process_io(self._h, _ptr(inp), in_format, _ptr(info), _ptr(stereo_ops), _ptr(tns_blob),
                                           _ptr(tns_offsets), _ptr(pcm), pcm_format, T))
        return pcm

    def process_device_io(self, d_input: int, in_format: int, d_info: int, d_pcm: int, pcm_format: int, n_frames: int,
                          stream: int = 0, d_tns_blob: int = 0, d_tns_offsets: int = 0, tns_blob_bytes: int = 0,
                          d_stereo_ops: int = 0):
        """aacfb_process_device_io with raw device addresses."""
        self._check(lib().aacfb_process_device_io(
            self._h, C.c_void_p(d_input), in_format, C.c_void_p(d_info), C.c_void_p(d_stereo_ops or None),
            C.c_void_p(d_tns_blob or None), C.c_void_p(d_tns_offsets or None), tns_blob_bytes, C.c_void_p(d_pcm),
            pcm_format, n_frames, C.c_void_p(stream or None)))

    def process_device(self, d_spectra: int, d_info: int, d_pcm: int, n_frames: int, stream: int = 0,
                       d_tns_blob: int = 0, d_tns_offsets: int = 0, tns_blob_bytes: int = 0, d_stereo_ops: int = 0):
        """Same with raw device addresses (e.g. torch.Tensor.data_ptr()), enqueued on `stream`."""
        if d_stereo_ops:
            self._check(lib().aacfb_process_device_stereo(
                self._h, C.c_void_p(d_spectra), C.c_void_p(d_info), C.c_void_p(d_stereo_ops),
                C.c_void_p(d_tns_blob or None), C.c_void_p(d_tns_offsets or None), tns_blob_bytes, C.c_void_p(d_pcm),
                n_frames, C.c_void_p(stream or None)))
            return
        self._check(lib().aacfb_process_device(self._h, C.c_void_p(d_spectra), C.c_void_p(d_info),
                                               C.c_void_p(d_tns_blob or None), C.c_void_p(d_tns_offsets or None),
                                               tns_blob_bytes, C.c_void_p(d_pcm), n_frames, C.c_void_p(stream or None)))


# ---------------------------------------------------------------------------
# Mirrors of the reference's JS classes for this path
# ---------------------------------------------------------------------------
def _info_record(info) -> np.ndarray:
    """Accept a JS-like ICSInfo (attributes or dict: windowSequence, windowShape[2], maxSFB)."""
    get = (lambda k, d=0: info.get(k, d)) if isinstance(info, dict) else (lambda k, d=0: getattr(info, k, d))
    r = np.zeros((), INFO_DTYPE)
    shape = get("windowShape", (0, 0))
    r["window_sequence"] = get("windowSequence")
    r["shape_prev"], r["shape_cur"] = shape[0], shape[1]
    r["max_sfb"] = get("maxSFB", 0)
    return r


class FilterBank:
    """`new FilterBank(smallFrames, channels)` -- filter_bank.js:24-44."""

    def __init__(self, smallFrames, channels, *, device=0, sample_index=4, ctx: Context | None = None, stream=0):
        if smallFrames:
            raise AacfbError(-2, "WHA?? No small frames allowed.")  # filter_bank.js:26
        self.ctx = ctx or Context(1, channels, sample_index, TNS_AS_SHIPPED, device)
        self.stream = stream
        self.length, self.shortLength = 1024, 128

    @property
    def overlaps(self):
        return self.ctx.get_overlap()[self.stream]

    def process(self, info, input, output, channel):
        """filterBank.process(info, input, output, channel) -- filter_bank.js:88-204.

        Reads `input` (1024 floats), fully overwrites `output`, updates overlaps[channel]."""
        rec = _info_record(info)
        x = np.ascontiguousarray(input, np.float32)
        assert x.size == 1024 and output.dtype == np.float32 and output.size == 1024 and output.flags.c_contiguous
        self.ctx._check(lib().aacfb_filterbank_process(self.ctx._h, self.stream, channel, _ptr(rec), _ptr(x),
                                                       _ptr(output)))


class TNS:
    """`new TNS(config)` -- tns.js:22-44.  decode() (the bit parse) stays in the JS host."""

    def __init__(self, config=None, *, ctx: Context | None = None, mode=None):
        cfg = config or {}
        self.sampleIndex = cfg.get("sampleIndex", 4) if isinstance(cfg, dict) else getattr(cfg, "sampleIndex", 4)
        self.nFilt = np.zeros(8, np.int32)
        self.length = np.zeros((8, 4), np.int32)
        self.direction = np.zeros((8, 4), bool)
        self.order = np.zeros((8, 4), np.int32)
        self.coef = np.zeros((8, 4, TNS_MAX_ORDER), np.float32)
        self._ctx = ctx
        self.mode = mode  # None: "tmp -> top" fixed, branch chosen by `decode`

    def block(self) -> bytes:
        """Serialise to one block of the aacfb.h TNS blob."""
        out = bytearray(int(v) for v in self.nFilt)
        for w in range(8):
            for f in range(int(self.nFilt[w])):
                order = int(self.order[w][f])
                if order > TNS_MAX_ORDER:
                    raise AacfbError(-4, f"TNS filter out of range: {order}")  # tns.js:85
                out += bytes([int(self.length[w][f]), order, int(bool(self.direction[w][f])), 0])
                out += np.asarray(self.coef[w][f][:order], np.float32).tobytes()
        return bytes(out)

    def process(self, ics, data, decode):
        """tns.process(ics, data, decode) -- tns.js:105-177, in place on `data` (Float32, 1024).

        mode=TNS_AS_SHIPPED reproduces the reference as it stands (identity, tns.js:122);
        otherwise `decode` selects the AR (:156-162) or MA (:163-174) branch."""
        mode = self.mode if self.mode is not None else (TNS_FIXED_AR if decode else TNS_FIXED_MA)
        info = ics["info"] if isinstance(ics, dict) else ics.info
        rec = _info_record(info)
        rec["max_sfb"] = ics["maxSFB"] if isinstance(ics, dict) else ics.maxSFB  # tns.js:106 reads ics.maxSFB
        ctx = self._ctx or Context(1, 1, self.sampleIndex, TNS_AS_SHIPPED, 0)
        assert data.dtype == np.float32 and data.size == 1024 and data.flags.c_contiguous
        blk = np.frombuffer(self.block(), np.uint8)
        ctx._check(lib().aacfb_tns_process(ctx._h, _ptr(rec), _ptr(blk), blk.size, _ptr(data), mode))


def pack_tns(blocks):
    """[bytes|None per channel-frame] -> (blob u8, offsets u32[n+1]) in the aacfb.h layout."""
    offs, blob = [0], bytearray()
    for b in blocks:
        if b:
            blob += b
            blob += b"\0" * (-len(blob) % 4)
        offs.append(len(blob))
    return (np.frombuffer(bytes(blob), np.uint8).copy() if blob else np.zeros(4, np.uint8)), np.asarray(offs, np.uint32)


def pack_stereo(element, sample_index: int, out=None):
    """The host's share of processMS / processIS (decoder.js:337-404): walk the bands of one channel
    pair element exactly as the reference does, but instead of touching the 2 x 1024 coefficients
    write WHAT to do per group of 4 coefficients into an aacfb_stereo_ops record (Python twin of
    js/stereo_pack.js).  `element`: record with the fields processMS / processIS read (common_window,
    mask_present, ms_used[128], and per channel window_sequence, group_count, group_length[8],
    max_sfb, band_types[120], sect_end[120], scale_factors[120]; tools/workloads.CPE_DTYPE).

    Returns (record, present): present = False when the element has nothing to apply."""
    rec = out if out is not None else np.zeros((), STEREO_DTYPE)
    op, scale = rec["op"], rec["scale"]
    op[...] = STEREO_NONE
    present = False
    if element["common_window"] and element["mask_present"]:   # decoder.js:295-296
        off_t = swb_offsets(sample_index, element["window_sequence"][0] == EIGHT_SHORT_SEQUENCE)
        btl, btr, glen = element["band_types"][0], element["band_types"][1], element["group_length"][0]
        group_off = idx = 0
        for g in range(int(element["group_count"][0])):
            for i in range(int(element["max_sfb"][0])):
                if element["ms_used"][idx] and btl[idx] < NOISE_BT and btr[idx] < NOISE_BT:
                    for w in range(int(glen[g])):
                        a = group_off + w * 128
                        op[(a + off_t[i]) // 4:(a + off_t[i + 1]) // 4] = STEREO_MS
                        present = True
                idx += 1
            group_off += int(glen[g]) * 128
    off_t = swb_offsets(sample_index, element["window_sequence"][1] == EIGHT_SHORT_SEQUENCE)
    bt, se, sf, glen = (element["band_types"][1], element["sect_end"][1], element["scale_factors"][1],
                        element["group_length"][1])
    idx = group_off = n_scales = 0
    for g in range(int(element["group_count"][1])):
        i, max_sfb = 0, int(element["max_sfb"][1])
        while i < max_sfb:
            end = int(se[idx])
            if bt[idx] in (INTENSITY_BT, INTENSITY_BT2):
                while i < end:
                    c = 1.0 if bt[idx] == INTENSITY_BT else -1.0
                    if element["mask_present"]:
                        c *= -1.0 if element["ms_used"][idx] else 1.0
                    scale[n_scales] = np.float32(c) * sf[idx]   # exact: c = +-1
                    for w in range(int(glen[g])):
                        a = group_off + w * 128
                        op[(a + off_t[i]) // 4:(a + off_t[i + 1]) // 4] = STEREO_IS + n_scales
                        present = True
                    n_scales += 1
                    i += 1
                    idx += 1
            else:
                idx += end - i
                i = end
        group_off += int(glen[g]) * 128
    return rec, present


_SF_INDEX = None


def scalefactor_index(value) -> int:
    """Index i with SCALEFACTOR_TABLE[i] == |value| (tables.js:168-176: distinct powers 2^((i-200)/4)), or
    BAND_UNDEFINED when the value is not in the table (NaN: the reference read outside it)."""
    global _SF_INDEX
    if _SF_INDEX is None:
        tab = np.empty(8192, np.float32)
        n = lib().aacfb_get_table(9, _ptr(tab), tab.size)
        _SF_INDEX = {float(v): i for i, v in enumerate(tab[:n])}
    return _SF_INDEX.get(abs(float(np.float32(value))), BAND_UNDEFINED)


def pack_qframe(ics, quant, out=None):
    """The host's share of ICStream.decodeSpectralData (ics.js:203-266) when the device dequantises:
    `ics` = what decodeBandTypes / decodeScaleFactors left behind (info.windowSequence, groupCount,
    groupLength, maxSFB; bandTypes[idx], scaleFactors[idx]) and `quant` = the 1024 Huffman-decoded
    integers in data[] order.  Python twin of js/quant_pack.js.  Returns one QFRAME_DTYPE record."""
    rec = out if out is not None else np.zeros((), QFRAME_DTYPE)
    get = (lambda o, k: o[k]) if isinstance(ics, dict) else getattr
    info = get(ics, "info")
    geti = (lambda k: info[k]) if isinstance(info, dict) else (lambda k: getattr(info, k))
    groups, max_sfb = int(geti("groupCount")), int(geti("maxSFB"))
    rec["group_len"][...] = 0
    rec["group_len"][:groups] = np.asarray(geti("groupLength"))[:groups]
    rec["band"][...] = BAND_ZERO
    bt, sf = get(ics, "bandTypes"), get(ics, "scaleFactors")
    for idx in range(groups * max_sfb):
        t = int(bt[idx])
        if t == ZERO_BT or t == INTENSITY_BT or t == INTENSITY_BT2:      # ics.js:222
            rec["band"][idx] = BAND_ZERO
        elif t == NOISE_BT:                                                # scaleFactors = -TABLE[i], ics.js:158
            rec["band"][idx] = BAND_NOISE | scalefactor_index(sf[idx])
        else:
            rec["band"][idx] = BAND_SPECTRAL | scalefactor_index(sf[idx])
    rec["reserved"][...] = 0
    rec["q"][...] = np.asarray(quant, np.int16)
    return rec


class AACDecoder:
    """The slice of AACDecoder this path replaces: `process(elements)` + the
    interleave of `readChunk` (decoder.js:204-213, 218-334), batched over K frames.

    `frames` is a list of frames, each a list of per-channel dicts
    {"info": ICSInfo-like, "data": Float32[1024], "tnsPresent": bool, "tns": TNS, "maxSFB": int}
    -- what ICStream.decode leaves behind (ics.js:56-81) after M/S and IS."""

    def __init__(self, channels, sample_index=4, *, device=0, tns_mode=TNS_AS_SHIPPED):
        self.config = {"chanConfig": channels, "sampleIndex": sample_index, "frameLength": 1024}
        self.ctx = Context(1, channels, sample_index, tns_mode, device)
        self.filter_bank = self.ctx  # decoder.js:112

    def readChunks(self, frames) -> np.ndarray:
        T, Cn = len(frames), self.config["chanConfig"]
        spectra = np.empty((1, T, Cn, 1024), np.float32)
        info = np.zeros((1, T, Cn), INFO_DTYPE)
        blocks = []
        for t, fr in enumerate(frames):
            assert len(fr) == Cn
            for c, ch in enumerate(fr):
                spectra[0, t, c] = ch["data"]
                rec = _info_record(ch["info"])
                rec["max_sfb"] = ch.get("maxSFB", 0)
                rec["tns_present"] = 1 if ch.get("tnsPresent") else 0
                info[0, t, c] = rec
                blocks.append(ch["tns"].block() if ch.get("tnsPresent") else None)
        blob, offs = pack_tns(blocks)
        pcm = self.ctx.process(spectra, info, blob, offs)
        return pcm.reshape(T * 1024 * Cn)  # K frames of readChunk output back to back
