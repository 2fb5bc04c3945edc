"""Run the reference's own JavaScript for the hot path through tools/jsmini.py.

    python tools/js_reference.py            # regenerate tests/golden/jsref_*.npz (needs /root/reference)

What is executed is the reference's unmodified source text: src/filter_bank.js (which requires
src/ics.js, src/mdct.js, src/fft.js, src/mdct_tables.js, src/tables.js, src/huffman.js, src/tns.js),
src/tns.js `TNS.prototype.process`, and lines 204-213 of src/decoder.js (the interleave, cut out of
the file by its comment markers at run time).  The only substitution ever made is the documented
one-token fix of reference defect C1 (`tmp` -> `top` at tns.js:122), applied in memory to obtain
the FIXED_AR / FIXED_MA behaviour; the as-shipped file is run too and shows the identity.

Used by tests/test_oracle_pin.py (live, when /root/reference exists) and to make the committed
golden vectors the GPU box checks against.
"""
from __future__ import annotations

import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import jsmini as J  # noqa: E402
from tools import workloads as W  # noqa: E402

REF_SRC = "/root/reference/src"


AV_STUB = """
function Base() {}
function makeExtend(Parent) {           // Aurora's class helper: Base.extend(function() { this.prototype.x = ... })
    return function(init) {
        function K() {}
        function P() {}
        P.prototype = Parent.prototype;
        K.prototype = new P();
        K.extend = makeExtend(K);
        K.register = function() {};
        init.call(K);
        return K;
    };
}
var AV = {Decoder: {extend: makeExtend(Base), register: function() {}},
          Demuxer: {extend: makeExtend(Base), register: function() {}}};
"""


class DecoderReference:
    """The reference's UNMODIFIED src/decoder.js (and everything it requires) loaded on a stand-in
    for its one missing dependency, the `av` peer package (only `AV.Decoder.extend/register` and
    `AV.Demuxer.extend/register` are needed to load the modules).  Elements are built with the
    reference's own constructors (`new ICStream(config)`, `new CPEElement(config)`), filled with what
    the bit parse would have left behind, and handed to `AACDecoder.prototype.process` -- processPair
    / processSingle, processMS, processIS, tns.process, filter_bank.process exactly as decoder.js
    :218-334 chains them -- followed by the interleave lines of readChunk (:204-213)."""

    def __init__(self, channels, sample_index=4, src_dir=REF_SRC):
        self.rt = J.Runtime(src_dir)
        self.rt.stubs["av"] = self.rt.run(AV_STUB)["AV"]
        self.Decoder = self.rt.require("./decoder")
        self.ICStream = self.rt.require("./ics")
        self.CPEElement = self.rt.require("./cpe")
        self.FilterBank = self.rt.require("./filter_bank")
        self.tables = self.rt.require("./tables")
        self.sample_index, self.channels = sample_index, channels
        self.config = J.obj(profile=2, chanConfig=channels, frameLength=1024, sampleIndex=sample_index)  # AOT_AAC_LC
        self.dec = self.Decoder.construct([])
        self.dec.put("config", self.config)
        self.dec.put("filter_bank", self.FilterBank.construct([False, float(channels)]))   # decoder.js:112
        self.dec.put("cces", J.JSArray([]))
        text = open(os.path.join(src_dir, "decoder.js")).read()
        m = re.search(r"// Interleave channels\n(.*?)\n\s*return output;", text, re.S)
        self.interleave_src = m.group(1)

    def ics(self, fi, data, cpe=None, c=0):
        """An ICStream as ICStream.decode would leave it (ics.js:56-81) for one channel-frame."""
        s = self.ICStream.construct([self.config])
        info = s.get("info")
        seq = int(fi["window_sequence"])
        short = seq == 2
        info.put("windowSequence", float(seq))
        info.get("windowShape").a[:] = [int(fi["shape_prev"]), int(fi["shape_cur"])]
        which = ("SWB_OFFSET_128", "SWB_SHORT_WINDOW_COUNT") if short else ("SWB_OFFSET_1024", "SWB_LONG_WINDOW_COUNT")
        info.put("swbOffsets", J.get_member(self.tables.get(which[0]), float(self.sample_index)))
        info.put("swbCount", J.get_member(self.tables.get(which[1]), float(self.sample_index)))
        info.put("windowCount", 8.0 if short else 1.0)
        if cpe is not None:
            info.put("groupCount", float(cpe["group_count"][c]))
            info.get("groupLength").a[:8] = cpe["group_length"][c]
            info.put("maxSFB", float(cpe["max_sfb"][c]))
            s.get("bandTypes").a[:120] = cpe["band_types"][c]
            s.get("sectEnd").a[:120] = cpe["sect_end"][c]
            s.get("scaleFactors").a[:120] = cpe["scale_factors"][c]
        else:
            info.put("groupCount", 1.0)
            info.put("maxSFB", float(fi["max_sfb"]))
        s.get("data").a[:] = data
        s.put("tnsPresent", False)
        return s

    def process_frame(self, elements):
        self.dec.get("process").call(self.dec, [J.JSArray(elements)])
        scope = self.rt.run(self.interleave_src, {"this": self.dec, "frameLength": 1024.0})
        return scope["output"].a.copy()

    def run_stereo_stream(self, spectra, info, cpe):
        """[T][2][1024] spectra before the stereo tools -> (pcm [T][1024][2], overlaps [2][1024])."""
        T = spectra.shape[0]
        pcm = np.empty((T, 1024, 2), np.float32)
        for t in range(T):
            e = self.CPEElement.construct([self.config])
            left = self.ics(info[t, 0], spectra[t, 0], cpe[t], 0)
            right = self.ics(info[t, 1], spectra[t, 1], cpe[t], 1)
            if cpe[t]["common_window"]:
                right.put("info", left.get("info"))     # cpe.js:41
            e.put("left", left); e.put("right", right)
            e.put("commonWindow", bool(cpe[t]["common_window"]))
            e.put("maskPresent", bool(cpe[t]["mask_present"]))
            e.put("ms_used", J.JSArray([bool(v) for v in cpe[t]["ms_used"]]))
            pcm[t] = self.process_frame([e]).reshape(1024, 2)
        ovl = np.stack([o.a.copy() for o in self.dec.get("filter_bank").get("overlaps").items])
        return pcm, ovl

    def run_mono_stream(self, spectra, info):
        T = spectra.shape[0]
        pcm = np.empty((T, 1024, 1), np.float32)
        for t in range(T):
            pcm[t] = self.process_frame([self.ics(info[t, 0], spectra[t, 0])]).reshape(1024, 1)
        return pcm, np.stack([o.a.copy() for o in self.dec.get("filter_bank").get("overlaps").items])


class PyBitstream:
    """What the reference needs of AV.Bitstream (read / peek / advance / align, plus offset / seek /
    available for hosts), over a bytes object, MSB first.  Reading past the end throws the
    `UnderflowError` of the AV stand-in, as Aurora's bitstream does."""

    def __init__(self, data: bytes, underflow_ctor):
        self.data, self.pos, self.underflow = data, 0, underflow_ctor

    def _bits(self, n, at):
        if at + n > 8 * len(self.data):
            raise J.JSThrow(self.underflow.construct([]))
        v = 0
        for i in range(at, at + n):
            v = (v << 1) | ((self.data[i >> 3] >> (7 - (i & 7))) & 1)
        return v

    def js(self):
        def read(this, a):
            n = int(J.to_number(a[0]))
            v = self._bits(n, self.pos)
            self.pos += n
            return float(v)

        def peek(this, a):
            return float(self._bits(int(J.to_number(a[0])), self.pos))

        def advance(this, a):
            self.pos += int(J.to_number(a[0]))
            return J.UNDEF

        def align(this, a):
            self.pos += -self.pos % 8
            return J.UNDEF

        def seek(this, a):
            self.pos = int(J.to_number(a[0]))
            return J.UNDEF

        def peek_buffer(this, a):          # AV.Stream.peekBuffer(offset, length) -> AV.Buffer {data: Uint8Array}
            off, n = int(J.to_number(a[0])), int(J.to_number(a[1]))
            at = self.pos // 8 + off
            return J.obj(data=J.JSTyped("Uint8Array", np.frombuffer(self.data[at:at + n], np.uint8).copy()))

        byte_stream = J.obj(remainingBytes=J.native(lambda this, a: float(len(self.data) - self.pos // 8)),
                            peekBuffer=J.native(peek_buffer))
        return J.obj(read=J.native(read), peek=J.native(peek), advance=J.native(advance), align=J.native(align),
                     seek=J.native(seek), offset=J.native(lambda this, a: float(self.pos)),
                     available=J.native(lambda this, a: self.pos + int(J.to_number(a[0])) <= 8 * len(self.data)),
                     stream=byte_stream)


STREAM_AV_STUB = AV_STUB + """
function UnderflowError() {}
AV.UnderflowError = UnderflowError;
AV.Stream = {fromBuffer: function(b) { return b; }};
"""


# The two tokens that keep TNS from doing anything in the reference as shipped (defects):
TNS_FIXES = {"tns.js": [(r"bottom = Math\.max\(0, tmp - length_w\[filt\]\)", "bottom = Math.max(0, top - length_w[filt])"),
                        (r"Math\.min\(this\.maxBands, ics\.maxSFB\)", "Math.min(this.maxBands, ics.info.maxSFB)")]}


def preload_patched(rt, patches):
    """Load modules of rt.src_dir with regex substitutions applied in memory (each must match once)."""
    for fname, subs in patches.items():
        path = os.path.normpath(os.path.join(rt.src_dir, fname))
        text = open(path).read()
        for pat, repl in subs:
            text, n = re.subn(pat, repl, text)
            assert n == 1, (fname, pat)
        module, exports = J.JSObject(J.OBJECT_PROTO), J.JSObject(J.OBJECT_PROTO)
        module.put("exports", exports)
        rt.modules[path] = module
        ast = J.Parser(J.tokenize(text)).program()
        names = set()
        J.hoisted_names(ast, names)
        scope = {n_: J.UNDEF for n_ in names}
        scope.update(module=module, exports=exports, this=exports,
                     require=J.native(lambda this, args: rt.require(J.to_string(args[0]))))
        J.compile_node(ast)((scope, (rt.globals, None)))


class StreamReference:
    """The reference's unmodified decoder.js end to end: setCookie on the 2-byte AudioSpecificConfig the
    ADTS demuxer makes (adts_demuxer.js:66-69), then readChunk -- ADTS header, the whole bit parse
    (ics.js, cpe.js, tns.js, huffman.js), process, interleave -- per access unit of a byte stream."""

    def __init__(self, data: bytes, profile=2, sample_index=4, channels=2, src_dir=REF_SRC, decoder_module="./decoder",
                 patches=None):
        self.rt = J.Runtime(src_dir)
        if patches:
            preload_patched(self.rt, patches)
        self.av = self.rt.run(STREAM_AV_STUB)["AV"]
        self.stream = PyBitstream(data, self.av.get("UnderflowError"))
        js_stream = self.stream.js()
        self.av.put("Bitstream", J.JSFunction(native=lambda this, args, new=False: args[0]))   # new AV.Bitstream(x) -> x
        self.rt.stubs["av"] = self.av
        self.Decoder = self.rt.require(decoder_module)
        self.dec = self.Decoder.construct([])
        self.dec.put("format", J.obj())
        cookie = bytes([(profile << 3) | ((sample_index >> 1) & 7), ((sample_index & 1) << 7) | (channels << 3), 0])
        self.Decoder.get("prototype").get("setCookie").call(self.dec, [PyBitstream(cookie, self.av.get("UnderflowError")).js()])
        self.dec.put("bitstream", js_stream)

    def read_chunk(self):
        return self.Decoder.get("prototype").get("readChunk").call(self.dec, []).a.copy()

    def decode_all(self):
        out = []
        while self.stream.pos + 56 <= 8 * len(self.stream.data):
            out.append(self.read_chunk())
        return np.concatenate(out) if out else np.zeros(0, np.float32)


class B200DecoderHarness:
    """aac.js_b200/js/decoder_b200.js -- the batching decoder a Node host registers instead of the
    reference's -- run by the interpreter ON TOP OF the unmodified reference (its base class is the
    real src/decoder.js; ICStream / CPEElement are the reference's) with a stand-in for the N-API addon.
    `compute(call)` plays the library: it receives what the JS staged for one aacfb_process[_stereo] call
    as numpy arrays (dict: spectra [T][C][1024], info, stereo_ops | None, tns_blob / tns_offsets | None)
    and returns the PCM [T][1024][C] that is written into the typed array readChunk returns."""

    def __init__(self, data: bytes, compute, channels=2, sample_index=4, profile=2, frames_per_chunk=4,
                 src_dir=REF_SRC, stereo_on_device=True, quant_on_device=False, pcm_format="f32", adts_index=True,
                 force_cpu_frames=()):
        import aacjs_b200 as A

        js_dir = os.path.join(ROOT, "aac.js_b200", "js")
        ref = J.Runtime(src_dir)
        av = ref.run(STREAM_AV_STUB)["AV"]
        av.put("Bitstream", J.JSFunction(native=lambda this, args, new=False: args[0]))
        ref.stubs["av"] = av
        self.calls, self.channels, self.index_calls, self.cpu_frames = [], channels, [], 0
        C = channels

        def finish(call, pcm_arr):
            if call["tns_blob"] is not None:
                call["tns_blob"] = call["tns_blob"][: int(call["tns_offsets"][-1])]
            self.calls.append(call)
            res = compute(call)
            pcm_arr.a[:] = np.ascontiguousarray(res, pcm_arr.a.dtype).reshape(-1)
            return J.UNDEF

        def opt(v, n=None):
            if v is None or v is J.UNDEF:
                return None
            return v.a.copy() if n is None else v.a[:n].copy()

        def run(name, a, stereo):          # process / processStereo: Float32 spectra in, Float32 PCM out
            k = 1 if stereo else 0
            n = int(J.to_number(a[6 + k]))
            call = {"spectra": a[1].a[: n * C * 1024].reshape(n, C, 1024).copy(),
                    "info": a[2].a[: n * C * 8].copy().view(W.INFO_DTYPE).reshape(n, C),
                    "stereo_ops": a[3].a[: n * 768].copy() if stereo else None,
                    "tns_blob": opt(a[3 + k]), "tns_offsets": opt(a[4 + k], n * C + 1), "entry": name, "pcm_format": 0}
            return finish(call, a[5 + k])

        def run_io(this, a):               # processIo(handle, input, inFormat, info, stereo, tnsBlob, tnsOffsets, pcm, pcmFormat, n)
            in_fmt, pcm_fmt, n = int(J.to_number(a[2])), int(J.to_number(a[8])), int(J.to_number(a[9]))
            call = {"info": a[3].a[: n * C * 8].copy().view(W.INFO_DTYPE).reshape(n, C),
                    "stereo_ops": opt(a[4], n * 768), "tns_blob": opt(a[5]), "tns_offsets": opt(a[6], n * C + 1),
                    "entry": "aacfb_process_io", "pcm_format": pcm_fmt}
            if in_fmt == 1:
                call["qframes"] = a[1].a[: n * C * 2304].copy().view(A.QFRAME_DTYPE).reshape(n, C)
            else:
                call["spectra"] = a[1].a[: n * C * 1024].reshape(n, C, 1024).copy()
            return finish(call, a[7])

        def adts(this, a):                 # adtsIndex(bytes, out u32 triples) -> complete frames in the buffer
            frames, consumed = A.adts_index(a[0].a)
            cap = a[1].a.size // 3 - 1
            k = min(len(frames), cap)
            for i in range(k):
                a[1].a[3 * i: 3 * i + 3] = [frames[i]["offset"], frames[i]["frame_length"], frames[i]["header_bytes"]]
            a[1].a[3 * k] = frames[k]["offset"] if k < len(frames) else consumed
            self.index_calls.append(k)
            return float(k)

        def get_overlap(this, a):
            a[1].a[:] = compute.overlap.reshape(-1)
            return J.UNDEF

        def set_overlap(this, a):
            compute.overlap[...] = a[1].a.reshape(compute.overlap.shape)
            return J.UNDEF

        fns = dict(create=J.native(lambda this, a: J.obj()),
                   process=J.native(lambda this, a: run("aacfb_process", a, False)),
                   processStereo=J.native(lambda this, a: run("aacfb_process_stereo", a, True)),
                   processIo=J.native(run_io), getOverlap=J.native(get_overlap), setOverlap=J.native(set_overlap))
        if adts_index:
            fns["adtsIndex"] = J.native(adts)
        addon = J.obj(**fns)
        stubs = {"av": av, "aac/src/decoder": ref.require("./decoder"), "aac/src/ics": ref.require("./ics"),
                 "aac/src/cpe": ref.require("./cpe"), "aac/src/huffman": ref.require("./huffman"),
                 "aac/src/tables": ref.require("./tables"), "./build/Release/aacfb.node": addon}
        js = J.Runtime(js_dir, stubs=stubs)
        self.Decoder = js.require("./decoder_b200")
        self.stream = PyBitstream(data, av.get("UnderflowError"))
        self.dec = self.Decoder.construct([])
        self.dec.put("format", J.obj())
        self.dec.put("framesPerChunk", float(frames_per_chunk))
        self.dec.put("stereoOnDevice", bool(stereo_on_device))
        self.dec.put("quantOnDevice", bool(quant_on_device))
        self.dec.put("pcmFormat", pcm_format)
        cookie = bytes([(profile << 3) | ((sample_index >> 1) & 7), ((sample_index & 1) << 7) | (channels << 3), 0])
        self.Decoder.get("prototype").get("setCookie").call(self.dec, [PyBitstream(cookie, av.get("UnderflowError")).js()])
        self.dec.put("bitstream", self.stream.js())
        if force_cpu_frames:
            # make stage() see a coupling element on the given access units: the frame then has to take the
            # reference's CPU path (cpuFrame), which re-parses it and finds none -- same PCM as the stock decoder
            proto_stage = self.Decoder.get("prototype").get("stage")
            count = [0]

            def stage(this, a):
                count[0] += 1
                if count[0] - 1 in force_cpu_frames:
                    this.put("cces", J.JSArray([J.obj()]))
                    self.cpu_frames += 1
                return proto_stage.call(this, a)

            self.dec.put("stage", J.native(stage))

    def decode_all(self):
        out = []
        while self.stream.pos + 56 <= 8 * len(self.stream.data):
            out.append(self.Decoder.get("prototype").get("readChunk").call(self.dec, []).a.copy())
        return np.concatenate(out) if out else np.zeros(0, np.float32)


class B200PoolHarness:
    """aac.js_b200/js/decoder_pool.js -- S pooled decoders, one library context -- run by the interpreter on
    top of the unmodified reference with a stand-in for the addon.  `compute(call)` plays the library for a
    batch: arrays shaped [S][T][C]... as aacfb_process* takes them; it returns PCM [S][T][1024][C]."""

    def __init__(self, streams, compute, channels=2, sample_index=4, profile=2, frames_per_chunk=4, src_dir=REF_SRC,
                 stereo_on_device=True, quant_on_device=True, pcm_format="f32", adts_index=True):
        import aacjs_b200 as A

        js_dir = os.path.join(ROOT, "aac.js_b200", "js")
        ref = J.Runtime(src_dir)
        av = ref.run(STREAM_AV_STUB)["AV"]
        av.put("Bitstream", J.JSFunction(native=lambda this, args, new=False: args[0]))
        ref.stubs["av"] = av
        S, C = len(streams), channels
        self.calls, self.S, self.C, self.created = [], S, C, []

        def opt(v, n=None):
            if v is None or v is J.UNDEF:
                return None
            return v.a.copy() if n is None else v.a[:n].copy()

        def finish(call, pcm_arr):
            if call["tns_blob"] is not None:
                call["tns_blob"] = call["tns_blob"][: int(call["tns_offsets"][-1])]
            self.calls.append(call)
            pcm_arr.a[:] = np.ascontiguousarray(compute(call), pcm_arr.a.dtype).reshape(-1)
            return J.UNDEF

        def run(name, a, stereo):
            k = 1 if stereo else 0
            T = int(J.to_number(a[6 + k]))
            n = S * T * C
            call = {"spectra": a[1].a[: n * 1024].reshape(S, T, C, 1024).copy(),
                    "info": a[2].a[: n * 8].copy().view(W.INFO_DTYPE).reshape(S, T, C),
                    "stereo_ops": a[3].a[: S * T * 768].copy() if stereo else None,
                    "tns_blob": opt(a[3 + k]), "tns_offsets": opt(a[4 + k], n + 1), "entry": name, "pcm_format": 0}
            return finish(call, a[5 + k])

        def run_io(this, a):
            in_fmt, pcm_fmt, T = int(J.to_number(a[2])), int(J.to_number(a[8])), int(J.to_number(a[9]))
            n = S * T * C
            call = {"info": a[3].a[: n * 8].copy().view(W.INFO_DTYPE).reshape(S, T, C),
                    "stereo_ops": opt(a[4], S * T * 768), "tns_blob": opt(a[5]), "tns_offsets": opt(a[6], n + 1),
                    "entry": "aacfb_process_io", "pcm_format": pcm_fmt}
            if in_fmt == 1:
                call["qframes"] = a[1].a[: n * 2304].copy().view(A.QFRAME_DTYPE).reshape(S, T, C)
            else:
                call["spectra"] = a[1].a[: n * 1024].reshape(S, T, C, 1024).copy()
            return finish(call, a[7])

        def adts(this, a):
            frames, consumed = A.adts_index(a[0].a)
            cap = a[1].a.size // 3 - 1
            k = min(len(frames), cap)
            for i in range(k):
                a[1].a[3 * i: 3 * i + 3] = [frames[i]["offset"], frames[i]["frame_length"], frames[i]["header_bytes"]]
            a[1].a[3 * k] = frames[k]["offset"] if k < len(frames) else consumed
            return float(k)

        def create(this, a):
            self.created.append([int(J.to_number(v)) for v in a])
            return J.obj()

        fns = dict(create=J.native(create), process=J.native(lambda this, a: run("aacfb_process", a, False)),
                   processStereo=J.native(lambda this, a: run("aacfb_process_stereo", a, True)),
                   processIo=J.native(run_io))
        if adts_index:
            fns["adtsIndex"] = J.native(adts)
        addon = J.obj(**fns)
        stubs = {"av": av, "aac/src/decoder": ref.require("./decoder"), "aac/src/ics": ref.require("./ics"),
                 "aac/src/cpe": ref.require("./cpe"), "aac/src/huffman": ref.require("./huffman"),
                 "aac/src/tables": ref.require("./tables"), "./build/Release/aacfb.node": addon}
        js = J.Runtime(js_dir, stubs=stubs)
        self.Pool = js.require("./decoder_pool")
        opts = J.obj(framesPerChunk=float(frames_per_chunk), stereoOnDevice=bool(stereo_on_device),
                     quantOnDevice=bool(quant_on_device), pcmFormat=pcm_format)
        self.pool = self.Pool.construct([float(S), opts])
        proto = self.Pool.get("prototype")
        cookie = bytes([(profile << 3) | ((sample_index >> 1) & 7), ((sample_index & 1) << 7) | (channels << 3), 0])
        self.streams = []
        for data in streams:
            dec = proto.get("createDecoder").call(self.pool, [])
            dec.put("format", J.obj())
            J.get_member(dec, "setCookie").call(dec, [PyBitstream(cookie, av.get("UnderflowError")).js()])
            st = PyBitstream(data, av.get("UnderflowError"))
            dec.put("bitstream", st.js())
            self.streams.append(st)

    def read_chunks(self):
        """One round: list of S arrays, or None (some stream has no complete frame left)."""
        res = self.Pool.get("prototype").get("readChunks").call(self.pool, [])
        if res is None or res is J.UNDEF:
            return None
        return [x.a.copy() for x in res.items]

    def decode_all(self):
        out = [[] for _ in range(self.S)]
        while True:
            r = self.read_chunks()
            if r is None:
                break
            for s in range(self.S):
                out[s].append(r[s])
        return [np.concatenate(o) if o else np.zeros(0, np.float32) for o in out]


class OracleLibrary:
    """CPU stand-in for libaacfb behind B200DecoderHarness (tests only): the op semantics of
    include/aacfb.h for the stereo records, then the oracle's TNS / filterbank / interleave, with the
    overlap state carried from call to call like an aacfb_ctx does."""

    def __init__(self, channels, sample_index=4, flags=0, n_streams=None):
        from oracle import oracle as O

        self.O, self.C, self.si, self.flags = O, channels, sample_index, flags
        self.batched = n_streams is not None          # calls carry a leading stream axis (B200PoolHarness)
        self.overlap = np.zeros((n_streams or 1, channels, 1024), np.float32)

    def __call__(self, call):
        info = call["info"] if self.batched else call["info"][None]
        if "qframes" in call:   # inverse quantisation first (ics.js:203-266), like the device
            q = call["qframes"] if self.batched else call["qframes"][None]
            sp = np.stack([self.O.dequant_batch(q[s], info[s], self.si) for s in range(q.shape[0])])
        else:
            sp = (call["spectra"] if self.batched else call["spectra"][None]).copy()
        if call["stereo_ops"] is not None:
            recs = call["stereo_ops"].view(np.dtype([("op", "u1", (256,)), ("scale", "f4", (128,))]))
            recs = recs.reshape(sp.shape[0], sp.shape[1])
            for s in range(sp.shape[0]):
                for t in range(sp.shape[1]):
                    if not info[s, t, 0]["stereo_present"]:
                        continue
                    op = np.repeat(recs[s, t]["op"], 4)
                    l, r = sp[s, t, 0].copy(), sp[s, t, 1].copy()
                    ms, it = op == 1, op >= 2
                    sp[s, t, 0][ms] = l[ms] + r[ms]
                    sp[s, t, 1][ms] = l[ms] - r[ms]
                    sp[s, t, 1][it] = l[it] * recs[s, t]["scale"][op[it] - 2]
        pcm, self.overlap = self.O.process_io(sp, 0, info, call["tns_blob"], call["tns_offsets"],
                                              self.overlap, pcm_format=call.get("pcm_format", 0), sample_index=self.si,
                                              flags=self.flags)
        return pcm if self.batched else pcm[0]


class AdtsReference:
    """ADTSDemuxer.readHeader (adts_demuxer.js:28-52) cut out of the file -- the module itself needs
    the `av` peer dependency -- and run by the interpreter on a Python-side bit reader that offers
    the two AV.Bitstream methods it calls (read(n) MSB first, advance(n))."""

    def __init__(self, src_dir=REF_SRC):
        text = open(os.path.join(src_dir, "adts_demuxer.js")).read()
        m = re.search(r"this\.readHeader = (function\(stream\) \{\n.*?\n    \});", text, re.S)
        assert m, "adts_demuxer.js readHeader not found"
        self.fn = J.Runtime(src_dir).run("var readHeader = %s;\n" % m.group(1))["readHeader"]

    def read_header(self, data: bytes):
        pos = [0]

        def read(this, args):
            n, v = int(J.to_number(args[0])), 0
            for _ in range(n):
                byte = data[pos[0] >> 3] if (pos[0] >> 3) < len(data) else 0
                v = (v << 1) | ((byte >> (7 - (pos[0] & 7))) & 1)
                pos[0] += 1
            return float(v)

        def advance(this, args):
            pos[0] += int(J.to_number(args[0]))
            return J.UNDEF

        stream = J.obj(read=J.native(read), advance=J.native(advance))
        try:
            ret = self.fn.call(J.UNDEF, [stream])
        except J.JSThrow as e:   # `throw new Error('Invalid ADTS header.')`
            return None, str(e)   # JSThrow carries the Error's message
        out = {k: int(J.to_number(ret.get(k))) for k in ("profile", "samplingIndex", "chanConfig", "frameLength", "numFrames")}
        out["bits"] = pos[0]
        return out, None


class Reference:
    def __init__(self, src_dir=REF_SRC, fix_tns=False):
        self.rt = J.Runtime(src_dir)
        self.fix_tns = fix_tns
        if fix_tns:
            # load tns.js with the one-token fix (reference defect C1) BEFORE anything requires it
            path = os.path.join(src_dir, "tns.js")
            text = open(path).read()
            fixed, n = re.subn(r"bottom = Math\.max\(0, tmp - length_w\[filt\]\)",
                               "bottom = Math.max(0, top - length_w[filt])", text)
            assert n == 1, "tns.js:122 not found"
            self.rt.modules[os.path.normpath(path)] = self._load_text(fixed)
        self.FilterBank = self.rt.require("./filter_bank")
        self.TNS = self.rt.require("./tns")
        self.tables = self.rt.require("./tables")
        dec = open(os.path.join(src_dir, "decoder.js")).read()
        m = re.search(r"// Interleave channels\n(.*?)\n\s*return output;", dec, re.S)
        assert m, "decoder.js interleave block not found"
        self.interleave_src = m.group(1)
        # processIS / processMS (decoder.js:337-404): the two prototype methods, cut out of the file
        # (decoder.js itself cannot be required: it needs the `av` peer dependency at load time)
        fns = {}
        for name in ("processIS", "processMS"):
            m = re.search(r"this\.prototype\.%s = (function\(element, left, right\) \{\n.*?\n    \});" % name, dec, re.S)
            assert m, f"decoder.js {name} not found"
            fns[name] = m.group(1)
        self.stereo_scope = self.rt.run("var ICStream = require('./ics');\nvar processIS = %s;\nvar processMS = %s;\n"
                                        % (fns["processIS"], fns["processMS"]))

    def _load_text(self, text):
        module, exports = J.JSObject(J.OBJECT_PROTO), J.JSObject(J.OBJECT_PROTO)
        module.put("exports", exports)
        ast = J.Parser(J.tokenize(text)).program()
        names = set()
        J.hoisted_names(ast, names)
        scope = {n: J.UNDEF for n in names}
        scope.update(module=module, exports=exports, this=exports,
                     require=J.native(lambda this, args: self.rt.require(J.to_string(args[0]))))
        J.compile_node(ast)((scope, (self.rt.globals, None)))
        return module

    def new_filterbank(self, channels):
        return self.FilterBank.construct([False, float(channels)])

    def filterbank_process(self, fb, seq, shape_prev, shape_cur, x, channel):
        info = J.obj(windowSequence=int(seq), windowShape=J.int32array([shape_prev, shape_cur]))
        out = J.float32array(np.zeros(1024))
        self.FilterBank.get("prototype").get("process").call(fb, [info, J.float32array(x), out, float(channel)])
        return out.a.copy()

    def overlaps(self, fb):
        return np.stack([o.a.copy() for o in fb.get("overlaps").items])

    def make_tns(self, sample_index, block: bytes):
        """A reference TNS object (tns.js:22-44) filled from one aacfb.h TNS block."""
        tns = self.TNS.construct([J.obj(sampleIndex=sample_index)])
        n_filt, pos = block[:8], 8
        for w in range(8):
            J.set_member(tns.get("nFilt"), float(w), float(n_filt[w]))
            for f in range(n_filt[w]):
                length, order, direction = block[pos], block[pos + 1], block[pos + 2]
                coef = np.frombuffer(block[pos + 4:pos + 4 + 4 * order], np.float32)
                pos += 4 + 4 * order
                J.set_member(tns.get("length").items[w], float(f), float(length))
                J.set_member(tns.get("order").items[w], float(f), float(order))
                J.set_member(tns.get("direction").items[w], float(f), bool(direction))
                dst = tns.get("coef").items[w].items[f]
                for i, c in enumerate(coef):
                    J.set_member(dst, float(i), float(c))
        return tns

    def tns_process(self, sample_index, seq, max_sfb, block, data, decode):
        tns = self.make_tns(sample_index, block)
        short = seq == 2
        info = J.obj(windowSequence=int(seq), windowCount=8 if short else 1,
                     swbCount=J.get_member(self.tables.get("SWB_SHORT_WINDOW_COUNT" if short else "SWB_LONG_WINDOW_COUNT"),
                                           float(sample_index)),
                     swbOffsets=J.get_member(self.tables.get("SWB_OFFSET_128" if short else "SWB_OFFSET_1024"),
                                             float(sample_index)))
        ics = J.obj(maxSFB=int(max_sfb), info=info)
        d = J.float32array(data)
        self.TNS.get("prototype").get("process").call(tns, [ics, d, bool(decode)])
        return d.a.copy()

    def stereo(self, cpe, sample_index, left, right):
        """processPair's stereo part, decoder.js:294-301: processMS if commonWindow && maskPresent, then
        processIS, on copies of left / right; `cpe` is a tools/workloads.CPE_DTYPE record."""
        def ics(c):
            short = int(cpe["window_sequence"][c]) == 2
            info = J.obj(windowSequence=int(cpe["window_sequence"][c]), groupCount=int(cpe["group_count"][c]),
                         maxSFB=int(cpe["max_sfb"][c]), groupLength=J.int32array(cpe["group_length"][c]),
                         swbOffsets=J.get_member(self.tables.get("SWB_OFFSET_128" if short else "SWB_OFFSET_1024"),
                                                 float(sample_index)))
            return J.obj(info=info, bandTypes=J.int32array(cpe["band_types"][c]), sectEnd=J.int32array(cpe["sect_end"][c]),
                         scaleFactors=J.float32array(cpe["scale_factors"][c]))
        element = J.obj(commonWindow=bool(cpe["common_window"]), maskPresent=bool(cpe["mask_present"]),
                        ms_used=J.JSArray([bool(v) for v in cpe["ms_used"]]), left=ics(0), right=ics(1))
        if cpe["common_window"]:
            element.props["right"].props["info"] = element.props["left"].props["info"]  # cpe.js:41
        l, r = J.float32array(left), J.float32array(right)
        if cpe["common_window"] and cpe["mask_present"]:
            self.stereo_scope["processMS"].call(J.UNDEF, [element, l, r])
        self.stereo_scope["processIS"].call(J.UNDEF, [element, l, r])
        return l.a.copy(), r.a.copy()

    def interleave(self, chans):
        """decoder.js:204-213 on this.data = chans (list of Float32Array rows)."""
        data = J.JSArray([J.float32array(c) for c in chans])
        scope = self.rt.run(self.interleave_src, {"this": J.obj(data=data), "frameLength": 1024.0})
        return scope["output"].a.copy()

    def process(self, spectra, info, tns_blob, tns_offsets, sample_index, mode):
        """The whole path the way decoder.js drives it (tns.process, filter_bank.process per channel,
        interleave), for a [S][T][C][1024] batch.  mode: 0 as shipped (decode=false, decoder.js:264),
        1 = decode true, 2 = decode false (both meaningful only with fix_tns)."""
        S, T, C, _ = spectra.shape
        pcm = np.empty((S, T, 1024, C), np.float32)
        ovl = np.empty((S, C, 1024), np.float32)
        for s in range(S):
            fb = self.new_filterbank(C)
            for t in range(T):
                chans = []
                for c in range(C):
                    fi = info[s, t, c]
                    data = spectra[s, t, c]
                    cf = (s * T + t) * C + c
                    if fi["tns_present"] and tns_blob is not None and tns_offsets[cf + 1] > tns_offsets[cf]:
                        blk = bytes(tns_blob[tns_offsets[cf]:tns_offsets[cf + 1]])
                        data = self.tns_process(sample_index, fi["window_sequence"], fi["max_sfb"], blk, data,
                                                decode=(mode == 1))
                    chans.append(self.filterbank_process(fb, fi["window_sequence"], fi["shape_prev"], fi["shape_cur"],
                                                         data, c))
                pcm[s, t] = self.interleave(chans).reshape(1024, C)
            ovl[s] = self.overlaps(fb)
        return pcm, ovl


CASES = {
    # name: (config, S, T, C, seed, shape_prev_mode, fix_tns, mode)
    "config1_mono_long": (1, 1, 1, 1, 0, "as_shipped", False, 0),
    "config2_long": (2, 1, 4, 2, 11, "as_shipped", False, 0),
    "config3_short": (3, 1, 3, 2, 12, "carried", False, 0),
    "config4_tns_as_shipped": (4, 1, 3, 2, 13, "as_shipped", False, 0),
    "config4_tns_fixed_ar": (4, 1, 3, 2, 13, "as_shipped", True, 1),
    "config4_tns_fixed_ma": (4, 1, 3, 2, 13, "as_shipped", True, 2),
    "config5_mixed": (5, 1, 18, 2, 14, "carried", False, 0),
}


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    refs = {}
    for name, (cfg, S, T, C, seed, spm, fix, mode) in CASES.items():
        ref = refs.setdefault(fix, Reference(fix_tns=fix))
        w = W.make(cfg, S, T, C, seed, spm)
        if cfg == 4 and not fix:
            pass  # the as-shipped reference ignores TNS (tns.js:122): same inputs, TNS side info present
        pcm, ovl = ref.process(w["spectra"], w["info"], w["tns_blob"], w["tns_offsets"], w["sample_index"], mode)
        np.savez_compressed(os.path.join(out_dir, f"jsref_{name}.npz"), pcm=pcm, overlap=ovl,
                            meta=np.array([cfg, S, T, C, seed, int(spm == "carried"), int(fix), mode]))
        print(name, pcm.shape, float(np.abs(pcm).max()))


STEREO_SEQS = [(0, 0), (2, 2), (0, 2), (2, 0), (1, 1), (3, 3), (1, 2), (3, 0)]


def stereo_elements(n=48, seed=77, sample_index=4):
    """The seeded channel pair elements of tests/golden/stereo/jsref_stereo.npz."""
    rng = np.random.default_rng(seed)
    cpe = np.zeros(n, W.CPE_DTYPE)
    for k in range(n):
        a, b = STEREO_SEQS[k % len(STEREO_SEQS)]
        cpe[k] = W.random_cpe(rng, a, b, sample_index, common_window=None if k % 3 else True, p_intensity=0.4)
    left = (rng.standard_normal((n, 1024)) * 1e4).astype(np.float32)
    right = (rng.standard_normal((n, 1024)) * 1e4).astype(np.float32)
    return cpe, left, right


def main_stereo():
    """processMS / processIS (decoder.js:337-404) run by the interpreter: (i) per element, (ii) a whole
    stereo stream: stereo tools -> filter_bank.process -> interleave."""
    out_dir = os.path.join(ROOT, "tests", "golden", "stereo")
    os.makedirs(out_dir, exist_ok=True)
    ref = Reference()
    cpe, left, right = stereo_elements()
    ol, orr = np.empty_like(left), np.empty_like(right)
    for k in range(len(cpe)):
        ol[k], orr[k] = ref.stereo(cpe[k], 4, left[k], right[k])
    S, T, seed = 2, 12, 78
    case = W.random_stereo_case(S, T, np.random.default_rng(seed))
    sp = case["spectra"].copy()
    for s in range(S):
        for t in range(T):
            sp[s, t, 0], sp[s, t, 1] = ref.stereo(case["cpe"][s, t], 4, sp[s, t, 0], sp[s, t, 1])
    pcm, ovl = ref.process(sp, case["info"], None, None, 4, 0)
    np.savez_compressed(os.path.join(out_dir, "jsref_stereo.npz"), out_left=ol, out_right=orr, pcm=pcm, overlap=ovl,
                        meta=np.array([len(cpe), 77, S, T, seed]))
    print("stereo", ol.shape, pcm.shape, int((ol != left).any(axis=1).sum()), int((orr != right).any(axis=1).sum()))


def decoder_case(seed=81, T=10):
    """The seeded stereo stream of tests/golden/stereo/jsref_decoder_process.npz: common-window pairs keep
    one window sequence/shape for both channels (cpe.js:40-42: right.info IS left.info)."""
    case = W.random_stereo_case(1, T, np.random.default_rng(seed), sigma=2e4)
    for t in range(T):
        if case["cpe"][0, t]["common_window"]:
            case["info"][0, t, 1] = case["info"][0, t, 0]
    return case


def main_decoder():
    """AACDecoder.process + interleave of the unmodified decoder.js on a stereo and a mono stream."""
    out_dir = os.path.join(ROOT, "tests", "golden", "stereo")
    os.makedirs(out_dir, exist_ok=True)
    case = decoder_case()
    pcm, ovl = DecoderReference(2).run_stereo_stream(case["spectra"][0], case["info"][0], case["cpe"][0])
    mono = W.make(5, 1, 18, 1, seed=82, shape_prev_mode="carried")
    mpcm, movl = DecoderReference(1).run_mono_stream(mono["spectra"][0], mono["info"][0])
    np.savez_compressed(os.path.join(out_dir, "jsref_decoder_process.npz"), pcm=pcm, overlap=ovl, mono_pcm=mpcm,
                        mono_overlap=movl, meta=np.array([81, 10, 82, 18]))
    print("decoder.process", pcm.shape, float(np.abs(pcm).max()), mpcm.shape)


def stream_cases():
    """(name, channels, n_frames, seed, frames_per_chunk) of tests/golden/stream/jsref_stream_*.npz"""
    return [("stereo", 2, 14, 301, 4), ("mono", 1, 9, 302, 3), ("surround", 6, 5, 303, 2),
            ("stereo_tnsfixed", 2, 10, 304, 5)]


def main_stream():
    """A synthetic ADTS stream through (a) the reference's unmodified readChunk and (b) decoder_b200.js on
    top of it with the oracle as the library; stores the stream, the reference PCM and, per addon call,
    the arrays the JS staged -- the GPU box replays those through the real library."""
    from tools import aac_bitstream as B

    out_dir = os.path.join(ROOT, "tests", "golden", "stream")
    os.makedirs(out_dir, exist_ok=True)
    for name, C, n, seed, K in stream_cases():
        data = B.write_adts_stream(B.random_frames(np.random.default_rng(seed), n, channels=C), B.codebooks(), channels=C)
        fixed = name.endswith("_tnsfixed")   # the reference with its two TNS tokens fixed: decode=false -> MA branch
        ref = StreamReference(data, channels=C, patches=TNS_FIXES if fixed else None).decode_all()
        h = B200DecoderHarness(data, OracleLibrary(C, flags=2 if fixed else 0), channels=C, frames_per_chunk=K)
        got = h.decode_all()
        assert np.array_equal(ref.view(np.uint32), got.view(np.uint32))
        arrays = {"adts": np.frombuffer(data, np.uint8), "pcm": ref, "meta": np.array([C, n, seed, K, len(h.calls)])}
        for i, c in enumerate(h.calls):
            arrays[f"c{i}_spectra"] = c["spectra"]
            arrays[f"c{i}_info"] = c["info"].view(np.uint8).reshape(-1)
            arrays[f"c{i}_stereo"] = c["stereo_ops"] if c["stereo_ops"] is not None else np.zeros(0, np.uint8)
            arrays[f"c{i}_tns_blob"] = c["tns_blob"] if c["tns_blob"] is not None else np.zeros(0, np.uint8)
            arrays[f"c{i}_tns_offsets"] = c["tns_offsets"] if c["tns_offsets"] is not None else np.zeros(0, np.uint32)
        np.savez_compressed(os.path.join(out_dir, f"jsref_stream_{name}.npz"), **arrays)
        print(name, len(data), "bytes,", n, "frames,", len(h.calls), "calls, peak", float(np.abs(ref).max()))


def main_stream_quant():
    """tests/golden/stream/jsref_streamq_*.npz: the same experiment with quantOnDevice -- the batching decoder
    swaps ICStream.decodeSpectralData for quant_pack.js's walk while the reference parses, stages
    aacfb_qframe records (Huffman integers + band codes) and, for the s16 case, asks for int16 PCM.  With
    the oracle as the library (its decode_spectral_data restates ics.js:203-266) the output equals the
    stock decoder's bit for bit; the staged records are replayed through the CUDA library on the GPU box."""
    from tools import aac_bitstream as B
    from oracle import oracle as O

    out_dir = os.path.join(ROOT, "tests", "golden", "stream")
    for name, C, n, seed, K, fmt in (("stereo", 2, 12, 401, 5, "f32"), ("mono_s16", 1, 8, 402, 3, "s16")):
        data = B.write_adts_stream(B.random_frames(np.random.default_rng(seed), n, channels=C), B.codebooks(), channels=C)
        ref = StreamReference(data, channels=C).decode_all()
        h = B200DecoderHarness(data, OracleLibrary(C), channels=C, frames_per_chunk=K, quant_on_device=True, pcm_format=fmt)
        got = h.decode_all()
        want = O.pcm_s16(ref * 32768) if fmt == "s16" else ref
        assert np.array_equal(want.view(np.uint16 if fmt == "s16" else np.uint32), got.view(np.uint16 if fmt == "s16" else np.uint32))
        arrays = {"adts": np.frombuffer(data, np.uint8), "pcm": ref, "meta": np.array([C, n, seed, K, len(h.calls), int(fmt == "s16")])}
        for i, c in enumerate(h.calls):
            arrays[f"c{i}_qframes"] = c["qframes"].view(np.uint8).reshape(-1)
            arrays[f"c{i}_info"] = c["info"].view(np.uint8).reshape(-1)
            arrays[f"c{i}_stereo"] = c["stereo_ops"] if c["stereo_ops"] is not None else np.zeros(0, np.uint8)
        np.savez_compressed(os.path.join(out_dir, f"jsref_streamq_{name}.npz"), **arrays)
        print(name, len(data), "bytes,", n, "frames,", len(h.calls), "calls, peak", float(np.abs(ref).max()))


def main_adts():
    """ADTSDemuxer.readHeader run by the interpreter on every frame of a seeded synthetic ADTS stream."""
    out_dir = os.path.join(ROOT, "tests", "golden", "adts")
    os.makedirs(out_dir, exist_ok=True)
    data, _ = W.adts_stream(np.random.default_rng(90), 64)
    ref, p, rows = AdtsReference(), 0, []
    while p + 7 <= len(data):
        h, err = ref.read_header(data[p:])
        assert err is None
        rows.append([p, h["frameLength"], h["bits"], h["profile"], h["samplingIndex"], h["chanConfig"], h["numFrames"]])
        p += h["frameLength"]
    _, msg = ref.read_header(b"\xff\xe1" + bytes(8))
    np.savez_compressed(os.path.join(out_dir, "jsref_adts.npz"), headers=np.asarray(rows, np.int64),
                        meta=np.array([90, 64]), error=np.frombuffer(msg.encode(), np.uint8))
    print("adts", len(rows), "frames;", msg)


class DequantReference:
    """ICStream.prototype.decodeSpectralData (src/ics.js:203-266), unmodified, run by the interpreter: the
    ICStream is built with the reference's own constructor and filled with what decodeBandTypes /
    decodeScaleFactors would have left behind; the one stand-in is the entropy decoder it calls --
    `Huffman.decodeSpectralData(stream, hcb, buf, 0)` is replaced by a function that hands out the integers
    of an aacfb_qframe record in the order the loop asks for them.  Inverse quantisation, scalefactor
    multiplication, perceptual noise substitution (with its generator, ics.js:234) and the zero bands are
    the reference's own statements, its lookup tables the ones src/tables.js builds."""

    FIRST_PAIR_BT, NOISE_BT, ZERO_BT = 5, 13, 0

    def __init__(self, sample_index=4, src_dir=REF_SRC):
        self.rt = J.Runtime(src_dir)
        self.ICStream = self.rt.require("./ics")
        self.tables = self.rt.require("./tables")
        self.huffman = self.rt.require("./huffman")
        self.sample_index = sample_index
        self.config = J.obj(profile=2, chanConfig=2, frameLength=1024, sampleIndex=sample_index)
        self.sf_table = self.tables.get("SCALEFACTOR_TABLE").a
        self.iq_table = self.tables.get("IQ_TABLE").a

    def run(self, qframe, fi, rng=None):
        """One aacfb_qframe record + its aacfb_frame_info -> ics.data (1024 f32) as the reference computes it."""
        s = self.ICStream.construct([self.config])
        info = s.get("info")
        short = int(fi["window_sequence"]) == 2
        which = "SWB_OFFSET_128" if short else "SWB_OFFSET_1024"
        offsets_js = J.get_member(self.tables.get(which), float(self.sample_index))
        offsets = np.asarray(offsets_js.a)
        info.put("windowSequence", float(fi["window_sequence"]))
        info.put("swbOffsets", offsets_js)
        groups = int(np.count_nonzero(np.cumsum(qframe["group_len"] == 0) == 0))
        max_sfb = int(fi["max_sfb"])
        info.put("groupCount", float(groups))
        info.get("groupLength").a[:8] = qframe["group_len"]
        info.put("maxSFB", float(max_sfb))
        feed = []
        group_off = 0
        for g in range(groups):
            glen = int(qframe["group_len"][g])
            for sfb in range(max_sfb):
                idx = g * max_sfb + sfb
                code = int(qframe["band"][idx])
                kind, ti = code & 0xc000, code & 0x1ff
                tab = float(self.sf_table[ti]) if ti < 428 else float("nan")
                if kind == 0x0000:
                    s.get("bandTypes").a[idx] = self.ZERO_BT
                    s.get("scaleFactors").a[idx] = 0.0
                elif kind == 0x8000:
                    s.get("bandTypes").a[idx] = self.NOISE_BT
                    s.get("scaleFactors").a[idx] = -tab                       # ics.js:158
                else:
                    # any spectral codebook: quads below FIRST_PAIR_BT, pairs from it on (ics.js:245)
                    hcb = 1 + (idx % 11) if rng is None else int(rng.integers(1, 12))
                    s.get("bandTypes").a[idx] = hcb
                    s.get("scaleFactors").a[idx] = tab                        # ics.js:171
                    lo, hi = int(offsets[sfb]), int(offsets[sfb + 1])
                    for w in range(glen):
                        base = group_off + 128 * w
                        feed.extend(int(v) for v in qframe["q"][base + lo: base + hi])
            group_off += glen << 7
        pos = [0]

        def decode_spectral_data(this, a):   # Huffman.decodeSpectralData(stream, hcb, buf, off): huffman.js:1426-1455
            hcb, buf = int(J.to_number(a[1])), a[2]
            num = 2 if hcb >= self.FIRST_PAIR_BT else 4
            buf.a[:num] = feed[pos[0]: pos[0] + num]
            pos[0] += num
            return J.UNDEF

        self.huffman.put("decodeSpectralData", J.native(decode_spectral_data))
        s.get("decodeSpectralData").call(s, [J.obj()])
        assert pos[0] == len(feed), (pos[0], len(feed))
        return s.get("data").a.copy()


def dequant_cases(seed=91, n=40):
    rng = np.random.default_rng(seed)
    case = W.random_q_case(n, 1, 1, rng, p_noise=0.12)
    return case["qframes"].reshape(n), case["info"].reshape(n)


def main_dequant():
    """tests/golden/dequant/jsref_dequant.npz: 40 records through the reference's decodeSpectralData."""
    ref = DequantReference()
    q, info = dequant_cases()
    data = np.stack([ref.run(q[i], info[i]) for i in range(len(q))])
    out = os.path.join(ROOT, "tests", "golden", "dequant")
    os.makedirs(out, exist_ok=True)
    np.savez_compressed(os.path.join(out, "jsref_dequant.npz"), qframes=q, info=info, data=data,
                        iq_table=ref.iq_table.copy(), sf_table=ref.sf_table.copy(), sample_index=4)
    print("wrote jsref_dequant.npz", data.shape, "NaN rows:", int(np.isnan(data).any(axis=1).sum()))


if __name__ == "__main__":
    {"stereo": main_stereo, "adts": main_adts, "decoder": main_decoder, "stream": main_stream, "streamq": main_stream_quant, "dequant": main_dequant}.get((sys.argv[1:] or [""])[0], main)()
