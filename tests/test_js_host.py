"""The JavaScript host files under aac.js_b200/js/ cannot run here as a whole (no Node, and the
reference's decoder.js needs the `av` peer dependency), but their pure packing logic can: it is
executed by tools/jsmini.py and compared with the Python twins the GPU tests exercise.
(js/stereo_pack.js: tests/test_stereo.py.)"""
import os

import numpy as np
import pytest

import aacjs_b200 as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_js_tns_packer_writes_the_same_block_as_the_python_twin():
    """js/tns_pack.js (what the Node host runs) executed by the interpreter: blockSize / writeBlock
    produce byte for byte the block that TNS.block() hands to the library."""
    import struct

    from tools import jsmini as J

    js_dir = os.path.join(ROOT, "aac.js_b200", "js")
    rt = J.Runtime(js_dir, stubs={"./build/Release/aacfb.node": J.obj()})
    pack = rt.require("./tns_pack")
    rng = np.random.default_rng(21)
    for trial in range(6):
        t = A.TNS({"sampleIndex": 4}, mode=A.TNS_FIXED_AR)
        n_win = 8 if trial % 2 else 1
        for w in range(n_win):
            t.nFilt[w] = rng.integers(0, 4)
            for f in range(int(t.nFilt[w])):
                t.length[w][f] = rng.integers(1, 40)
                t.order[w][f] = rng.integers(0, 21)
                t.direction[w][f] = bool(rng.integers(0, 2))
                t.coef[w][f][: t.order[w][f]] = rng.standard_normal(int(t.order[w][f])).astype(np.float32)
        want = t.block()
        jt = J.obj(nFilt=J.int32array(t.nFilt),
                   length=J.JSArray([J.int32array(r) for r in t.length]),
                   order=J.JSArray([J.int32array(r) for r in t.order]),
                   direction=J.JSArray([J.JSArray([bool(v) for v in r]) for r in t.direction]),
                   coef=J.JSArray([J.JSArray([J.float32array(c) for c in r]) for r in t.coef]))
        assert int(J.to_number(pack.get("blockSize").call(J.UNDEF, [jt]))) == len(want)
        raw = np.full(len(want) + 8, 0xAA, np.uint8)
        view = J.obj(setFloat32=J.native(lambda this, a: raw.__setitem__(
            slice(int(J.to_number(a[0])), int(J.to_number(a[0])) + 4),
            np.frombuffer(struct.pack("<f" if J.to_bool(a[2]) else ">f", J.to_number(a[1])), np.uint8)) or J.UNDEF))
        end = pack.get("writeBlock").call(J.UNDEF, [jt, J.JSTyped("Uint8Array", raw), view, 4.0])
        assert int(J.to_number(end)) == 4 + len(want)
        assert bytes(raw[4:4 + len(want)]) == want and (raw[:4] == 0xAA).all() and (raw[4 + len(want):] == 0xAA).all()


def test_js_filter_bank_dropin_drives_the_addon_like_the_c_abi_expects():
    """js/filter_bank.js with a recording stand-in for the N-API addon: `new FilterBank(smallFrames,
    channels)` -> create(device 0, 1 stream, channels, sampleIndex 4, smallFrames, flags 0);
    `process(info, input, output, channel)` -> filterbankProcess(handle, stream 0, channel,
    aacfb_frame_info bytes, input, output) with the same typed arrays (no copies)."""
    from tools import jsmini as J

    calls = []

    def create(this, a):
        calls.append(("create", [int(J.to_number(v)) for v in a]))
        return J.obj(tag="handle")

    def fb_process(this, a):
        calls.append(("process", a))
        a[5].a[:] = a[4].a[::-1]          # the "library" writes into the caller's output array
        return J.UNDEF

    addon = J.obj(create=J.native(create), filterbankProcess=J.native(fb_process))
    rt = J.Runtime(os.path.join(ROOT, "aac.js_b200", "js"), stubs={"./build/Release/aacfb.node": addon})
    FilterBank = rt.require("./filter_bank")
    fb = FilterBank.construct([False, 2.0])
    assert calls[0] == ("create", [0, 1, 2, 4, 0, 0])
    assert J.to_number(fb.get("length")) == 1024 and J.to_number(fb.get("shortLength")) == 128
    x, out = J.float32array(np.arange(1024)), J.float32array(np.zeros(1024))
    info = J.obj(windowSequence=3, windowShape=J.int32array([1, 0]), maxSFB=49)
    FilterBank.get("prototype").get("process").call(fb, [info, x, out, 1.0])
    name, a = calls[1]
    assert name == "process" and a[0] is fb.get("handle") and J.to_number(a[1]) == 0 and J.to_number(a[2]) == 1
    assert list(a[3].a) == [3, 1, 0, 49, 0, 0, 0, 0]          # aacfb_frame_info, include/aacfb.h
    assert a[4] is x and a[5] is out and out.a[0] == 1023.0   # borrowed arrays, output filled in place
    FilterBank.construct([True, 2.0])                          # smallFrames is forwarded; the library throws
    assert calls[2] == ("create", [0, 1, 2, 4, 1, 0])


JS_STUBS = """
function AACDecoder() {}
AACDecoder.prototype.setCookie = function(cfg) { this.config = cfg; };
AACDecoder.prototype.processMS = function(e, l, r) { this.hostStereoCalls.push('ms'); };
AACDecoder.prototype.processIS = function(e, l, r) { this.hostStereoCalls.push('is'); };
AACDecoder.extend = function(init) {              // Aurora's class helper, as decoder.js uses it
    function K() { this.hostStereoCalls = []; }
    function P() {}
    P.prototype = AACDecoder.prototype;
    K.prototype = new P();
    init.call(K);
    return K;
};
function ICStream() {}
ICStream.NOISE_BT = 13; ICStream.INTENSITY_BT2 = 14; ICStream.INTENSITY_BT = 15;
function CPEElement() {}
function UnderflowError() {}
var AV = {Decoder: {register: function() {}}, UnderflowError: UnderflowError};
"""


def _js_tns(J, t):
    return J.obj(nFilt=J.int32array(t.nFilt), length=J.JSArray([J.int32array(r) for r in t.length]),
                 order=J.JSArray([J.int32array(r) for r in t.order]),
                 direction=J.JSArray([J.JSArray([bool(v) for v in r]) for r in t.direction]),
                 coef=J.JSArray([J.JSArray([J.float32array(c) for c in r]) for r in t.coef]))


def _random_tns(rng, short):
    t = A.TNS({"sampleIndex": 4}, mode=A.TNS_FIXED_AR)
    for w in range(8 if short else 1):
        t.nFilt[w] = rng.integers(0, 3)
        for f in range(int(t.nFilt[w])):
            t.length[w][f], t.order[w][f] = rng.integers(1, 12), rng.integers(0, 8)
            t.direction[w][f] = bool(rng.integers(0, 2))
            t.coef[w][f][: t.order[w][f]] = rng.standard_normal(int(t.order[w][f])).astype(np.float32)
    return t


@pytest.mark.parametrize("stereo_on_device", [True, False])
def test_js_batching_decoder_stages_frames_as_the_library_expects(stereo_on_device):
    """js/decoder_b200.js run by the interpreter on stand-ins for `av`, the reference's AACDecoder /
    ICStream / CPEElement and the addon: readChunk parses frames until the bitstream underflows,
    rewinds to the last frame boundary, and hands ONE call to the addon whose typed arrays are,
    byte for byte, what the Python host mirror builds for the same frames (spectra before the stereo
    tools, aacfb_frame_info records with stereo_present, aacfb_stereo_ops records, TNS blob/offsets)."""
    from tools import jsmini as J
    from tools import workloads as W

    js_dir = os.path.join(ROOT, "aac.js_b200", "js")
    sc = J.Runtime(js_dir).run(JS_STUBS)
    calls = []
    addon = J.obj(create=J.native(lambda this, a: (calls.append(("create", [J.to_number(v) for v in a])), J.obj())[1]),
                  process=J.native(lambda this, a: (calls.append(("process", a)), J.UNDEF)[1]),
                  processStereo=J.native(lambda this, a: (calls.append(("processStereo", a)), J.UNDEF)[1]))
    rt = J.Runtime(js_dir, stubs={"av": sc["AV"], "aac/src/decoder": sc["AACDecoder"], "aac/src/ics": sc["ICStream"],
                                  "aac/src/cpe": sc["CPEElement"], "aac/src/huffman": J.obj(), "aac/src/tables": J.obj(),
                                  "./build/Release/aacfb.node": addon})
    B200 = rt.require("./decoder_b200")
    proto = B200.get("prototype")
    dec = B200.construct([])
    dec.put("stereoOnDevice", stereo_on_device)
    dec.put("quantOnDevice", False)         # Float32 staging (quantised staging: test_js_quant_packer_..., test_stream_e2e)
    dec.put("framesPerChunk", 8.0)
    proto.get("setCookie").call(dec, [J.obj(chanConfig=2, sampleIndex=4)])
    assert calls[0] == ("create", [0, 1, 2, 4, 0, 0])

    rng = np.random.default_rng(31)
    T = 5                                   # fewer than framesPerChunk: the 6th parse underflows
    case = W.random_stereo_case(1, T, rng, sigma=1e4)
    tns = [[_random_tns(rng, case["info"]["window_sequence"][0, t, c] == 2) if rng.random() < 0.6 else None
            for c in range(2)] for t in range(T)]
    frames = []
    for t in range(T):
        e = case["cpe"][0, t]
        cpe = sc["CPEElement"].construct([])
        for name, c in (("left", 0), ("right", 1)):
            ics = sc["ICStream"].construct([])
            fi = case["info"][0, t, c]
            short = int(e["window_sequence"][c]) == 2
            ics.put("info", J.obj(windowSequence=int(fi["window_sequence"]), maxSFB=int(e["max_sfb"][c]),
                                  windowShape=J.int32array([fi["shape_prev"], fi["shape_cur"]]),
                                  groupCount=int(e["group_count"][c]), groupLength=J.int32array(e["group_length"][c]),
                                  swbOffsets=J.JSTyped("Uint16Array", A.swb_offsets(4, short))))
            ics.put("data", J.float32array(case["spectra"][0, t, c]))
            ics.put("bandTypes", J.int32array(e["band_types"][c]))
            ics.put("sectEnd", J.int32array(e["sect_end"][c]))
            ics.put("scaleFactors", J.float32array(e["scale_factors"][c]))
            ics.put("tnsPresent", tns[t][c] is not None)
            if tns[t][c] is not None:
                ics.put("tns", _js_tns(J, tns[t][c]))
            cpe.put(name, ics)
        cpe.put("commonWindow", bool(e["common_window"]))
        cpe.put("maskPresent", bool(e["mask_present"]))
        cpe.put("ms_used", J.JSArray([bool(v) for v in e["ms_used"]]))
        frames.append(cpe)

    pos, seeks = [0], []

    def parse(this, a):
        if pos[0] >= T:
            raise J.JSThrow(sc["UnderflowError"].construct([]))
        pos[0] += 1
        return J.JSArray([frames[pos[0] - 1]])

    dec.put("parseElements", J.native(parse))
    dec.put("bitstream", J.obj(offset=J.native(lambda this, a: float(100 * pos[0])),
                               seek=J.native(lambda this, a: (seeks.append(J.to_number(a[0])), J.UNDEF)[1])))
    pcm = proto.get("readChunk").call(dec, [])
    assert seeks == [100.0 * T]             # rewound to the boundary of the frame that underflowed
    assert pcm.a.size == T * 1024 * 2

    # what the Python mirror stages for the same frames
    want_ops = np.zeros(T, A.STEREO_DTYPE)
    want_info = np.zeros((T, 2), A.INFO_DTYPE)
    blocks = []
    for t in range(T):
        _, present = A.pack_stereo(case["cpe"][0, t], 4, out=want_ops[t])
        for c in range(2):
            fi = case["info"][0, t, c]
            want_info[t, c]["window_sequence"], want_info[t, c]["shape_prev"] = fi["window_sequence"], fi["shape_prev"]
            want_info[t, c]["shape_cur"], want_info[t, c]["max_sfb"] = fi["shape_cur"], case["cpe"][0, t]["max_sfb"][c]
            want_info[t, c]["tns_present"] = tns[t][c] is not None
            blocks.append(tns[t][c].block() if tns[t][c] is not None else None)
        want_info[t, 0]["stereo_present"] = int(present and stereo_on_device)
    blob, offs = A.pack_tns(blocks)

    name, a = calls[-1]
    any_stereo = bool(want_info["stereo_present"].any())
    assert name == ("processStereo" if any_stereo else "process")
    if any_stereo:
        spectra, info, ops, tb, to, out, n = a[1], a[2], a[3], a[4], a[5], a[6], a[7]
        assert np.array_equal(ops.a[:T * 768].view(A.STEREO_DTYPE)["op"], want_ops["op"])
        used = want_ops["op"].max(axis=1).astype(int) - 1          # scales referenced per frame
        got_ops = ops.a[:T * 768].view(A.STEREO_DTYPE)
        for t in range(T):
            k = max(int(used[t]), 0)
            assert np.array_equal(got_ops["scale"][t][:k].view(np.uint32), want_ops["scale"][t][:k].view(np.uint32))
    else:
        spectra, info, tb, to, out, n = a[1], a[2], a[3], a[4], a[5], a[6]
    # the library writes into the decoder's (page-locked) staging buffer; readChunk hands out a copy
    assert J.to_number(n) == T and out.a.size == pcm.a.size and out is not pcm
    assert np.array_equal(spectra.a[:T * 2048].view(np.uint32), case["spectra"][0].reshape(-1).view(np.uint32))
    assert np.array_equal(info.a[:T * 16], want_info.view(np.uint8).reshape(-1))
    if any(b is not None for b in blocks):
        assert np.array_equal(to.a[:2 * T + 1], offs) and bytes(tb.a[:offs[-1]]) == bytes(blob[:offs[-1]])
    else:
        assert tb is None and to is None
    host_calls = [J.to_string(v) for v in dec.get("hostStereoCalls").items]
    if stereo_on_device:
        assert host_calls == []             # the reference's processMS / processIS were not run on the CPU
    else:
        assert host_calls.count("is") == T and host_calls.count("ms") == int(sum(
            bool(case["cpe"][0, t]["common_window"] and case["cpe"][0, t]["mask_present"]) for t in range(T)))


def test_js_quant_packer_writes_the_same_record_as_the_python_twin():
    """js/quant_pack.js `pack` executed by the interpreter: the aacfb_qframe record it writes for an ICStream
    equals aacjs_b200.pack_qframe's byte for byte (group lengths, band codes incl. the scalefactor-table
    index recovered from the Float32 value, the Huffman integers)."""
    from oracle import oracle as O
    from tools import jsmini as J

    js_dir = os.path.join(ROOT, "aac.js_b200", "js")
    sc = J.Runtime(js_dir).run(JS_STUBS + "ICStream.ZERO_BT = 0; ICStream.FIRST_PAIR_BT = 5;")
    sf_tab = O.dequant_table(1)
    tables = J.obj(SCALEFACTOR_TABLE=J.float32array(sf_tab))
    rt = J.Runtime(js_dir, stubs={"aac/src/ics": sc["ICStream"], "aac/src/huffman": J.obj(), "aac/src/tables": tables})
    pack = rt.require("./quant_pack")
    assert J.to_number(pack.get("RECORD_BYTES")) == A.QFRAME_DTYPE.itemsize
    rng = np.random.default_rng(8)
    for trial in range(6):
        short = trial % 2 == 1
        groups = int(rng.integers(1, 9)) if short else 1
        glen = np.zeros(8, np.int32)
        if short:
            cuts = np.sort(rng.choice(np.arange(1, 8), size=groups - 1, replace=False))
            glen[:groups] = np.diff(np.concatenate([[0], cuts, [8]]))
        else:
            glen[0] = 1
        max_sfb = int(rng.integers(0, 15 if short else 50))
        bt = rng.choice([0, 1, 3, 5, 11, 13, 14, 15], size=120).astype(np.int32)
        sf = sf_tab[rng.integers(0, 428, 120)] * np.where(bt == 13, -1, 1).astype(np.float32)
        sf[5] = np.float32("nan")
        quant = rng.integers(-8191, 8192, 1024).astype(np.int16)
        ics_py = {"info": {"windowSequence": 2 if short else 0, "groupCount": groups, "groupLength": glen, "maxSFB": max_sfb},
                  "bandTypes": bt, "scaleFactors": sf}
        want = A.pack_qframe(ics_py, quant)
        ics_js = J.obj(info=J.obj(windowSequence=2 if short else 0, groupCount=groups, groupLength=J.int32array(glen), maxSFB=max_sfb),
                       bandTypes=J.int32array(bt), scaleFactors=J.float32array(sf), quant=J.JSTyped("Int16Array", quant.copy()))
        buf = J.JSArrayBuffer(2304 * 2)
        rt2 = J.Runtime(js_dir)
        env = rt2.run("var bytes = new Uint8Array(buf), view = new DataView(buf);", {"buf": buf})
        buf.b[:] = 0xAA
        pack.get("pack").call(J.UNDEF, [ics_js, env["bytes"], env["view"], 2304.0])
        assert (buf.b[:2304] == 0xAA).all()
        assert bytes(buf.b[2304:]) == want.tobytes()
