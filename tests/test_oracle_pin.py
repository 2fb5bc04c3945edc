"""Pinning the oracle to the reference itself.

No JS engine exists in this image, so the reference's *own source files* for the path are executed
by tools/jsmini.py (a small generic ES5-subset interpreter) -- see tools/js_reference.py.  The
vectors it produced are committed as tests/golden/jsref_*.npz (made here, where /root/reference
exists); the oracle must reproduce them BIT FOR BIT.  When the reference tree is present the
interpreter is also run live on fresh random frames."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as O
from tools import jsmini as J
from tools import workloads as W

GOLD = os.path.join(os.path.dirname(__file__), "golden")
HAVE_REF = os.path.isdir("/root/reference/src")


def load_jsref(path):
    z = np.load(path)
    cfg, S, T, C, seed, carried, fix, mode = (int(v) for v in z["meta"])
    w = W.make(cfg, S, T, C, seed, "carried" if carried else "as_shipped")
    w["flags"] = mode if fix else 0  # the as-shipped reference: TNS is the identity (tns.js:122)
    return w, z["pcm"], z["overlap"]


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "jsref_*.npz"))), ids=os.path.basename)
def test_oracle_equals_interpreted_reference_bit_for_bit(path):
    w, pcm, ovl = load_jsref(path)
    got, gov = O.process(w["spectra"], w["info"], w["tns_blob"], w["tns_offsets"], sample_index=4, flags=w["flags"])
    assert np.array_equal(got.view(np.uint32), pcm.view(np.uint32))
    assert np.array_equal(gov.view(np.uint32), ovl.view(np.uint32))


def test_golden_set_covers_every_sequence_and_tns_mode():
    names = {os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "jsref_*.npz"))}
    for need in ("config1_mono_long", "config2_long", "config3_short", "config4_tns_as_shipped", "config4_tns_fixed_ar",
                 "config4_tns_fixed_ma", "config5_mixed"):
        assert f"jsref_{need}.npz" in names
    w, pcm, _ = load_jsref(os.path.join(GOLD, "jsref_config4_tns_as_shipped.npz"))
    # the reference as shipped ignores its TNS side info: same PCM as with TNS switched off
    plain, _ = O.process(w["spectra"], w["info"], None, None, sample_index=4, flags=0)
    assert np.array_equal(plain, pcm)
    w2, pcm_ar, _ = load_jsref(os.path.join(GOLD, "jsref_config4_tns_fixed_ar.npz"))
    assert not np.array_equal(pcm_ar, pcm)


# ---------------------------------------------------------------- the interpreter itself
def run(src):
    return J.Runtime(os.path.dirname(__file__)).run(src)


def test_jsmini_number_and_typed_array_semantics():
    s = run("""
        var f = new Float32Array(3); f[0] = 0.1; f[1] = 16777217; f[5] = 7;
        var a = f[0], b = f[1], c = f[5], d = f[NaN], n = f.length;
        var t = new Float32Array(20) - 3;          // ToNumber("0,0,...") -> NaN   (tns.js:122)
        var one = new Float32Array(1) - 3;         // ToNumber("0") = 0
        var mx = Math.max(0, t), cmp = (t <= 0), cmp2 = !(t > 0);
        var u = [1, 2, 3][7], e = 5 - undefined;
        var i32 = new Int32Array(2); i32[0] = 3.9; i32[1] = -1 >>> 0;
        var sh = (1 << 31), ush = (-1 >>> 28), ii = 7; ii <<= 2;
        var str = "x" + 1 + 2, num = 1 + 2 + "x";
    """)
    assert s["a"] == float(np.float32(0.1)) and s["b"] == 16777216.0 and s["c"] is J.UNDEF and s["d"] is J.UNDEF
    assert s["n"] == 3.0 and np.isnan(s["t"]) and s["one"] == -3.0
    assert np.isnan(s["mx"]) and s["cmp"] is False and s["cmp2"] is True
    assert s["u"] is J.UNDEF and np.isnan(s["e"])
    assert s["i32"].a.tolist() == [3, -1] and s["sh"] == -2147483648.0 and s["ush"] == 15.0 and s["ii"] == 28.0
    assert s["str"] == "x12" and s["num"] == "3x"


def test_jsmini_control_flow_scoping_and_objects():
    s = run("""
        function F(x) { this.x = x; }
        F.prototype.get = function() { return this.x + inc; };
        var inc = 10;
        var o = new F(5), g = o.get(), inst = o instanceof F;
        function hoist() { var r = typeof later; var later = 1; return r + "," + typeof fn; function fn() {} }
        var h = hoist();
        var sw = 0; switch (2) { case 1: sw += 1; case 2: sw += 2; case 3: sw += 4; break; case 4: sw += 8; }
        var acc = 0; for (var i = 0, j = 10; i < j; i++, j--) { if (i === 2) continue; if (i === 4) break; acc += i; }
        var w = 0; while (w < 5) w++;
        var top = 3, bottom = 2; for (var q = 0; q < 2; q++) { var top = bottom, bottom = top - 1; }
        var caught = (function() { return (0, arguments.length); })(1, 2, 3);
        var tern = inc > 5 ? "big" : "small";
    """)
    assert s["g"] == 15.0 and s["inst"] is True and s["h"] == "undefined,function"
    assert s["sw"] == 6.0 and s["acc"] == 0 + 1 + 3 and s["w"] == 5.0
    assert (s["top"], s["bottom"]) == (1.0, 0.0) and s["caught"] == 3.0 and s["tern"] == "big"
    with pytest.raises(J.JSThrow, match="boom"):
        run('throw new Error("boom");')


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present (GPU box)")
def test_live_interpreted_reference_random_frames():
    """Fresh random frames through the reference's own JS (all four sequences, both shapes) vs the oracle."""
    from tools.js_reference import Reference

    ref = Reference(fix_tns=True)
    rng = np.random.default_rng(2024)
    case = W.random_case(1, 6, 2, rng, tns_mode=1)
    case["info"]["window_sequence"][0, :, 0] = [0, 1, 2, 2, 3, 0]
    case["info"]["window_sequence"][0, :, 1] = [0, 0, 1, 2, 3, 0]
    pcm, ovl = ref.process(case["spectra"], case["info"], case["tns_blob"], case["tns_offsets"], 4, 1)
    got, gov = O.process(case["spectra"], case["info"], case["tns_blob"], case["tns_offsets"], sample_index=4, flags=1)
    assert np.array_equal(got.view(np.uint32), pcm.view(np.uint32))
    assert np.array_equal(gov.view(np.uint32), ovl.view(np.uint32))


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present (GPU box)")
def test_reference_tables_through_the_interpreter():
    """Window and FFT-root tables as the reference's own code builds them == the oracle's."""
    rt = J.Runtime("/root/reference/src")
    FFT = rt.require("./fft")
    for length, which in ((512.0, 0), (64.0, 1)):
        roots = FFT.construct([length]).get("roots").items
        js = np.array([[float(r.a[0]), float(r.a[1])] for r in roots], np.float32).reshape(-1)
        assert np.array_equal(js.view(np.uint32), O.table(which).view(np.uint32))
    with pytest.raises(J.JSThrow, match="No small frames allowed"):
        rt.require("./filter_bank").construct([True, 2.0])


def test_jsmini_host_side_features():
    """What the JS host files under aac.js_b200/js need beyond the reference's own sources:
    ArrayBuffer + typed-array views + DataView, TypedArray.set, try / catch / finally with
    instanceof dispatch, Function.prototype.call / apply, Array.prototype.push."""
    s = run("""
        var buf = new ArrayBuffer(32), b = new Uint8Array(buf), v = new DataView(buf);
        v.setFloat32(4, 1.5, true); v.setUint16(0, 0x1234, false); b[2] = 7;
        var f = new Float32Array(buf, 4, 2), x = f[0]; f[1] = 2.25;
        var g = v.getFloat32(8, true), n = buf.byteLength, b0 = b[0], b1 = b[1];
        var big = new Float32Array(8); big.set(f, 3); var y = big[4];
        function E1() {} function E2() {}
        var log = [];
        function t(k) { try { if (k == 1) throw new E1(); if (k == 2) throw new E2(); log.push('ok'); return 1; }
                        catch (err) { if (!(err instanceof E1)) throw err; log.push('caught'); return 2; }
                        finally { log.push('fin'); } }
        var r0 = t(0), r1 = t(1), r2;
        try { t(2); } catch (e) { r2 = e instanceof E2; }
        function who(a, b) { return this.tag + a + b; }
        var c1 = who.call({tag: 10}, 1, 2), c2 = who.apply({tag: 20}, [3, 4]);
    """)
    assert (s["x"], s["g"], s["n"], s["y"], s["b0"], s["b1"]) == (1.5, 2.25, 32.0, 2.25, 0x12, 0x34)
    assert (s["r0"], s["r1"], s["r2"], s["c1"], s["c2"]) == (1.0, 2.0, True, 13.0, 27.0)
    assert [J.to_string(v) for v in s["log"].items] == ["ok", "fin", "caught", "fin", "fin"]
