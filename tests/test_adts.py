"""ADTS frame index: aacfb_adts_index locates the access units of a
byte buffer from their headers alone.  Host-side integer work, no GPU needed: bit-exact against the
oracle's bit-reader restatement of ADTSDemuxer.readHeader (adts_demuxer.js:28-52), which is pinned
to the reference's own function run by tools/jsmini.py (tests/golden/adts/jsref_adts.npz; live as
well where /root/reference exists)."""
import os

import numpy as np
import pytest

import aacjs_b200 as A
from oracle import oracle as O
from tools import workloads as W

GOLD = os.path.join(os.path.dirname(__file__), "golden", "adts", "jsref_adts.npz")
HAVE_REF = os.path.isdir("/root/reference/src")
FIELDS = ("frameLength", "bits", "profile", "samplingIndex", "chanConfig", "numFrames")


def fixture():
    z = np.load(GOLD)
    seed, n = (int(v) for v in z["meta"])
    data, meta = W.adts_stream(np.random.default_rng(seed), n)
    return z, data, meta


def test_oracle_equals_interpreted_reference():
    z, data, _ = fixture()
    assert len(z["headers"]) == 64
    for row in z["headers"]:
        h = O.adts_header(data[int(row[0]):])
        assert [h[k] for k in FIELDS] == [int(v) for v in row[1:]]
    assert bytes(z["error"]).decode() == "Invalid ADTS header."   # adts_demuxer.js:30
    assert O.adts_header(b"\xff\xe1" + bytes(8)) is None


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference is not on this machine")
def test_oracle_equals_reference_live():
    from tools.js_reference import AdtsReference

    ref, rng = AdtsReference(), np.random.default_rng(5)
    data, meta = W.adts_stream(rng, 40, crc_every=2, sampling_index=11, chan_config=7)
    for off, *_ in meta:
        j, err = ref.read_header(data[off:])
        assert err is None and j == O.adts_header(data[off:])
    for _ in range(50):   # random bytes: both accept or both reject
        junk = bytes(rng.integers(0, 256, 12, dtype=np.uint8))
        j, err = ref.read_header(junk)
        o = O.adts_header(junk)
        assert (j is None) == (o is None) and (j is None or j == o)


def test_index_matches_the_oracle_walk():
    z, data, meta = fixture()
    frames, consumed = A.adts_index(data)
    assert len(frames) == len(meta) == len(z["headers"]) and consumed == len(data)
    for f, row in zip(frames, z["headers"]):
        h = O.adts_header(data[int(f["offset"]):])
        assert int(f["offset"]) == int(row[0])
        assert (int(f["frame_length"]), 8 * int(f["header_bytes"]), int(f["profile"]), int(f["sampling_index"]),
                int(f["chan_config"]), int(f["num_frames"])) == tuple(h[k] for k in FIELDS)


def test_index_edge_cases():
    rng = np.random.default_rng(3)
    data, meta = W.adts_stream(rng, 9)
    # empty / shorter than a header: nothing, nothing consumed
    assert A.adts_index(b"")[1] == 0 and len(A.adts_index(data[:6])[0]) == 0
    # a cut inside the last frame (even inside its header): the rewind point is that frame's syncword
    for cut in (1, 3, 7, meta[-1][1] - 1):
        frames, consumed = A.adts_index(data[:meta[-1][0] + cut])
        assert len(frames) == len(meta) - 1 and consumed == meta[-1][0]
    # capacity limits the walk; resuming at `consumed` finds the rest
    first, c1 = A.adts_index(data, capacity=4)
    rest, c2 = A.adts_index(data[c1:])
    assert len(first) == 4 and len(rest) == 5 and c1 + c2 == len(data)
    assert [int(o) + c1 for o in rest["offset"]] == [m[0] for m in meta[4:]]
    # bad syncword / a frame shorter than its header: the reference's error
    with pytest.raises(A.AacfbError, match="Invalid ADTS header"):
        A.adts_index(b"\xff\xe0" + bytes(10))
    bad = bytearray(data)
    bad[meta[2][0] + 1] &= 0x0f
    with pytest.raises(A.AacfbError, match="AACFB_ERR_ADTS"):
        A.adts_index(bytes(bad))
    short = bytearray(data[:16])
    short[3] &= 0xfc; short[4] = 0; short[5] = (3 << 5) | (short[5] & 0x1f)   # frameLength = 3
    with pytest.raises(A.AacfbError, match="shorter than its header"):
        A.adts_index(bytes(short))


def test_index_throughput_is_not_the_bottleneck():
    """One header hop per access unit: the index of a 30 MB stream takes milliseconds (the oracle's
    bit-by-bit reader is there for fidelity, not speed)."""
    import time

    rng = np.random.default_rng(1)
    data, meta = W.adts_stream(rng, 2000)
    big = np.frombuffer(data * 20, np.uint8)
    t0 = time.perf_counter()
    frames, consumed = A.adts_index(big)
    dt = time.perf_counter() - t0
    assert len(frames) == 40000 and consumed == big.size
    assert dt < 0.5, dt
