"""SURVEY 8f rank 1 pinned to the TEXT of the reference's src/ts.js: the hand-derived cases of
tools/make_ts_cases.py (tests/fixtures/ts_cases.json; every expected delivery is a list of (packet, first
payload byte) pairs worked out from ts.js by hand -- the reference's JS cannot run in this image).

  * the host mirror jsmpeg_b200/ts.py must make exactly those destination.write(pts, buffers) calls
    (chunking, pts, bytes) and keep exactly that many leftover bytes;
  * the fixture file must be what the generator produces (nobody edits the JSON by hand);
  * (-m gpu) the device demuxer must append exactly the concatenation of the video deliveries and report the
    PES starts; any chunking of the input must give the same result.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers
from jsmpeg_b200 import ts

CASES = json.load(open(os.path.join(helpers.ROOT, "tests", "fixtures", "ts_cases.json")))["cases"]


class Recorder:
    def __init__(self):
        self.calls = []

    def write(self, pts, buffers):
        self.calls.append((pts, b"".join(buffers)))


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_host_mirror_makes_the_hand_derived_deliveries(case):
    demux = ts.TS()
    rec = {sid: Recorder() for sid in case["connect"]}
    for sid, r in rec.items():
        demux.connect(sid, r)
    for w in case["writes"]:
        demux.write(bytes.fromhex(w))
    for sid in case["connect"]:
        want = case["expect"].get(str(sid), [])
        got = rec[sid].calls
        assert len(got) == len(want), (case["name"], sid, len(got), len(want))
        for (pts, payload), w in zip(got, want):
            assert payload == bytes.fromhex(w["payload"]), (case["name"], sid, w["parts"])
            assert abs(pts - w["pts_ticks"] / 90000.0) < 1e-9
    assert len(demux.leftover) == case["leftover_bytes"]


def test_fixture_file_is_what_the_generator_writes():
    before = open(os.path.join(helpers.ROOT, "tests", "fixtures", "ts_cases.json")).read()
    sys.path.insert(0, os.path.join(helpers.ROOT, "tools"))
    import make_ts_cases
    now = json.dumps({"generator": "tools/make_ts_cases.py (hand-derived from src/ts.js; see the comments there)",
                      "cases": make_ts_cases.build()}, indent=1)
    assert now == before


def _device_demux(writes, stream_id=0xE0):
    from jsmpeg_b200.batch import BatchDecoder
    bd = BatchDecoder(1)
    total, pes = 0, []
    for w in writes:
        n, p = bd.write_ts(0, w, stream_id)
        total += n
        pes += p
    es = bytes(helpers_es(bd, total))
    bd.close()
    return es, pes


def helpers_es(bd, total):
    """the stream's bit buffer as the host sees it (the write pointer is at its end)"""
    import ctypes
    ptr = bd.lib.jsmpeg_b200_batch_get_write_ptr(bd.handle, 0, 0)
    return ctypes.string_at(ptr - total, total) if total else b""


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_device_demux_appends_the_hand_derived_video_bytes(case):
    want = b"".join(bytes.fromhex(w["payload"]) for w in case["expect"].get(str(0xE0), []))
    starts = []
    # PES starts of the VIDEO stream: a delivery's first part that begins behind a PES header (offset 18 in these cases)
    off = 0
    for w in case["expect"].get(str(0xE0), []):
        if w["parts"] and w["parts"][0][1] == 18:
            starts.append((off, w["pts_ticks"] / 90000.0))
        off += len(w["payload"]) // 2
    writes = [bytes.fromhex(w) for w in case["writes"]]
    es, pes = _device_demux(writes)
    # the reference delivers what is COMPLETE; the device appends what has ARRIVED.  The hand-derived bytes are therefore a
    # prefix of the device's (the rest being payload the reference still holds in pi.buffers) -- and equal when the case ends complete.
    assert es[:len(want)] == want, case["name"]
    assert [(o, round(t, 6)) for o, t in pes][:len(starts)] == [(o, round(t, 6)) for o, t in starts], (case["name"], pes, starts)


@pytest.mark.gpu
def test_device_demux_equals_host_mirror_for_any_chunking_garbage_and_rebinding():
    """Encoder-made clip with an audio-style second PID spliced in, a PID re-bound mid-stream, garbage between packets;
    written whole, in 188-byte multiples and in odd chunk sizes: the ES must equal the host mirror's (held + delivered)."""
    import gen_streams
    clip = gen_streams.make_clip_ts(176, 144, 12, seed=3, noise=4)
    packets = [clip[i:i + 188] for i in range(0, len(clip), 188)]
    rng = np.random.default_rng(11)
    # garbage: 1..186 random bytes (no 0x47) at three places, each followed by >= 6 packets
    data = bytearray()
    for k, pkt in enumerate(packets):
        if k in (5, len(packets) // 2, len(packets) - 20):
            junk = rng.integers(0, 256, int(rng.integers(1, 187)), dtype=np.uint8)
            junk[junk == 0x47] = 0x46
            data += junk.tobytes()
        data += pkt
    data = bytes(data)

    def host(writes):
        demux = ts.TS()
        col = ts.ESCollector()
        demux.connect(0xE0, col)
        for w in writes:
            demux.write(w)
        pi = demux.pes_packet_info[0xE0]
        return col.es + b"".join(pi.buffers)  # delivered + still held

    for chunk in (len(data), 188 * 7, 1000, 333):
        writes = [data[i:i + chunk] for i in range(0, len(data), chunk)]
        want = host(writes)
        es, pes = _device_demux(writes)
        assert es == want, f"chunk size {chunk}: {len(es)} vs {len(want)} bytes"
        assert len(pes) >= 10
