#!/usr/bin/env python
"""Generate tests/golden/: small elementary streams plus the outputs of the REFERENCE itself on them.

Run in the build container (needs oracle/_ref/libjsmpeg_ref.so, i.e. /root/reference):

    python tools/make_golden.py

For every case it writes
    tests/golden/<name>.es     the elementary stream (synthetic syntax-corner streams of
                               tools/synth_es.py, and two small cv2/FFmpeg-encoded clips)
    tests/golden/<name>.json   per decode() call: the bit index afterwards and the sha256 of the
                               Y / Cr / Cb planes the reference exposes through get_{y,cr,cb}_ptr,
                               plus width/height/coded size/frame rate
all produced by the unmodified reference C (src/wasm/mpeg1.c + buffer.c compiled by
oracle/Makefile) driven through its 15-function ABI.  The reference ships no golden vectors of
its own; these are the vectors that pin the oracle and the CUDA path.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import gen_streams  # noqa: E402
import helpers  # noqa: E402
import synth_es  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

FFMPEG_CASES = {
    "ffmpeg_320x240_ip": dict(width=320, height=240, frames=14, seed=1234, noise=9),
    "ffmpeg_176x144_ip": dict(width=176, height=144, frames=26, seed=7, noise=4),
}


def sha(a):
    return hashlib.sha256(a.tobytes()).hexdigest()


def describe(lib, es):
    frames, idx, d = helpers.decode_all(lib, [(0.0, es)])
    out = {"width": d.width, "height": d.height, "coded_size": d.codedSize, "frame_rate": round(float(d.frameRate), 3),
           "es_sha256": hashlib.sha256(es).hexdigest(),
           "pictures": [{"index": i, "y": sha(y), "cr": sha(cr), "cb": sha(cb)} for i, (y, cr, cb) in zip(idx, frames)]}
    d.destroy()
    return out


def main():
    ref = helpers.ref_lib()
    if ref is None:
        raise SystemExit("oracle/_ref/libjsmpeg_ref.so missing: run `make -C oracle` where /root/reference exists")
    os.makedirs(GOLDEN, exist_ok=True)
    cases = {name: synth_es.make_case(name) for name in synth_es.CASES}
    for name, kw in FFMPEG_CASES.items():
        packets = gen_streams.make_clip_es(kw["width"], kw["height"], kw["frames"], kw["seed"], kw["noise"])
        cases[name] = b"".join(p for _, p in packets)
    for name, es in cases.items():
        with open(os.path.join(GOLDEN, name + ".es"), "wb") as f:
            f.write(es)
        info = describe(ref, es)
        info["generator"] = "tools/synth_es.py" if name in synth_es.CASES else "tools/gen_streams.py (cv2/FFmpeg mpeg1video)"
        info["produced_by"] = "reference src/wasm/mpeg1.c + buffer.c (unmodified), oracle/_ref build"
        with open(os.path.join(GOLDEN, name + ".json"), "w") as f:
            json.dump(info, f, indent=1)
        print(f"{name}: {len(es)} bytes, {len(info['pictures'])} pictures")


if __name__ == "__main__":
    main()
