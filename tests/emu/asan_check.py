"""Run both walks of jsmpeg_b200/csrc/walk.cuh (host emulation) under AddressSanitizer on exact-size
heap buffers: clean, bit-flipped and truncated streams.  Started by tests/test_walk_emu.py with
libasan preloaded; prints "asan clean over N pictures" on success (ASan aborts the process otherwise)."""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "tools"))
import test_walk_emu as T  # noqa: E402
import synth_es  # noqa: E402

lib = ctypes.CDLL(sys.argv[1])
lib.emu_walk_picture.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]


def walk_exact(es, start, mbw, mbh, lanes):
    mb = mbw * mbh
    n = len(es)
    buf = np.zeros((n + 3) // 4 * 4, dtype=np.uint8)  # the walk reads whole words, like the device mirror
    buf[:n] = np.frombuffer(es, dtype=np.uint8)
    hdr = np.zeros(mb * 4, dtype=np.uint32)
    park = np.zeros(mb * 6 * 2, dtype=np.uint32)  # exact size: the dense {bit offset, dc} side array
    info = np.zeros(12, dtype=np.int32)
    lib.emu_walk_picture(buf.ctypes.data, n, start, mbw, mbh, hdr.ctypes.data, park.ctypes.data, info.ctypes.data, lanes)


rng = np.random.default_rng(3)
streams = [open(os.path.join(T.HERE, "golden", g + ".es"), "rb").read() for g in T.GOLDEN]
streams += [synth_es.make_case(c) for c in synth_es.CASES]
for es in list(streams):
    bad = bytearray(es)
    for pos in rng.integers(64, len(es), size=6):
        bad[pos] ^= 1 << int(rng.integers(0, 8))
    streams.append(bytes(bad))
    streams.append(es[: len(es) * 2 // 3])
count = 0
limit = int(os.environ.get("ASAN_CHECK_PICTURES", "0")) or None  # the test keeps the CPU suite short; 0 = everything
for es in streams:
    if es.find(b"\x00\x00\x01\xb3") < 0:
        continue
    mbw, mbh = T.stream_geometry(es)
    if mbw == 0 or mbh == 0 or mbw * mbh > 20000:
        continue
    starts = T.picture_starts(es)
    for s in (starts if limit is None else starts[:3]):
        if limit is not None and count >= limit:
            break
        walk_exact(es, s, mbw, mbh, 1)  # staged records + fix-up, exact-size staging area
        walk_exact(es, s, mbw, mbh, 2)  # staging area of 40 entries: lanes run out, second pass
        walk_exact(es, s, mbw, mbh, 0)
        count += 1
print("asan clean over", count, "pictures")
