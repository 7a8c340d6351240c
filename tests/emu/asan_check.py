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
        walk_exact(es, s, mbw, mbh, 4)  # the slice walk: a lane per slice, serial fall-back
        count += 1

# ---- the B-picture extension's device code (walk_b.cuh, stage 1b, recon.cuh<BIDIR>) on exact-size buffers:
# the syntax generator's B streams clean, bit-flipped and truncated; references filled with noise
vp = ctypes.c_void_p
lib.emu_walk_picture_b.argtypes = [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, vp, vp, vp]
lib.emu_expand_picture.argtypes = [vp, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]
lib.emu_reconstruct_picture_b.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int]
lib.emu_set_quant(bytes(range(8, 72)), bytes([16] * 64))
b_streams = [synth_es.make_case(c) for c in synth_es.B_CASES]
for es in list(b_streams):
    for trial in range(2):
        bad = bytearray(es)
        for pos in rng.integers(64, len(es), size=5):
            bad[pos] ^= 1 << int(rng.integers(0, 8))
        b_streams.append(bytes(bad))
    b_streams.append(es[: len(es) * 3 // 5])
b_count = 0
for es in b_streams:
    if es.find(b"\x00\x00\x01\xb3") < 0:
        continue
    mbw, mbh = T.stream_geometry(es)
    if mbw == 0 or mbh == 0 or mbw * mbh > 20000:
        continue
    mb = mbw * mbh
    n = len(es)
    buf = np.zeros((n + 3) // 4 * 4, dtype=np.uint8)
    buf[:n] = np.frombuffer(es, dtype=np.uint8)
    size = mb * 384 + mbw * 16 + 64  # a plane set as the product allocates it (engine.cu)
    fwd = rng.integers(0, 256, size, dtype=np.uint8)
    bwd = rng.integers(0, 256, size, dtype=np.uint8)
    b_starts = [s for s in T.picture_starts(es) if s + 1 < n and ((es[s + 1] >> 3) & 7) == 3]
    for s in (b_starts if limit is None else b_starts[:2]):
        hdr = np.zeros(mb * 4, dtype=np.uint32)
        park = np.zeros(mb * 6 * 2, dtype=np.uint32)
        coef = np.zeros(mb * 6 * 32, dtype=np.uint32)
        info = np.zeros(12, dtype=np.int32)
        cur = np.zeros(size, dtype=np.uint8)
        lib.emu_walk_picture_b(buf.ctypes.data, n, s, mbw, mbh, hdr.ctypes.data, park.ctypes.data, info.ctypes.data)
        if info[2] == 1:
            lib.emu_expand_picture(buf.ctypes.data, n, mbw, mbh, hdr.ctypes.data, park.ctypes.data, coef.ctypes.data, info.ctypes.data)
            lib.emu_reconstruct_picture_b(hdr.ctypes.data, coef.ctypes.data, cur.ctypes.data, fwd.ctypes.data, bwd.ctypes.data, mbw, mbh)
        b_count += 1
print("asan clean over", count, "pictures and", b_count, "B pictures (walk + expand + two-reference reconstruct)")
