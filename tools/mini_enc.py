#!/usr/bin/env python
"""A small MPEG-1 video ENCODER (numpy) that writes I, P and B pictures with real image content.

Why it exists: the B-picture extension has no reference implementation to run -- the
reference skips B pictures (src/mpeg1.js:181-184) and the FFmpeg build behind cv2 writes none for
mpeg1video.  The syntax-level generator (tools/synth_es.py) covers the bitstream corners, but its random
coefficients overflow what a conforming decoder accepts, so FFmpeg's decoder conceals "errors" in them and
cannot serve as a cross-check.  This encoder produces conforming streams of natural content (the moving
scene of tools/gen_streams.py) that FFmpeg decodes cleanly: the oracle's B-picture path is compared with
FFmpeg's mpeg1video decoder on them (tests/test_b_pictures.py; PSNR, the two IDCTs differ).

Scope: fixed quantiser scale per picture, one slice per macroblock row, full + half-pel block matching in
a small window, forward / backward / interpolated prediction chosen by SAD, skipped macroblocks where the
syntax allows them, intra fallback.  Vectors keep their whole footprint inside the coded planes (the
reference reads references by flat index).  Closed loop (predicts from its own reconstruction,
float IDCT): good enough to keep the content clean, not bit-exact with any decoder.

    pictures in DISPLAY order -> encode(...) -> elementary stream bytes (pictures in CODED order)
"""
from __future__ import annotations

import os
import sys

import numpy as np
from scipy.fft import dctn, idctn

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import vlc_tables as V  # noqa: E402
from synth_es import (BitWriter, BitWriter_bits, CBP_CODE, DC_CHROMA_CODE, DC_LUMA_CODE, MBA_CODE,  # noqa: E402
                      MOTION_CODE, TYPE_B_CODE, TYPE_I_CODE, TYPE_P_CODE)

ZZ = np.array(V.ZIG_ZAG)
INTRA_Q = np.array(V.DEFAULT_INTRA_QUANT_MATRIX, dtype=np.float64)
NON_INTRA_Q = np.full(64, 16.0)


def scene(width, height, frames, seed=1234, noise=4, step=(3, 2)):
    """The moving test scene of tools/gen_streams.py as Y, Cb, Cr planes (4:2:0), display order."""
    import cv2
    rng = np.random.default_rng(seed)
    lo = rng.integers(0, 256, (height // 32 + 3, width // 32 + 3, 3), dtype=np.uint8)
    base = cv2.resize(lo, (width + 64, height + 64), interpolation=cv2.INTER_CUBIC).astype(np.int16)
    out = []
    for i in range(frames):
        dx, dy = (i * step[0]) % 64, (i * step[1]) % 64
        f = base[dy:dy + height, dx:dx + width].copy()
        cx = width // 2 + int(width / 3 * np.sin(i / 7.0))
        cy = height // 2 + int(height / 3 * np.cos(i / 9.0))
        cv2.circle(f, (cx, cy), max(height // 8, 4), (255, 64, 32), -1)
        if noise:
            f += rng.integers(-noise, noise + 1, f.shape, dtype=np.int16)
        bgr = np.clip(f, 0, 255).astype(np.uint8)
        yuv = cv2.cvtColor(bgr, cv2.COLOR_BGR2YUV_I420).reshape(-1)
        y = yuv[:width * height].reshape(height, width)
        cb = yuv[width * height:width * height * 5 // 4].reshape(height // 2, width // 2)
        cr = yuv[width * height * 5 // 4:].reshape(height // 2, width // 2)
        # keep away from the extremes: reconstruction overshoot must not clip differently in two decoders
        out.append(tuple(np.clip(p.astype(np.int16), 24, 232).astype(np.uint8) for p in (y, cb, cr)))
    return out


def predict(ref, x, y, size, mx, my):
    """size x size block at (x, y) of plane `ref`, vector (mx, my) in half-pel units (ISO 11172-2 2.4.4.2)."""
    fx, fy = x + (mx >> 1), y + (my >> 1)
    ox, oy = mx & 1, my & 1
    a = ref[fy:fy + size + oy, fx:fx + size + ox].astype(np.int32)
    if ox and oy:
        return (a[:-1, :-1] + a[:-1, 1:] + a[1:, :-1] + a[1:, 1:] + 2) >> 2
    if ox:
        return (a[:, :-1] + a[:, 1:] + 1) >> 1
    if oy:
        return (a[:-1, :] + a[1:, :] + 1) >> 1
    return a


class Encoder:
    def __init__(self, width, height, b_frames=2, gop_refs=4, qscale=(6, 8, 10), search=7, seed=1,
                 f_codes=(3, 3), full_pel=(0, 0), one_slice=False):
        assert width % 16 == 0 and height % 16 == 0
        self.w, self.h = width, height
        self.mbw, self.mbh = width // 16, height // 16
        self.b_frames, self.gop_refs = b_frames, gop_refs
        self.qs = dict(zip((1, 2, 3), qscale))
        self.search = search
        self.f_codes = f_codes    # (forward_f_code, backward_f_code): vectors in [-16 f, 16 f - 1] of their unit
        self.full_pel = full_pel  # (full_pel_forward_vector, full_pel_backward_vector): unit = a whole sample
        self.one_slice = one_slice  # one slice per picture (what FFmpeg writes) instead of one per macroblock row
        self.rng = np.random.default_rng(seed)
        self.out = BitWriter()
        self.stats = {"intra": 0, "fwd": 0, "bwd": 0, "bi": 0, "skipped": 0}
        self.trace = []  # (picture type, display index, macroblock, mode, forward vector, backward vector, cbp, skipped)

    # ------------------------------------------------------------------ vectors
    def inside(self, mb_col, mb_row, mx, my):
        x0, y0 = mb_col * 16 + (mx >> 1), mb_row * 16 + (my >> 1)
        if x0 < 0 or y0 < 0 or x0 + 16 + (mx & 1) > self.w or y0 + 16 + (my & 1) > self.h:
            return False
        cx, cy = int(mx / 2), int(my / 2)
        x0, y0 = mb_col * 8 + (cx >> 1), mb_row * 8 + (cy >> 1)
        return not (x0 < 0 or y0 < 0 or x0 + 8 + (cx & 1) > self.w // 2 or y0 + 8 + (cy & 1) > self.h // 2)

    def best_vector(self, cur_y, ref_y, mb_col, mb_row, limit, full_only=False):
        """Luma block matching: full-pel in +-search, then (unless full_only) the eight half-pel neighbours.
        Vector in half-pel units, |component| <= limit, footprint inside the planes."""
        x, y = mb_col * 16, mb_row * 16
        blk = cur_y[y:y + 16, x:x + 16].astype(np.int32)
        best = (None, 1 << 30)
        for dy in range(-self.search, self.search + 1):
            for dx in range(-self.search, self.search + 1):
                mx, my = 2 * dx, 2 * dy
                if abs(mx) > limit or abs(my) > limit or not self.inside(mb_col, mb_row, mx, my):
                    continue
                sad = int(np.abs(blk - predict(ref_y, x, y, 16, mx, my)).sum()) + abs(dx) + abs(dy)
                if sad < best[1]:
                    best = ((mx, my), sad)
        (bx, by), _ = best
        if full_only:
            return best
        for hy in (-1, 0, 1):
            for hx in (-1, 0, 1):
                mx, my = bx + hx, by + hy
                if abs(mx) > limit or abs(my) > limit or not self.inside(mb_col, mb_row, mx, my):
                    continue
                sad = int(np.abs(blk - predict(ref_y, x, y, 16, mx, my)).sum())
                if sad < best[1]:
                    best = ((mx, my), sad)
        return best

    def predict_mb(self, ref, mb_col, mb_row, mv):
        mx, my = mv
        cx, cy = int(mx / 2), int(my / 2)  # chroma: truncation toward zero (mpeg1.js:562-565)
        return (predict(ref[0], mb_col * 16, mb_row * 16, 16, mx, my),
                predict(ref[1], mb_col * 8, mb_row * 8, 8, cx, cy),
                predict(ref[2], mb_col * 8, mb_row * 8, 8, cx, cy))

    # ------------------------------------------------------------------ blocks
    @staticmethod
    def blocks_of(y, cb, cr):
        """The six 8x8 blocks of a macroblock given as (16x16, 8x8, 8x8) arrays."""
        return [y[:8, :8], y[:8, 8:], y[8:, :8], y[8:, 8:], cb, cr]

    def quantise(self, block, intra, qs):
        c = dctn(block.astype(np.float64), norm="ortho").reshape(64)
        lv = np.zeros(64, dtype=np.int32)
        if intra:
            lv[1:] = np.round(c[1:] * 8.0 / (qs * INTRA_Q[1:])).astype(np.int32)
            lv[0] = int(np.clip(round(c[0] / 8.0), 0, 255))
        else:
            lv = (np.sign(c) * np.floor(np.abs(c) * 8.0 / (qs * NON_INTRA_Q))).astype(np.int32)
        return np.clip(lv, -255, 255)

    @staticmethod
    def dequantise(lv, intra, qs):
        """What a decoder rebuilds (ISO form; the float IDCT of the closed loop)."""
        q = INTRA_Q if intra else NON_INTRA_Q
        lv = lv.astype(np.int64)
        if intra:
            r = (2 * lv * qs * q).astype(np.int64) // 16
        else:
            r = np.where(lv == 0, 0, np.sign(lv) * (((2 * np.abs(lv) + 1) * qs * q).astype(np.int64) // 16))
        r = np.where((r & 1) == 0, r - np.sign(r), r)
        r = np.clip(r, -2048, 2047).astype(np.float64)
        if intra:
            r[0] = lv[0] * 8.0
        return idctn(r.reshape(8, 8), norm="ortho")

    def write_block(self, w, lv, intra, luma, dc_pred):
        """Levels in raster order.  Returns the new intra DC predictor."""
        z = lv[ZZ]
        start = 0
        if intra:
            diff = int(z[0]) - dc_pred
            size = 0 if diff == 0 else int(abs(diff)).bit_length()
            w.code((DC_LUMA_CODE if luma else DC_CHROMA_CODE)[size])
            if size:
                w.put(diff if diff > 0 else diff + (1 << size) - 1, size)
            dc_pred = int(z[0])
            start = 1
        run = 0
        first = not intra
        for n in range(start, 64):
            level = int(z[n])
            if level == 0:
                run += 1
                continue
            code = V.DCT_RUN_LEVEL_CODE.get((run, abs(level)))
            if code is not None:
                if code == "1":
                    w.code("1" if first else "11")
                else:
                    w.code(code)
                w.put(1 if level < 0 else 0, 1)
            else:
                w.code(V.DCT_ESCAPE_CODE)
                w.put(run, 6)
                if -127 <= level <= 127:
                    w.put(level & 0xFF, 8)
                elif level > 0:
                    w.put(0, 8)
                    w.put(level, 8)
                else:
                    w.put(128, 8)
                    w.put(level + 256, 8)
            first = False
            run = 0
        w.code(V.DCT_EOB_CODE)
        return dc_pred

    # ------------------------------------------------------------------ motion syntax
    @staticmethod
    def write_motion(w, target, prev, f, r_size):
        d = target - prev
        if d > 16 * f - 1:
            d -= 32 * f
        elif d < -16 * f:
            d += 32 * f
        if f == 1 or d == 0:
            w.code(MOTION_CODE[d])
        else:
            a = abs(d) - 1
            code = (a >> r_size) + 1
            w.code(MOTION_CODE[code if d > 0 else -code])
            w.put(a & (f - 1), r_size)
        return target

    # ------------------------------------------------------------------ pictures
    def picture(self, ptype, temporal, cur, fwd, bwd):
        """Encodes `cur` (Y, Cb, Cr); fwd / bwd = reconstructed references.  Returns the reconstruction."""
        w = self.out
        (fc_f, fc_b), (fp_f, fp_b) = self.f_codes, self.full_pel
        f, r_size = 1 << (fc_f - 1), fc_f - 1
        fb, r_size_b = 1 << (fc_b - 1), fc_b - 1
        sc_f, sc_b = (2 if fp_f else 1), (2 if fp_b else 1)  # half-pel units per bitstream unit
        limit, limit_b = (16 * f - 2) * sc_f, (16 * fb - 2) * sc_b
        w.start_code(0x00)
        w.put(temporal & 1023, 10)
        w.put(ptype, 3)
        w.put(0xFFFF, 16)
        if ptype >= 2:
            w.put(fp_f, 1)
            w.put(fc_f, 3)
        if ptype == 3:
            w.put(fp_b, 1)
            w.put(fc_b, 3)
        w.put(0, 1)
        w.align()
        qs = self.qs[ptype]
        rec = [np.zeros_like(p) for p in cur]
        sw = None
        for row in range(self.mbh):
            if sw is None:  # a slice starts: every predictor is reset (mpeg1.js:255-266)
                sw = BitWriter()
                sw.start_code(row + 1)
                sw.put(qs, 5)
                sw.put(0, 1)
                pf, pb = [0, 0], [0, 0]
                dc = [128, 128, 128]
                last_mode = None
                addr = row * self.mbw - 1
            first_mb = 0 if self.one_slice else row * self.mbw                            # of the slice
            last_mb = self.mbw * self.mbh - 1 if self.one_slice else (row + 1) * self.mbw - 1
            for col in range(self.mbw):
                mb = row * self.mbw + col
                src = (cur[0][row * 16:row * 16 + 16, col * 16:col * 16 + 16].astype(np.int32),
                       cur[1][row * 8:row * 8 + 8, col * 8:col * 8 + 8].astype(np.int32),
                       cur[2][row * 8:row * 8 + 8, col * 8:col * 8 + 8].astype(np.int32))
                mode, mvf, mvb, pred = "intra", (0, 0), (0, 0), None
                if ptype >= 2:
                    cands = []
                    vf, sf = self.best_vector(cur[0], fwd[0], col, row, limit, bool(fp_f))
                    cands.append(("fwd", vf, (0, 0), sf))
                    if ptype == 3:
                        vb, sb = self.best_vector(cur[0], bwd[0], col, row, limit_b, bool(fp_b))
                        cands.append(("bwd", (0, 0), vb, sb))
                        bi = (predict(fwd[0], col * 16, row * 16, 16, *vf) + predict(bwd[0], col * 16, row * 16, 16, *vb) + 1) >> 1
                        cands.append(("bi", vf, vb, int(np.abs(src[0] - bi).sum())))
                    mode, mvf, mvb, sad = min(cands, key=lambda c: c[3])
                    flat = int(np.abs(src[0] - int(src[0].mean())).sum())
                    if flat + 512 < sad or self.rng.random() < 0.02:
                        mode = "intra"
                    else:
                        a = self.predict_mb(fwd, col, row, mvf) if mode in ("fwd", "bi") else None
                        b = self.predict_mb(bwd, col, row, mvb) if mode in ("bwd", "bi") else None
                        pred = a if b is None else (b if a is None else tuple((p + q + 1) >> 1 for p, q in zip(a, b)))
                intra = mode == "intra"
                resid = src if intra else tuple(s - p for s, p in zip(src, pred))
                levels = [self.quantise(blk, intra, qs) for blk in self.blocks_of(*resid)]
                cbp = 0x3F if intra else sum((0x20 >> k) for k in range(6) if np.any(levels[k]))
                # skipped macroblock: nothing coded, not first / last of the slice, and -- P: zero vector,
                # B: the same prediction and vectors as the macroblock before (which must not be intra)
                can_skip = not intra and cbp == 0 and first_mb < mb < last_mb
                if can_skip and ptype == 2:
                    can_skip = mvf == (0, 0)
                elif can_skip:
                    can_skip = last_mode == (mode, mvf if mode != "bwd" else None, mvb if mode != "fwd" else None)
                    # No skipped macroblock in a direction coded with full_pel vectors: FFmpeg's decoder -- the cross-check
                    # these streams are written for -- repeats the vector of the macroblock before WITHOUT the full-pel
                    # doubling there (it keeps last_mv in bitstream units), where ISO 11172-2 2.4.4.2 says "the same
                    # vector".  (The syntax generator, tools/synth_es.py, does write that case: oracle vs CUDA.)
                    if (mode != "bwd" and fp_f) or (mode != "fwd" and fp_b):
                        can_skip = False
                self.trace.append((ptype, temporal, mb, mode, mvf, mvb, cbp, bool(can_skip)))
                if can_skip:
                    self.stats["skipped"] += 1
                    dc = [128, 128, 128]
                    if ptype == 2:
                        pf = [0, 0]
                    rec_mb = pred
                else:
                    self.stats[mode] += 1
                    inc = mb - addr
                    mw = BitWriter()  # the macroblock's bits, appended to the slice below
                    while inc > 33:  # macroblock_escape (a long skipped run in a one-slice picture)
                        mw.code(MBA_CODE[35])
                        inc -= 33
                    mw.code(MBA_CODE[inc])
                    addr = mb
                    if ptype == 1:
                        mw.code(TYPE_I_CODE[0x01])
                    elif ptype == 2:
                        mw.code(TYPE_P_CODE[0x01 if intra else (0x0A if cbp else 0x08)])
                    else:
                        t = 0x01 if intra else ({"fwd": 0x08, "bwd": 0x04, "bi": 0x0C}[mode] | (0x02 if cbp else 0))
                        mw.code(TYPE_B_CODE[t])
                    if intra:
                        pf, pb = [0, 0], [0, 0]
                        last_mode = None
                    else:
                        dc = [128, 128, 128]
                        if mode in ("fwd", "bi"):  # (the bitstream carries the vector in its own unit: mpeg1.js:421-424)
                            pf[0] = self.write_motion(mw, mvf[0] // sc_f, pf[0], f, r_size)
                            pf[1] = self.write_motion(mw, mvf[1] // sc_f, pf[1], f, r_size)
                        elif ptype == 2:
                            pf = [0, 0]
                        if mode in ("bwd", "bi"):
                            pb[0] = self.write_motion(mw, mvb[0] // sc_b, pb[0], fb, r_size_b)
                            pb[1] = self.write_motion(mw, mvb[1] // sc_b, pb[1], fb, r_size_b)
                        last_mode = (mode, mvf if mode != "bwd" else None, mvb if mode != "fwd" else None)
                        if cbp:
                            mw.code(CBP_CODE[cbp])
                    for k in range(6):
                        if cbp & (0x20 >> k):
                            which = 0 if k < 4 else k - 3
                            dc[which] = self.write_block(mw, levels[k], intra, k < 4, dc[which])
                    # The reference ends a slice as soon as the next BYTE-ALIGNED bytes are a start code prefix
                    # (buffer.js:141-150): a last macroblock whose bits all fit into the byte the
                    # previous one ended in would never be decoded by it (a conforming decoder decodes it).
                    # macroblock_stuffing in front pushes it over the byte boundary.
                    if mb == last_mb and len(sw.bits) % 8 and len(sw.bits) % 8 + len(mw.bits) <= 8:
                        sw.code(MBA_CODE[34])
                    sw.bits.extend(mw.bits)
                    # closed loop: what a decoder rebuilds
                    rb = [self.dequantise(levels[k], intra, qs) if cbp & (0x20 >> k) else np.zeros((8, 8)) for k in range(6)]
                    ry = np.block([[rb[0], rb[1]], [rb[2], rb[3]]])
                    if intra:
                        rec_mb = (ry, rb[4], rb[5])
                    else:
                        rec_mb = (pred[0] + ry, pred[1] + rb[4], pred[2] + rb[5])
                rec[0][row * 16:row * 16 + 16, col * 16:col * 16 + 16] = np.clip(np.rint(rec_mb[0]), 0, 255)
                rec[1][row * 8:row * 8 + 8, col * 8:col * 8 + 8] = np.clip(np.rint(rec_mb[1]), 0, 255)
                rec[2][row * 8:row * 8 + 8, col * 8:col * 8 + 8] = np.clip(np.rint(rec_mb[2]), 0, 255)
            if not self.one_slice or row == self.mbh - 1:
                sw.align()
                w.bits.extend(BitWriter_bits(sw.tobytes()))
                sw = None
        return rec

    def encode(self, frames):
        """frames: [(Y, Cb, Cr)] in display order.  Returns (ES bytes, picture types in coded order, display
        index of every coded picture)."""
        w = self.out
        w.start_code(0xB3)
        w.put(self.w, 12)
        w.put(self.h, 12)
        w.put(1, 4)
        w.put(5, 4)
        w.put(0x3FFFF, 18)
        w.put(1, 1)
        w.put(20, 10)
        w.put(0, 1)
        w.put(0, 1)
        w.put(0, 1)
        w.start_code(0xB8)
        w.put(0, 25)
        w.put(1, 1)
        w.put(0, 1)
        m = self.b_frames + 1
        refs = list(range(0, len(frames), m))
        order, types = [], []
        prev_ref = None
        for k, r in enumerate(refs):
            order.append(r)
            types.append(1 if k % self.gop_refs == 0 else 2)
            if prev_ref is not None:
                for d in range(prev_ref + 1, r):
                    order.append(d)
                    types.append(3)
            prev_ref = r
        older = newer = None
        for idx, t in zip(order, types):
            cur = frames[idx]
            if t == 3:
                self.picture(3, idx, cur, older, newer)
            else:
                rec = self.picture(t, idx, cur, newer, None)
                older, newer = newer, rec
        w.start_code(0xB7)
        return w.tobytes(), types, order


def make_b_clip(width=176, height=144, frames=13, b_frames=2, seed=1234, **kw):
    """ES bytes, picture types (coded order), display index per coded picture.  Trailing display pictures that
    have no closing reference are dropped."""
    n = ((frames - 1) // (b_frames + 1)) * (b_frames + 1) + 1
    enc = Encoder(width, height, b_frames=b_frames, **kw)
    es, types, order = enc.encode(scene(width, height, n, seed))
    return es, types, order, enc.stats


if __name__ == "__main__":
    es, types, order, stats = make_b_clip()
    print(len(es), "bytes", types, order, stats)
