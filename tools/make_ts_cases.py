#!/usr/bin/env python
"""Hand-derived MPEG-TS demux cases -> tests/fixtures/ts_cases.json (pins SURVEY 8f rank 1 to the TEXT of
the reference's src/ts.js, which cannot be executed in this image).

Every case is a handful of 188-byte packets built field by field below, one or more write() calls, and
the deliveries `destination.write(pts, buffers)` the reference makes -- written down BY HAND as lists of
(packet number, first payload byte) pairs, each with the ts.js lines that decide it.  A reviewer replays a
case against ts.js with pencil and paper: payload bytes are a ramp ((7 * packet + offset) & 255) so a
slice that is off by one byte is a different byte string.

The expected byte strings in the JSON are nothing but `packet[k][start:188]` concatenated in the stated
order; no demuxer of ours is involved in producing them (tools/make_ts_cases.py imports none).

    python tools/make_ts_cases.py        # rewrites tests/fixtures/ts_cases.json
"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VIDEO, AUDIO = 0xE0, 0xC0


def packet(k, pid, pusi, afc, af_len=None, head=b""):
    """Packet number k.  Bytes 0..3: sync 0x47, pusi / PID, adaptation_field_control / continuity (ts.js:45-58).
    afc & 2: adaptation field = one length byte + af_len bytes (ts.js:73-77).  `head` = bytes placed right
    after it (a PES header, or anything else); the rest of the packet is the ramp."""
    b = bytearray((7 * k + i) & 255 for i in range(188))
    b[0] = 0x47
    b[1] = (pusi << 6) | (pid >> 8)
    b[2] = pid & 255
    b[3] = (afc << 4) | (k & 15)
    at = 4
    if afc & 2:
        b[4] = af_len
        for i in range(af_len):
            if 5 + i < 188:
                b[5 + i] = 0xFF if i else 0x00
        at = 5 + af_len
    b[at:at + len(head)] = head[:max(0, 188 - at)]
    return bytes(b)


def pes_header(stream_id, pts_ticks=None, packet_length=0, header_length=None):
    """00 00 01 sid | length(16) | one skipped byte | PTS_DTS_flags(2) + 6 skipped bits | header_length(8) | PTS
    (ts.js:79-116).  9 + header_length bytes; the payload begins right after (ts.js:90, :126)."""
    opt = b""
    if pts_ticks is not None:
        p = pts_ticks
        opt = bytes([0x21 | ((p >> 29) & 0x0E), (p >> 22) & 0xFF, 0x01 | ((p >> 14) & 0xFE), (p >> 7) & 0xFF, 0x01 | ((p << 1) & 0xFE)])
    hl = len(opt) if header_length is None else header_length
    opt = opt + b"\xff" * (hl - len(opt))
    return b"\x00\x00\x01" + bytes([stream_id, packet_length >> 8, packet_length & 255, 0x80, 0x80 if pts_ticks is not None else 0x00, hl]) + opt


def slices(packets, parts):
    return b"".join(packets[k][start:188] for k, start in parts)


def case(name, note, packets, writes, connect, expect, leftover=0, prefix=b""):
    """writes: list of byte strings handed to TS.write() in order; expect: {stream id: [(pts ticks, [(k, start)...])]}"""
    return {
        "name": name, "note": note, "connect": connect, "writes": [w.hex() for w in writes],
        "expect": {str(sid): [{"pts_ticks": pts, "parts": parts, "payload": slices(packets, parts).hex()} for pts, parts in lst]
                   for sid, lst in expect.items()},
        "leftover_bytes": leftover,
    }


def build():
    cases = []
    PID = 0x100

    # A -- one video PES over three packets; the third is padded with an adaptation field.
    #   P0 pusi, payload only: PES header at byte 4 (ts.js:79), header_length 5 -> payload begins at 4 + 9 + 5 = 18 (:90, :126)
    #   P1 continuation: payload = bytes 4..187 (:132-136)
    #   P2 adaptation field of 1 + 100 bytes (:73-77): payload = bytes 105..187; !payloadStart && (afc & 2) -> "padded", the
    #      video frame end is guessed (:143-146) -> packetComplete: ONE write(pts = 90000 ticks, 170 + 184 + 83 = 437 bytes)
    p = [packet(0, PID, 1, 1, head=pes_header(VIDEO, 90000)), packet(1, PID, 0, 1), packet(2, PID, 0, 3, af_len=100)]
    allp = b"".join(p)
    cases.append(case("padded_packet_ends_video_pes", "ts.js:79-126, 132-146", p, [allp], [VIDEO],
                      {VIDEO: [(90000, [(0, 18), (1, 4), (2, 105)])]}))
    # the same bytes in writes of 100 bytes: leftoverBytes carries partial packets (:25-41); same delivery
    cases.append(case("padded_packet_ends_video_pes_100_byte_writes", "ts.js:25-41", p,
                      [allp[i:i + 100] for i in range(0, len(allp), 100)], [VIDEO],
                      {VIDEO: [(90000, [(0, 18), (1, 4), (2, 105)])]}))

    # B -- a PES with a length (audio style): packet_length 208 = 3 + header_length 5 + 200 payload bytes
    #   -> totalLength = 208 - 5 - 3 = 200 (:119-122).  P0 adds 170 bytes, P1 184: currentLength 354 >= 200 -> complete (:201),
    #   and the reference delivers ALL 354 bytes it collected (no trimming: packetAddData pushes whole subarrays, :191-197).
    p = [packet(0, 0x101, 1, 1, head=pes_header(AUDIO, 180000, packet_length=208)), packet(1, 0x101, 0, 1), packet(2, 0x101, 0, 1)]
    #   P2: after the completion totalLength is 0 again (:207), so its 184 bytes just accumulate and are never delivered here
    cases.append(case("pes_length_reached", "ts.js:119-122, 191-202", p, [b"".join(p)], [AUDIO],
                      {AUDIO: [(180000, [(0, 18), (1, 4)])]}))

    # C -- a PID re-bound to another stream id.
    #   P0, P1: video PES on PID 0x100 (170 + 184 bytes).  P2: payload start on the same PID while video has currentLength
    #   -> the open video chunk is completed first (:62-70): write(90000, 354 bytes).  P2's PES header says stream id 0xC0:
    #   pidsToStreamIds[0x100] = 0xC0 (:81-83), its 170 payload bytes go to AUDIO.  P3 (continuation, same PID) follows the
    #   new binding (:60, :128): 184 bytes to AUDIO.  P4 padded -> audio chunk complete: write(180000, 170 + 184 + 83 bytes).
    p = [packet(0, PID, 1, 1, head=pes_header(VIDEO, 90000)), packet(1, PID, 0, 1),
         packet(2, PID, 1, 1, head=pes_header(AUDIO, 180000)), packet(3, PID, 0, 1), packet(4, PID, 0, 3, af_len=100)]
    cases.append(case("pid_rebound_to_another_stream", "ts.js:60-70, 81-83, 128", p, [b"".join(p)], [VIDEO, AUDIO],
                      {VIDEO: [(90000, [(0, 18), (1, 4)])], AUDIO: [(180000, [(2, 18), (3, 4), (4, 105)])]}))
    #   with only VIDEO connected the audio PES is parsed (the PID is re-bound all the same) and dropped (:92, :131)
    cases.append(case("pid_rebound_audio_not_connected", "ts.js:81-83, 92, 131", p, [b"".join(p)], [VIDEO],
                      {VIDEO: [(90000, [(0, 18), (1, 4)])]}))

    # D -- resync.  Three garbage bytes, then six packets.  parsePacket reads 0x12 != 0x47 (:45) -> resync (:155-189): at least
    #   6 packets of data from the NEXT byte on? 3 + 6 * 188 - 1 = 1130 >= 1128, yes (:157); first 0x47 in the next 187 bytes
    #   with four more 188 apart: index 3 (:165-176); parsing resumes behind it (:179).  P0 PES (170), P1..P4 184 each, P5 padded (83).
    p = [packet(0, PID, 1, 1, head=pes_header(VIDEO, 90000))] + [packet(k, PID, 0, 1) for k in range(1, 5)] + [packet(5, PID, 0, 3, af_len=100)]
    cases.append(case("resync_after_garbage", "ts.js:45-50, 155-189", p, [b"\x12\x34\x56" + b"".join(p)], [VIDEO],
                      {VIDEO: [(90000, [(0, 18), (1, 4), (2, 4), (3, 4), (4, 4), (5, 105)])]}))
    #   only five packets behind the garbage: has(188 * 6) fails (:157), resync returns false, nothing is parsed; the bad byte was
    #   consumed by read(8) (:45), everything else stays as leftover: 3 + 5 * 188 - 1 = 942 bytes (:36-40)
    cases.append(case("resync_needs_six_packets", "ts.js:157-159, 36-40", p[:5], [b"\x12\x34\x56" + b"".join(p[:5])], [VIDEO],
                      {VIDEO: []}, leftover=942))

    # E -- payload start WITHOUT a start code on a bound PID (a section, a broken header).  P0 video PES (170).  P1 pusi = 1 but
    #   its payload begins FF FF FF: streamId is still 0xE0 from the table (:60), payloadStart && streamId with currentLength > 0
    #   completes the open chunk (:62-70): write(90000, 170 bytes); no PES header (:79) so pts / totalLength stay; its 184 bytes are
    #   added (:128-136).  P2 padded: complete -> write(90000, 184 + 83 bytes) -- same pts, pi.pts was not touched.
    p = [packet(0, PID, 1, 1, head=pes_header(VIDEO, 90000)), packet(1, PID, 1, 1, head=b"\xff\xff\xff"), packet(2, PID, 0, 3, af_len=100)]
    cases.append(case("payload_start_without_start_code", "ts.js:60-70, 79, 128-146", p, [b"".join(p)], [VIDEO],
                      {VIDEO: [(90000, [(0, 18)]), (90000, [(1, 4), (2, 105)])]}))

    # F -- packets that carry nothing for us: adaptation field only (afc = 2, :72), a PID never bound (:128), then a video PES.
    p = [packet(0, 0x200, 0, 2, af_len=183), packet(1, 0x300, 0, 1), packet(2, PID, 1, 1, head=pes_header(VIDEO, 270000)),
         packet(3, PID, 0, 3, af_len=0)]
    #   P3: adaptation field of length 0 (1 byte): payload 5..187, and it counts as padding (:143) -> write(270000, 170 + 183)
    cases.append(case("adaptation_only_and_unbound_pid", "ts.js:72, 128, 143", p, [b"".join(p)], [VIDEO],
                      {VIDEO: [(270000, [(2, 18), (3, 5)])]}))

    # G -- the end of the BUFFER counts as a start code (buffer.js:141-150: i >= byteLength).  P0, P1 video (170 + 184).  P2:
    #   payload start, adaptation field of 1 + 183 bytes -> the "payload" begins at byte 188, i.e. at the end of this write():
    #   nextBytesAreStartCode() is true there (:79), the stream id is read past the end = 0 and the PID is bound to 0 (:81-83).
    #   First, though, the open chunk is completed (:62-70): write(90000, 354).  Second write: P3 continuation on the PID: stream
    #   id 0 is falsy (:128) -> dropped.
    p = [packet(0, PID, 1, 1, head=pes_header(VIDEO, 90000)), packet(1, PID, 0, 1), packet(2, PID, 1, 3, af_len=183), packet(3, PID, 0, 1)]
    cases.append(case("end_of_buffer_is_a_start_code", "buffer.js:141-150, ts.js:79-83, 128", p, [b"".join(p[:3]), p[3]], [VIDEO],
                      {VIDEO: [(90000, [(0, 18), (1, 4)])]}))
    #   the same four packets in ONE write: behind P2's adaptation field lie P3's bytes 47 01 00 ..., no start code, the PID
    #   stays with video; P2 adds nothing (its payload starts at 188 = end), P3 adds 184 bytes; nothing completes them here
    #   (P2 completed the first chunk, :62-70).
    cases.append(case("same_packets_in_one_write", "buffer.js:141-150", p, [b"".join(p)], [VIDEO],
                      {VIDEO: [(90000, [(0, 18), (1, 4)])]}))
    return cases


def main():
    out = os.path.join(ROOT, "tests", "fixtures", "ts_cases.json")
    with open(out, "w") as f:
        json.dump({"generator": "tools/make_ts_cases.py (hand-derived from src/ts.js; see the comments there)", "cases": build()}, f, indent=1)
    print(out)


if __name__ == "__main__":
    main()
