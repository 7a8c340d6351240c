"""MPEG-TS demuxer -- host-side mirror of ``JSMpeg.Demuxer.TS`` (reference src/ts.js:1-226).

This is the step *before* the hot path.  It stays on the host,
as it does in the reference; it exists here so tests and bench.py can feed MPEG-TS clips through
the same ``destination.write(pts, buffers)`` boundary (src/ts.js:205-206) the reference uses.

Semantics kept from the reference:
  * 188-byte packets, sync byte 0x47, resync needs 5 sync bytes 188 apart (src/ts.js:155-189)
  * adaptation field skip (src/ts.js:73-77), PES header parse + 33-bit PTS / 90 kHz (src/ts.js:79-116)
  * a PES packet is complete on the next payload_unit_start of the same stream (src/ts.js:65-70),
    when its announced length is reached (src/ts.js:201), or -- video, length 0 -- when a
    non-start TS packet carries an adaptation field, i.e. was padded (src/ts.js:143-146)
"""
from __future__ import annotations


class _PesInfo:
    __slots__ = ("destination", "current_length", "total_length", "pts", "pts_ticks", "buffers")

    def __init__(self, destination):
        self.destination = destination
        self.current_length = 0
        self.total_length = 0
        self.pts = 0.0
        self.pts_ticks = 0  # the same in 90 kHz ticks (what the device demuxer reports)
        self.buffers = []


class TS:
    """``demuxer = TS(); demuxer.connect(TS.STREAM_VIDEO_1, decoder); demuxer.write(bytes)``"""

    STREAM_AUDIO_1 = 0xC0
    STREAM_VIDEO_1 = 0xE0

    def __init__(self, options=None):
        self.leftover = b""
        self.guess_video_frame_end = True
        self.pids_to_stream_ids = {}
        self.pes_packet_info = {}
        self.start_time = 0.0
        self.current_time = 0.0

    def connect(self, stream_id, destination):
        self.pes_packet_info[stream_id] = _PesInfo(destination)

    # -- src/ts.js:25-41
    def write(self, buffer):
        data = bytes(buffer)
        if self.leftover:
            data = self.leftover + data
        self._bytes = memoryview(data)
        self._pos = 0  # byte position (the reference tracks bits; every access here is byte aligned)
        n = len(data)
        while n - self._pos >= 188 and self._parse_packet():
            pass
        self.leftover = data[self._pos:] if self._pos < n else b""

    def _rd(self, i):
        """bits.read past the end of the buffer: `undefined & mask` = 0 (src/buffer.js:152-170)"""
        return self._bytes[i] if i < len(self._bytes) else 0

    # -- src/ts.js:43-153.  Positions are byte positions in the whole buffer (the reference reads through
    # its bit buffer, which knows nothing of packet boundaries: a header field that lies behind the packet's
    # end is read from the next packet's bytes, one behind the buffer's end reads as 0).
    def _parse_packet(self):
        b, rd = self._bytes, self._rd
        p = self._pos
        if b[p] != 0x47:
            self._pos = p + 1
            if not self._resync():
                return False
            p = self._pos
        else:
            p += 1
        end = p + 187
        payload_start = (rd(p) >> 6) & 1
        pid = ((rd(p) & 0x1F) << 8) | rd(p + 1)
        adaptation_field = (rd(p + 2) >> 4) & 3
        p += 3

        stream_id = self.pids_to_stream_ids.get(pid)
        if payload_start and stream_id:
            pi = self.pes_packet_info.get(stream_id)
            if pi is not None and pi.current_length:
                self._packet_complete(pi)

        if adaptation_field & 1:
            if adaptation_field & 2:
                p += 1 + rd(p)
            # nextBytesAreStartCode (src/buffer.js:141-150): true at the very end of the BUFFER as well
            if payload_start and (p >= len(b) or (rd(p) == 0 and p + 2 < len(b) and b[p + 1] == 0 and b[p + 2] == 1)):
                stream_id = rd(p + 3)
                self.pids_to_stream_ids[pid] = stream_id
                packet_length = (rd(p + 4) << 8) | rd(p + 5)
                pts_dts_flag = rd(p + 7) >> 6
                header_length = rd(p + 8)
                payload_begin = p + 9 + header_length
                pi = self.pes_packet_info.get(stream_id)
                if pi is not None:
                    pts = 0.0
                    pts_ticks = 0
                    if pts_dts_flag & 2:
                        q = p + 9
                        p32_30 = (rd(q) >> 1) & 7
                        p29_15 = (rd(q + 1) << 7) | (rd(q + 2) >> 1)
                        p14_0 = (rd(q + 3) << 7) | (rd(q + 4) >> 1)
                        pts_ticks = p32_30 * 1073741824 + p29_15 * 32768 + p14_0
                        pts = pts_ticks / 90000.0
                        self.current_time = pts
                        if self.start_time == -1:
                            self.start_time = pts
                    payload_length = packet_length - header_length - 3 if packet_length else 0
                    pi.total_length = payload_length
                    pi.current_length = 0
                    pi.pts = pts
                    pi.pts_ticks = pts_ticks
                p = payload_begin
            if stream_id:
                pi = self.pes_packet_info.get(stream_id)
                if pi is not None:
                    start = p
                    pi.buffers.append(bytes(b[start:end]) if start < end else b"")  # subarray(start, end)
                    pi.current_length += end - start
                    complete = pi.total_length != 0 and pi.current_length >= pi.total_length
                    has_padding = (not payload_start) and (adaptation_field & 2)
                    if complete or (self.guess_video_frame_end and has_padding):
                        self._packet_complete(pi)
        self._pos = end
        return True

    # -- src/ts.js:155-189
    def _resync(self):
        b = self._bytes
        byte_index = self._pos
        if len(b) - byte_index < 188 * 6:
            return False
        for i in range(187):
            if b[byte_index + i] == 0x47:
                if all(b[byte_index + i + 188 * j] == 0x47 for j in range(1, 5)):
                    self._pos = byte_index + i + 1
                    return True
        self._pos = byte_index + 187
        return False

    # -- src/ts.js:205-210
    def _packet_complete(self, pi):
        pi.destination.write(pi.pts, pi.buffers)
        pi.total_length = 0
        pi.current_length = 0
        pi.buffers = []

    def flush(self):
        """Deliver a trailing, still-open PES packet (the reference only does this implicitly when
        the next PES header arrives; whole-file tools call it once at end of input)."""
        for pi in self.pes_packet_info.values():
            if pi.current_length:
                self._packet_complete(pi)


class ESCollector:
    """A ``destination`` that just records what the demuxer delivers (pts, concatenated payload)."""

    def __init__(self):
        self.packets = []

    def write(self, pts, buffers):
        self.packets.append((pts, b"".join(buffers)))

    @property
    def es(self):
        return b"".join(p for _, p in self.packets)


def demux_video_es(ts_bytes):
    """Whole MPEG-TS clip -> list of (pts, payload bytes) video PES payloads, in order."""
    demux = TS()
    col = ESCollector()
    demux.connect(TS.STREAM_VIDEO_1, col)
    demux.write(ts_bytes)
    demux.flush()
    return col.packets
