"""Test-side writer of NuppelVideo files (the reference has a demuxer, lib/demux_nuv.c, but no muxer).

Layout as lib/demux_nuv.c reads it: 72-byte file header (:58-105), then frames with 12-byte headers
(:246-260): type, comptype / subtype, keyframe, filters, timecode (u32le), size (u32le & 0xffffff).
The first frame is 'D' / 'R' with the 128 raw RTjpeg tables (:157-176).  RTjpeg video frames carry the
block stream WITHOUT the 12-byte RTjpeg_frameheader."""
import struct

import numpy as np


def frame(ftype, comptype, keyframe, timecode, payload=b""):
    return struct.pack("<ccBBII", ftype.encode(), comptype.encode(), keyframe, 0, timecode, len(payload)) + bytes(payload)


def write_nuv(w, h, fps, raw_tables, video_payloads, keyframes=None, audio_every=0, seek_every=0, extra=()):
    """video_payloads: block streams (bytes / uint8 arrays), or ('L',) / ('N',) / ('0', bytes) for other frame types."""
    out = bytearray()
    out += b"NuppelVideo\0" + b"0.07\0" + b"\0\0\0"
    out += struct.pack("<II", w, h) + struct.pack("<II", w, h)
    out += b"P" + b"\0\0\0"
    out += struct.pack("<dd", 1.0, fps)
    n_audio = (len(video_payloads) // audio_every) if audio_every else 0
    out += struct.pack("<II", len(video_payloads), n_audio)
    out += struct.pack("<II", 0, 30)
    assert len(out) == 72
    out += frame("D", "R", 0, 0, np.asarray(raw_tables, dtype="<u4").tobytes())
    for i, pl in enumerate(video_payloads):
        tc = int(round(i * 1000.0 / fps))
        if seek_every and i % seek_every == 0:
            out += b"S" + b"RTjjjjjjjjj"                      # seek point: twelve bytes in all, size field meaningless
        key = 1 if (keyframes is None or keyframes[i]) else 0
        if isinstance(pl, tuple):
            out += frame("V", pl[0], key, tc, pl[1] if len(pl) > 1 else b"")
        else:
            out += frame("V", "1", key, tc, pl if isinstance(pl, (bytes, bytearray)) else np.asarray(pl, dtype=np.uint8).tobytes())
        if audio_every and (i + 1) % audio_every == 0:
            out += frame("A", "0", 0, tc, bytes(4096))
    for e in extra:
        out += e
    return np.frombuffer(bytes(out), dtype=np.uint8).copy()
