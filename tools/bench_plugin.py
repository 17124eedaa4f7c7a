#!/usr/bin/env python3
"""Frames/s through the 'RTJ0' plugin (the bgav video-decoder contract), driven by the test host stub,
with and without the look-ahead batcher.  Usage: bench_plugin.py [frames]"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gmerlin_avdecoder_b200 as g  # noqa: E402
from oracle import oracle as O  # noqa: E402

FOURCC = (ord('R') << 24) | (ord('T') << 16) | (ord('J') << 8) | ord('0')


def host_lib():
    build = os.path.join(ROOT, "tests", "_build")
    os.makedirs(build, exist_ok=True)
    so = os.path.join(build, "libbgav_host_stub.so")
    subprocess.check_call(["gcc", "-O1", "-fPIC", "-shared", "-o", so, os.path.join(ROOT, "tests", "bgav_host_stub.c")])
    H = C.CDLL(so, mode=C.RTLD_GLOBAL)
    P = C.CDLL(g.PLUGIN_PATH, mode=C.RTLD_GLOBAL)
    vp = C.c_void_p
    H.stub_find_decoder.restype = vp; H.stub_find_decoder.argtypes = [C.c_uint32]
    H.stub_stream_create.restype = vp; H.stub_stream_create.argtypes = [C.c_int, C.c_int]
    H.stub_stream_set_packets.argtypes = [vp, vp, vp, vp, C.c_int]
    H.stub_init.argtypes = [vp, vp]; H.stub_close.argtypes = [vp, vp]
    H.stub_decode.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.POINTER(C.c_int64)]
    H.stub_stream_destroy.argtypes = [vp]
    P.bgav_init_video_decoders_rtjpeg.restype = None
    P.bgav_init_video_decoders_rtjpeg()
    return H


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    w, h = 720, 576
    clip = O.make_clip(w, h, 128, key_rate=-1, noise_y=2, noise_c=0, seed=1)
    s, o = O.encode_clip(clip, F, threads=min(os.cpu_count() or 1, 64))
    s = np.ascontiguousarray(s)
    sizes = O.packet_sizes(s, o).astype(np.uint32)
    offs = np.ascontiguousarray(o[:-1], dtype=np.uint64)
    H = host_lib()
    dec = H.stub_find_decoder(FOURCC)
    Y = np.zeros((h, w), np.uint8); U = np.zeros((h // 2, w // 2), np.uint8); V = np.zeros_like(U)
    for K in (1, 8, 32, 128):
        os.environ["RTJPEG_B200_LOOKAHEAD"] = str(K)
        st = H.stub_stream_create(w, h)
        H.stub_stream_set_packets(st, s.ctypes.data, offs.ctypes.data, sizes.ctypes.data, F)
        assert H.stub_init(dec, st) == 1
        for _ in range(min(F, 2 * K)):                          # warm-up: allocations, first launches
            H.stub_decode(dec, st, Y.ctypes.data, U.ctypes.data, V.ctypes.data, w, w // 2, None)
        n = 0
        t0 = time.perf_counter()
        while H.stub_decode(dec, st, Y.ctypes.data, U.ctypes.data, V.ctypes.data, w, w // 2, None) == 1:
            n += 1
        dt = time.perf_counter() - t0
        print(json.dumps({"case": "plugin decode(), 720x576 Q128 intra, frame copied to the caller", "lookahead": K,
                          "frames": n, "frames_per_s": n / dt if dt > 0 else None}), flush=True)
        H.stub_close(dec, st)
        H.stub_stream_destroy(st)


if __name__ == "__main__":
    main()
