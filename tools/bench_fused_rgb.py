#!/usr/bin/env python3
"""Device-resident decode -> packed pixels on the configs[1] batch, two ways: rtjgpu_decode_device then
rtjgpu_convert_device (planes written to HBM and read back), and rtjgpu_decode_device_rgb (the converter fused into
K2).  One JSON line per converter kind."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gmerlin_avdecoder_b200 as g  # noqa: E402
from gmerlin_avdecoder_b200 import device as D  # noqa: E402
from oracle import oracle as O  # noqa: E402

w, h, q = 720, 576, 128
F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = 10
peak = 6453.4
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
clip = O.make_clip(w, h, q, noise_y=2)
stream, offsets = O.encode_clip(clip, F, threads=min(os.cpu_count() or 1, 64))
desc, _ = g.plan(stream, offsets)
b = D.upload(stream, desc, w, h)
ctx = g.BatchContext(0)
st = torch.cuda.current_stream().cuda_stream


def timed(fn):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for kind, name in ((g.capi.CONV_RGB32, "rgb32"), (g.capi.CONV_BGR24, "bgr24"), (g.capi.CONV_RGB16, "rgb16")):
    bpp = g.CONV_BPP[kind]
    out = torch.empty((F, w * h * bpp), dtype=torch.uint8, device="cuda")
    fused = torch.empty_like(out)

    def two_pass():
        D.decode(ctx, b)
        ctx.convert_device(kind, b.out.data_ptr(), w * h * 3 // 2, F, w, h, out.data_ptr(), w * bpp, w * h * bpp, 255, st)

    def one_pass():
        ctx.decode_device_rgb(b.stream.data_ptr(), b.desc.data_ptr(), F, w, h, kind, fused.data_ptr(), w * bpp, w * h * bpp, 255,
                              None, None, st)

    ms2, ms1 = timed(two_pass), timed(one_pass)
    assert torch.equal(out, fused)
    algo = b.payload_bytes + F * w * h * bpp                      # payload read + packed pixels written
    print(json.dumps({"kind": name, "frames": F, "decode_then_convert_ms": ms2, "fused_ms": ms1, "speedup": ms2 / ms1,
                      "fused_frames_per_s": F / ms1 * 1e3, "fused_algorithmic_GBps": algo / ms1 / 1e6,
                      "fused_frac_of_measured_peak": algo / ms1 / 1e6 / peak}), flush=True)
    del out, fused
ctx.close()
