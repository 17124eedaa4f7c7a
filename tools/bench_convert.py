#!/usr/bin/env python3
"""Device-resident rate of the colour converters (rtjgpu_convert_device) on the configs[1] geometry: 4096 pictures of
720x576, one JSON line per converter with the HBM traffic it stands for (1.5 bytes read and bpp written per pixel)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gmerlin_avdecoder_b200 as g  # noqa: E402

NAMES = ["yuv420rgb32", "yuv420bgr32", "yuv420rgb24", "yuv420bgr24", "yuv420rgb16", "yuv420rgb8", "yuv422rgb24"]
w, h = 720, 576
F = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
peak = 6453.4
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
ctx = g.BatchContext(0)
for kind, name in enumerate(NAMES):
    bpp = g.CONV_BPP[kind]
    src_fb = w * h * 2 if kind == 6 else w * h * 3 // 2
    src = torch.randint(16, 236, (F, src_fb), dtype=torch.uint8, device="cuda")
    out = torch.empty((F, w * h * bpp), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    run = lambda: ctx.convert_device(kind, src.data_ptr(), src_fb, F, w, h, out.data_ptr(), w * bpp, w * h * bpp, 255, st)
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps = 10
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    read = F * (w * h if kind == 5 else src_fb)
    algo = read + F * w * h * bpp
    print(json.dumps({"converter": name, "frames": F, "ms": ms, "frames_per_s": F / ms * 1e3, "algorithmic_GBps": algo / ms / 1e6,
                      "frac_of_measured_peak": algo / ms / 1e6 / peak}), flush=True)
ctx.close()
