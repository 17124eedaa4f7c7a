#!/usr/bin/env python3
"""Device-resident rate of the encoder (rtjgpu_encode_device) on the configs[1] geometry: pictures decoded from the bench
stream are encoded again, intra-only and with GOP 30; one JSON line per case."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gmerlin_avdecoder_b200 as g  # noqa: E402
from gmerlin_avdecoder_b200 import device as D  # noqa: E402
from oracle import oracle as O  # noqa: E402

w, h, Q = 720, 576, 128
F = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
stream, offsets = O.encode_clip(O.make_clip(w, h, Q, noise_y=2), F, threads=min(os.cpu_count() or 1, 64))
ctx = g.BatchContext(0)
desc, _ = g.plan(stream, offsets)
b = D.upload(stream, desc, w, h)
D.decode(ctx, b)
torch.cuda.synchronize()
cap = F * (12 + (w // 8) * (h // 8) * 2 * 64 + 16)
d_stream = torch.empty(min(cap, 4 << 30), dtype=torch.uint8, device="cuda")
d_off = torch.zeros(F + 1, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for name, kr, lm in (("intra", 0, 0), ("inter GOP 30, lm = cm = 2", 29, 2)):
    def run():
        ctx.encoder_config(Q, kr, lm, lm)
        ctx.encode_device(b.out.data_ptr(), F, w, h, d_stream.data_ptr(), d_stream.numel(), d_off.data_ptr(), st)
    for _ in range(2):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 5
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    nbytes, overflow = ctx.encode_info()
    print(json.dumps({"case": "encode 720x576 Q128 " + name, "frames": F, "ms": ms, "frames_per_s": F / ms * 1e3,
                      "packet_bytes_per_frame": nbytes / F, "overflow": overflow}), flush=True)
ctx.close()
