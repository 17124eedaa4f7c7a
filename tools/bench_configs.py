#!/usr/bin/env python3
"""Device-resident stage times of the other BASELINE.json configs (parity cases, not the bench line):
config 3 (720x576 inter, GOP 30), config 4 (1920x1088 Q255 dense).  Prints one JSON line per case."""
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


def run(name, w, h, Q, F, steps=10, **kw):
    clip = O.make_clip(w, h, Q, **kw)
    stream, offsets = O.encode_clip(clip, F, threads=min(os.cpu_count() or 1, 64))
    desc, _ = g.plan(stream, offsets)
    ctx = g.BatchContext(0)
    b = D.upload(stream, desc, w, h, device=0)
    ctx.enable_timing(True)
    for _ in range(3):
        D.decode(ctx, b)
        torch.cuda.synchronize()      # the library arranges a batch by what the batch before held (skipped blocks, raw prefixes)
    info = ctx.batch_info()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        D.decode(ctx, b)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    st = [ctx.timing_at(i) for i in range(steps)]
    fsz = w * h * 3 // 2
    algo = b.payload_bytes + F * fsz
    print(json.dumps({
        "case": name, "frames": F, "w": w, "h": h, "quality": Q, "payload_bytes_per_frame": b.payload_bytes / F,
        "skipped_blocks_frac": info.skipped_blocks / (F * (w // 16) * (h // 16) * 6),
        "ms_per_step": ms, "frames_per_s": F / (ms * 1e-3), "algorithmic_GBps": algo / (ms * 1e-3) / 1e9,
        "stage_ms": {"scan": sum(t.scan_ms for t in st) / steps, "resolve": sum(t.resolve_ms for t in st) / steps,
                     "idct": sum(t.idct_ms for t in st) / steps}}), flush=True)
    ctx.close()


if __name__ == "__main__":
    run("config3 720x576 inter GOP30 lm=cm=1", 720, 576, 128, 2048, key_rate=29, lm=1, cm=1, noise_y=2)
    run("config3 720x576 inter GOP30 lm=cm=4", 720, 576, 128, 2048, key_rate=29, lm=4, cm=4, noise_y=2)
    run("config2 720x576 intra Q32", 720, 576, 32, 2048, noise_y=2)
    run("config2 720x576 intra Q255", 720, 576, 255, 1024, noise_y=2)
    run("config4 1920x1088 Q255 dense", 1920, 1088, 255, 128, noise_y=60, noise_c=20, steps=5)
