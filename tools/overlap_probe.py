#!/usr/bin/env python3
"""Probe: do K1 and K2 of different sub-batches overlap usefully?  Decodes the configs[1] batch as N sub-batches,
(a) one after the other on one stream, (b) round-robin on two streams, and prints ms per 4096 frames."""
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

w, h, Q, F = 720, 576, 128, 4096
clip = O.make_clip(w, h, Q, noise_y=2)
stream, offsets = O.encode_clip(clip, F, threads=min(os.cpu_count() or 1, 64))
for nsub in (1, 2, 4, 8):
    per = F // nsub
    subs = []
    for k in range(nsub):
        o = offsets[k * per:(k + 1) * per + 1]
        s = stream[int(o[0]):int(o[-1])]
        desc, _ = g.plan(s, o - o[0])
        subs.append(D.upload(s, desc, w, h, device=0))
    for nstreams in (1, 2):
        ctxs = [g.BatchContext(0) for _ in range(nstreams)]
        streams = [torch.cuda.Stream() for _ in range(nstreams)]
        def run():
            for k, b in enumerate(subs):
                D.decode(ctxs[k % nstreams], b, stream=streams[k % nstreams])
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in streams:
            st.wait_event(e0)
        steps = 10
        for _ in range(steps):
            run()
        for st in streams:
            e = torch.cuda.Event()
            e.record(st)
            torch.cuda.current_stream().wait_event(e)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"sub_batches": nsub, "streams": nstreams, "ms_per_4096": e0.elapsed_time(e1) / steps}), flush=True)
        for c in ctxs:
            c.close()
