#!/usr/bin/env python3
"""Development aid (GPU box): how long does the one-lane-per-frame scan (rtj_scan_lane_kernel) take on a configs[1]-shaped
batch -- alone, and while another context keeps the device busy with K2-heavy work on a second stream?"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gmerlin_avdecoder_b200 as g  # noqa: E402
from gmerlin_avdecoder_b200 import capi  # noqa: E402
from gmerlin_avdecoder_b200 import device as D  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    w, h, q, F = 720, 576, 128, 4096
    clip = O.make_clip(w, h, q, noise_y=2)
    stream, offsets = O.encode_clip(clip, F, threads=min(os.cpu_count() or 1, 64))
    desc, _ = g.plan(stream, offsets)
    ba = D.upload(stream, desc, w, h, device=0)
    bb = D.upload(stream, desc, w, h, device=0)
    ca, cb = g.BatchContext(0), g.BatchContext(0)
    ca.set_pipeline(capi.PIPELINE_SERIAL)
    cb.set_pipeline(capi.PIPELINE_SERIAL)
    cb.set_scan_mode(capi.SCAN_LANE)
    cb.enable_timing(True)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        D.decode(ca, ba, stream=sa)
        D.decode(cb, bb, stream=sb)
    torch.cuda.synchronize()
    # alone
    for _ in range(5):
        D.decode(cb, bb, stream=sb)
    torch.cuda.synchronize()
    alone = [cb.timing_at(i) for i in range(5)]
    print(json.dumps({"case": "lane scan alone", "scan_ms": [round(t.scan_ms, 4) for t in alone],
                      "idct_ms": [round(t.idct_ms, 4) for t in alone]}), flush=True)
    # beside another context's full decode loop (chunk scan + K2, serial stages)
    for _ in range(12):
        D.decode(ca, ba, stream=sa)
    for _ in range(5):
        D.decode(cb, bb, stream=sb)
    for _ in range(12):
        D.decode(ca, ba, stream=sa)
    torch.cuda.synchronize()
    busy = [cb.timing_at(i) for i in range(5)]
    print(json.dumps({"case": "lane scan beside another context's decodes", "scan_ms": [round(t.scan_ms, 4) for t in busy],
                      "idct_ms": [round(t.idct_ms, 4) for t in busy]}), flush=True)
    assert torch.equal(ba.out, bb.out)
    ca.close(); cb.close()


if __name__ == "__main__":
    main()
