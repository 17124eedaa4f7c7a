#!/usr/bin/env python3
"""Development aid (GPU box): configs[1]-shaped batch under the arrangements of rtjgpu_set_pipeline --
serial stages, and pipelined with several slice sizes / first-slice sizes / K1 stream priorities.
One JSON line per arrangement; every arrangement's frames are compared with the serial ones."""
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
    w, h, q, F = 720, 576, 128, int(os.environ.get("PIPE_FRAMES", "4096"))
    kw = {}
    if os.environ.get("PIPE_INTER"):
        lm = int(os.environ["PIPE_INTER"])
        kw = dict(key_rate=29, lm=lm, cm=lm)
    steps = int(os.environ.get("PIPE_STEPS", "20"))
    clip = O.make_clip(w, h, q, noise_y=2, **kw)
    stream, offsets = O.encode_clip(clip, F, threads=min(os.cpu_count() or 1, 64))
    desc, _ = g.plan(stream, offsets)
    b = D.upload(stream, desc, w, h, device=0)
    ref = None
    cases = [("serial", dict(RTJPEG_B200_PIPELINE="1")),
             ]
    for sl in (576, 1184, 2048):
        for prio in (0, 1):
            cases.append((f"frame slices {sl} prio={prio}", dict(RTJPEG_B200_PIPELINE="2", RTJPEG_B200_SLICE=str(sl), RTJPEG_B200_SCAN_PRIO=str(prio))))
    for name, env in cases:
        for k in ("RTJPEG_B200_PIPELINE", "RTJPEG_B200_SLICE", "RTJPEG_B200_SLICE0", "RTJPEG_B200_SCAN_PRIO",
                  "RTJPEG_B200_WALK_MIN", "RTJPEG_B200_WALK_SLICES", "RTJPEG_B200_WALK_THREADS", "RTJPEG_B200_WALK_EXCLUSIVE"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ctx = g.BatchContext(0)
        ctx.enable_timing(True)
        b.out.fill_(0xCD)
        for _ in range(3):
            D.decode(ctx, b)
        torch.cuda.synchronize()
        assert ctx.batch_info().bad_frames == 0
        if ref is None:
            ref = b.out.clone()
        same = bool(torch.equal(ref, b.out))
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            D.decode(ctx, b)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / steps
        st = [ctx.timing_at(i) for i in range(steps)]
        print(json.dumps({"case": name, "ms_per_step": round(ms, 4), "frames_per_s": round(F / (ms * 1e-3)),
                          "equal_to_serial": same,
                          "scan_ms": round(sum(t.scan_ms for t in st) / steps, 4),
                          "idct_ms": round(sum(t.idct_ms for t in st) / steps, 4),
                          "total_ms": round(sum(t.total_ms for t in st) / steps, 4)}), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
