"""Helpers for the -m gpu tests: everything goes through the C ABI."""
import numpy as np
import torch

import gmerlin_avdecoder_b200 as g
from gmerlin_avdecoder_b200 import device as D


def gpu_decode(ctx, stream, offsets, w, h, carry=None, state=None, fmt=0):
    """rtjgpu_plan + rtjgpu_decode_device; returns frames [F, frame bytes] as numpy."""
    desc, st = g.plan(stream, offsets, state)
    b = D.upload(stream, desc, w, h, fmt=fmt)
    ct = None if carry is None else torch.from_numpy(np.ascontiguousarray(carry)).cuda()
    # poison the output so that an unwritten byte cannot pass by accident
    b.out.fill_(0xCD)
    D.decode(ctx, b, ct)
    torch.cuda.synchronize()
    return b.out.cpu().numpy(), st


def first_diff(a, b, w, h):
    idx = np.argwhere(a != b)
    if len(idx) == 0:
        return "equal"
    f, o = idx[0]
    plane = "Y" if o < w * h else ("U" if o < w * h * 5 // 4 else "V")
    return f"{len(idx)} bytes differ; first at frame {f}, plane {plane}, offset {o}: got {a[f, o]} want {b[f, o]}"
