"""Device-side plumbing for the device-resident entry point (torch owns the memory).

PyTorch is used for exactly three things here: allocating HBM, host<->device
copies, and handing the current CUDA stream to the library.  No torch op touches
the pixel data.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import capi


@dataclass
class DeviceBatch:
    """One batch resident in HBM: packets, descriptors and the output frames."""
    stream: torch.Tensor      # uint8, packets back to back + STREAM_SLACK_BYTES of 0x7F
    desc: torch.Tensor        # uint8 view of rtjgpu_frame_desc[F]
    out: torch.Tensor         # uint8 [F, w*h*3/2]
    F: int
    w: int
    h: int
    payload_bytes: int        # sum over frames of (packet length - 12)
    fmt: int = 0              # RTJ_YUV420 / RTJ_YUV422 / RTJ_RGB8 (grey)

    @property
    def frame_bytes(self) -> int:
        return capi.frame_bytes(self.fmt, self.w, self.h)


def upload(stream: np.ndarray, desc: np.ndarray, w: int, h: int, device: int | str = 0,
           out: torch.Tensor | None = None, fmt: int = 0) -> DeviceBatch:
    dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
    F = len(desc)
    host = np.full(stream.size + capi.STREAM_SLACK_BYTES, 0x7F, dtype=np.uint8)
    host[:stream.size] = stream
    d_stream = torch.from_numpy(host).to(dev)
    d_desc = torch.from_numpy(np.ascontiguousarray(desc).view(np.uint8).copy()).to(dev)
    if out is None:
        out = torch.empty((F, capi.frame_bytes(fmt, w, h)), dtype=torch.uint8, device=dev)
    payload = int(desc["length"].astype(np.int64).sum()) - 12 * F
    return DeviceBatch(d_stream, d_desc, out, F, w, h, payload, fmt)


def decode(ctx: capi.BatchContext, b: DeviceBatch, carry: torch.Tensor | None = None,
           stream: torch.cuda.Stream | None = None) -> None:
    """Launch K1/K3/K2 for the batch on `stream` (default: torch's current stream)."""
    st = stream if stream is not None else torch.cuda.current_stream(b.out.device)
    ctx.set_format(b.fmt)
    ctx.decode_device(b.stream.data_ptr(), b.desc.data_ptr(), b.F, b.w, b.h, b.out.data_ptr(),
                      None if carry is None else carry.data_ptr(), st.cuda_stream)
