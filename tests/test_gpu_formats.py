"""The other two branches of RTjpeg_decompress (lib/RTjpeg.c:3580-3585) on the CUDA path: YUV422
(RTjpeg_decompressYUV422 :2639-2686) and 8-bit grey (RTjpeg_decompress8 :2751-2772), bit-exact
against the reference through the C ABI."""
import numpy as np
import pytest

import gmerlin_avdecoder_b200 as g
from gmerlin_avdecoder_b200 import capi
from oracle import oracle as O
from gpu_util import gpu_decode
from streams import golden, reference_frames_fmt

pytestmark = pytest.mark.gpu

FMT_CLIPS = ["yuv422_inter_64x48_q200_gop4", "yuv422_intra_96x32_q128", "grey_inter_64x48_q255_gop4",
             "grey_intra_96x32_q64"]


@pytest.fixture(scope="module", params=["auto", "chunk", "lane", "warp", "sync", "runs", "segment"])
def ctx(request):
    c = g.BatchContext(0)
    c.set_scan_mode({"auto": capi.SCAN_AUTO, "chunk": capi.SCAN_CHUNK, "lane": capi.SCAN_LANE,
                     "warp": capi.SCAN_WARP, "sync": capi.SCAN_SYNC, "runs": capi.SCAN_AUTO, "segment": capi.SCAN_SEGMENT}[request.param])
    if request.param == "runs":
        c.set_frame_runs(3)              # K2 works through runs of three frames, the strip staying on chip
    c.flavour = request.param
    yield c
    c.close()


def _diff(a, b):
    idx = np.argwhere(a != b)
    return "equal" if len(idx) == 0 else f"{len(idx)} bytes differ; first at frame {idx[0][0]}, offset {idx[0][1]}"


@pytest.mark.parametrize("name", FMT_CLIPS)
def test_golden_fixtures_other_formats(ctx, name):
    gd = golden(name)
    w, h, fmt = int(gd["w"]), int(gd["h"]), int(gd["fmt"])
    init = np.full(O.frame_bytes(fmt, w, h), int(gd["init_fill"]), dtype=np.uint8)
    got, _ = gpu_decode(ctx, gd["stream"], gd["offsets"], w, h, carry=init, fmt=fmt)
    assert np.array_equal(got, gd["frames"]), _diff(got, gd["frames"])
    assert ctx.batch_info().bad_frames == 0


@pytest.mark.parametrize("fmt", [1, 2])
@pytest.mark.parametrize("w,h,Q,F,kr,lm", [
    (720, 576, 128, 12, -1, 0),        # configs[1] geometry, intra
    (720, 576, 128, 33, 29, 2),        # configs[2] geometry: block-skip inter frames, GOP 30
    (720, 576, 255, 6, 2, 1),          # raw prefix 9 on luma
    (320, 240, 32, 8, 3, 4),           # dark flat regions: key frames with skips
    (16, 16, 200, 3, 1, 1),            # one unit wide
    (2064, 16, 128, 2, -1, 0),         # wider than one strip
    (1920, 1088, 255, 2, -1, 0),       # configs[3] geometry, dense
])
def test_clips_other_formats(ctx, fmt, w, h, Q, F, kr, lm):
    noise = 60 if (w, h) == (1920, 1088) else 6
    s, o = O.encode_clip_fmt(w, h, Q, F, fmt, kr, lm, lm, noise_y=noise, noise_c=3, dark=1 if Q == 32 else 0)
    init = np.full(O.frame_bytes(fmt, w, h), 0x3C, dtype=np.uint8)
    want = reference_frames_fmt(s, o, w, h, fmt, init)
    got, _ = gpu_decode(ctx, s, o, w, h, carry=init, fmt=fmt)
    assert np.array_equal(got, want), _diff(got, want)
    bi = ctx.batch_info()
    assert bi.bad_frames == 0
    if kr >= 0 and lm and w >= 320:
        assert bi.skipped_blocks > 0


@pytest.mark.parametrize("fmt", [1, 2])
def test_rtjpeg_decompress_other_formats(fmt):
    # Level 1: RTjpeg_set_format + RTjpeg_decompress, skipped blocks keep the caller's pixels
    gd = golden("yuv422_inter_64x48_q200_gop4" if fmt == 1 else "grey_inter_64x48_q255_gop4")
    s, o, w, h = gd["stream"], gd["offsets"], int(gd["w"]), int(gd["h"])
    r = g.RTjpeg()
    r.set_format(fmt)
    planes = np.full(O.frame_bytes(fmt, w, h), int(gd["init_fill"]), dtype=np.uint8)
    sizes = O.packet_sizes(s, o)
    for f in range(len(o) - 1):
        r.decompress(s[int(o[f]):int(o[f]) + int(sizes[f])], planes)
        assert np.array_equal(planes, gd["frames"][f]), f
    r.close()


@pytest.mark.parametrize("fmt", [1, 2])
def test_decode_host_other_formats(ctx, fmt):
    w, h, F = 320, 240, 40
    s, o = O.encode_clip_fmt(w, h, 128, F, fmt, 9, 2, 2, noise_y=4)
    fsz = O.frame_bytes(fmt, w, h)
    init = np.full(fsz, 0x21, dtype=np.uint8)
    want = reference_frames_fmt(s, o, w, h, fmt, init)
    out = np.empty((F, fsz), dtype=np.uint8)
    ctx.set_format(fmt)
    try:
        ctx.decode_host(s, o, out, carry=init.copy())
    finally:
        ctx.set_format(0)
    assert np.array_equal(out, want), _diff(out, want)
