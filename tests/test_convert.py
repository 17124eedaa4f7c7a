"""The colour converters of lib/RTjpeg.c:3071-3486 (RTjpeg_yuv420rgb32 ... RTjpeg_yuv422rgb24): the oracle's
restatement against the unmodified reference and a committed fixture (CPU), the CUDA converters against
both (GPU, through the C ABI)."""
import hashlib

import numpy as np
import pytest

from oracle import oracle as O
from streams import golden

KINDS = list(range(7))
NAMES = ["rgb32", "bgr32", "rgb24", "bgr24", "rgb16", "rgb8", "yuv422rgb24"]


def _picture(kind, w, h, seed):
    """Full-range bytes: every clamp of the converter is exercised (the decoder itself only emits 16..235)."""
    rng = np.random.default_rng(seed)
    n = w * h * 2 if kind == O.CONV_YUV422_RGB24 else w * h * 3 // 2
    pic = rng.integers(0, 256, n).astype(np.uint8)
    pic[:64] = np.arange(64) * 4            # a ramp through the luma range
    return pic


@pytest.mark.parametrize("kind", KINDS)
def test_restatement_matches_golden_converters(kind):
    g = golden("convert_48x32")
    out = O.convert(kind, g["pic422"] if kind == O.CONV_YUV422_RGB24 else g["pic420"], 48, 32, fill=0x5A)
    assert np.array_equal(out, g[NAMES[kind]])


@pytest.mark.skipif(not O.have_ref(), reason="reference build absent (oracle/_ref)")
@pytest.mark.parametrize("kind", KINDS)
def test_restatement_matches_live_reference_converters(kind):
    for (w, h) in [(16, 16), (64, 48), (720, 576)]:
        pic = _picture(kind, w, h, 7 + kind)
        assert np.array_equal(O.convert(kind, pic, w, h, fill=0xA5), O.ref_convert(kind, pic, w, h, fill=0xA5))
        pitch = w * 4 + 48
        assert np.array_equal(O.convert(kind, pic, w, h, pitch=pitch, fill=3), O.ref_convert(kind, pic, w, h, pitch=pitch, fill=3))


def _want(kind, pic, w, h, pitch=None, fill=0):
    return O.ref_convert(kind, pic, w, h, pitch, fill) if O.have_ref() else O.convert(kind, pic, w, h, pitch, fill)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_convert_device_matches_reference(kind):
    import torch
    import gmerlin_avdecoder_b200 as g
    bpp = g.CONV_BPP[kind]
    with g.BatchContext(0) as ctx:
        for (w, h, F, pad) in [(16, 16, 3, 0), (64, 48, 5, 32), (720, 576, 4, 0), (1920, 1088, 2, 16)]:
            pics = np.stack([_picture(kind, w, h, 100 * kind + f) for f in range(F)])
            src_fb = pics.shape[1]
            row_pitch = w * bpp + pad
            frame_pitch = row_pitch * h + 64
            alpha = 0x7E
            d_src = torch.from_numpy(pics).cuda()
            d_out = torch.full((F, frame_pitch), alpha, dtype=torch.uint8, device="cuda")
            ctx.convert_device(kind, d_src.data_ptr(), src_fb, F, w, h, d_out.data_ptr(), row_pitch, frame_pitch, alpha,
                               torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            got = d_out.cpu().numpy()
            for f in range(F):
                # the reference leaves byte 3 of a 32-bit pixel and the row padding alone: pre-filled with alpha
                want = _want(kind, pics[f], w, h, pitch=row_pitch, fill=alpha)
                assert np.array_equal(got[f, :row_pitch * h].reshape(h, row_pitch), want), (w, h, f)
                assert (got[f, row_pitch * h:] == alpha).all()


@pytest.mark.gpu
def test_convert_device_after_decode():
    """Decode a clip and convert it where it lies: what a player does with RTjpeg_decompress + RTjpeg_yuv420rgb32."""
    import torch
    import gmerlin_avdecoder_b200 as g
    from gmerlin_avdecoder_b200 import device as D
    from streams import clip, reference_frames
    w, h, F = 320, 240, 12
    s, o = clip(w, h, 128, F, key_rate=5, lm=2, cm=2, noise_y=6, noise_c=3)
    want_yuv = reference_frames(s, o, w, h)
    with g.BatchContext(0) as ctx:
        desc, _ = g.plan(s, o)
        b = D.upload(s, desc, w, h)
        D.decode(ctx, b)
        rgb = torch.zeros((F, h, w * 4), dtype=torch.uint8, device="cuda")
        ctx.convert_device(g.capi.CONV_BGR32, b.out.data_ptr(), b.frame_bytes, F, w, h, rgb.data_ptr(), w * 4, w * 4 * h, 255,
                           torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = rgb.cpu().numpy()
    for f in range(F):
        assert np.array_equal(got[f], _want(O.CONV_BGR32, want_yuv[f], w, h, fill=255)), f


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_rtjpeg_converters_drop_in(kind):
    """Level 1: RTjpeg_yuv420rgb32 & co. with host planes and host rows; byte 3 of 32-bit pixels stays the caller's."""
    import gmerlin_avdecoder_b200 as g
    w, h = 96, 64
    pic = _picture(kind, w, h, 55 + kind)
    pitch = w * g.CONV_BPP[kind] + 8
    out = np.full((h, pitch), 0xC3, dtype=np.uint8)
    r = g.RTjpeg()
    assert r.set_size(w, h) == 0
    r.convert(kind, pic, w, h, out)
    assert r.last_error() == 0
    r.close()
    assert np.array_equal(out, _want(kind, pic, w, h, pitch=pitch, fill=0xC3))


def test_sha_of_fixture_is_stable():
    g = golden("convert_48x32")
    assert hashlib.sha256(g["pic420"].tobytes()).hexdigest()[:16] == str(g["sha420"])


@pytest.mark.gpu
def test_convert_and_encode_refuse_bad_arguments():
    import torch
    import gmerlin_avdecoder_b200 as g
    from gmerlin_avdecoder_b200 import capi
    w, h = 64, 48
    src = torch.zeros(w * h * 3 // 2, dtype=torch.uint8, device="cuda")
    out = torch.zeros(w * h * 4, dtype=torch.uint8, device="cuda")
    with g.BatchContext(0) as ctx:
        def conv(kind=0, src_fb=w * h * 3 // 2, ww=w, hh=h, rp=w * 4, fp=w * h * 4):
            ctx.convert_device(kind, src.data_ptr(), src_fb, 1, ww, hh, out.data_ptr(), rp, fp)
        for kwargs, code in [(dict(kind=9), capi.E_FORMAT), (dict(ww=60), capi.E_SIZE), (dict(rp=w * 4 - 16), capi.E_ARG),
                             (dict(rp=w * 4 + 8), capi.E_ARG), (dict(fp=w * 4 * (h - 1)), capi.E_ARG),
                             (dict(src_fb=w * h), capi.E_ARG)]:
            with pytest.raises(g.RTjpegError) as e:
                conv(**kwargs)
            assert e.value.code == code, kwargs
        conv()                                               # and the well-formed call goes through
        off = torch.zeros(2, dtype=torch.int64, device="cuda")
        ctx.set_format(2)                                    # the 8-bit encoder is not offered
        with pytest.raises(g.RTjpegError) as e:
            ctx.encode_device(src.data_ptr(), 1, w, h, out.data_ptr(), out.numel(), off.data_ptr())
        assert e.value.code == capi.E_FORMAT
        ctx.set_format(0)
        with pytest.raises(g.RTjpegError) as e:
            ctx.encode_device(src.data_ptr(), 1, 50, h, out.data_ptr(), out.numel(), off.data_ptr())
        assert e.value.code == capi.E_SIZE
        torch.cuda.synchronize()
