"""The oracle (oracle/rtjpeg_oracle.c) against the golden fixtures made by the
unmodified reference, and against the reference itself when its build is present."""
import numpy as np
import pytest

from oracle import oracle as O
from streams import golden, sha

CLIPS = ["intra_64x48_q128", "intra_320x240_q128", "inter_64x48_q200_gop6",
         "inter_dark_96x64_q32_gop4", "dense_48x32_q255", "random_48x32"]


@pytest.mark.parametrize("name", CLIPS)
def test_restatement_matches_golden(name):
    g = golden(name)
    w, h = int(g["w"]), int(g["h"])
    init = np.full(w * h * 3 // 2, int(g["init_fill"]), dtype=np.uint8)
    frames = O.decode_stream(g["stream"], g["offsets"], w, h, init=init)
    assert [sha(f) for f in frames] == list(g["sha"])
    if "frames" in g:
        assert np.array_equal(frames, g["frames"])


FMT_CLIPS = ["yuv422_inter_64x48_q200_gop4", "yuv422_intra_96x32_q128", "grey_inter_64x48_q255_gop4",
             "grey_intra_96x32_q64"]


@pytest.mark.parametrize("name", FMT_CLIPS)
def test_restatement_matches_golden_other_formats(name):
    # YUV422 and 8-bit grey: RTjpeg_decompressYUV422 / RTjpeg_decompress8 (lib/RTjpeg.c:2639-2686, :2751-2772)
    g = golden(name)
    w, h, fmt = int(g["w"]), int(g["h"]), int(g["fmt"])
    init = np.full(O.frame_bytes(fmt, w, h), int(g["init_fill"]), dtype=np.uint8)
    frames = O.decode_stream_fmt(g["stream"], g["offsets"], w, h, fmt, init=init)
    assert [sha(f) for f in frames] == list(g["sha"])
    assert np.array_equal(frames, g["frames"])


@pytest.mark.skipif(not O.have_ref(), reason="reference build absent (oracle/_ref)")
@pytest.mark.parametrize("fmt", [1, 2])
def test_restatement_matches_live_reference_other_formats(fmt):
    for (w, h, Q, kr, lm) in [(96, 64, 57, -1, 0), (64, 48, 171, 4, 1), (48, 32, 255, 3, 4), (160, 32, 32, 2, 6)]:
        s, o = O.encode_clip_fmt(w, h, Q, 7, fmt, kr, lm, lm, noise_y=12, noise_c=4, dark=1, seed=Q)
        init = np.full(O.frame_bytes(fmt, w, h), 0xA5, dtype=np.uint8)
        assert np.array_equal(O.ref_decode_seq_fmt(s, o, w, h, fmt, init=init),
                              O.decode_stream_fmt(s, o, w, h, fmt, init=init)), (w, h, Q)


def test_tables_match_golden():
    tabs = golden("tables")["tables"]
    for q in range(1, 256):
        t = O.tables_from_quality(q)
        assert np.array_equal(t.liqt.astype(np.uint32), tabs[q - 1][:64]), q
        assert np.array_equal(t.ciqt.astype(np.uint32), tabs[q - 1][64:]), q
    # raw-prefix classes the survey probed: lb8 0 (Q<=170), 4, 8, 9; cb8 always 0
    assert {(O.tables_from_quality(q).lb8, O.tables_from_quality(q).cb8) for q in range(1, 256)} == \
        {(0, 0), (4, 0), (8, 0), (9, 0)}
    assert O.tables_from_quality(170).lb8 == 0 and O.tables_from_quality(171).lb8 == 4
    assert O.tables_from_quality(200).lb8 == 8 and O.tables_from_quality(228).lb8 == 9
    # clamping of set_quality
    assert np.array_equal(O.tables_from_quality(0).liqt, O.tables_from_quality(1).liqt)
    assert np.array_equal(O.tables_from_quality(999).liqt, O.tables_from_quality(255).liqt)


def test_set_tables_path_matches_golden():
    g = golden("set_tables")
    w, h = int(g["w"]), int(g["h"])
    for i in range(3):
        raw, pkt, want = g[f"raw{i}"], g[f"pkt{i}"], g[f"planes{i}"]
        t = O.tables_from_raw(raw)
        # decode with the restatement's primitives: custom tables + macroblock walk
        n, offs, eob = O.walk_payload(pkt[12:], (w // 16) * (h // 16), t.lb8, t.cb8)
        assert n == pkt.size - 12
        planes = np.full(w * h * 3 // 2, int(g["init_fill"]), dtype=np.uint8)
        import ctypes as C
        L = O.oracle_lib()
        blk = (C.c_int16 * 64)()
        cw = w // 2
        for b, off in enumerate(offs):
            mb, sub = divmod(b, 6)
            my, mx = divmod(mb, w // 16)
            iq = (t.liqt if sub < 4 else t.ciqt).astype(np.int32)
            payload = np.ascontiguousarray(pkt[12 + int(off):])
            L.rtjo_unpack_block(payload.ctypes.data_as(C.POINTER(C.c_uint8)), t.lb8 if sub < 4 else t.cb8,
                                iq.ctypes.data_as(C.POINTER(C.c_int32)), blk)
            if sub < 4:
                o, pitch = (my * 16 + (sub >> 1) * 8) * w + mx * 16 + (sub & 1) * 8, w
            else:
                o, pitch = w * h + (sub - 4) * cw * (h // 2) + my * 8 * cw + mx * 8, cw
            L.rtjo_idct_block(blk, C.cast(planes.ctypes.data + o, C.POINTER(C.c_uint8)), pitch)
        assert np.array_equal(planes, want), i


def test_walker_consumes_exactly_the_payload():
    for name in CLIPS:
        g = golden(name)
        s, o = g["stream"], g["offsets"]
        w, h = int(g["w"]), int(g["h"])
        sizes = O.packet_sizes(s, o)
        for f in range(len(o) - 1):
            q = int(s[int(o[f]) + 10])
            t = O.tables_from_quality(q)
            n, offs, eob = O.walk_payload(s[int(o[f]) + 12:int(o[f]) + int(sizes[f])], (w // 16) * (h // 16), t.lb8, t.cb8)
            assert n == int(sizes[f]) - 12
            assert ((eob == 0) == (offs == 0xFFFFFFFF)).all()


def test_truncated_packet_is_refused():
    g = golden("intra_64x48_q128")
    s, o = g["stream"], g["offsets"]
    pkt = s[int(o[0]):int(o[1])][:60]
    with pytest.raises(ValueError):
        O.OracleDecoder().decode(pkt)


def test_dc_only_block_is_flat():
    # DC-only: every pixel = clamp((dc*iq0 + 4) >> 3)
    blk = np.zeros(64, dtype=np.int16)
    for dc, want in [(0, 16), (8 * 100, 100), (8 * 300, 235), (-500, 16), (8 * 16 - 4, 16), (8 * 17 - 4, 17)]:
        blk[0] = dc
        assert (O.idct_block(blk) == want).all(), dc


@pytest.mark.skipif(not O.have_ref(), reason="reference build absent (oracle/_ref)")
def test_restatement_matches_live_reference():
    rng = np.random.default_rng(3)
    # encoder-made streams with the quality changing mid-stream
    parts = []
    for q in (1, 57, 170, 171, 255):
        parts.append(O.encode_clip(O.make_clip(96, 64, q, key_rate=4, lm=1, cm=1, noise_y=12, noise_c=4, seed=q), 7))
    pk = []
    for s, o in parts:
        sz = O.packet_sizes(s, o)
        pk += [s[int(o[f]):int(o[f]) + int(sz[f])] for f in range(len(o) - 1)]
    stream, offs = O.pack_packets(pk)
    init = rng.integers(0, 256, 96 * 64 * 3 // 2).astype(np.uint8)
    assert np.array_equal(O.ref_decode_seq(stream, offs, 96, 64, init=init),
                          O.decode_stream(stream, offs, 96, 64, init=init))
    # grammar-only random packets: int16 wrap, every raw-prefix class, long blocks, skips
    pk = [O.random_wellformed_packet(rng, 64, 32, int(q)) for q in (1, 2, 32, 170, 171, 199, 200, 227, 228, 255)
          for _ in range(2)]
    stream, offs = O.pack_packets(pk)
    init = np.full(64 * 32 * 3 // 2, 0xA5, dtype=np.uint8)
    assert np.array_equal(O.ref_decode_seq(stream, offs, 64, 32, init=init),
                          O.decode_stream(stream, offs, 64, 32, init=init))
    # quality byte 0 on a fresh decoder: all-zero tables, coded blocks become flat 16
    pkt = O.random_wellformed_packet(rng, 32, 32, 1, skip_prob=0.3)
    pkt[10] = 0
    stream, offs = O.pack_packets([pkt, pkt])
    a = O.ref_decode_seq(stream, offs, 32, 32, init=np.full(1536, 0x55, dtype=np.uint8))
    b = O.decode_stream(stream, offs, 32, 32, init=np.full(1536, 0x55, dtype=np.uint8))
    assert np.array_equal(a, b) and set(np.unique(a)) <= {16, 0x55}
