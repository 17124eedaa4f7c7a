"""The two drop-in boundaries on the GPU: the RTjpeg.h-compatible codec API and the
bgav 'RTJ0' video-decoder plugin (driven through a stub host, tests/bgav_host_stub.c)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import gmerlin_avdecoder_b200 as g
from gmerlin_avdecoder_b200 import capi
from oracle import oracle as O
from streams import clip, golden, reference_frames

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _packets(s, o):
    sz = O.packet_sizes(s, o)
    return [s[int(o[f]):int(o[f]) + int(sz[f])] for f in range(len(o) - 1)]


def test_rtjpeg_decompress_sequence_with_persistent_planes():
    gd = golden("inter_64x48_q200_gop6")
    w, h = 64, 48
    planes = np.full(w * h * 3 // 2, 0x55, dtype=np.uint8)
    r = g.RTjpeg()
    for f, pkt in enumerate(_packets(gd["stream"], gd["offsets"])):
        r.decompress(pkt, planes)
        assert r.last_error() == 0
        assert np.array_equal(planes, gd["frames"][f]), f
    r.close()


def test_rtjpeg_skipped_blocks_keep_the_callers_pixels():
    """The reference leaves skipped blocks untouched in the CALLER's memory: scribbling on the
    planes between calls must show through exactly where the stream skips."""
    s, o = clip(160, 96, 128, 12, key_rate=11, lm=3, cm=3, noise_y=3)
    w, h = 160, 96
    rng = np.random.default_rng(1)
    mine = np.zeros(w * h * 3 // 2, dtype=np.uint8)
    ref = mine.copy()
    r, od = g.RTjpeg(), O.OracleDecoder()
    for f, pkt in enumerate(_packets(s, o)):
        if f % 3 == 1:
            at = rng.integers(0, mine.size, 500)
            mine[at] = ref[at] = 0xEE
        r.decompress(pkt, mine)
        od.decode(pkt, ref)
        assert np.array_equal(mine, ref), f
    r.close()


def test_rtjpeg_set_tables_path():
    gt = golden("set_tables")
    w, h = int(gt["w"]), int(gt["h"])
    for i in range(3):
        r = g.RTjpeg()
        assert r.set_size(w, h) == 0
        r.set_tables(gt[f"raw{i}"])
        planes = np.full(w * h * 3 // 2, int(gt["init_fill"]), dtype=np.uint8)
        r.decompress(gt[f"pkt{i}"], planes)
        assert r.last_error() == 0
        assert np.array_equal(planes, gt[f"planes{i}"]), i
        want = np.concatenate([O.tables_from_raw(gt[f"raw{i}"]).liqt, O.tables_from_raw(gt[f"raw{i}"]).ciqt])
        assert np.array_equal(r.get_tables(), want.astype(np.uint32))
        r.close()


def test_rtjpeg_setters_and_errors():
    r = g.RTjpeg()
    assert (r.get_tables() == 0).all()                       # fresh instance: zero tables
    assert r.set_quality(0) == 1 and r.set_quality(999) == 255 and r.set_quality(128) == 128
    assert np.array_equal(r.get_tables(), golden("tables")["tables"][127])
    assert r.set_size(-1, 16) == -1 and r.set_size(16, 65536) == -1 and r.set_size(65535, 0) == 0
    assert r.set_intra(300, -2, 99) == (255, 0, 16)
    gd = golden("intra_64x48_q128")
    pkt = _packets(gd["stream"], gd["offsets"])[0]
    planes = np.zeros(64 * 48 * 3 // 2, dtype=np.uint8)
    assert r.decompress_n(pkt[:8], planes) == capi.E_HEADER
    assert r.decompress_n(pkt[:pkt.size // 2], planes) == capi.E_OVERRUN
    assert r.last_error() == capi.E_OVERRUN and r.last_error() == 0
    r.set_format(7)                                          # the reference's switch has no such case (lib/RTjpeg.c:3580-3585)
    assert r.decompress_n(pkt, planes) == capi.E_FORMAT
    r.set_format(0)
    assert r.decompress_n(pkt, planes) == 0
    assert np.array_equal(planes, gd["frames"][0])
    bad = pkt.copy()
    bad[6] = 70
    assert r.decompress_n(bad, planes) == capi.E_SIZE
    r.close()


# --------------------------------------------------------------------------
# plugin
# --------------------------------------------------------------------------

@pytest.fixture(scope="module")
def host():
    build = os.path.join(HERE, "_build")
    os.makedirs(build, exist_ok=True)
    so = os.path.join(build, "libbgav_host_stub.so")
    subprocess.check_call(["gcc", "-O1", "-fPIC", "-shared", "-Wall", "-Wno-unused-parameter",
                           "-o", so, os.path.join(HERE, "bgav_host_stub.c")])
    H = C.CDLL(so, mode=C.RTLD_GLOBAL)                       # the plugin resolves bgav_*/gavl_* against it
    P = C.CDLL(g.PLUGIN_PATH, mode=C.RTLD_GLOBAL)
    vp = C.c_void_p
    H.stub_find_decoder.restype = vp
    H.stub_find_decoder.argtypes = [C.c_uint32]
    H.stub_decoder_name.restype = C.c_char_p
    H.stub_decoder_name.argtypes = [vp]
    H.stub_stream_create.restype = vp
    H.stub_stream_create.argtypes = [C.c_int, C.c_int]
    H.stub_stream_set_packets.argtypes = [vp, vp, vp, vp, C.c_int]
    H.stub_init.argtypes = [vp, vp]
    H.stub_close.argtypes = [vp, vp]
    H.stub_decode.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.POINTER(C.c_int64)]
    H.stub_stream_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                   C.POINTER(C.c_int), C.c_char_p, C.c_char_p]
    H.stub_stream_destroy.argtypes = [vp]
    H.stub_seek.argtypes = [vp, vp, C.c_int]
    H.stub_seek.restype = None
    H.stub_has_resync.argtypes = [vp]
    H.stub_skipto_intra.argtypes = [vp, C.c_int64]
    P.bgav_init_video_decoders_rtjpeg.restype = None
    P.bgav_init_video_decoders_rtjpeg()                      # what bgav_codecs_init does, lib/codecs.c:176
    return H


FOURCC_RTJ0 = (ord('R') << 24) | (ord('T') << 16) | (ord('J') << 8) | ord('0')


def test_plugin_registers_rtj0(host):
    assert host.stub_decoder_count() == 1
    dec = host.stub_find_decoder(FOURCC_RTJ0)
    assert dec
    assert host.stub_decoder_name(dec) == b"rtjpeg video decoder"
    assert not host.stub_find_decoder((ord('N') << 24) | (ord('U') << 16) | (ord('V') << 8) | ord(' '))


def test_plugin_decodes_with_strides_and_frame_skips(host):
    gd = golden("inter_64x48_q200_gop6")
    s, o = np.ascontiguousarray(gd["stream"]), gd["offsets"]
    sizes = O.packet_sizes(s, o).astype(np.uint32)
    F = len(o) - 1
    iw, ih = 60, 40                                          # image smaller than the padded 64x48 frame
    dec = host.stub_find_decoder(FOURCC_RTJ0)
    st = host.stub_stream_create(iw, ih)
    offs = np.ascontiguousarray(o[:-1], dtype=np.uint64)
    host.stub_stream_set_packets(st, s.ctypes.data, offs.ctypes.data, sizes.ctypes.data, F)
    assert host.stub_init(dec, st) == 1
    fw, fh, pf, done = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    key, val = C.create_string_buffer(64), C.create_string_buffer(64)
    host.stub_stream_info(st, fw, fh, pf, done, key, val)
    assert (fw.value, fh.value) == (64, 48) and key.value == b"Format" and val.value == b"RTjpeg"

    sy, sc = 80, 48                                          # padded strides
    Y = np.zeros((ih, sy), dtype=np.uint8)
    U = np.zeros(((ih + 1) // 2, sc), dtype=np.uint8)
    V = np.zeros_like(U)
    # the plugin's persistent picture starts zeroed; frames 3 and 7 are dropped undecoded
    dropped = {3, 7}
    od = O.OracleDecoder()
    ref = np.zeros(64 * 48 * 3 // 2, dtype=np.uint8)
    pts = C.c_int64()
    for f in range(F):
        pkt = s[int(o[f]):int(o[f]) + int(sizes[f])]
        if f in dropped:
            assert host.stub_decode(dec, st, None, None, None, 0, 0, None) == 1
            continue
        assert host.stub_decode(dec, st, Y.ctypes.data, U.ctypes.data, V.ctypes.data, sy, sc, C.byref(pts)) == 1
        od.decode(pkt, ref)
        assert pts.value == 1000 + 40 * f
        ry = ref[:64 * 48].reshape(48, 64)
        ru = ref[64 * 48:64 * 48 * 5 // 4].reshape(24, 32)
        rv = ref[64 * 48 * 5 // 4:].reshape(24, 32)
        assert np.array_equal(Y[:, :iw], ry[:ih, :iw]), f
        assert np.array_equal(U[:, :iw // 2], ru[:ih // 2, :iw // 2]), f
        assert np.array_equal(V[:, :iw // 2], rv[:ih // 2, :iw // 2]), f
        assert (Y[:, iw:] == 0).all()                        # nothing written outside the image
    assert host.stub_decode(dec, st, Y.ctypes.data, U.ctypes.data, V.ctypes.data, sy, sc, None) == 0   # EOF propagates
    host.stub_stream_info(st, fw, fh, pf, done, key, val)
    assert done.value == F                                   # every packet handed back exactly once
    host.stub_close(dec, st)
    host.stub_stream_destroy(st)


def test_plugin_ends_stream_on_malformed_packet(host):
    gd = golden("intra_64x48_q128")
    s, o = np.ascontiguousarray(gd["stream"]), gd["offsets"]
    sizes = O.packet_sizes(s, o).astype(np.uint32)
    sizes[1] = sizes[1] // 2                                 # second packet arrives truncated
    dec = host.stub_find_decoder(FOURCC_RTJ0)
    st = host.stub_stream_create(64, 48)
    offs = np.ascontiguousarray(o[:-1], dtype=np.uint64)
    host.stub_stream_set_packets(st, s.ctypes.data, offs.ctypes.data, sizes.ctypes.data, 3)
    assert host.stub_init(dec, st) == 1
    Y = np.zeros((48, 64), dtype=np.uint8); U = np.zeros((24, 32), dtype=np.uint8); V = np.zeros_like(U)
    assert host.stub_decode(dec, st, Y.ctypes.data, U.ctypes.data, V.ctypes.data, 64, 32, None) == 1
    assert np.array_equal(Y.ravel(), gd["frames"][0][:64 * 48])
    assert host.stub_decode(dec, st, Y.ctypes.data, U.ctypes.data, V.ctypes.data, 64, 32, None) == 0
    host.stub_close(dec, st)
    host.stub_stream_destroy(st)


@pytest.mark.parametrize("lookahead", [1, 5, 32])
def test_plugin_lookahead_is_invisible(host, lookahead, monkeypatch):
    """The plugin reads packets ahead and decodes them in batches; what the caller sees must be what
    lib/video_rtjpeg.c shows it: frames in order with their packets' timestamps, and a dropped packet
    (NULL frame) leaving the picture -- and every skipped block of later frames -- untouched."""
    from streams import clip
    monkeypatch.setenv("RTJPEG_B200_LOOKAHEAD", str(lookahead))
    w, h, F = 96, 64, 41
    s, o = clip(w, h, 128, F, key_rate=9, lm=2, cm=2, noise_y=4)
    s = np.ascontiguousarray(s)
    sizes = O.packet_sizes(s, o).astype(np.uint32)
    dec = host.stub_find_decoder(FOURCC_RTJ0)
    st = host.stub_stream_create(w, h)
    offs = np.ascontiguousarray(o[:-1], dtype=np.uint64)
    host.stub_stream_set_packets(st, s.ctypes.data, offs.ctypes.data, sizes.ctypes.data, F)
    assert host.stub_init(dec, st) == 1
    Y = np.zeros((h, w), dtype=np.uint8); U = np.zeros((h // 2, w // 2), dtype=np.uint8); V = np.zeros_like(U)
    dropped = {0, 4, 5, 6, 17, 31, 32, 40}                   # at ring borders, in runs, first and last
    od = O.OracleDecoder()
    ref = np.zeros(w * h * 3 // 2, dtype=np.uint8)
    pts = C.c_int64()
    for f in range(F):
        if f in dropped:
            assert host.stub_decode(dec, st, None, None, None, 0, 0, None) == 1
            continue
        assert host.stub_decode(dec, st, Y.ctypes.data, U.ctypes.data, V.ctypes.data, w, w // 2, C.byref(pts)) == 1
        od.decode(s[int(o[f]):int(o[f]) + int(sizes[f])], ref)   # the reference never saw the dropped packets
        assert pts.value == 1000 + 40 * f
        got = np.concatenate([Y.ravel(), U.ravel(), V.ravel()])
        assert np.array_equal(got, ref), f
    assert host.stub_decode(dec, st, Y.ctypes.data, U.ctypes.data, V.ctypes.data, w, w // 2, None) == 0
    fw, fh, pf, done = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    key, val = C.create_string_buffer(64), C.create_string_buffer(64)
    host.stub_stream_info(st, fw, fh, pf, done, key, val)
    assert done.value == F
    host.stub_close(dec, st)
    host.stub_stream_destroy(st)


def test_plugin_serves_frames_before_a_damaged_packet(host, monkeypatch):
    monkeypatch.setenv("RTJPEG_B200_LOOKAHEAD", "8")
    from streams import clip, reference_frames
    w, h, F = 320, 240, 6
    s, o = clip(w, h, 128, F, noise_y=3)
    s = np.ascontiguousarray(s)
    want = reference_frames(s, o, w, h, np.zeros(w * h * 3 // 2, dtype=np.uint8))
    sizes = O.packet_sizes(s, o).astype(np.uint32)
    sizes[2] = sizes[2] // 3                                 # third packet arrives truncated, inside the first batch
    dec = host.stub_find_decoder(FOURCC_RTJ0)
    st = host.stub_stream_create(w, h)
    offs = np.ascontiguousarray(o[:-1], dtype=np.uint64)
    host.stub_stream_set_packets(st, s.ctypes.data, offs.ctypes.data, sizes.ctypes.data, F)
    assert host.stub_init(dec, st) == 1
    Y = np.zeros((h, w), dtype=np.uint8); U = np.zeros((h // 2, w // 2), dtype=np.uint8); V = np.zeros_like(U)
    for f in range(2):
        assert host.stub_decode(dec, st, Y.ctypes.data, U.ctypes.data, V.ctypes.data, w, w // 2, None) == 1
        assert np.array_equal(np.concatenate([Y.ravel(), U.ravel(), V.ravel()]), want[f]), f
    assert host.stub_decode(dec, st, Y.ctypes.data, U.ctypes.data, V.ctypes.data, w, w // 2, None) == 0
    assert host.stub_decode(dec, st, Y.ctypes.data, U.ctypes.data, V.ctypes.data, w, w // 2, None) == 0
    host.stub_close(dec, st)
    host.stub_stream_destroy(st)



def _open(host, s, o, w, h):
    sizes = O.packet_sizes(s, o).astype(np.uint32)
    dec = host.stub_find_decoder(FOURCC_RTJ0)
    st = host.stub_stream_create(w, h)
    offs = np.ascontiguousarray(o[:-1], dtype=np.uint64)
    host.stub_stream_set_packets(st, s.ctypes.data, offs.ctypes.data, sizes.ctypes.data, len(o) - 1)
    assert host.stub_init(dec, st) == 1
    return dec, st, sizes, offs


@pytest.mark.parametrize("lookahead", [1, 7, 32])
def test_plugin_resync_drops_what_it_read_ahead(host, lookahead, monkeypatch):
    """A seek (bgav_video_resync, lib/video.c:525-562): the stream's queue is flushed and repositioned, then .resync
    is called.  The reference holds no packet between calls, so after a seek it decodes whatever comes next into
    the picture it has (lib/video_rtjpeg.c:81).  The plugin, which reads ahead, must do the same: drop the packets
    and frames it holds, keep picture and decoder state."""
    monkeypatch.setenv("RTJPEG_B200_LOOKAHEAD", str(lookahead))
    w, h, F = 96, 64, 60
    s, o = clip(w, h, 128, F, key_rate=9, lm=2, cm=2, noise_y=4)
    s = np.ascontiguousarray(s)
    dec, st, sizes, offs = _open(host, s, o, w, h)
    assert host.stub_has_resync(dec) == 1
    Y = np.zeros((h, w), dtype=np.uint8); U = np.zeros((h // 2, w // 2), dtype=np.uint8); V = np.zeros_like(U)
    od = O.OracleDecoder()
    ref = np.zeros(w * h * 3 // 2, dtype=np.uint8)
    pts = C.c_int64()

    def play(frames):
        for f in frames:
            assert host.stub_decode(dec, st, Y.ctypes.data, U.ctypes.data, V.ctypes.data, w, w // 2, C.byref(pts)) == 1
            od.decode(s[int(o[f]):int(o[f]) + int(sizes[f])], ref)
            assert pts.value == 1000 + 40 * f, (f, pts.value)
            assert np.array_equal(np.concatenate([Y.ravel(), U.ravel(), V.ravel()]), ref), f

    play(range(0, 10))                       # the ring now holds packets far beyond 10
    host.stub_seek(dec, st, 43)              # forward, into the middle of a GOP: inter frames land on the old picture
    play(range(43, 50))
    host.stub_seek(dec, st, 20)              # backward, onto a key frame
    play(range(20, 33))
    host.stub_seek(dec, st, 58)              # near the end
    play(range(58, 60))
    assert host.stub_decode(dec, st, Y.ctypes.data, U.ctypes.data, V.ctypes.data, w, w // 2, None) == 0
    host.stub_seek(dec, st, 5)               # and a seek after the end of the stream was reported
    play(range(5, 8))
    host.stub_close(dec, st)
    host.stub_stream_destroy(st)


def test_plugin_follows_the_hosts_intra_skip(host, monkeypatch):
    """bgav_video_skipto's shortcut for streams without P frames (lib/video.c:613-630) takes packets from the
    stream without asking the decoder and leaves s->out_time at the first packet it kept.  The packets the plugin
    holds lie before that one: they are dropped undecoded, and the next frame is the one the host expects."""
    monkeypatch.setenv("RTJPEG_B200_LOOKAHEAD", "8")
    w, h, F = 96, 64, 40
    s, o = clip(w, h, 128, F, noise_y=4)
    s = np.ascontiguousarray(s)
    want = reference_frames(s, o, w, h, np.zeros(w * h * 3 // 2, dtype=np.uint8))
    dec, st, sizes, offs = _open(host, s, o, w, h)
    Y = np.zeros((h, w), dtype=np.uint8); U = np.zeros((h // 2, w // 2), dtype=np.uint8); V = np.zeros_like(U)
    pts = C.c_int64()

    def next_frame():
        assert host.stub_decode(dec, st, Y.ctypes.data, U.ctypes.data, V.ctypes.data, w, w // 2, C.byref(pts)) == 1
        f = (pts.value - 1000) // 40
        assert np.array_equal(np.concatenate([Y.ravel(), U.ravel(), V.ravel()]), want[f]), f
        return f

    assert [next_frame() for _ in range(3)] == [0, 1, 2]          # packets 3..7 are held
    assert host.stub_skipto_intra(st, 1000 + 40 * 20 + 1) == 12    # the host eats packets 8..19 itself
    assert next_frame() == 20
    assert [next_frame() for _ in range(2)] == [21, 22]            # packets 23..27 are held
    # a target inside the held packets: the host finds the stream's next packet already behind it and eats nothing;
    # the plugin cannot know the time asked for and lands on that packet (documented: up to K - 1 frames late)
    assert host.stub_skipto_intra(st, 1000 + 40 * 25 + 1) == 0
    assert next_frame() == 28
    host.stub_close(dec, st)
    host.stub_stream_destroy(st)


def test_plugin_redecodes_only_up_to_the_next_clean_frame(host, monkeypatch):
    """A dropped packet voids the frames decoded ahead of it only up to the next frame without skip markers."""
    monkeypatch.setenv("RTJPEG_B200_LOOKAHEAD", "32")
    w, h, F = 96, 64, 32
    s, o = clip(w, h, 128, F, key_rate=7, lm=2, cm=2, noise_y=4)
    s = np.ascontiguousarray(s)
    dec, st, sizes, offs = _open(host, s, o, w, h)
    Y = np.zeros((h, w), dtype=np.uint8); U = np.zeros((h // 2, w // 2), dtype=np.uint8); V = np.zeros_like(U)
    od = O.OracleDecoder()
    ref = np.zeros(w * h * 3 // 2, dtype=np.uint8)
    L = g.load_library()
    for f in range(F):
        if f in (2, 3, 12):
            assert host.stub_decode(dec, st, None, None, None, 0, 0, None) == 1
            continue
        assert host.stub_decode(dec, st, Y.ctypes.data, U.ctypes.data, V.ctypes.data, w, w // 2, None) == 1
        od.decode(s[int(o[f]):int(o[f]) + int(sizes[f])], ref)
        assert np.array_equal(np.concatenate([Y.ravel(), U.ravel(), V.ravel()]), ref), f
    host.stub_close(dec, st)
    host.stub_stream_destroy(st)
