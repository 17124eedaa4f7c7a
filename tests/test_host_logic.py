"""Host-side logic of the product library that needs no GPU: the C ABI loads and
exports every declared symbol, header planning, table derivation, shard split,
and the loud failure when no CUDA device exists."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gmerlin_avdecoder_b200 as g
from gmerlin_avdecoder_b200 import capi
from oracle import oracle as O
from streams import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#\s*define.*$", "", text, flags=re.M)          # RTjpeg_yuv422rgb8 is a macro, as in the reference
    return sorted(set(re.findall(r"\b((?:RTjpeg|rtjgpu|rtjnuv|bgav_init)_\w+)\s*\(", text)))


def _exported(lib):
    return sorted(l.split()[-1] for l in os.popen(f"nm -D --defined-only {lib}").read().splitlines() if l.strip())


def test_abi_exports_exactly_the_declared_symbols():
    """Header -> library: every prototype is exported.  Library -> header: nothing else leaves the DSO (it is meant to be
    linked into libgmerlin_avdec): every exported name is a prototype of include/rtjpeg_b200.h, and the reference's own
    entry points (include/RTjpeg.h:115-138) are all among them."""
    L = g.load_library()
    names = _declared("rtjpeg_b200.h")
    assert len(names) >= 55
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/rtjpeg_b200.h but not exported"
    exported = _exported(g.LIB_PATH)
    assert exported == names, sorted(set(exported) ^ set(names))
    for n in ("RTjpeg_init", "RTjpeg_close", "RTjpeg_set_quality", "RTjpeg_set_format", "RTjpeg_set_size", "RTjpeg_set_intra",
              "RTjpeg_compress", "RTjpeg_decompress", "RTjpeg_yuv420rgb32", "RTjpeg_yuv420bgr32", "RTjpeg_yuv420rgb24",
              "RTjpeg_yuv420bgr24", "RTjpeg_yuv420rgb16", "RTjpeg_yuv420rgb8", "RTjpeg_yuv422rgb24", "RTjpeg_get_tables",
              "RTjpeg_set_tables"):
        assert n in names, n
    assert _exported(g.PLUGIN_PATH) == ["bgav_init_video_decoders_rtjpeg"]


def test_plugin_library_exports_registration_symbol():
    out = os.popen(f"nm -D --defined-only {g.PLUGIN_PATH}").read()
    assert "bgav_init_video_decoders_rtjpeg" in out
    # the symbols the host has to provide, exactly the reference's (lib/video_rtjpeg.c)
    und = os.popen(f"nm -D --undefined-only {g.PLUGIN_PATH}").read()
    for n in ("bgav_stream_get_packet_read", "bgav_stream_done_packet_read",
              "bgav_set_video_frame_from_packet", "bgav_video_decoder_register", "gavl_dictionary_set_string"):
        assert n in und


def test_product_never_links_the_oracle():
    for lib in (g.LIB_PATH, g.PLUGIN_PATH):
        syms = os.popen(f"nm -D {lib}").read()
        assert "rtjo_" not in syms and "refdrv_" not in syms and "ref_RTjpeg" not in syms
    src = os.path.join(ROOT, "gmerlin-avdecoder_b200")
    for dirpath, _, files in os.walk(src):
        for f in files:
            if f.endswith((".py", ".c", ".cpp", ".cu", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), f"{f} mentions the oracle"


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(g.RTjpegError):
        g.BatchContext(0)
    with pytest.raises(g.RTjpegError):
        g.RTjpeg()


def test_tables_match_reference_golden():
    tabs = golden("tables")["tables"]
    for q in range(1, 256):
        scaled, lb8, cb8 = g.tables_for_quality(q)
        assert np.array_equal(scaled, tabs[q - 1]), q
        t = O.tables_from_quality(q)
        assert (lb8, cb8) == (t.lb8, t.cb8)
    gt = golden("set_tables")
    for i in range(3):
        scaled, lb8, cb8 = g.tables_from_raw(gt[f"raw{i}"])
        t = O.tables_from_raw(gt[f"raw{i}"])
        assert np.array_equal(scaled, np.concatenate([t.liqt, t.ciqt]).astype(np.uint32))
        assert (lb8, cb8) == (t.lb8, t.cb8)
    assert g.tables_from_raw(gt["raw1"])[1:] == (62, 62)
    # a table that never exceeds 8 is undefined behaviour in the reference (unbounded scan,
    # lib/RTjpeg.c:2388-2393); the library caps the prefix at 63 instead of reading out of bounds
    assert g.tables_from_raw(np.full(128, 3, dtype=np.uint32))[1:] == (63, 63)


def test_plan_follows_lazy_reconfiguration():
    gd = golden("inter_64x48_q200_gop6")
    s, o = gd["stream"], gd["offsets"]
    desc, st = g.plan(s, o)
    sizes = O.packet_sizes(s, o)
    assert (desc["offset"] == o[:-1]).all()
    assert (desc["length"] == sizes).all()               # framesize bounds the packet, not the padding
    assert (desc["table"] == 200).all()
    assert (st.width, st.height, st.quality, st.table) == (64, 48, 200, 200)

    # quality byte 0: fresh instance keeps the all-zero table; a configured one falls to Q=1
    pkt = s[int(o[0]):int(o[0]) + int(sizes[0])].copy()
    pkt[10] = 0
    s2, o2 = O.pack_packets([pkt, pkt])
    d2, st2 = g.plan(s2, o2)
    assert list(d2["table"]) == [g.TABLE_ZERO, g.TABLE_ZERO] and st2.quality == 0
    d3, st3 = g.plan(s2, o2, capi.State(64, 48, 200, 200))
    assert list(d3["table"]) == [1, 1] and st3.quality == 1

    # custom tables stay in force while the quality byte equals the instance's quality
    d4, st4 = g.plan(s2, o2, capi.State(64, 48, g.TABLE_CUSTOM, 0))
    assert list(d4["table"]) == [g.TABLE_CUSTOM] * 2
    d5, st5 = g.plan(s, o, capi.State(64, 48, g.TABLE_CUSTOM, 0))
    assert (d5["table"] == 200).all()


def test_plan_rejects_bad_input():
    gd = golden("intra_64x48_q128")
    s, o = gd["stream"].copy(), gd["offsets"]
    with pytest.raises(g.RTjpegError) as e:
        g.plan(s[:int(o[0]) + 8], np.array([o[0], o[0] + 8], dtype=np.uint64))
    assert e.value.code == capi.E_HEADER
    bad = s.copy()
    bad[int(o[1]) + 6] = 80                              # second frame claims another width
    with pytest.raises(g.RTjpegError) as e:
        g.plan(bad, o)
    assert e.value.code == capi.E_SIZE
    bad = s.copy()
    bad[int(o[0]) + 6] = 70                              # not a multiple of 16
    with pytest.raises(g.RTjpegError) as e:
        g.plan(bad, o)
    assert e.value.code == capi.E_SIZE
    with pytest.raises(g.RTjpegError) as e:
        g.plan(s, o + np.uint64(2))                      # unaligned packet start
    assert e.value.code == capi.E_ARG
    d, _ = g.plan(s, o[:1])                              # empty batch is fine
    assert len(d) == 0


def test_split_shards_cuts_only_on_clean_frames():
    F = 120
    clean = np.zeros(F, dtype=np.uint8)
    clean[::30] = 1                                      # GOP 30, every key frame clean
    first = g.split_shards(clean, 4)
    assert list(first) == [0, 30, 60, 90, 120]
    first = g.split_shards(clean, 3)                     # ideal cuts 40, 80 -> the nearest clean frames 30, 90
    assert list(first) == [0, 30, 90, 120]
    clean[:] = 0
    clean[0] = 1                                         # nothing clean after frame 0: the path does not shard ...
    assert list(g.split_shards(clean, 4)) == [0, 120, 120, 120, 120]
    with pytest.raises(ValueError):                      # ... and says so when asked to
        g.split_shards(clean, 4, allow_empty=False)
    clean[:] = 1                                         # intra-only: equal split
    assert list(g.split_shards(clean, 8)) == [0, 15, 30, 45, 60, 75, 90, 105, 120]
    assert list(g.split_shards(np.zeros(0, dtype=np.uint8), 2)) == [0, 0, 0]


def test_split_shards_lead_never_leaves_a_shard_empty():
    F = 120
    clean = np.zeros(F, dtype=np.uint8)
    clean[::30] = 1
    first, lead = g.split_shards_lead(clean, 4)
    assert list(first) == [0, 30, 60, 90, 120] and list(lead) == [0, 0, 0, 0]
    first, lead = g.split_shards_lead(clean, 8)          # ideal cuts 15, 30, 45 ...: half of them have no clean frame within 7
    assert list(first) == [0, 15, 30, 45, 60, 75, 90, 105, 120]
    assert list(lead) == [0, 15, 0, 15, 0, 15, 0, 15]
    clean[:] = 0                                         # no clean frame at all, not even frame 0: everything from the start
    first, lead = g.split_shards_lead(clean, 4)
    assert list(first) == [0, 30, 60, 90, 120] and list(lead) == [0, 30, 60, 90]
    clean[:] = 0
    clean[50] = 1                                        # one clean frame: cuts near it snap to it, later ones lead back to it
    first, lead = g.split_shards_lead(clean, 4)
    assert list(first) == [0, 30, 50, 90, 120] and list(lead) == [0, 30, 0, 40]
    rng = np.random.default_rng(3)
    for _ in range(200):                                 # every shard has work while F >= n; leads end on a clean frame or 0
        F = int(rng.integers(1, 300))
        n = int(rng.integers(1, 9))
        clean = (rng.random(F) < rng.choice([0.0, 0.02, 0.2, 1.0])).astype(np.uint8)
        first, lead = g.split_shards_lead(clean, n)
        assert first[0] == 0 and first[-1] == F and (np.diff(first) >= 0).all()
        if F >= n:
            assert (np.diff(first) > 0).all()
        for i in range(n):
            a = int(first[i] - lead[i])
            assert a >= 0 and (lead[i] == 0 or a == 0 or clean[a])
            if i and lead[i] == 0 and first[i] < F and first[i + 1] > first[i]:
                assert clean[first[i]]


def test_split_shards_agrees_with_oracle_decode():
    """Decoding each shard independently from zeroed planes equals the sequential decode,
    because every shard starts on a frame without skipped blocks."""
    gd = golden("inter_64x48_q200_gop6")
    s, o = gd["stream"], gd["offsets"]
    w, h = 64, 48
    F = len(o) - 1
    t = O.tables_from_quality(200)
    sizes = O.packet_sizes(s, o)
    clean = np.array([(O.walk_payload(s[int(o[f]) + 12:int(o[f]) + int(sizes[f])], 12, t.lb8, t.cb8)[2] != 0).all()
                      for f in range(F)], dtype=np.uint8)
    assert clean[0] and clean.sum() >= 2
    whole = O.decode_stream(s, o, w, h)
    first = g.split_shards(clean, 2)
    for a, b in zip(first[:-1], first[1:]):
        if b > a:
            part = O.decode_stream(s, o[a:b + 1], w, h)
            assert np.array_equal(part, whole[a:b])
