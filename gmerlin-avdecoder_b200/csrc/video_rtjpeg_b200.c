/*
 * video_rtjpeg_b200.c -- bgav video-decoder plugin for fourcc 'RTJ0' backed by
 * the B200 decoder.  Takes the place of lib/video_rtjpeg.c in
 * gmerlin-avdecoder: same registration symbol, same decoder name and fourcc
 * list, same init / decode / close behaviour (one packet in, one frame out,
 * NULL frame = drop the packet undecoded).
 *
 * Differences a maintainer should know about:
 *   - packets are read ahead and decoded in batches (see below); the persistent
 *     picture lives in pinned host memory owned by the plugin (tight pitch)
 *     instead of a gavl_video_frame_create() frame; the copy into the caller's
 *     frame honours the caller's strides;
 *   - init fails (returns 0) when no CUDA device is usable -- there is no CPU
 *     path behind this plugin;
 *   - a packet that is truncated or malformed ends the stream (GAVL_SOURCE_EOF)
 *     instead of being read past (the reference has no check at all).
 */
#include <stdlib.h>
#include <string.h>

#ifdef RTJ_B200_IN_TREE
#include <avdec_private.h>
#include <codecs.h>
#else
#include "../../include/bgav_rtjpeg_plugin.h"
#endif
#include "../../include/rtjpeg_b200.h"

#define MB_SIZE 16
#define ROUND_UP_MB(x) ((((x) + MB_SIZE - 1) / MB_SIZE) * MB_SIZE)
#define LOOKAHEAD_DEFAULT 32
#define LOOKAHEAD_MAX 1024

/*
 * The reference decodes one packet per call (lib/video_rtjpeg.c:69-81).  One frame per GPU round trip
 * wastes the device, so this plugin reads ahead: when it runs dry it pulls up
 * to K packets from the stream (copying them, handing each back at once), decodes them as ONE batch
 * and then serves the calls from the decoded frames.  What the caller can observe stays the reference's:
 *   - frames come out in packet order with their packets' timestamps;
 *   - a call with a NULL frame drops one packet UNDECODED (:75-79) -- the picture, and with it every
 *     skipped block of later inter-coded frames, stays as it was; frames decoded ahead across a dropped
 *     packet are therefore decoded again, from the picture last delivered, before they are served;
 *   - a packet that is truncated, malformed or of the wrong size ends the stream when ITS turn comes,
 *     after every frame before it was served;
 *   - a seek (bgav_video_resync, lib/video.c:525-562, calls .resync after the stream's packet queue was flushed)
 *     drops every packet and frame held: what comes next is what the stream hands out next.  The picture and the
 *     decoder state stay, as they do in the reference's RTjpeg_t and priv->frame;
 *   - bgav_video_skipto's shortcut for intra-only streams (lib/video.c:613-630) consumes packets from the stream
 *     behind the decoder's back and leaves s->out_time at the first packet it kept.  Held packets that end at or
 *     before s->out_time are therefore dropped undecoded before the next frame is served -- the packets the
 *     shortcut would have consumed had they still been in the stream's queue.  (It cannot know the time asked
 *     for: a target INSIDE the held packets lands on the first packet behind them, up to K - 1 frames late.)
 * RTJPEG_B200_LOOKAHEAD sets K (1 = the reference's one-packet behaviour, exact in every respect).
 */
typedef struct {
    rtjgpu_ctx    *ctx;
    int            fw, fh, K;
    size_t         fsz;
    /* packets held: copies, 16-byte aligned, in pinned memory */
    uint8_t       *pk;   size_t pk_cap;
    uint64_t      *off;                  /* K + 1 */
    bgav_packet_t *meta;                 /* K: the packets' metadata (timestamps ...), payload pointer cleared */
    int            n, head;              /* packets held / next one to serve */
    int            bad;                  /* the packet after the held ones was unusable: the stream ends there */
    gavl_source_status_t tail;           /* why the last fill stopped early (EOF / AGAIN), else OK */
    /* frames decoded ahead */
    uint8_t       *frames;               /* pinned, K * fsz */
    uint8_t       *ok;                   /* K: frames[i] is the picture after held packet i */
    uint8_t       *clean;                /* K: held packet i carries no skip marker (known once it was decoded) */
    uint32_t      *skips;                /* K: scratch for the per-frame skip counts of a batch */
    uint64_t      *rel;                  /* K + 1: scratch for packet offsets relative to a batch */
    uint32_t      *lens;                 /* K: scratch for the packets' true lengths */
    int            slow;                 /* a batch failed: the rest of the held packets decode one by one */
    uint8_t       *picture;              /* pinned, fsz: the picture last delivered ... */
    const uint8_t *last;                 /* ... or where it still sits among the decoded frames */
    uint8_t       *carry;                /* pinned, fsz: scratch for rtjgpu_decode_host's in/out picture */
    rtjgpu_state   st;                   /* decoder state in front of packet `head` */
} rtjpeg_b200_priv_t;

static void free_priv(rtjpeg_b200_priv_t *priv)
{
    if (!priv) return;
    rtjgpu_host_free(priv->pk);
    rtjgpu_host_free(priv->frames);
    rtjgpu_host_free(priv->picture);
    rtjgpu_host_free(priv->carry);
    free(priv->off);
    free(priv->meta);
    free(priv->ok);
    free(priv->clean);
    free(priv->skips);
    free(priv->rel);
    free(priv->lens);
    if (priv->ctx) rtjgpu_destroy(priv->ctx);
    free(priv);
}

/* lib/video_rtjpeg.c:41-60 */
static int init_rtjpeg_b200(bgav_stream_t *s)
{
    rtjpeg_b200_priv_t *priv = calloc(1, sizeof(*priv));
    if (!priv) return 0;
    int dev = 0;
    const char *e;
    if ((e = getenv("RTJPEG_B200_DEVICE"))) dev = atoi(e);
    if (rtjgpu_create(dev, &priv->ctx) != RTJGPU_OK) { free(priv); return 0; }     /* no CUDA device: no decoder */
    priv->K = LOOKAHEAD_DEFAULT;
    if ((e = getenv("RTJPEG_B200_LOOKAHEAD"))) priv->K = atoi(e);
    if (priv->K < 1) priv->K = 1;
    if (priv->K > LOOKAHEAD_MAX) priv->K = LOOKAHEAD_MAX;

    gavl_video_format_t *fmt = s->data.video.format;
    fmt->frame_width = ROUND_UP_MB(fmt->image_width);
    fmt->frame_height = ROUND_UP_MB(fmt->image_height);
    fmt->pixelformat = GAVL_YUV_420_P;

    priv->fw = fmt->frame_width;
    priv->fh = fmt->frame_height;
    priv->fsz = (size_t)priv->fw * priv->fh * 3 / 2;
    const size_t fsz = priv->fsz ? priv->fsz : 1;
    priv->frames = rtjgpu_host_alloc(fsz * (size_t)priv->K);
    priv->picture = rtjgpu_host_alloc(fsz);
    priv->carry = rtjgpu_host_alloc(fsz);
    priv->off = calloc((size_t)priv->K + 1, sizeof(*priv->off));
    priv->meta = calloc((size_t)priv->K, sizeof(*priv->meta));
    priv->ok = calloc((size_t)priv->K, 1);
    priv->clean = calloc((size_t)priv->K, 1);
    priv->skips = calloc((size_t)priv->K, sizeof(*priv->skips));
    priv->rel = calloc((size_t)priv->K + 1, sizeof(*priv->rel));
    priv->lens = calloc((size_t)priv->K, sizeof(*priv->lens));
    if (!priv->frames || !priv->picture || !priv->carry || !priv->off || !priv->meta || !priv->ok || !priv->clean
        || !priv->skips || !priv->rel || !priv->lens) { free_priv(priv); return 0; }
    memset(priv->picture, 0, fsz);               /* gavl_video_frame_create hands out a cleared frame */
    priv->last = priv->picture;
    priv->st.width = priv->st.height = 0;
    priv->st.table = RTJGPU_TABLE_ZERO;
    priv->st.quality = 0;
    priv->tail = GAVL_SOURCE_OK;
    s->decoder_priv = priv;

    gavl_dictionary_set_string(s->m, GAVL_META_FORMAT, "RTjpeg");
    return 1;
}

static void copy_plane(uint8_t *dst, int dst_stride, const uint8_t *src, int src_stride, int bytes, int rows)
{
    for (int r = 0; r < rows; r++) memcpy(dst + (size_t)r * dst_stride, src + (size_t)r * src_stride, (size_t)bytes);
}

/* Pull packets until K are held or the stream has nothing more right now. */
static void fill_ring(bgav_stream_t *s, rtjpeg_b200_priv_t *priv)
{
    if (priv->last != priv->picture) {           /* the frames are about to be overwritten: keep the picture */
        memcpy(priv->picture, priv->last, priv->fsz);
        priv->last = priv->picture;
    }
    priv->n = priv->head = 0;
    memset(priv->ok, 0, (size_t)priv->K);
    memset(priv->clean, 0, (size_t)priv->K);
    priv->slow = 0;
    priv->off[0] = 0;
    priv->tail = GAVL_SOURCE_OK;
    while (priv->n < priv->K && !priv->bad) {
        bgav_packet_t *p = NULL;
        gavl_source_status_t st = bgav_stream_get_packet_read(s, &p);
        if (st != GAVL_SOURCE_OK) { priv->tail = st; break; }
        /* The packet's own dimensions must be the stream's padded dimensions, or the persistent
         * picture would be addressed with the wrong pitch. */
        const int ok = p->buf.len >= RTJPEG_B200_HEADER_BYTES
                    && (p->buf.buf[6] | p->buf.buf[7] << 8) == priv->fw
                    && (p->buf.buf[8] | p->buf.buf[9] << 8) == priv->fh;
        if (!ok) {
            priv->bad = 1;
        } else {
            const size_t at = (size_t)priv->off[priv->n], len = (size_t)p->buf.len;
            const size_t need = at + ((len + 15) & ~(size_t)15) + RTJGPU_STREAM_SLACK_BYTES;
            if (need > priv->pk_cap) {
                size_t cap = priv->pk_cap ? priv->pk_cap : (size_t)1 << 16;
                while (cap < need) cap *= 2;
                uint8_t *nb = rtjgpu_host_alloc(cap);
                if (!nb) { priv->bad = 1; bgav_stream_done_packet_read(s, p); break; }
                if (at) memcpy(nb, priv->pk, at);
                rtjgpu_host_free(priv->pk);
                priv->pk = nb;
                priv->pk_cap = cap;
            }
            memcpy(priv->pk + at, p->buf.buf, len);
            priv->meta[priv->n] = *p;
            priv->meta[priv->n].buf.buf = NULL;
            priv->meta[priv->n].buf.len = (int)len;
            priv->off[priv->n + 1] = at + ((len + 15) & ~(size_t)15);
            priv->n++;
        }
        bgav_stream_done_packet_read(s, p);
    }
}

/* Decode held packets [from, to) from the picture last delivered.  0 on success. */
static int decode_range(rtjpeg_b200_priv_t *priv, int from, int to)
{
    uint64_t *rel = priv->rel;
    for (int i = from; i <= to; i++) rel[i - from] = priv->off[i] - priv->off[from];
    /* the packets' true lengths (gavl_packet_t.buf.len), not their aligned slots and not the header's framesize,
     * which the reference never reads (lib/RTjpeg.c:3565-3586) */
    for (int i = from; i < to; i++) priv->lens[i - from] = (uint32_t)priv->meta[i].buf.len;
    rtjgpu_state st = priv->st;
    memcpy(priv->carry, priv->last, priv->fsz);
    const int rc = rtjgpu_decode_host_n(priv->ctx, priv->pk + priv->off[from], rel, priv->lens, to - from, &st,
                                        priv->frames + (size_t)from * priv->fsz, priv->carry,
                                        RTJGPU_HOST_IN_PINNED | RTJGPU_HOST_OUT_PINNED);
    if (rc != RTJGPU_OK) return rc;
    /* a frame without skip markers rewrites the whole picture: what follows it does not depend on what precedes it */
    if (rtjgpu_get_host_skip_counts(priv->ctx, priv->skips, to - from) == RTJGPU_OK)
        for (int i = from; i < to; i++) priv->clean[i] = priv->skips[i - from] == 0;
    memset(priv->ok + from, 1, (size_t)(to - from));
    return RTJGPU_OK;
}

/* Held packets the host has moved past behind the decoder's back (bgav_video_skipto's intra-only shortcut,
 * lib/video.c:613-630, leaves s->out_time at the first packet it kept): dropped undecoded, like its own. */
static void drop_passed_packets(bgav_stream_t *s, rtjpeg_b200_priv_t *priv)
{
    if (s->out_time == GAVL_TIME_UNDEFINED) return;
    int dropped = 0;
    while (priv->head < priv->n) {
        const bgav_packet_t *m = &priv->meta[priv->head];
        if (!(m->pts < s->out_time && m->pts + m->duration <= s->out_time)) break;
        priv->head++;
        dropped = 1;
    }
    if (dropped)                               /* what was decoded ahead assumed these packets had been decoded */
        for (int i = priv->head; i < priv->n && !priv->clean[i]; i++) priv->ok[i] = 0;
}

/* lib/video_rtjpeg.c:62-90 */
static gavl_source_status_t decode_rtjpeg_b200(bgav_stream_t *s, gavl_video_frame_t *f)
{
    rtjpeg_b200_priv_t *priv = s->decoder_priv;

    drop_passed_packets(s, priv);
    if (priv->head == priv->n) {
        if (priv->bad) return GAVL_SOURCE_EOF;   /* the unusable packet's turn */
        fill_ring(s, priv);
        if (priv->n == 0) return priv->bad ? GAVL_SOURCE_EOF : priv->tail;
    }

    if (!f) {                                   /* skip this frame: its packet is dropped undecoded */
        priv->head++;
        /* what was decoded ahead assumed this packet had been decoded: the frames up to the next clean one are void */
        for (int i = priv->head; i < priv->n && !priv->clean[i]; i++) priv->ok[i] = 0;
        return GAVL_SOURCE_OK;
    }

    if (!priv->ok[priv->head]) {
        int rc = RTJGPU_E_ARG;
        if (!priv->slow) {
            int to = priv->head + 1;            /* up to the next frame that is still good (a clean frame or the end) */
            while (to < priv->n && !priv->ok[to]) to++;
            rc = decode_range(priv, priv->head, to);
            if (rc != RTJGPU_OK) priv->slow = 1;  /* some packet ahead is damaged: find it one by one */
        }
        if (rc != RTJGPU_OK) {
            rc = decode_range(priv, priv->head, priv->head + 1);
            if (rc != RTJGPU_OK) {              /* this one: the stream ends here, as for a wrong-sized packet */
                priv->head = priv->n;
                priv->bad = 1;
                return GAVL_SOURCE_EOF;
            }
        }
    }

    const uint8_t *pic = priv->frames + (size_t)priv->head * priv->fsz;
    const gavl_video_format_t *fmt = s->data.video.format;
    const int iw = fmt->image_width, ih = fmt->image_height;
    const size_t ysz = (size_t)priv->fw * priv->fh;
    copy_plane(f->planes[0], f->strides[0], pic, priv->fw, iw, ih);
    copy_plane(f->planes[1], f->strides[1], pic + ysz, priv->fw / 2, (iw + 1) / 2, (ih + 1) / 2);
    copy_plane(f->planes[2], f->strides[2], pic + ysz + ysz / 4, priv->fw / 2, (iw + 1) / 2, (ih + 1) / 2);
    bgav_set_video_frame_from_packet(&priv->meta[priv->head], f);

    /* the decoder state moves past this packet (lib/RTjpeg.c:3568-3579); a dropped packet never touches it */
    {
        const uint64_t one[2] = {0, (uint64_t)priv->meta[priv->head].buf.len};
        const uint32_t len = (uint32_t)priv->meta[priv->head].buf.len;
        rtjgpu_frame_desc d;
        rtjgpu_plan_n(priv->pk + priv->off[priv->head], one, &len, 1, &priv->st, &d);
    }
    priv->last = pic;
    priv->head++;
    return GAVL_SOURCE_OK;
}

/* include/avdec_private.h:90-118 (.resync), called by bgav_video_resync (lib/video.c:561-562) after a seek: the
 * stream's packet queue was flushed, so every packet held here -- and every frame decoded from them -- is stale.
 * The reference registers no resync because it holds nothing between calls; its RTjpeg_t (size, quality, tables)
 * and its picture survive a seek, and so do they here. */
static void resync_rtjpeg_b200(bgav_stream_t *s)
{
    rtjpeg_b200_priv_t *priv = s->decoder_priv;
    if (!priv) return;
    if (priv->last != priv->picture) {
        memcpy(priv->picture, priv->last, priv->fsz);
        priv->last = priv->picture;
    }
    priv->n = priv->head = 0;
    memset(priv->ok, 0, (size_t)priv->K);
    memset(priv->clean, 0, (size_t)priv->K);
    priv->slow = 0;
    priv->bad = 0;
    priv->tail = GAVL_SOURCE_OK;
}

/* lib/video_rtjpeg.c:93-101 */
static void close_rtjpeg_b200(bgav_stream_t *s)
{
    free_priv(s->decoder_priv);
    s->decoder_priv = NULL;
}

static const uint32_t rtjpeg_b200_fourccs[] = { BGAV_MK_FOURCC('R', 'T', 'J', '0'), 0x00 };

/* writable: the registry links decoders through ->next (lib/codecs.c:201-215) */
static bgav_video_decoder_t rtjpeg_b200_decoder = {
    .fourccs = rtjpeg_b200_fourccs,
    .name    = "rtjpeg video decoder",
    .init    = init_rtjpeg_b200,
    .decode  = decode_rtjpeg_b200,
    .close   = close_rtjpeg_b200,
    .resync  = resync_rtjpeg_b200,
};

/* include/codecs.h:97, lib/video_rtjpeg.c:112 */
void bgav_init_video_decoders_rtjpeg(void)
{
    bgav_video_decoder_register(&rtjpeg_b200_decoder);
}
