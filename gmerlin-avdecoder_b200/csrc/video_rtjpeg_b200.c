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
 *     after every frame before it was served.
 * RTJPEG_B200_LOOKAHEAD sets K (1 = the reference's one-packet behaviour).
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
    /* frames decoded ahead: [dec_lo, dec_hi) of the held packets */
    uint8_t       *frames;               /* pinned, K * fsz */
    int            dec_lo, dec_hi, slow; /* slow: a batch failed, the rest of the held packets decode one by one */
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
    if (!priv->frames || !priv->picture || !priv->carry || !priv->off || !priv->meta) { free_priv(priv); return 0; }
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
    priv->dec_lo = priv->dec_hi = 0;
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
    uint64_t rel[LOOKAHEAD_MAX + 1];
    for (int i = from; i <= to; i++) rel[i - from] = priv->off[i] - priv->off[from];
    /* the end of the last packet is its true length, not the aligned slot */
    rel[to - from] = priv->off[to - 1] - priv->off[from] + (uint64_t)priv->meta[to - 1].buf.len;
    rtjgpu_state st = priv->st;
    memcpy(priv->carry, priv->last, priv->fsz);
    return rtjgpu_decode_host(priv->ctx, priv->pk + priv->off[from], rel, to - from, &st,
                              priv->frames + (size_t)from * priv->fsz, priv->carry,
                              RTJGPU_HOST_IN_PINNED | RTJGPU_HOST_OUT_PINNED);
}

/* lib/video_rtjpeg.c:62-90 */
static gavl_source_status_t decode_rtjpeg_b200(bgav_stream_t *s, gavl_video_frame_t *f)
{
    rtjpeg_b200_priv_t *priv = s->decoder_priv;

    if (priv->head == priv->n) {
        if (priv->bad) return GAVL_SOURCE_EOF;   /* the unusable packet's turn */
        fill_ring(s, priv);
        if (priv->n == 0) return priv->bad ? GAVL_SOURCE_EOF : priv->tail;
    }

    if (!f) {                                   /* skip this frame: its packet is dropped undecoded */
        priv->head++;
        priv->dec_lo = priv->dec_hi = 0;        /* what was decoded ahead assumed this packet had been decoded */
        return GAVL_SOURCE_OK;
    }

    if (priv->head < priv->dec_lo || priv->head >= priv->dec_hi) {
        int rc = RTJGPU_E_ARG;
        if (!priv->slow) {
            rc = decode_range(priv, priv->head, priv->n);
            if (rc == RTJGPU_OK) { priv->dec_lo = priv->head; priv->dec_hi = priv->n; }
            else priv->slow = 1;                /* some packet ahead is damaged: find it one by one */
        }
        if (rc != RTJGPU_OK) {
            rc = decode_range(priv, priv->head, priv->head + 1);
            if (rc != RTJGPU_OK) {              /* this one: the stream ends here, as for a wrong-sized packet */
                priv->head = priv->n;
                priv->bad = 1;
                return GAVL_SOURCE_EOF;
            }
            priv->dec_lo = priv->head;
            priv->dec_hi = priv->head + 1;
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
        rtjgpu_frame_desc d;
        rtjgpu_plan(priv->pk + priv->off[priv->head], one, 1, &priv->st, &d);
    }
    priv->last = pic;
    priv->head++;
    return GAVL_SOURCE_OK;
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
};

/* include/codecs.h:97, lib/video_rtjpeg.c:112 */
void bgav_init_video_decoders_rtjpeg(void)
{
    bgav_video_decoder_register(&rtjpeg_b200_decoder);
}
