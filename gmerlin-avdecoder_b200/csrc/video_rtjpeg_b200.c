/*
 * video_rtjpeg_b200.c -- bgav video-decoder plugin for fourcc 'RTJ0' backed by
 * the B200 decoder.  Takes the place of lib/video_rtjpeg.c in
 * gmerlin-avdecoder: same registration symbol, same decoder name and fourcc
 * list, same init / decode / close behaviour (one packet in, one frame out,
 * NULL frame = drop the packet undecoded).
 *
 * Differences a maintainer should know about:
 *   - the persistent picture lives in plain host memory owned by the plugin
 *     (tight pitch, as RTjpeg_decompress requires) instead of a
 *     gavl_video_frame_create() frame; the copy into the caller's frame
 *     honours the caller's strides;
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

typedef struct {
    RTjpeg_t *rtjpeg;
    uint8_t  *picture;          /* frame_width * frame_height * 3 / 2, persists between packets */
    uint8_t  *planes[3];
    int       fw, fh;
} rtjpeg_b200_priv_t;

/* lib/video_rtjpeg.c:41-60 */
static int init_rtjpeg_b200(bgav_stream_t *s)
{
    rtjpeg_b200_priv_t *priv = calloc(1, sizeof(*priv));
    if (!priv) return 0;
    priv->rtjpeg = RTjpeg_init();
    if (!priv->rtjpeg) { free(priv); return 0; }
    s->decoder_priv = priv;

    gavl_video_format_t *fmt = s->data.video.format;
    fmt->frame_width = ROUND_UP_MB(fmt->image_width);
    fmt->frame_height = ROUND_UP_MB(fmt->image_height);
    fmt->pixelformat = GAVL_YUV_420_P;

    priv->fw = fmt->frame_width;
    priv->fh = fmt->frame_height;
    const size_t ysz = (size_t)priv->fw * priv->fh;
    priv->picture = calloc(ysz * 3 / 2 ? ysz * 3 / 2 : 1, 1);
    if (!priv->picture) { RTjpeg_close(priv->rtjpeg); free(priv); s->decoder_priv = NULL; return 0; }
    priv->planes[0] = priv->picture;
    priv->planes[1] = priv->picture + ysz;
    priv->planes[2] = priv->picture + ysz + ysz / 4;

    gavl_dictionary_set_string(s->m, GAVL_META_FORMAT, "RTjpeg");
    return 1;
}

static void copy_plane(uint8_t *dst, int dst_stride, const uint8_t *src, int src_stride, int bytes, int rows)
{
    for (int r = 0; r < rows; r++) memcpy(dst + (size_t)r * dst_stride, src + (size_t)r * src_stride, (size_t)bytes);
}

/* lib/video_rtjpeg.c:62-90 */
static gavl_source_status_t decode_rtjpeg_b200(bgav_stream_t *s, gavl_video_frame_t *f)
{
    rtjpeg_b200_priv_t *priv = s->decoder_priv;
    bgav_packet_t *p = NULL;
    gavl_source_status_t st;

    if ((st = bgav_stream_get_packet_read(s, &p)) != GAVL_SOURCE_OK)
        return st;

    if (!f) {                                   /* skip this frame: the packet is dropped undecoded */
        bgav_stream_done_packet_read(s, p);
        return GAVL_SOURCE_OK;
    }

    /* The packet's own dimensions must be the stream's padded dimensions, or the
     * persistent picture would be addressed with the wrong pitch. */
    int ok = p->buf.len >= RTJPEG_B200_HEADER_BYTES
          && (p->buf.buf[6] | p->buf.buf[7] << 8) == priv->fw
          && (p->buf.buf[8] | p->buf.buf[9] << 8) == priv->fh;
    if (ok) ok = RTjpeg_b200_decompress_n(priv->rtjpeg, p->buf.buf, (size_t)p->buf.len, priv->planes) == 0;
    if (!ok) {
        bgav_stream_done_packet_read(s, p);
        return GAVL_SOURCE_EOF;
    }

    const gavl_video_format_t *fmt = s->data.video.format;
    const int iw = fmt->image_width, ih = fmt->image_height;
    copy_plane(f->planes[0], f->strides[0], priv->planes[0], priv->fw, iw, ih);
    copy_plane(f->planes[1], f->strides[1], priv->planes[1], priv->fw / 2, (iw + 1) / 2, (ih + 1) / 2);
    copy_plane(f->planes[2], f->strides[2], priv->planes[2], priv->fw / 2, (iw + 1) / 2, (ih + 1) / 2);

    bgav_set_video_frame_from_packet(p, f);
    bgav_stream_done_packet_read(s, p);
    return GAVL_SOURCE_OK;
}

/* lib/video_rtjpeg.c:93-101 */
static void close_rtjpeg_b200(bgav_stream_t *s)
{
    rtjpeg_b200_priv_t *priv = s->decoder_priv;
    if (!priv) return;
    RTjpeg_close(priv->rtjpeg);
    free(priv->picture);
    free(priv);
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
