/*
 * bgav_host_stub.c -- TEST INFRASTRUCTURE: the handful of libgmerlin_avdec /
 * gavl functions the 'RTJ0' plugin calls, backed by a trivial packet queue, so
 * that the plugin contract (include/bgav_rtjpeg_plugin.h) can be driven
 * without gavl.  Behaviour follows the reference where it matters to the
 * plugin: lib/codecs.c:201-215 (register appends, clears ->next),
 * lib/video.c:861-871 (frame metadata from the packet), lib/stream.c:597
 * (done_packet_read is a no-op).
 */
#include <stdlib.h>
#include <string.h>

#include "../include/bgav_rtjpeg_plugin.h"

static bgav_video_decoder_t *decoders;
static int n_decoders;

void bgav_video_decoder_register(bgav_video_decoder_t *dec)
{
    if (!decoders) decoders = dec;
    else {
        bgav_video_decoder_t *b = decoders;
        while (b->next) b = b->next;
        b->next = dec;
    }
    dec->next = NULL;
    n_decoders++;
}

typedef struct {
    bgav_packet_t *pk;
    int n, rd, done_calls;
    char meta_key[64], meta_val[64];
    gavl_video_format_t fmt;
} host_t;

gavl_source_status_t bgav_stream_get_packet_read(bgav_stream_t *s, bgav_packet_t **ret)
{
    host_t *h = s->host_priv;
    if (h->rd >= h->n) return GAVL_SOURCE_EOF;
    *ret = &h->pk[h->rd++];
    return GAVL_SOURCE_OK;
}

void bgav_stream_done_packet_read(bgav_stream_t *s, bgav_packet_t *p) { ((host_t *)s->host_priv)->done_calls++; }

void bgav_set_video_frame_from_packet(const bgav_packet_t *p, gavl_video_frame_t *f)
{
    f->timestamp = p->pts;
    f->duration = p->duration;
    f->timecode = p->timecode;
    f->dst_x = p->dst_x;
    f->dst_y = p->dst_y;
    f->src_rect = p->src_rect;
}

void gavl_dictionary_set_string(gavl_dictionary_t *d, const char *key, const char *val)
{
    host_t *h = (host_t *)d;
    strncpy(h->meta_key, key, sizeof(h->meta_key) - 1);
    strncpy(h->meta_val, val, sizeof(h->meta_val) - 1);
}

/* ---- entry points for the Python test ---- */

int stub_decoder_count(void) { return n_decoders; }

bgav_video_decoder_t *stub_find_decoder(uint32_t fourcc)
{
    for (bgav_video_decoder_t *d = decoders; d; d = d->next)
        for (const uint32_t *f = d->fourccs; *f; f++)
            if (*f == fourcc) return d;
    return NULL;
}

const char *stub_decoder_name(bgav_video_decoder_t *d) { return d->name; }

bgav_stream_t *stub_stream_create(int image_w, int image_h)
{
    bgav_stream_t *s = calloc(1, sizeof(*s));
    host_t *h = calloc(1, sizeof(*h));
    h->fmt.image_width = image_w;
    h->fmt.image_height = image_h;
    s->host_priv = h;
    s->m = (gavl_dictionary_t *)h;
    s->data.video.format = &h->fmt;
    s->fourcc = BGAV_MK_FOURCC('R', 'T', 'J', '0');
    s->out_time = GAVL_TIME_UNDEFINED;
    return s;
}

void stub_stream_set_packets(bgav_stream_t *s, uint8_t *buf, const uint64_t *offsets, const uint32_t *sizes, int n)
{
    host_t *h = s->host_priv;
    free(h->pk);
    h->pk = calloc((size_t)n, sizeof(bgav_packet_t));
    for (int i = 0; i < n; i++) {
        h->pk[i].buf.buf = buf + offsets[i];
        h->pk[i].buf.len = (int)sizes[i];
        h->pk[i].pts = 1000 + 40 * i;
        h->pk[i].duration = 40;
        h->pk[i].timecode = (uint32_t)i;
    }
    h->n = n;
    h->rd = 0;
}

int stub_init(bgav_video_decoder_t *d, bgav_stream_t *s) { return d->init(s); }
void stub_close(bgav_video_decoder_t *d, bgav_stream_t *s) { d->close(s); }

/* decode into planes with the given strides; y == NULL means "skip this frame" */
int stub_decode(bgav_video_decoder_t *d, bgav_stream_t *s, uint8_t *y, uint8_t *u, uint8_t *v,
                int sy, int sc, int64_t *pts_out)
{
    gavl_video_frame_t f;
    memset(&f, 0, sizeof(f));
    f.planes[0] = y; f.planes[1] = u; f.planes[2] = v;
    f.strides[0] = sy; f.strides[1] = sc; f.strides[2] = sc;
    int st = (int)d->decode(s, y ? &f : NULL);
    if (pts_out) *pts_out = f.timestamp;
    /* lib/video.c:274,295: the stream's clock follows the frames handed out */
    if (y && st == (int)GAVL_SOURCE_OK) s->out_time = f.timestamp + f.duration;
    return st;
}

/* a seek as the decoder sees it (bgav_video_resync, lib/video.c:525-562): the stream's queue now starts at packet
 * `next`, then the decoder's resync -- when it registered one -- is called */
void stub_seek(bgav_video_decoder_t *d, bgav_stream_t *s, int next)
{
    host_t *h = s->host_priv;
    h->rd = next;
    s->out_time = h->pk[next < h->n ? next : h->n - 1].pts;
    if (d->resync) d->resync(s);
}

int stub_has_resync(bgav_video_decoder_t *d) { return d->resync != NULL; }

/* bgav_video_skipto for a stream without P frames (lib/video.c:613-630): packets are taken from the stream, not
 * through the decoder, until the next one ends behind `time`; returns how many it consumed, -1 at the end */
int stub_skipto_intra(bgav_stream_t *s, int64_t time)
{
    host_t *h = s->host_priv;
    int consumed = 0;
    for (;;) {
        if (h->rd >= h->n) return -1;
        bgav_packet_t *p = &h->pk[h->rd];
        if (p->pts + p->duration > time) { s->out_time = p->pts; return consumed; }
        h->rd++;
        h->done_calls++;
        consumed++;
    }
}

void stub_stream_info(bgav_stream_t *s, int *fw, int *fh, int *pixfmt, int *done_calls, char *key, char *val)
{
    host_t *h = s->host_priv;
    *fw = h->fmt.frame_width; *fh = h->fmt.frame_height; *pixfmt = h->fmt.pixelformat;
    *done_calls = h->done_calls;
    strcpy(key, h->meta_key); strcpy(val, h->meta_val);
}

void stub_stream_destroy(bgav_stream_t *s)
{
    host_t *h = s->host_priv;
    free(h->pk); free(h); free(s);
}
