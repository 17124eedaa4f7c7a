/*
 * bgav_rtjpeg_plugin.h -- the slice of gmerlin-avdecoder's private plugin
 * contract that the 'RTJ0' video decoder touches, restated so that the plugin
 * (gmerlin-avdecoder_b200/csrc/video_rtjpeg_b200.c) compiles and can be driven
 * WITHOUT gavl and without the rest of libgmerlin_avdec (neither is installed
 * here).
 *
 * When the plugin is built inside the gmerlin-avdecoder tree, define
 * RTJ_B200_IN_TREE: this header then vanishes and the real <avdec_private.h>
 * and <codecs.h> are used instead.  Field and function
 * NAMES below are the reference's; layouts are minimal stand-ins and are not
 * ABI-compatible with a real gavl build -- they only have to agree with the
 * test host (tests/bgav_host_stub.c).
 *
 * Reference declarations this mirrors:
 *   bgav_video_decoder_s            include/avdec_private.h:90-118
 *   bgav_stream_s (fields used)     include/avdec_private.h:231 (data.video.format),
 *                                   :263 (decoder_priv), :272 (fourcc), :317 (m)
 *   BGAV_MK_FOURCC                  include/avdec_private.h:47
 *   bgav_stream_get_packet_read     include/avdec_private.h:470, lib/stream.c:539
 *   bgav_stream_done_packet_read    include/avdec_private.h:474, lib/stream.c:597
 *   bgav_set_video_frame_from_packet include/avdec_private.h:1438, lib/video.c:861
 *   bgav_video_decoder_register     include/avdec_private.h:1501, lib/codecs.c:201
 *   bgav_init_video_decoders_rtjpeg include/codecs.h:97, lib/video_rtjpeg.c:112
 */
#ifndef BGAV_RTJPEG_PLUGIN_H
#define BGAV_RTJPEG_PLUGIN_H

#ifndef RTJ_B200_IN_TREE

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- gavl stand-ins -------------------------------------------------------- */

typedef enum {
    GAVL_SOURCE_EOF   = 0,
    GAVL_SOURCE_OK    = 1,
    GAVL_SOURCE_AGAIN = 2
} gavl_source_status_t;

#define GAVL_YUV_420_P       0x0501     /* opaque tag for the stub; planar 4:2:0, 8 bit */
#define GAVL_MAX_PLANES      4
#define GAVL_META_FORMAT     "Format"

typedef struct gavl_dictionary_s gavl_dictionary_t;   /* owned by the host */

typedef struct {
    int x, y, w, h;
} gavl_rectangle_i_t;

typedef struct {
    int image_width, image_height;
    int frame_width, frame_height;
    int pixelformat;
} gavl_video_format_t;

#define GAVL_TIME_UNDEFINED ((int64_t)0x8000000000000000LL)     /* gavl/gavltime.h */

typedef struct {
    uint8_t *planes[GAVL_MAX_PLANES];
    int      strides[GAVL_MAX_PLANES];
    int64_t  timestamp;
    int64_t  duration;
    uint32_t timecode;
    int      dst_x, dst_y;
    gavl_rectangle_i_t src_rect;
} gavl_video_frame_t;

typedef struct {
    struct { uint8_t *buf; int len; } buf;
    int64_t  pts;
    int64_t  duration;
    uint32_t timecode;
    int      dst_x, dst_y;
    gavl_rectangle_i_t src_rect;
} gavl_packet_t;

void gavl_dictionary_set_string(gavl_dictionary_t *d, const char *key, const char *val);

/* ---- bgav stand-ins -------------------------------------------------------- */

#define bgav_packet_t gavl_packet_t              /* include/avdec_private.h:56 */
#define BGAV_MK_FOURCC(a, b, c, d) ((uint32_t)(((uint32_t)(a) << 24) | ((b) << 16) | ((c) << 8) | (d)))

typedef struct bgav_stream_s bgav_stream_t;
typedef struct bgav_video_decoder_s bgav_video_decoder_t;

struct bgav_video_decoder_s {
    const uint32_t *fourccs;                      /* zero-terminated */
    const char     *name;
    int  (*probe)(const gavl_dictionary_t *stream);
    int  (*init)(bgav_stream_t *);                /* 1 = ok, 0 = failure */
    gavl_source_status_t (*decode)(bgav_stream_t *, gavl_video_frame_t *);   /* frame NULL: skip */
    void (*close)(bgav_stream_t *);
    void (*resync)(bgav_stream_t *);
    int  (*skipto)(bgav_stream_t *, int64_t dest);
    bgav_video_decoder_t *next;
};

struct bgav_stream_s {
    void              *decoder_priv;
    uint32_t           fourcc;
    gavl_dictionary_t *m;
    struct {
        struct {
            gavl_video_format_t *format;
        } video;
    } data;
    int64_t            out_time;                  /* include/avdec_private.h:300: end of the last frame delivered, or where a skip landed */
    void              *host_priv;                 /* stub only: the test host's packet queue */
};

gavl_source_status_t bgav_stream_get_packet_read(bgav_stream_t *s, bgav_packet_t **ret);
void bgav_stream_done_packet_read(bgav_stream_t *s, bgav_packet_t *p);
void bgav_set_video_frame_from_packet(const bgav_packet_t *p, gavl_video_frame_t *f);
void bgav_video_decoder_register(bgav_video_decoder_t *dec);

#ifdef __cplusplus
}
#endif

#endif /* !RTJ_B200_IN_TREE */

#ifdef __cplusplus
extern "C" {
#endif

/* Registration entry point, same symbol as the reference's
 * (include/codecs.h:97, called from bgav_codecs_init, lib/codecs.c:176). */
void bgav_init_video_decoders_rtjpeg(void);

#ifdef __cplusplus
}
#endif

#endif /* BGAV_RTJPEG_PLUGIN_H */
