/*
 * rtj_nuv.c -- NuppelVideo / MythTV container reader in front of the RTjpeg decoder
 * (SURVEY.md section 8f-1).  Host-side byte shuffling only, no CUDA.
 *
 * Follows the reference demuxer lib/demux_nuv.c: probe_nuv (:44-55), the file header and codec
 * data of open_nuv (:58-243) and the frame walk of next_packet_nuv (:246-318).  In the reference
 * these packets are tagged 'NUV ' and leave for libavcodec (lib/video_ffmpeg.c:1880-1883); here
 * the RTjpeg-coded video frames are rewrapped as 'RTJ0' packets -- a 12-byte RTjpeg_frameheader
 * (include/RTjpeg.h:100-109) in front of the block stream -- so that the in-tree arithmetic of
 * lib/RTjpeg.c (and this library) decodes them.  The tables do not come from a quality byte but
 * from the file's 'D'/'R' packet (128 x u32, the form RTjpeg_set_tables takes, lib/RTjpeg.c:2380);
 * the headers written here carry quality 0, which on an instance whose quality is 0 leaves the
 * tables loaded with RTjpeg_set_tables / rtjgpu_set_custom_tables in force (lib/RTjpeg.c:3575).
 */
#include <string.h>

#include "rtj_common.h"

#define NUV_SIG_LEN 12
#define NUV_HDRSIZE 12
#define NUV_FILE_HEADER 72       /* signature 12, version 8, w/h 8, desired w/h 8, 'P' + pad 4, aspect/fps 16, packs 8, text/keydist 8 */

static const char nuppel_sig[NUV_SIG_LEN] = "NuppelVideo";
static const char mythtv_sig[NUV_SIG_LEN] = "MythTVVideo";

static uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static void wr32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static void wr16(uint8_t *p, unsigned v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static double rd_f64le(const uint8_t *p) { double d; memcpy(&d, p, 8); return d; }   /* little-endian hosts only, like the CUDA side */

/* lib/demux_nuv.c:44-55 */
int rtjnuv_probe(const uint8_t *data, size_t len)
{
    if (!data || len < NUV_SIG_LEN) return 0;
    return !memcmp(data, nuppel_sig, NUV_SIG_LEN) || !memcmp(data, mythtv_sig, NUV_SIG_LEN);
}

/* lib/demux_nuv.c:58-243: the fixed header, then frames until the codec data has been seen */
int rtjnuv_open(const uint8_t *data, size_t len, rtjnuv_header *out)
{
    if (!data || !out) return RTJGPU_E_ARG;
    memset(out, 0, sizeof(*out));
    if (len < NUV_FILE_HEADER || !rtjnuv_probe(data, len)) return RTJGPU_E_HEADER;
    out->is_mythtv = !memcmp(data, mythtv_sig, NUV_SIG_LEN);
    out->width = (int)rd32(data + 20);
    out->height = (int)rd32(data + 24);
    out->interlaced = data[36] != 'P';
    out->aspect = rd_f64le(data + 40);
    out->fps = rd_f64le(data + 48);
    out->video_packets = rd32(data + 56);
    out->audio_packets = rd32(data + 60);

    uint64_t pos = NUV_FILE_HEADER;
    int done = !out->is_mythtv && !out->video_packets;
    while (!done) {
        if (pos + 1 > len) return RTJGPU_E_HEADER;
        const uint8_t type = data[pos];
        uint32_t size = 0;
        if (type == 'S') {                                   /* seek point: eleven more bytes, no size field */
            pos += 12;
            continue;
        }
        if (pos + NUV_HDRSIZE > len) return RTJGPU_E_HEADER;
        size = rd32(data + pos + 8) & 0xffffffu;
        if (pos + NUV_HDRSIZE + size > len) return RTJGPU_E_HEADER;
        if (type == 'D') {
            if (out->video_packets && data[pos + 1] == 'R' && size >= 512) {
                for (int i = 0; i < 128; i++) out->tables[i] = rd32(data + pos + NUV_HDRSIZE + 4 * i);
                out->has_tables = 1;
                if (!out->is_mythtv) done = 1;
            }
        } else if (type == 'X') {                            /* MythTV extended header closes the codec data */
            done = 1;
        }
        pos += NUV_HDRSIZE + size;
    }
    out->data_start = pos;
    return RTJGPU_OK;
}

/* lib/demux_nuv.c:246-318: one frame header at *pos.  Returns 1 and advances, 0 at the end of the data. */
int rtjnuv_next(const uint8_t *data, size_t len, uint64_t *pos, rtjnuv_packet *out)
{
    if (!data || !pos || !out) return 0;
    if (*pos + NUV_HDRSIZE > len) return 0;
    const uint8_t *h = data + *pos;
    out->type = h[0];
    out->comptype = h[1];
    out->keyframe = h[2];
    out->filters = h[3];
    out->timecode = rd32(h + 4);
    out->size = h[0] == 'S' ? 0 : (rd32(h + 8) & 0xffffffu);       /* a seek point's size field is not a size */
    out->payload_offset = *pos + NUV_HDRSIZE;
    if (out->payload_offset + out->size > len) return 0;
    *pos = out->payload_offset + out->size;
    return 1;
}

int rtjnuv_extract_rtj0(const uint8_t *data, size_t len, const rtjnuv_header *hdr, uint8_t *out, size_t out_cap,
                        uint64_t *offsets, uint32_t *timecodes, int max_frames, int *nframes, int *unsupported)
{
    if (!data || !hdr || !offsets || !nframes || max_frames < 0) return RTJGPU_E_ARG;
    const int w = hdr->width, h = hdr->height;
    if (w <= 0 || h <= 0 || (w & 15) || (h & 15) || w > 65535 || h > 65535) return RTJGPU_E_SIZE;
    const uint32_t nblk = (uint32_t)(w >> 4) * (uint32_t)(h >> 4) * 6u;
    uint64_t pos = hdr->data_start, o = 0;
    int n = 0, skipped = 0;
    rtjnuv_packet p;
    while (rtjnuv_next(data, len, &pos, &p)) {
        if (p.type != 'V') continue;                         /* audio, extradata, seek points */
        uint32_t payload;
        int repeat = 0;
        if (p.comptype == '1') payload = p.size;            /* RTjpeg block stream */
        else if (p.comptype == 'L') { payload = nblk; repeat = 1; }   /* "same as the last frame": every block skipped */
        else { skipped++; continue; }                        /* raw, LZO-packed and black frames are libavcodec's business */
        if (n >= max_frames) return RTJGPU_E_TOOBIG;
        const uint64_t need = RTJPEG_B200_HEADER_BYTES + (uint64_t)payload;
        offsets[n] = o;
        if (out) {
            if (o + need > out_cap) return RTJGPU_E_TOOBIG;
            uint8_t *q = out + o;
            wr32(q, (uint32_t)need);                         /* framesize, include/RTjpeg.h:102 */
            q[4] = RTJPEG_B200_HEADER_BYTES;                 /* headersize */
            q[5] = 0;                                        /* version */
            wr16(q + 6, (unsigned)w);
            wr16(q + 8, (unsigned)h);
            q[10] = 0;                                       /* quality 0: keep the tables that were loaded */
            q[11] = p.keyframe ? 0 : 1;                      /* key: 0 on key frames (ignored by the decoder) */
            if (repeat) memset(q + RTJPEG_B200_HEADER_BYTES, 0xFF, payload);
            else memcpy(q + RTJPEG_B200_HEADER_BYTES, data + p.payload_offset, payload);
        }
        if (timecodes) timecodes[n] = p.timecode;
        o += (need + 15) & ~(uint64_t)15;                    /* packets 16-byte aligned, as rtjgpu_plan likes them */
        n++;
    }
    offsets[n] = o;
    *nframes = n;
    if (unsupported) *unsupported = skipped;
    return RTJGPU_OK;
}
