/*
 * rtj_shim.cpp -- Level 1 of include/rtjpeg_b200.h: the reference's codec API
 * (include/RTjpeg.h:115-139) on top of the batch context.
 *
 * One RTjpeg_t owns one rtjgpu_ctx (tables, workspace, streams) on the device
 * named by $RTJPEG_B200_DEVICE.  RTjpeg_decompress is the one-frame batch:
 * packet -> pinned staging -> device -> K1/K3/K2 -> pinned frame -> caller's
 * planes.  The reference leaves skipped blocks untouched in the caller's
 * planes (lib/RTjpeg.c:2704); the shim reproduces that on the host side by
 * copying back only the blocks the frame coded, so the caller's memory -- not a
 * device copy of it -- stays the single source of truth, exactly as in the
 * reference.
 */
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <algorithm>

#include "rtj_common.h"

extern "C" const rtj_host_table *rtj_ctx_host_table(const rtjgpu_ctx *ctx, int table);

namespace {

struct Instance {
    rtjgpu_ctx  *ctx = nullptr;
    rtjgpu_state st = {0, 0, RTJGPU_TABLE_ZERO, 0};
    int          format = RTJ_YUV420;
    int          key = 0, lm = 0, cm = 0;
    int          enc_quality = 0;              /* the quality the context's encoder tables stand for */
    int          err = 0;
    /* one-frame staging */
    uint8_t           *h_pkt = nullptr;  size_t h_pkt_cap = 0;    /* pinned */
    uint8_t           *h_frame = nullptr; size_t h_frame_cap = 0; /* pinned */
    rtjgpu_frame_desc *h_desc = nullptr;                          /* pinned */
    uint8_t           *d_pkt = nullptr;  size_t d_pkt_cap = 0;
    uint8_t           *d_frame = nullptr; size_t d_frame_cap = 0;
    rtjgpu_frame_desc *d_desc = nullptr;
    cudaStream_t       stream = nullptr;
    std::vector<uint32_t> entries;
};

int fail(Instance *in, int code) { in->err = code; return code; }

int ensure(Instance *in, size_t pkt_bytes, size_t frame_bytes)
{
    if (pkt_bytes > in->h_pkt_cap) {
        if (in->h_pkt) cudaFreeHost(in->h_pkt);
        if (in->d_pkt) cudaFree(in->d_pkt);
        in->h_pkt = nullptr; in->d_pkt = nullptr; in->h_pkt_cap = in->d_pkt_cap = 0;
        const size_t cap = pkt_bytes + pkt_bytes / 2;
        if (cudaMallocHost(&in->h_pkt, cap) != cudaSuccess) return RTJGPU_E_CUDA;
        if (cudaMalloc(&in->d_pkt, cap) != cudaSuccess) return RTJGPU_E_CUDA;
        in->h_pkt_cap = in->d_pkt_cap = cap;
    }
    if (frame_bytes > in->h_frame_cap) {
        if (in->h_frame) cudaFreeHost(in->h_frame);
        if (in->d_frame) cudaFree(in->d_frame);
        in->h_frame = nullptr; in->d_frame = nullptr; in->h_frame_cap = in->d_frame_cap = 0;
        if (cudaMallocHost(&in->h_frame, frame_bytes) != cudaSuccess) return RTJGPU_E_CUDA;
        if (cudaMalloc(&in->d_frame, frame_bytes) != cudaSuccess) return RTJGPU_E_CUDA;
        in->h_frame_cap = in->d_frame_cap = frame_bytes;
    }
    return RTJGPU_OK;
}

void copy_block(uint8_t *dst, const uint8_t *src, int pitch)
{
    for (int r = 0; r < 8; r++) memcpy(dst + (size_t)r * pitch, src + (size_t)r * pitch, 8);
}

} // namespace

extern "C" {

RTjpeg_t *RTjpeg_init(void)
{
    Instance *in = new (std::nothrow) Instance();
    if (!in) return nullptr;
    int dev = 0;
    if (const char *e = getenv("RTJPEG_B200_DEVICE")) dev = atoi(e);
    if (rtjgpu_create(dev, &in->ctx) != RTJGPU_OK) { delete in; return nullptr; }
    bool ok = cudaStreamCreateWithFlags(&in->stream, cudaStreamNonBlocking) == cudaSuccess
           && cudaMallocHost(&in->h_desc, sizeof(rtjgpu_frame_desc)) == cudaSuccess
           && cudaMalloc(&in->d_desc, sizeof(rtjgpu_frame_desc)) == cudaSuccess;
    if (!ok) { RTjpeg_close(in); return nullptr; }
    return in;
}

void RTjpeg_close(RTjpeg_t *rtj)
{
    Instance *in = static_cast<Instance *>(rtj);
    if (!in) return;
    if (in->stream) { cudaStreamSynchronize(in->stream); cudaStreamDestroy(in->stream); }
    if (in->h_pkt) cudaFreeHost(in->h_pkt);
    if (in->h_frame) cudaFreeHost(in->h_frame);
    if (in->h_desc) cudaFreeHost(in->h_desc);
    if (in->d_pkt) cudaFree(in->d_pkt);
    if (in->d_frame) cudaFree(in->d_frame);
    if (in->d_desc) cudaFree(in->d_desc);
    rtjgpu_destroy(in->ctx);
    delete in;
}

int RTjpeg_set_quality(RTjpeg_t *rtj, int *quality)
{
    Instance *in = static_cast<Instance *>(rtj);
    if (*quality < 1) *quality = 1;
    if (*quality > 255) *quality = 255;
    in->st.quality = *quality;
    in->st.table = *quality;
    return 0;
}

int RTjpeg_set_format(RTjpeg_t *rtj, int *format)
{
    static_cast<Instance *>(rtj)->format = *format;
    return 0;
}

int RTjpeg_set_size(RTjpeg_t *rtj, int *w, int *h)
{
    Instance *in = static_cast<Instance *>(rtj);
    if (*w < 0 || *w > 65535) return -1;
    if (*h < 0 || *h > 65535) return -1;
    in->st.width = *w;
    in->st.height = *h;
    return 0;
}

int RTjpeg_set_intra(RTjpeg_t *rtj, int *key, int *lm, int *cm)
{
    Instance *in = static_cast<Instance *>(rtj);
    if (*key < 0) *key = 0;
    if (*key > 255) *key = 255;
    if (*lm < 0) *lm = 0;
    if (*lm > 16) *lm = 16;
    if (*cm < 0) *cm = 0;
    if (*cm > 16) *cm = 16;
    in->key = *key; in->lm = *lm; in->cm = *cm;
    rtjgpu_encoder_set_intra(in->ctx, in->key, in->lm, in->cm);
    return 0;
}

void RTjpeg_get_tables(RTjpeg_t *rtj, uint32_t *tables)
{
    Instance *in = static_cast<Instance *>(rtj);
    const rtj_host_table *t = rtj_ctx_host_table(in->ctx, in->st.table);
    for (int i = 0; i < 64; i++) {
        tables[i] = (uint32_t)t->liqt[i];
        tables[64 + i] = (uint32_t)t->ciqt[i];
    }
}

void RTjpeg_set_tables(RTjpeg_t *rtj, uint32_t *tables)
{
    Instance *in = static_cast<Instance *>(rtj);
    if (in->stream) cudaStreamSynchronize(in->stream);
    if (rtjgpu_set_custom_tables(in->ctx, tables) != RTJGPU_OK) { in->err = RTJGPU_E_CUDA; return; }
    in->st.table = RTJGPU_TABLE_CUSTOM;      /* quality (rtj->Q) is left alone, lib/RTjpeg.c:2380-2395 */
}

} // extern "C"

namespace {

/* One picture through a device converter: planes -> pinned staging -> device -> converter -> pinned -> rows.
 * The 32-bit kinds run as their 24-bit twins and are spread out here, so that the fourth byte of the caller's
 * pixels stays as it is (lib/RTjpeg.c:3147). */
void convert_rows(Instance *in, int kind, uint8_t **planes, uint8_t **rows)
{
    if (!in || !planes || !rows) return;
    const int w = in->st.width, h = in->st.height;
    if (w <= 0 || h <= 0 || (w & 15) || (h & 15)) { in->err = RTJGPU_E_SIZE; return; }
    const bool wide = kind == RTJ_CONV_RGB32 || kind == RTJ_CONV_BGR32;
    const int dkind = kind == RTJ_CONV_RGB32 ? RTJ_CONV_RGB24 : kind == RTJ_CONV_BGR32 ? RTJ_CONV_BGR24 : kind;
    const int bpp = rtjgpu_convert_bpp(dkind);
    const size_t ysz = (size_t)w * h;
    const size_t csz = kind == RTJ_CONV_RGB8 ? 0 : kind == RTJ_CONV_YUV422_RGB24 ? ysz / 2 : ysz / 4;
    const size_t src_bytes = ysz + 2 * csz, row_bytes = (size_t)w * bpp, out_bytes = row_bytes * h;
    /* the staging pair holds source and result side by side */
    if (ensure(in, 16, src_bytes + out_bytes)) { in->err = RTJGPU_E_CUDA; return; }
    memcpy(in->h_frame, planes[0], ysz);
    if (csz) {
        memcpy(in->h_frame + ysz, planes[1], csz);
        memcpy(in->h_frame + ysz + csz, planes[2], csz);
    }
    cudaStream_t s = in->stream;
    uint8_t *d_rgb = in->d_frame + src_bytes, *h_rgb = in->h_frame + src_bytes;       /* src_bytes is a multiple of 64 */
    if (cudaMemcpyAsync(in->d_frame, in->h_frame, src_bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) { in->err = RTJGPU_E_CUDA; return; }
    const int rc = rtjgpu_convert_device(in->ctx, dkind, in->d_frame, src_bytes, 1, w, h, d_rgb, row_bytes, out_bytes, 0, s);
    if (rc) { in->err = rc; return; }
    if (cudaMemcpyAsync(h_rgb, d_rgb, out_bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess
        || cudaStreamSynchronize(s) != cudaSuccess) { in->err = RTJGPU_E_CUDA; return; }
    for (int r = 0; r < h; r++) {
        const uint8_t *src = h_rgb + (size_t)r * row_bytes;
        uint8_t *dst = rows[r];
        if (!wide) { memcpy(dst, src, row_bytes); continue; }
        for (int x = 0; x < w; x++) {
            dst[4 * x] = src[3 * x];
            dst[4 * x + 1] = src[3 * x + 1];
            dst[4 * x + 2] = src[3 * x + 2];
        }
    }
}

} // namespace

extern "C" {

/* include/RTjpeg.h:125, lib/RTjpeg.c:3488: one picture (tight planes of the size and format last set) -> one packet at
 * sp; returns its size.  The packet is byte for byte the reference's; 0 and a sticky error when the format is the
 * 8-bit one (see rtjgpu_encode_device) or the device fails. */
int RTjpeg_compress(RTjpeg_t *rtj, uint8_t *sp, uint8_t **planes)
{
    Instance *in = static_cast<Instance *>(rtj);
    if (!in || !sp || !planes) return 0;
    const int w = in->st.width, h = in->st.height, fmt = in->format;
    if (fmt != RTJ_YUV420 && fmt != RTJ_YUV422) { in->err = RTJGPU_E_FORMAT; return 0; }
    if (w <= 0 || h <= 0 || (w & 15) || (h & 15)) { in->err = RTJGPU_E_SIZE; return 0; }
    if (in->enc_quality != in->st.quality && in->st.quality) {          /* RTjpeg_set_quality, or a packet decoded since */
        if (rtjgpu_encoder_set_quality(in->ctx, in->st.quality)) { in->err = RTJGPU_E_CUDA; return 0; }
        in->enc_quality = in->st.quality;
    }
    const size_t ysz = (size_t)w * h, csz = fmt == RTJ_YUV420 ? ysz / 4 : ysz / 2, fsz = ysz + 2 * csz;
    const size_t cap = (RTJPEG_B200_HEADER_BYTES + (size_t)RTJ_FMT_NBLK(fmt, w, h) * 64 + 15) & ~(size_t)15;
    /* the staging pair holds the picture, then the packet, then the two offsets */
    if (ensure(in, 16, fsz + cap + 16)) { in->err = RTJGPU_E_CUDA; return 0; }
    memcpy(in->h_frame, planes[0], ysz);
    memcpy(in->h_frame + ysz, planes[1], csz);
    memcpy(in->h_frame + ysz + csz, planes[2], csz);
    cudaStream_t s = in->stream;
    uint8_t *d_pkt = in->d_frame + fsz;
    uint64_t *d_off = reinterpret_cast<uint64_t *>(in->d_frame + fsz + cap);
    if (cudaMemcpyAsync(in->d_frame, in->h_frame, fsz, cudaMemcpyHostToDevice, s) != cudaSuccess) { in->err = RTJGPU_E_CUDA; return 0; }
    rtjgpu_set_format(in->ctx, fmt);
    int rc = rtjgpu_encode_device(in->ctx, in->d_frame, 1, w, h, d_pkt, cap, d_off, s);
    if (rc) { in->err = rc; return 0; }
    uint64_t bytes = 0;
    int overflow = 0;
    if ((rc = rtjgpu_get_encode_info(in->ctx, &bytes, &overflow)) || overflow) { in->err = rc ? rc : RTJGPU_E_TOOBIG; return 0; }
    if (cudaMemcpyAsync(in->h_frame + fsz, d_pkt, (size_t)bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess
        || cudaStreamSynchronize(s) != cudaSuccess) { in->err = RTJGPU_E_CUDA; return 0; }
    const uint8_t *pk = in->h_frame + fsz;
    const uint32_t ds = (uint32_t)pk[0] | (uint32_t)pk[1] << 8 | (uint32_t)pk[2] << 16 | (uint32_t)pk[3] << 24;
    memcpy(sp, pk, ds);
    return (int)ds;
}

void RTjpeg_yuv420rgb32(RTjpeg_t *rtj, uint8_t **planes, uint8_t **rows) { convert_rows(static_cast<Instance *>(rtj), RTJ_CONV_RGB32, planes, rows); }
void RTjpeg_yuv420bgr32(RTjpeg_t *rtj, uint8_t **planes, uint8_t **rows) { convert_rows(static_cast<Instance *>(rtj), RTJ_CONV_BGR32, planes, rows); }
void RTjpeg_yuv420rgb24(RTjpeg_t *rtj, uint8_t **planes, uint8_t **rows) { convert_rows(static_cast<Instance *>(rtj), RTJ_CONV_RGB24, planes, rows); }
void RTjpeg_yuv420bgr24(RTjpeg_t *rtj, uint8_t **planes, uint8_t **rows) { convert_rows(static_cast<Instance *>(rtj), RTJ_CONV_BGR24, planes, rows); }
void RTjpeg_yuv420rgb16(RTjpeg_t *rtj, uint8_t **planes, uint8_t **rows) { convert_rows(static_cast<Instance *>(rtj), RTJ_CONV_RGB16, planes, rows); }
void RTjpeg_yuv420rgb8(RTjpeg_t *rtj, uint8_t **planes, uint8_t **rows) { convert_rows(static_cast<Instance *>(rtj), RTJ_CONV_RGB8, planes, rows); }
void RTjpeg_yuv422rgb24(RTjpeg_t *rtj, uint8_t **planes, uint8_t **rows) { convert_rows(static_cast<Instance *>(rtj), RTJ_CONV_YUV422_RGB24, planes, rows); }

int RTjpeg_b200_last_error(RTjpeg_t *rtj)
{
    Instance *in = static_cast<Instance *>(rtj);
    const int e = in->err;
    in->err = 0;
    return e;
}

int RTjpeg_b200_decompress_n(RTjpeg_t *rtj, const uint8_t *sp, size_t len, uint8_t **planes)
{
    Instance *in = static_cast<Instance *>(rtj);
    if (!in || !sp || !planes) return RTJGPU_E_ARG;
    const int fmt = in->format;
    if (fmt < RTJ_YUV420 || fmt > RTJ_RGB8) return fail(in, RTJGPU_E_FORMAT);
    if (len < RTJPEG_B200_HEADER_BYTES) return fail(in, RTJGPU_E_HEADER);

    rtjgpu_state st = in->st;
    const uint64_t offs[2] = {0, (uint64_t)len};
    rtjgpu_frame_desc desc;
    if (len > (size_t)RTJGPU_MAX_PAYLOAD_BYTES + RTJPEG_B200_HEADER_BYTES) return fail(in, RTJGPU_E_TOOBIG);
    const uint32_t len32 = (uint32_t)len;
    int rc = rtjgpu_plan_n(sp, offs, &len32, 1, &st, &desc);     /* sp may be unaligned: offset 0 is what is planned */
    if (rc) return fail(in, rc);
    const int w = st.width, h = st.height;
    const size_t fsz = RTJ_FMT_FRAME_BYTES(fmt, w, h);
    const int nblk = RTJ_FMT_NBLK(fmt, w, h);
    /* never more than the grammar can consume (64 bytes a block): a header whose framesize is far too large must not
     * make the host read far past the caller's packet before any block is looked at */
    const size_t plen = std::min<size_t>(desc.length, RTJPEG_B200_HEADER_BYTES + (size_t)nblk * 64);
    desc.length = (uint32_t)plen;

    if ((rc = ensure(in, plen + RTJGPU_STREAM_SLACK_BYTES, fsz))) return fail(in, rc);
    memcpy(in->h_pkt, sp, plen);
    memset(in->h_pkt + plen, 0x7F, RTJGPU_STREAM_SLACK_BYTES);
    *in->h_desc = desc;
    cudaStream_t s = in->stream;
    if (cudaMemcpyAsync(in->d_pkt, in->h_pkt, plen + RTJGPU_STREAM_SLACK_BYTES, cudaMemcpyHostToDevice, s) != cudaSuccess
        || cudaMemcpyAsync(in->d_desc, in->h_desc, sizeof(desc), cudaMemcpyHostToDevice, s) != cudaSuccess)
        return fail(in, RTJGPU_E_CUDA);
    rtjgpu_set_format(in->ctx, fmt);
    if ((rc = rtjgpu_decode_device(in->ctx, in->d_pkt, in->d_desc, 1, w, h, in->d_frame, nullptr, s)))
        return fail(in, rc);
    if (cudaMemcpyAsync(in->h_frame, in->d_frame, fsz, cudaMemcpyDeviceToHost, s) != cudaSuccess
        || cudaStreamSynchronize(s) != cudaSuccess)
        return fail(in, RTJGPU_E_CUDA);

    rtjgpu_batch_info bi;
    if ((rc = rtjgpu_get_batch_info(in->ctx, &bi))) return fail(in, rc);
    in->st = st;        /* the reference reconfigures before it decodes, so state advances even on a bad stream */

    /* plane p of the tight frame: Y, then U and V (quarter size in YUV420, half size in YUV422, none in grey) */
    const size_t ysz = (size_t)w * h, csz = fmt == RTJ_YUV420 ? ysz / 4 : fmt == RTJ_YUV422 ? ysz / 2 : 0;
    const int nplanes = fmt == RTJ_RGB8 ? 1 : 3;
    if (bi.skipped_blocks == 0) {
        memcpy(planes[0], in->h_frame, ysz);
        if (nplanes == 3) {
            memcpy(planes[1], in->h_frame + ysz, csz);
            memcpy(planes[2], in->h_frame + ysz + csz, csz);
        }
    } else {
        /* copy back only what this frame coded; skipped blocks keep the caller's pixels */
        in->entries.resize((size_t)nblk);
        if ((rc = rtjgpu_get_entries(in->ctx, in->entries.data(), (size_t)nblk))) return fail(in, rc);
        const int cw = w >> 1;
        const int unit = RTJ_FMT_UNIT_BLOCKS(fmt), unit_luma = RTJ_FMT_UNIT_LUMA(fmt), ux = RTJ_FMT_UNITS_X(fmt, w);
        for (int b = 0; b < nblk; b++) {
            if (RTJ_ENT_IS_SKIP(in->entries[(size_t)b])) continue;
            const int un = b / unit, sub = b - un * unit;
            const int uy = un / ux, uxx = un - uy * ux;
            if (sub < unit_luma) {
                /* YUV420: four luma blocks of a 16x16 macroblock; YUV422: two of a 16x8 unit; grey: the 8x8 block */
                const int rows = fmt == RTJ_YUV420 ? 16 : 8, uw = fmt == RTJ_RGB8 ? 8 : 16;
                const size_t o = (size_t)(uy * rows + (fmt == RTJ_YUV420 ? (sub >> 1) * 8 : 0)) * w
                               + (size_t)uxx * uw + (size_t)(fmt == RTJ_YUV420 ? (sub & 1) : sub) * 8;
                copy_block(planes[0] + o, in->h_frame + o, w);
            } else {
                const int pl = sub - unit_luma;                  /* 0 = U, 1 = V */
                const size_t o = (size_t)(uy * 8) * cw + (size_t)uxx * 8;
                copy_block(planes[1 + pl] + o, in->h_frame + ysz + (pl ? csz : 0) + o, cw);
            }
        }
    }
    if (bi.bad_frames) return fail(in, RTJGPU_E_OVERRUN);
    return RTJGPU_OK;
}

void RTjpeg_decompress(RTjpeg_t *rtj, uint8_t *sp, uint8_t **planes)
{
    /* the reference trusts the packet (lib/RTjpeg.c:3565-3586 reads no length);
     * the header's framesize field is the only length available here */
    uint32_t framesize = (uint32_t)sp[0] | (uint32_t)sp[1] << 8 | (uint32_t)sp[2] << 16 | (uint32_t)sp[3] << 24;
    if (framesize < RTJPEG_B200_HEADER_BYTES) {
        static_cast<Instance *>(rtj)->err = RTJGPU_E_HEADER;
        return;
    }
    RTjpeg_b200_decompress_n(rtj, sp, framesize, planes);
}

} // extern "C"
