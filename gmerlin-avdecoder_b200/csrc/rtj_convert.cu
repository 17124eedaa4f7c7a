/*
 * rtj_convert.cu -- the colour converters of lib/RTjpeg.c (:3071-3486) for sm_100a: planar YUV frames that
 * K2 left in HBM become packed RGB there, without a trip through the host.
 *
 *   RTjpeg_yuv420rgb32 / bgr32   (:3123, :3192)   4 bytes per pixel, R G B x / B G R x
 *   RTjpeg_yuv420rgb24 / bgr24   (:3261, :3326)   3 bytes per pixel
 *   RTjpeg_yuv420rgb16           (:3391)          5-6-5, low byte first
 *   RTjpeg_yuv420rgb8            (:3477)          the luma plane, row by row
 *   RTjpeg_yuv422rgb24           (:3077)          as rgb24, chroma of full height
 *
 * Arithmetic is the reference's, bit for bit: 16 fractional bits, (Y - 16) * 76284, Cr/Cb - 128 times
 * 76284 / 53281 / 25625 / 132252 (:3071-3075), arithmetic shift, clamp to 0..255.  The work is elementwise
 * and bound by HBM: 1.5 bytes read and up to 4 written per pixel.  A thread owns 8 pixels of two picture rows
 * (one row for 4:2:2 and for the grey copy), so that every load is an aligned 8- or 4-byte word and every
 * store an aligned 8- or 16-byte vector, contiguous across the warp.
 *
 * One difference from the reference, by necessity of a batch interface that owns its output: the fourth byte
 * of a 32-bit pixel, which the reference steps over (:3147), is written with the caller's `alpha`.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "rtj_common.h"

namespace {

constexpr int KY = 76284, KCRR = 76284, KCRG = 53281, KCBG = 25625, KCBB = 132252;
constexpr int CV_THREADS = 256;

/* sat(a) << 8 | sat(b) in the low half, c's low half above it: two clamps to 0..255 and the packing in one I2IP */
__device__ __forceinline__ uint32_t pack_sat(int a, int b, uint32_t c)
{
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

struct Chroma { int r, g, b; };          /* the three chroma terms of a 2x1 / 2x2 neighbourhood */

__device__ __forceinline__ Chroma chroma_terms(uint32_t cb, uint32_t cr)
{
    Chroma c;
    c.r = ((int)cr - 128) * KCRR;
    c.g = ((int)cr - 128) * KCRG + ((int)cb - 128) * KCBG;
    c.b = ((int)cb - 128) * KCBB;
    return c;
}

/* 8 pixels of one row: luma bytes in yw (little endian, pixel 0 lowest), chroma terms per pixel pair.
 * KIND as in include/rtjpeg_b200.h (RTJ_CONV_*). */
template <int KIND>
__device__ __forceinline__ void row8(const uint2 yw, const Chroma (&c)[4], uint32_t alpha, uint8_t *__restrict__ o)
{
    int R[8], G[8], B[8];                /* before the clamp */
#pragma unroll
    for (int x = 0; x < 8; x++) {
        const uint32_t yb = ((x < 4 ? yw.x : yw.y) >> (8 * (x & 3))) & 0xFFu;
        const int y = ((int)yb - 16) * KY;
        R[x] = (y + c[x >> 1].r) >> 16;
        G[x] = (y - c[x >> 1].g) >> 16;
        B[x] = (y + c[x >> 1].b) >> 16;
    }
    if (KIND == RTJ_CONV_RGB32 || KIND == RTJ_CONV_BGR32) {
        uint32_t p[8];
#pragma unroll
        for (int x = 0; x < 8; x++) {
            const bool rgb = KIND == RTJ_CONV_RGB32;
            p[x] = pack_sat(G[x], rgb ? R[x] : B[x], pack_sat((int)alpha, rgb ? B[x] : R[x], 0u));
        }
        reinterpret_cast<uint4 *>(o)[0] = make_uint4(p[0], p[1], p[2], p[3]);
        reinterpret_cast<uint4 *>(o)[1] = make_uint4(p[4], p[5], p[6], p[7]);
    } else if (KIND == RTJ_CONV_RGB16) {
        uint32_t p[4];
#pragma unroll
        for (int x = 0; x < 8; x += 2) {
            const uint32_t gr0 = pack_sat(G[x], R[x], 0u), b0 = pack_sat(0, B[x], 0u);          /* G << 8 | R, B */
            const uint32_t gr1 = pack_sat(G[x + 1], R[x + 1], 0u), b1 = pack_sat(0, B[x + 1], 0u);
            const uint32_t lo = (b0 >> 3) | ((gr0 >> 10) << 5) | (((gr0 & 0xFFu) >> 3) << 11);
            const uint32_t hi = (b1 >> 3) | ((gr1 >> 10) << 5) | (((gr1 & 0xFFu) >> 3) << 11);
            p[x >> 1] = lo | hi << 16;
        }
        *reinterpret_cast<uint4 *>(o) = make_uint4(p[0], p[1], p[2], p[3]);
    } else {                             /* 24 bits: four pixels are twelve bytes, three words */
        const bool rgb = KIND != RTJ_CONV_BGR24;
        const int *F0 = rgb ? R : B, *F2 = rgb ? B : R;        /* first and third byte of a pixel */
        uint32_t wd[6];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int x = 4 * q;
            wd[3 * q] = pack_sat(G[x], F0[x], pack_sat(F0[x + 1], F2[x], 0u));
            wd[3 * q + 1] = pack_sat(F2[x + 1], G[x + 1], pack_sat(G[x + 2], F0[x + 2], 0u));
            wd[3 * q + 2] = pack_sat(F0[x + 3], F2[x + 2], pack_sat(F2[x + 3], G[x + 3], 0u));
        }
        reinterpret_cast<uint2 *>(o)[0] = make_uint2(wd[0], wd[1]);
        reinterpret_cast<uint2 *>(o)[1] = make_uint2(wd[2], wd[3]);
        reinterpret_cast<uint2 *>(o)[2] = make_uint2(wd[4], wd[5]);
    }
}

template <int KIND>
__global__ void __launch_bounds__(CV_THREADS)
rtj_convert_kernel(const uint8_t *__restrict__ src, size_t src_frame_bytes, int w, int h,
                   uint8_t *__restrict__ out, size_t row_pitch, size_t frame_pitch, uint32_t alpha)
{
    constexpr bool V422 = KIND == RTJ_CONV_YUV422_RGB24, COPY = KIND == RTJ_CONV_RGB8;
    constexpr int ROWS = (V422 || COPY) ? 1 : 2;                   /* picture rows per thread */
    constexpr int BPP = (KIND == RTJ_CONV_RGB32 || KIND == RTJ_CONV_BGR32) ? 4 : KIND == RTJ_CONV_RGB16 ? 2 : COPY ? 1 : 3;
    const int chunks = w >> 3;                                     /* 8-pixel chunks per row */
    const int item = blockIdx.x * CV_THREADS + threadIdx.x;
    if (item >= chunks * (h / ROWS)) return;
    const int rr = item / chunks, c = item - rr * chunks;
    const int row = rr * ROWS;
    const uint8_t *fy = src + (size_t)blockIdx.y * src_frame_bytes;
    uint8_t *fo = out + (size_t)blockIdx.y * frame_pitch + (size_t)row * row_pitch + (size_t)c * (8 * BPP);
    const uint2 y0 = *reinterpret_cast<const uint2 *>(fy + (size_t)row * w + c * 8);
    if (COPY) {
        *reinterpret_cast<uint2 *>(fo) = y0;
        return;
    }
    const int cw = w >> 1;
    const size_t ysz = (size_t)w * h, csz = V422 ? ysz >> 1 : ysz >> 2;
    const size_t co = (size_t)(V422 ? row : row >> 1) * cw + c * 4;
    const uint32_t ub = *reinterpret_cast<const uint32_t *>(fy + ysz + co);         /* planes[1] = Cb */
    const uint32_t vb = *reinterpret_cast<const uint32_t *>(fy + ysz + csz + co);   /* planes[2] = Cr */
    Chroma ct[4];
#pragma unroll
    for (int k = 0; k < 4; k++) ct[k] = chroma_terms((ub >> (8 * k)) & 0xFFu, (vb >> (8 * k)) & 0xFFu);
    row8<KIND>(y0, ct, alpha, fo);
    if (ROWS == 2) {
        const uint2 y1 = *reinterpret_cast<const uint2 *>(fy + (size_t)(row + 1) * w + c * 8);
        row8<KIND>(y1, ct, alpha, fo + row_pitch);
    }
}

template <int KIND>
cudaError_t convert_launch(const uint8_t *src, size_t sfb, int F, int w, int h, uint8_t *out, size_t rp, size_t fp,
                           uint32_t alpha, cudaStream_t st)
{
    const int rows = (KIND == RTJ_CONV_YUV422_RGB24 || KIND == RTJ_CONV_RGB8) ? h : h / 2;
    const int items = (w >> 3) * rows;
    dim3 grid((unsigned)((items + CV_THREADS - 1) / CV_THREADS), (unsigned)F);
    rtj_convert_kernel<KIND><<<grid, CV_THREADS, 0, st>>>(src, sfb, w, h, out, rp, fp, alpha);
    return cudaGetLastError();
}

} // namespace

/* Bytes per pixel of a converter's output, 0 for an unknown kind. */
extern "C" int rtj_convert_bpp(int kind)
{
    switch (kind) {
    case RTJ_CONV_RGB32: case RTJ_CONV_BGR32: return 4;
    case RTJ_CONV_RGB24: case RTJ_CONV_BGR24: case RTJ_CONV_YUV422_RGB24: return 3;
    case RTJ_CONV_RGB16: return 2;
    case RTJ_CONV_RGB8: return 1;
    default: return 0;
    }
}

extern "C" int rtj_launch_convert(int kind, const uint8_t *d_src, size_t src_frame_bytes, int F, int w, int h,
                                  uint8_t *d_out, size_t row_pitch, size_t frame_pitch, unsigned alpha, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t a = alpha & 0xFFu;
    switch (kind) {
    case RTJ_CONV_RGB32: return (int)convert_launch<RTJ_CONV_RGB32>(d_src, src_frame_bytes, F, w, h, d_out, row_pitch, frame_pitch, a, st);
    case RTJ_CONV_BGR32: return (int)convert_launch<RTJ_CONV_BGR32>(d_src, src_frame_bytes, F, w, h, d_out, row_pitch, frame_pitch, a, st);
    case RTJ_CONV_RGB24: return (int)convert_launch<RTJ_CONV_RGB24>(d_src, src_frame_bytes, F, w, h, d_out, row_pitch, frame_pitch, a, st);
    case RTJ_CONV_BGR24: return (int)convert_launch<RTJ_CONV_BGR24>(d_src, src_frame_bytes, F, w, h, d_out, row_pitch, frame_pitch, a, st);
    case RTJ_CONV_RGB16: return (int)convert_launch<RTJ_CONV_RGB16>(d_src, src_frame_bytes, F, w, h, d_out, row_pitch, frame_pitch, a, st);
    case RTJ_CONV_RGB8: return (int)convert_launch<RTJ_CONV_RGB8>(d_src, src_frame_bytes, F, w, h, d_out, row_pitch, frame_pitch, a, st);
    case RTJ_CONV_YUV422_RGB24: return (int)convert_launch<RTJ_CONV_YUV422_RGB24>(d_src, src_frame_bytes, F, w, h, d_out, row_pitch, frame_pitch, a, st);
    default: return (int)cudaErrorInvalidValue;
    }
}
