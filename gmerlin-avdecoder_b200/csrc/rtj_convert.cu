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
#include "rtj_convert.cuh"

namespace {

constexpr int CV_THREADS = 256;
using namespace rtjcv;

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
