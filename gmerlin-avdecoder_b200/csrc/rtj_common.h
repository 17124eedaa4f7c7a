/*
 * rtj_common.h -- internal definitions shared by the host code and the kernels.
 */
#ifndef RTJ_COMMON_H
#define RTJ_COMMON_H

#include <stdint.h>
#include "../../include/rtjpeg_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- K1 output: one 32-bit entry per 8x8 block ------------------------------
 *   bits  0..24  byte offset of the block inside its frame's payload
 *   bits 25..31  end-of-block bound E (1..64): every zig-zag position >= E is
 *                zero.  E == 0 marks a block the stream skipped (0xFF marker).
 */
#define RTJ_ENT_OFF_BITS 25
#define RTJ_ENT_OFF_MASK ((1u << RTJ_ENT_OFF_BITS) - 1u)
#define RTJ_ENT(off, eob) ((uint32_t)(off) | ((uint32_t)(eob) << RTJ_ENT_OFF_BITS))

/* K3 output for skipped blocks: index of the last frame of the batch that
 * coded the block, or RTJ_SRC_CARRY when none has yet. */
#define RTJ_SRC_CARRY 0xFFFFu

#define RTJ_NUM_TABLES 257

/* Dequantisation tables as the kernels read them: zig-zag order (entry k is
 * the AAN-scaled multiplier of the k-th coefficient of the stream), luma then
 * chroma, plus the raw-prefix lengths lb8/cb8 (lib/RTjpeg.c:2362-2367). */
typedef struct rtj_dev_table {
    int32_t iq[2][64];
    int32_t bt8[2];
    int32_t pad[2];
} rtj_dev_table;

/* Host-side tables in the reference's own (raster) order. */
typedef struct rtj_host_table {
    int32_t liqt[64];
    int32_t ciqt[64];
    int     lb8, cb8;
} rtj_host_table;

extern const uint8_t rtj_zigzag[64];   /* position -> raster, lib/RTjpeg.c:59-74 */

void rtj_table_from_quality(int Q, rtj_host_table *out);           /* lib/RTjpeg.c:2344-2369 + 1208-1217 */
void rtj_table_from_raw(const uint32_t raw[128], rtj_host_table *out); /* lib/RTjpeg.c:2380-2395 */
void rtj_table_to_device_layout(const rtj_host_table *in, rtj_dev_table *out);

/* Device counters of one batch (lives in device memory, mirrored on request). */
typedef struct rtj_dev_info {
    unsigned long long skipped_blocks;
    unsigned long long payload_bytes;
    unsigned int       bad_frames;
    int                first_bad_frame;
} rtj_dev_info;

/* ---- kernel launchers (rtj_kernels.cu); stream is a cudaStream_t ----------- */
typedef struct rtj_launch_args {
    const uint8_t           *d_stream;
    const rtjgpu_frame_desc *d_desc;
    const rtj_dev_table     *d_tables;
    int                      F, w, h;
    uint32_t                *d_ent;         /* [F][nblk] */
    uint16_t                *d_src;         /* [F][nblk] */
    uint32_t                *d_frame_skips; /* [F] */
    rtj_dev_info            *d_info;
    uint8_t                 *d_out;
    const uint8_t           *d_carry;
    int                      scan_mode;     /* RTJGPU_SCAN_* */
} rtj_launch_args;

int rtj_launch_scan(const rtj_launch_args *a, void *stream);
int rtj_launch_resolve(const rtj_launch_args *a, void *stream);
int rtj_launch_idct(const rtj_launch_args *a, void *stream);
int rtj_kernels_init(void);   /* one-time function attributes (dynamic shared memory opt-in) */

#ifdef __cplusplus
}
#endif
#endif
