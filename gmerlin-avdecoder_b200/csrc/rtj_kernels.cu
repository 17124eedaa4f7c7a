/*
 * rtj_kernels.cu -- sm_100a kernels of the RTjpeg YUV420 decoder.
 *
 *   K1  rtj_scan_kernel     one warp walks one frame's run-length stream and
 *                           emits a 32-bit entry (payload offset + end-of-block
 *                           bound, or "skipped") for every 8x8 block.  Replaces
 *                           the `sp += RTjpeg_s2b(...)` pointer chase of
 *                           RTjpeg_decompressYUV420 (lib/RTjpeg.c:2701-2745) and
 *                           the length logic of RTjpeg_s2b (:157-186).
 *   K3  rtj_resolve_kernel  per block position, a last-writer scan over the
 *                           frames of the batch: for every skipped block, which
 *                           earlier frame coded it last.  Replaces the implicit
 *                           "skipped blocks keep the previous picture" state of
 *                           the reference (lib/video_rtjpeg.c:81 decodes every
 *                           packet into the same persistent frame).
 *   K2  rtj_idct_kernel     one CTA per (frame, macroblock row): unpack +
 *                           dequantise (RTjpeg_s2b value path, :162-183) +
 *                           integer AAN IDCT and clamp (RTjpeg_idct, :2209-2332)
 *                           into a shared-memory picture strip that leaves as
 *                           128-bit stores.  Blocks are bucketed by sparsity
 *                           class inside the CTA so that a warp runs one
 *                           specialised flow graph without divergence.
 *
 * All arithmetic is 32-bit integer and bit-exact with the reference: no tensor
 * cores, no floating point.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "rtj_common.h"

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;

/* ------------------------------------------------------------------------ */
/* K1: block-offset scan                                                      */
/* ------------------------------------------------------------------------ */

constexpr int SCAN_WARPS = 4;

} // namespace

extern "C" __global__ void __launch_bounds__(SCAN_WARPS * 32)
rtj_scan_kernel(const uint8_t *__restrict__ stream, const rtjgpu_frame_desc *__restrict__ desc,
                const rtj_dev_table *__restrict__ tables, int F, int nblk,
                uint32_t *__restrict__ ent, uint32_t *__restrict__ frame_skips,
                rtj_dev_info *__restrict__ info)
{
    const int lane = threadIdx.x & 31;
    const int f = blockIdx.x * SCAN_WARPS + (threadIdx.x >> 5);
    if (f >= F) return;

    const rtjgpu_frame_desc d = desc[f];
    const uint8_t *pay = stream + d.offset + RTJPEG_B200_HEADER_BYTES;
    const int len = d.length > RTJPEG_B200_HEADER_BYTES ? (int)d.length - RTJPEG_B200_HEADER_BYTES : 0;
    const int lb8 = tables[d.table].bt8[0];
    const int cb8 = tables[d.table].bt8[1];
    uint32_t *out = ent + (size_t)f * nblk;

    /* warp-uniform parser state */
    int pos = 0;        /* payload offset of lane 0's byte */
    int blk = 0;        /* blocks emitted so far */
    int k6 = 0;         /* index of the current block inside its macroblock, 0..5 */
    int rawleft = 0;    /* DC/raw bytes of the current block still to pass */
    int need = 0;       /* zig-zag positions still to fill by tokens (0: not inside a block) */
    int cur_off = 0;    /* payload offset of the current block */
    int skips = 0;
    int consumed = 0;
    uint32_t held = 0;  /* entry of block (blk & ~31) + lane, flushed every 32 blocks */

    while (blk < nblk) {
        /* one byte per lane; past the packet a 0x7F run token ends any block */
        const int at = pos + lane;
        const int b = at < len ? (int)pay[at] : 0x7F;
        const int sb = (int)(signed char)b;
        const bool isrun = sb > 63;
        const int a = isrun ? sb - 63 : 1;         /* positions this byte fills when read as a token */
        int S = a;                                   /* inclusive scan of a over the window */
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int up = __shfl_up_sync(FULL, S, o);
            if (lane >= o) S += up;
        }
        const int T = S - a;                         /* exclusive scan */
        const unsigned runmask = __ballot_sync(FULL, isrun);
        const int Slast = __shfl_sync(FULL, S, 31);

        int s = 0;                                   /* first unread lane of the window */
        while (s < 32 && blk < nblk) {
            bool done = false;
            uint32_t e_out = 0;
            if (rawleft == 0 && need == 0) {         /* at a block boundary */
                const int first = __shfl_sync(FULL, b, s);
                if (first == 0xFF) {                 /* skipped block: one byte, lib/RTjpeg.c:2704 */
                    s += 1;
                    skips++;
                    done = true;
                    e_out = RTJ_ENT(0, 0);
                } else {
                    const int bt8 = k6 < 4 ? lb8 : cb8;
                    cur_off = pos + s;
                    rawleft = 1 + bt8;               /* DC byte + raw 8-bit coefficients */
                    need = 63 - bt8;
                }
            }
            if (!done && rawleft > 0) {
                const int adv = min(rawleft, 32 - s);
                s += adv;
                rawleft -= adv;
                if (rawleft > 0) break;              /* raw prefix continues in the next window */
                if (need == 0) {                     /* 63 raw coefficients: no token tail */
                    done = true;
                    e_out = RTJ_ENT(min(cur_off, len), 64);
                } else if (s == 32) {
                    break;
                }
            }
            if (!done) {
                /* token tail: first lane e >= s where the filled positions reach `need` */
                const int base = __shfl_sync(FULL, T, s);
                const unsigned m = __ballot_sync(FULL, S - base >= need) & (FULL << s);
                if (m == 0) {                        /* block continues in the next window */
                    need -= Slast - base;
                    s = 32;
                    break;
                }
                const int e = __ffs(m) - 1;
                const int Te = __shfl_sync(FULL, T, e);
                /* positions >= eob are zero: a final run token starts at the bound */
                int eob = ((runmask >> e) & 1u) ? 64 - need + (Te - base) : 64;
                eob = max(1, min(eob, 64));
                e_out = RTJ_ENT(min(cur_off, len), eob);
                s = e + 1;
                need = 0;
                done = true;
            }
            if (done) {
                if (lane == (blk & 31)) held = e_out;
                blk++;
                k6 = k6 == 5 ? 0 : k6 + 1;
                if ((blk & 31) == 0) out[blk - 32 + lane] = held;
            }
        }
        consumed = pos + s;
        pos += 32;
    }
    if (lane < (blk & 31)) out[(blk & ~31) + lane] = held;

    if (lane == 0) {
        frame_skips[f] = (uint32_t)skips;
        if (skips) atomicAdd(&info->skipped_blocks, (unsigned long long)skips);
        atomicAdd(&info->payload_bytes, (unsigned long long)min(consumed, len));
        if (consumed > len || len > (int)RTJGPU_MAX_PAYLOAD_BYTES) {
            atomicAdd(&info->bad_frames, 1u);
            atomicMin((unsigned int *)&info->first_bad_frame, (unsigned int)f);
        }
    }
}

/* ------------------------------------------------------------------------ */
/* K3: last-writer resolution of skipped blocks                               */
/* ------------------------------------------------------------------------ */

extern "C" __global__ void __launch_bounds__(128)
rtj_resolve_kernel(const uint32_t *__restrict__ ent, uint16_t *__restrict__ src, int F, int nblk,
                   const rtj_dev_info *__restrict__ info)
{
    if (info->skipped_blocks == 0) return;           /* intra-only batch: nothing to resolve */
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    unsigned last = RTJ_SRC_CARRY;
    int f = 0;
    for (; f + 8 <= F; f += 8) {
        uint32_t e[8];
#pragma unroll
        for (int j = 0; j < 8; j++) e[j] = ent[(size_t)(f + j) * nblk + b];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if ((e[j] >> RTJ_ENT_OFF_BITS) == 0) src[(size_t)(f + j) * nblk + b] = (uint16_t)last;
            else last = (unsigned)(f + j);
        }
    }
    for (; f < F; f++) {
        const uint32_t e = ent[(size_t)f * nblk + b];
        if ((e >> RTJ_ENT_OFF_BITS) == 0) src[(size_t)f * nblk + b] = (uint16_t)last;
        else last = (unsigned)f;
    }
}

/* ------------------------------------------------------------------------ */
/* K2: unpack + dequantise + IDCT + store                                     */
/* ------------------------------------------------------------------------ */

namespace {

/* MULTIPLY of the reference (lib/RTjpeg.c:1206): 8 fractional bits, +128, arithmetic shift. */
__device__ __forceinline__ int fxmul(int v, int c) { return (v * c + 128) >> 8; }

/* low 16 bits, sign-extended: the `int16_t` stores of RTjpeg_s2b and DESCALE */
__device__ __forceinline__ int wrap16(int v) { return (int)(short)v; }

/* 8-point AAN flow graph shared by both passes (lib/RTjpeg.c:2240-2283, :2289-2326).
 * Inputs that are literal zeros fold away at compile time: fxmul(0, c) == 0. */
__device__ __forceinline__ void aan8(int x0, int x1, int x2, int x3, int x4, int x5, int x6, int x7,
                                     int (&y)[8])
{
    const int s04 = x0 + x4, d04 = x0 - x4;
    const int s26 = x2 + x6;
    const int m26 = fxmul(x2 - x6, 362) - s26;
    const int e0 = s04 + s26, e3 = s04 - s26, e1 = d04 + m26, e2 = d04 - m26;

    const int z13 = x5 + x3, z10 = x5 - x3, z11 = x1 + x7, z12 = x1 - x7;
    const int o7 = z11 + z13;
    const int o11 = fxmul(z11 - z13, 362);
    const int z5 = fxmul(z10 + z12, 473);
    const int o10 = fxmul(z12, 277) - z5;
    const int o12 = fxmul(z10, -669) + z5;
    const int o6 = o12 - o7;
    const int o5 = o11 - o6;
    const int o4 = o10 + o5;

    y[0] = e0 + o7; y[7] = e0 - o7;
    y[1] = e1 + o6; y[6] = e1 - o6;
    y[2] = e2 + o5; y[5] = e2 - o5;
    y[4] = e3 + o4; y[3] = e3 - o4;
}

/* Four row outputs (already carrying the +4 rounding term) -> four clamped bytes.
 * DESCALE (lib/RTjpeg.c:1200) narrows to int16 before RL (:1204) clamps to 16..235;
 * packing the low halves reproduces that narrowing exactly. */
__device__ __forceinline__ uint32_t descale_pack4(int y0, int y1, int y2, int y3)
{
    uint32_t a = __byte_perm((uint32_t)(y0 >> 3), (uint32_t)(y1 >> 3), 0x5410);
    uint32_t b = __byte_perm((uint32_t)(y2 >> 3), (uint32_t)(y3 >> 3), 0x5410);
    a = __vmaxs2(__vmins2(a, 0x00EB00EBu), 0x00100010u);
    b = __vmaxs2(__vmins2(b, 0x00EB00EBu), 0x00100010u);
    return __byte_perm(a, b, 0x6420);
}

/* zig-zag position k sits at (row, col): lib/RTjpeg.c:59-74 */
#define RTJ_ZZ_LIST(X) \
    X(0,0,0) X(1,1,0) X(2,0,1) X(3,0,2) X(4,1,1) X(5,2,0) X(6,3,0) X(7,2,1) \
    X(8,1,2) X(9,0,3) X(10,0,4) X(11,1,3) X(12,2,2) X(13,3,1) X(14,4,0) X(15,5,0) \
    X(16,4,1) X(17,3,2) X(18,2,3) X(19,1,4) X(20,0,5) X(21,0,6) X(22,1,5) X(23,2,4) \
    X(24,3,3) X(25,4,2) X(26,5,1) X(27,6,0) X(28,7,0) X(29,6,1) X(30,5,2) X(31,4,3) \
    X(32,3,4) X(33,2,5) X(34,1,6) X(35,0,7) X(36,1,7) X(37,2,6) X(38,3,5) X(39,4,4) \
    X(40,5,3) X(41,6,2) X(42,7,1) X(43,7,2) X(44,6,3) X(45,5,4) X(46,4,5) X(47,3,6) \
    X(48,2,7) X(49,3,7) X(50,4,6) X(51,5,5) X(52,6,4) X(53,7,3) X(54,7,4) X(55,6,5) \
    X(56,5,6) X(57,4,7) X(58,5,7) X(59,6,6) X(60,7,5) X(61,7,6) X(62,6,7) X(63,7,7)

/*
 * Decode one block whose zig-zag positions >= K are known to be zero.
 * src points at the block's DC byte, iq at the 64 multipliers in zig-zag order,
 * bt8 is the raw-prefix length.  px receives 8 rows x 8 bytes.
 */
template <int K>
__device__ __forceinline__ void decode_block(const uint8_t *__restrict__ src, const int *__restrict__ iq,
                                             int bt8, uint32_t (&px)[16])
{
    int m[8][8];
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 8; c++) m[r][c] = 0;

    /* DC is an unsigned byte (lib/RTjpeg.c:163); +4 is DESCALE's rounding term, which
     * reaches every output unchanged because the DC path has no multiply. */
    m[0][0] = wrap16((int)__ldg(src) * iq[0]) + 4;

    const uint8_t *q = src + 1;
    int z = 0;          /* zero positions still owed by the last run token */
#define RTJ_STEP(k, r, c)                                                   \
    if ((k) > 0 && (k) < K) {                                               \
        int v = 0;                                                          \
        if (z == 0) {                                                       \
            const int bb = (int)(signed char)__ldg(q);                      \
            q++;                                                            \
            if ((k) > bt8 && bb > 63) z = bb - 64; else v = bb;             \
        } else {                                                            \
            z--;                                                            \
        }                                                                   \
        m[r][c] = wrap16(v * iq[k]);                                        \
    }
    RTJ_ZZ_LIST(RTJ_STEP)
#undef RTJ_STEP

    /* pass 1: columns (lib/RTjpeg.c:2221-2285) */
    int ws[8][8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        int y[8];
        aan8(m[0][c], m[1][c], m[2][c], m[3][c], m[4][c], m[5][c], m[6][c], m[7][c], y);
#pragma unroll
        for (int r = 0; r < 8; r++) ws[r][c] = y[r];
    }
    /* pass 2: rows, descale, clamp (lib/RTjpeg.c:2287-2330) */
#pragma unroll
    for (int r = 0; r < 8; r++) {
        int y[8];
        aan8(ws[r][0], ws[r][1], ws[r][2], ws[r][3], ws[r][4], ws[r][5], ws[r][6], ws[r][7], y);
        px[2 * r] = descale_pack4(y[0], y[1], y[2], y[3]);
        px[2 * r + 1] = descale_pack4(y[4], y[5], y[6], y[7]);
    }
}

/* sparsity classes, most expensive first so that the long chunks start early */
enum { CLS_FULLG = 0, CLS_FULL, CLS_T4, CLS_T2, CLS_DC, CLS_CARRY, NCLS };

constexpr int IDCT_MAX_MB = 128;     /* macroblocks per CTA strip */

struct IdctSmemHeader {
    int iq[2][64];
    int cnt[NCLS + 1];
    int base[NCLS + 1];
    int cursor[NCLS + 1];
    int nchunk_total;
};

} // namespace

extern "C" __global__ void __launch_bounds__(256)
rtj_idct_kernel(const uint8_t *__restrict__ stream, const rtjgpu_frame_desc *__restrict__ desc,
                const rtj_dev_table *__restrict__ tables, const uint32_t *__restrict__ ent,
                const uint16_t *__restrict__ srcf, int nblk, int w, int h, int seg_mb, int nstrips,
                uint8_t *__restrict__ out, const uint8_t *__restrict__ carry)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int f = blockIdx.y;
    const int strip = blockIdx.x % nstrips, my = blockIdx.x / nstrips;
    const int mbw = w >> 4;
    const int mx0 = strip * seg_mb;
    const int mbs = min(seg_mb, mbw - mx0);
    const int nb = mbs * 6;
    const int segW = mbs * 16, segC = mbs * 8;

    uint8_t *tileY = smem;
    uint8_t *tileU = tileY + 16 * segW;
    uint8_t *tileV = tileU + 8 * segC;
    IdctSmemHeader *hd = reinterpret_cast<IdctSmemHeader *>(tileV + 8 * segC);
    uint32_t *s_ent = reinterpret_cast<uint32_t *>(hd + 1);
    uint16_t *s_src = reinterpret_cast<uint16_t *>(s_ent + nb);
    uint16_t *s_ord = s_src + nb;
    uint8_t *s_cls = reinterpret_cast<uint8_t *>(s_ord + nb);

    const rtjgpu_frame_desc fd = desc[f];
    const int mytable = fd.table;
    if (tid < 128) hd->iq[tid >> 6][tid & 63] = tables[mytable].iq[tid >> 6][tid & 63];
    if (tid < NCLS + 1) { hd->cnt[tid] = 0; hd->cursor[tid] = 0; }
    __syncthreads();

    /* ---- gather the strip's entries, resolve skipped blocks, classify ---- */
    const size_t frame_blk0 = (size_t)f * nblk + (size_t)(my * mbw + mx0) * 6;
    for (int i0 = 0; i0 < nb; i0 += blockDim.x) {
        const int i = i0 + tid;
        int cls = NCLS;
        if (i < nb) {
            uint32_t e = ent[frame_blk0 + i];
            unsigned sf = (unsigned)f;
            if ((e >> RTJ_ENT_OFF_BITS) == 0) {
                const unsigned s = srcf[frame_blk0 + i];
                if (s != RTJ_SRC_CARRY) {
                    sf = s;
                    e = ent[(size_t)s * nblk + (size_t)(my * mbw + mx0) * 6 + i];
                }
            }
            const int eob = (int)(e >> RTJ_ENT_OFF_BITS);
            if (eob == 0) cls = CLS_CARRY;
            else if (sf != (unsigned)f && desc[sf].table != mytable) cls = CLS_FULLG;
            else if (eob == 1) cls = CLS_DC;
            else if (eob <= 3) cls = CLS_T2;
            else if (eob <= 10) cls = CLS_T4;
            else cls = CLS_FULL;
            s_ent[i] = e;
            s_src[i] = (uint16_t)sf;
            s_cls[i] = (uint8_t)cls;
        }
        const unsigned peers = __match_any_sync(FULL, cls);
        if (lane == __ffs(peers) - 1) atomicAdd(&hd->cnt[cls], __popc(peers));
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0, chunks = 0;
        for (int c = 0; c < NCLS; c++) {
            hd->base[c] = acc;
            acc += hd->cnt[c];
            chunks += (hd->cnt[c] + 31) >> 5;
        }
        hd->nchunk_total = chunks;
    }
    __syncthreads();
    for (int i0 = 0; i0 < nb; i0 += blockDim.x) {
        const int i = i0 + tid;
        const int cls = i < nb ? (int)s_cls[i] : NCLS;
        const unsigned peers = __match_any_sync(FULL, cls);
        const int leader = __ffs(peers) - 1;
        int slot = 0;
        if (lane == leader) slot = atomicAdd(&hd->cursor[cls], __popc(peers));
        slot = __shfl_sync(FULL, slot, leader);
        if (i < nb) s_ord[hd->base[cls] + slot + __popc(peers & ((1u << lane) - 1u))] = (uint16_t)i;
    }
    __syncthreads();

    /* ---- decode: one class-homogeneous chunk of 32 blocks per warp step ---- */
    const size_t fsz = (size_t)w * h * 3 / 2;
    const int total = hd->nchunk_total;
    for (int ch = warp; ch < total; ch += nwarps) {
        int cls = 0, rel = ch;
        for (; cls < NCLS; cls++) {
            const int nc = (hd->cnt[cls] + 31) >> 5;
            if (rel < nc) break;
            rel -= nc;
        }
        const int idx = rel * 32 + lane;
        const bool act = idx < hd->cnt[cls];
        if (!act) continue;
        const int i = s_ord[hd->base[cls] + idx];
        const int mb = i / 6, sub = i - mb * 6;
        const int chroma = sub >= 4;
        uint32_t px[16];

        if (cls == CLS_CARRY) {
            if (carry) {
                const uint8_t *cp;
                int pitch;
                if (!chroma) {
                    pitch = w;
                    cp = carry + (size_t)(my * 16 + (sub >> 1) * 8) * w + (mx0 + mb) * 16 + (sub & 1) * 8;
                } else {
                    pitch = w >> 1;
                    cp = carry + (size_t)w * h + (sub == 5 ? (size_t)(w >> 1) * (h >> 1) : 0)
                         + (size_t)(my * 8) * pitch + (mx0 + mb) * 8;
                }
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    const uint2 v = *reinterpret_cast<const uint2 *>(cp + (size_t)r * pitch);
                    px[2 * r] = v.x;
                    px[2 * r + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int r = 0; r < 16; r++) px[r] = 0;
            }
        } else {
            const uint32_t e = s_ent[i];
            const unsigned sf = s_src[i];
            const uint64_t foff = sf == (unsigned)f ? fd.offset : desc[sf].offset;
            const uint8_t *src = stream + foff + RTJPEG_B200_HEADER_BYTES + (e & RTJ_ENT_OFF_MASK);
            if (cls == CLS_FULLG) {
                const rtj_dev_table *t = &tables[desc[sf].table];
                decode_block<64>(src, t->iq[chroma], t->bt8[chroma], px);
            } else {
                const int *iq = hd->iq[chroma];
                const int bt8 = tables[mytable].bt8[chroma];
                if (cls == CLS_FULL) decode_block<64>(src, iq, bt8, px);
                else if (cls == CLS_T4) decode_block<10>(src, iq, bt8, px);
                else if (cls == CLS_T2) decode_block<3>(src, iq, bt8, px);
                else decode_block<1>(src, iq, bt8, px);
            }
        }
        uint8_t *dst;
        int pitch;
        if (!chroma) {
            pitch = segW;
            dst = tileY + ((sub >> 1) * 8) * segW + mb * 16 + (sub & 1) * 8;
        } else {
            pitch = segC;
            dst = (sub == 4 ? tileU : tileV) + mb * 8;
        }
#pragma unroll
        for (int r = 0; r < 8; r++)
            *reinterpret_cast<uint2 *>(dst + r * pitch) = make_uint2(px[2 * r], px[2 * r + 1]);
    }
    __syncthreads();

    /* ---- the strip leaves as wide stores ---- */
    uint8_t *oy = out + (size_t)f * fsz + (size_t)(my * 16) * w + mx0 * 16;
    const int vy = segW >> 4;                        /* 16-byte vectors per luma row */
    for (int v = tid; v < 16 * vy; v += blockDim.x) {
        const int r = v / vy, c = v - r * vy;
        *reinterpret_cast<uint4 *>(oy + (size_t)r * w + c * 16) =
            *reinterpret_cast<const uint4 *>(tileY + r * segW + c * 16);
    }
    const int cw = w >> 1;
    uint8_t *ou = out + (size_t)f * fsz + (size_t)w * h + (size_t)(my * 8) * cw + mx0 * 8;
    uint8_t *ov = ou + (size_t)cw * (h >> 1);
    const int vc = segC >> 3;                        /* 8-byte vectors per chroma row */
    for (int v = tid; v < 2 * 8 * vc; v += blockDim.x) {
        const int pl = v >= 8 * vc;
        const int vv = pl ? v - 8 * vc : v;
        const int r = vv / vc, c = vv - r * vc;
        *reinterpret_cast<uint2 *>((pl ? ov : ou) + (size_t)r * cw + c * 8) =
            *reinterpret_cast<const uint2 *>((pl ? tileV : tileU) + r * segC + c * 8);
    }
}

namespace {

inline int idct_seg_mb(int mbw, int *nstrips)
{
    const int n = (mbw + IDCT_MAX_MB - 1) / IDCT_MAX_MB;
    *nstrips = n;
    return (mbw + n - 1) / n;
}

inline size_t idct_smem_bytes(int seg_mb)
{
    const size_t nb = (size_t)seg_mb * 6;
    size_t s = (size_t)seg_mb * 16 * 24;             /* Y 16 rows + U,V 8 rows of half width */
    s += sizeof(IdctSmemHeader);
    s += nb * (4 + 2 + 2 + 1);
    return (s + 15) & ~(size_t)15;
}

} // namespace

extern "C" int rtj_kernels_init(void)
{
    int nstrips;
    const size_t worst = idct_smem_bytes(idct_seg_mb(IDCT_MAX_MB, &nstrips));
    cudaError_t e = cudaFuncSetAttribute(rtj_idct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)worst);
    return e == cudaSuccess ? 0 : (int)e;
}

extern "C" int rtj_launch_scan(const rtj_launch_args *a, void *stream)
{
    const int nblk = (a->w >> 4) * (a->h >> 4) * 6;
    const int grid = (a->F + SCAN_WARPS - 1) / SCAN_WARPS;
    rtj_scan_kernel<<<grid, SCAN_WARPS * 32, 0, (cudaStream_t)stream>>>(
        a->d_stream, a->d_desc, a->d_tables, a->F, nblk, a->d_ent, a->d_frame_skips, a->d_info);
    return (int)cudaGetLastError();
}

extern "C" int rtj_launch_resolve(const rtj_launch_args *a, void *stream)
{
    const int nblk = (a->w >> 4) * (a->h >> 4) * 6;
    rtj_resolve_kernel<<<(nblk + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        a->d_ent, a->d_src, a->F, nblk, a->d_info);
    return (int)cudaGetLastError();
}

extern "C" int rtj_launch_idct(const rtj_launch_args *a, void *stream)
{
    const int mbw = a->w >> 4, mbh = a->h >> 4;
    const int nblk = mbw * mbh * 6;
    int nstrips;
    const int seg_mb = idct_seg_mb(mbw, &nstrips);
    const int threads = seg_mb * 6 >= 192 ? 256 : 128;
    dim3 grid((unsigned)(nstrips * mbh), (unsigned)a->F);
    rtj_idct_kernel<<<grid, threads, idct_smem_bytes(seg_mb), (cudaStream_t)stream>>>(
        a->d_stream, a->d_desc, a->d_tables, a->d_ent, a->d_src, nblk, a->w, a->h, seg_mb, nstrips,
        a->d_out, a->d_carry);
    return (int)cudaGetLastError();
}
