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
rtj_scan_warp_kernel(const uint8_t *__restrict__ stream, const rtjgpu_frame_desc *__restrict__ desc,
                const rtj_dev_table *__restrict__ tables, int F, int nblk,
                uint32_t *__restrict__ ent, uint32_t *__restrict__ frame_skips,
                rtj_dev_info *__restrict__ info, int raw_only)
{
    const int lane = threadIdx.x & 31;
    const int f = blockIdx.x * SCAN_WARPS + (threadIdx.x >> 5);
    if (f >= F) return;

    const rtjgpu_frame_desc d = desc[f];
    const uint8_t *pay = stream + d.offset + RTJPEG_B200_HEADER_BYTES;
    const int len = d.length > RTJPEG_B200_HEADER_BYTES ? (int)d.length - RTJPEG_B200_HEADER_BYTES : 0;
    const int lb8 = tables[d.table].bt8[0];
    const int cb8 = tables[d.table].bt8[1];
    if (raw_only && (lb8 | cb8) == 0) return;        /* rtj_scan_chunk_kernel has done this frame */
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
                    e_out = RTJ_ENT_SKIP;
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
/* K1, lane-serial flavour: one THREAD walks one frame                        */
/* ------------------------------------------------------------------------ */
/*
 * The block grammar is a serial state machine (lib/RTjpeg.c:157-186): where a block
 * ends depends on every token before it.  One lane walks one frame, four tokens
 * per step with SIMD-within-a-register arithmetic:
 *
 *   t                     four token bytes
 *   r = t & ~(t>>1) & 0x40404040        bit 6 of every byte of the form 01xxxxxx (a run token, 64..127)
 *   x = t & ((r>>6) * 0x3F)             run length - 1 in run bytes, 0 in coefficient bytes
 *   P = x * 0x01010101 + 0x04030201     byte k = positions filled by tokens 0..k  (each token fills 1 + x_k)
 *   c = (P + (128-need) * 0x01010101) & 0x80808080
 *                                       bit 7 of byte k set  <=>  tokens 0..k fill >= need positions
 *
 * so the first set bit of c names the block's last token.  Byte overflows and carries
 * can only occur at or after that first crossing and never disturb it.  A lane costs
 * ~1 issue slot per block, so thousands of frames parse at a sliver of the machine;
 * the price is latency (one dependent chain per frame), which large batches hide.
 * The warp-cooperative flavour above serves batches with few, large frames.
 */
namespace {

__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint32_t *__restrict__ base4, int byte_off)
{
    const uint32_t *wp = base4 + (byte_off >> 2);
    return __funnelshift_r(__ldg(wp), __ldg(wp + 1), (unsigned)(byte_off & 3) * 8);
}

struct LaneResult { int blk, skips, consumed; };

/* SWAR pieces: see the comment above.  swar_x: run length - 1 in run bytes, 0 elsewhere. */
__device__ __forceinline__ uint32_t swar_runs(uint32_t t) { return t & ~(t >> 1) & 0x40404040u; }
__device__ __forceinline__ uint32_t swar_x(uint32_t t, uint32_t r) { return t & ((r >> 6) * 0x3Fu); }

template <bool RAW>
__device__ __forceinline__ LaneResult lane_scan_frame(const uint8_t *__restrict__ pay, int len, int lb8, int cb8,
                                                     uint32_t *__restrict__ out, int nblk)
{
    const uint32_t *base4 = reinterpret_cast<const uint32_t *>(pay);     /* packets start 4-byte aligned */
    int o = 0, blk = 0, skips = 0, k6 = 0;
    /* Lanes of a warp move in lockstep, so one lane missing L1 stalls all 32 (and with 128 sector
     * look-ups per step somebody always misses).  A real load touches the sector each lane will need
     * ~40 blocks from now; its value is consumed four steps later, by when even a DRAM miss is back. */
    uint32_t touch0 = 0, touch1 = 0, touch2 = 0, touch3 = 0, sink = 0;
    while (blk < nblk && o < len) {
        const int bt8 = RAW ? (k6 < 4 ? lb8 : cb8) : 0;
        if (RAW) k6 = k6 == 5 ? 0 : k6 + 1;
        sink ^= touch3;
        touch3 = touch2; touch2 = touch1; touch1 = touch0;
        touch0 = __ldg(base4 + ((o + 160) >> 2));

        /* the block's first byte and its first eight tokens */
        uint32_t first, t0, t1;
        {
            const uint32_t *wp = base4 + (o >> 2);
            const unsigned sh = (unsigned)(o & 3) * 8;
            const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3);
            const uint32_t u0 = __funnelshift_r(w0, w1, sh), u1 = __funnelshift_r(w1, w2, sh),
                           u2 = __funnelshift_r(w2, w3, sh);
            first = u0 & 0xFFu;
            t0 = __funnelshift_r(u0, u1, 8);          /* bytes o+1 .. o+4 */
            t1 = __funnelshift_r(u1, u2, 8);          /* bytes o+5 .. o+8 */
        }
        int tok = o + 1 + bt8;
        if (RAW) {
            const uint32_t *wp = base4 + (tok >> 2);
            const unsigned sh = (unsigned)(tok & 3) * 8;
            const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
            t0 = __funnelshift_r(w0, w1, sh);
            t1 = __funnelshift_r(w1, w2, sh);
        }
        const bool isff = first == 0xFFu;             /* skipped block: one byte, lib/RTjpeg.c:2704 */
        int need = 63 - bt8;                          /* positions the token tail has to fill */

        /* eight tokens at once, no branches: the common block (<= 9 bytes) resolves here */
        const uint32_t K = (uint32_t)(128 - need) * 0x01010101u;
        const uint32_t r0 = swar_runs(t0), r1 = swar_runs(t1);
        const uint32_t P0 = swar_x(t0, r0) * 0x01010101u + 0x04030201u;
        const uint32_t P1 = swar_x(t1, r1) * 0x01010101u + 0x04030201u + (P0 >> 24) * 0x01010101u;
        const uint32_t c0 = (P0 + K) & 0x80808080u, c1 = (P1 + K) & 0x80808080u;
        uint32_t c = c0 ? c0 : c1, t = c0 ? t0 : t1, r = c0 ? r0 : r1;
        int ntok = c0 ? 0 : 4;
        if (!isff && c == 0 && need > 0) {            /* long block: keep going four tokens at a time */
            need -= (int)(P1 >> 24);
            ntok = 8;
            for (;;) {
                if (tok + ntok >= len + 64) { c = 0x80u; r = 0; break; }       /* runaway on a truncated frame */
                t = ld_u32_unaligned(base4, tok + ntok);
                r = swar_runs(t);
                const uint32_t P = swar_x(t, r) * 0x01010101u + 0x04030201u;
                c = (P + (uint32_t)(128 - need) * 0x01010101u) & 0x80808080u;
                if (c) break;
                need -= (int)(P >> 24);
                ntok += 4;
            }
        }
        const int bit = __ffs((int)c) - 1;            /* 7, 15, 23 or 31 */
        ntok += (bit >> 3) + 1;
        const uint32_t bk = (t >> (bit - 7)) & 0xFFu;
        /* positions >= eob are zero: a final run of n zeros ends at 64, so it starts at 64 - n */
        int eob = ((r >> (bit - 1)) & 1u) ? 63 - (int)(bk & 0x3Fu) : 64;
        if (need <= 0) { ntok = 0; eob = 64; }        /* 63 raw coefficients: no token tail */
        out[blk++] = isff ? RTJ_ENT_SKIP : RTJ_ENT(min(o, len), max(eob, 1));
        skips += isff;
        o = isff ? o + 1 : tok + ntok;
    }
    if (sink == 0x5eed5eedu && (touch0 ^ touch1 ^ touch2) == 0x0badf00du) skips = -1;     /* keeps the touch loads alive; never true in effect */
    LaneResult res = {blk, skips, o};
    return res;
}

} // namespace

extern "C" __global__ void __launch_bounds__(32)
rtj_scan_lane_kernel(const uint8_t *__restrict__ stream, const rtjgpu_frame_desc *__restrict__ desc,
                     const rtj_dev_table *__restrict__ tables, int F, int nblk,
                     uint32_t *__restrict__ ent, uint32_t *__restrict__ frame_skips,
                     rtj_dev_info *__restrict__ info, int raw_only)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const rtjgpu_frame_desc d = desc[f];
    const uint8_t *pay = stream + d.offset + RTJPEG_B200_HEADER_BYTES;
    const int len = d.length > RTJPEG_B200_HEADER_BYTES ? (int)d.length - RTJPEG_B200_HEADER_BYTES : 0;
    const int lb8 = tables[d.table].bt8[0];
    const int cb8 = tables[d.table].bt8[1];
    if (raw_only && (lb8 | cb8) == 0) return;        /* rtj_scan_chunk_kernel has done this frame */
    uint32_t *out = ent + (size_t)f * nblk;

    const LaneResult res = (lb8 == 0 && cb8 == 0) ? lane_scan_frame<false>(pay, len, 0, 0, out, nblk)
                                                  : lane_scan_frame<true>(pay, len, lb8, cb8, out, nblk);

    /* a frame whose stream ends early or mid-block: flag it, give the missing blocks a harmless entry */
    const bool bad = res.blk < nblk || res.consumed > len;
    for (int b = res.blk; b < nblk; b++) out[b] = RTJ_ENT(len, 1);
    frame_skips[f] = (uint32_t)res.skips;
    if (res.skips) atomicAdd(&info->skipped_blocks, (unsigned long long)res.skips);
    atomicAdd(&info->payload_bytes, (unsigned long long)min(res.consumed, len));
    if (bad || len > (int)RTJGPU_MAX_PAYLOAD_BYTES) {
        atomicAdd(&info->bad_frames, 1u);
        atomicMin((unsigned int *)&info->first_bad_frame, (unsigned int)f);
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
            if (RTJ_ENT_IS_SKIP(e[j])) src[(size_t)(f + j) * nblk + b] = (uint16_t)last;
            else last = (unsigned)(f + j);
        }
    }
    for (; f < F; f++) {
        const uint32_t e = ent[(size_t)f * nblk + b];
        if (RTJ_ENT_IS_SKIP(e)) src[(size_t)f * nblk + b] = (uint16_t)last;
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
 * Byte source of one block for the sparse classes: the first 4*NW bytes sit in
 * registers (aligned 32-bit loads, funnel-shifted to the block's byte offset) and
 * leave through the low byte, so the token walk issues no dependent loads.
 */
template <int NW>
struct RegBytes {
    uint32_t u[NW];
    __device__ __forceinline__ explicit RegBytes(const uint8_t *__restrict__ src)
    {
        const uintptr_t a = reinterpret_cast<uintptr_t>(src);
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
        const unsigned sh = (unsigned)(a & 3) * 8;
        uint32_t w[NW + 1];
#pragma unroll
        for (int i = 0; i <= NW; i++) w[i] = __ldg(wp + i);      /* independent loads; slack bytes follow the stream */
#pragma unroll
        for (int i = 0; i < NW; i++) u[i] = __funnelshift_r(w[i], w[i + 1], sh);
    }
    __device__ __forceinline__ int peek_u8() const { return (int)(u[0] & 0xFFu); }
    __device__ __forceinline__ int peek_s8() const { return (int)(signed char)(u[0] & 0xFFu); }
    __device__ __forceinline__ void advance(bool take)
    {
        const unsigned sh = take ? 8u : 0u;
#pragma unroll
        for (int i = 0; i < NW - 1; i++) u[i] = __funnelshift_r(u[i], u[i + 1], sh);
        u[NW - 1] >>= sh;
    }
};

/* Byte source for dense blocks: straight from global memory, one byte at a time. */
struct MemBytes {
    const uint8_t *q;
    __device__ __forceinline__ explicit MemBytes(const uint8_t *__restrict__ src) : q(src) {}
    __device__ __forceinline__ int peek_u8() const { return (int)__ldg(q); }
    __device__ __forceinline__ int peek_s8() const { return (int)(signed char)__ldg(q); }
    __device__ __forceinline__ void advance(bool take) { q += take ? 1 : 0; }
};

/*
 * Decode one block whose zig-zag positions >= K are known to be zero.
 * iq holds the 64 multipliers in zig-zag order, bt8 is the raw-prefix length.
 * px receives 8 rows x 8 bytes.
 */
template <int K, typename Bytes>
__device__ __forceinline__ void decode_block(Bytes &by, const int *__restrict__ iq, int bt8, uint32_t (&px)[16])
{
    int m[8][8];
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 8; c++) m[r][c] = 0;

    /* DC is an unsigned byte (lib/RTjpeg.c:163); +4 is DESCALE's rounding term, which
     * reaches every output unchanged because the DC path has no multiply. */
    m[0][0] = wrap16(by.peek_u8() * iq[0]) + 4;
    by.advance(true);

    int z = 0;          /* zero positions still owed by the last run token */
#define RTJ_STEP(k, r, c)                                                   \
    if ((k) > 0 && (k) < K) {                                               \
        const bool take = z == 0;                                           \
        const int bb = by.peek_s8();                                        \
        const bool run = take && (k) > bt8 && bb > 63;                      \
        const int v = (take && !run) ? bb : 0;                              \
        z = take ? (run ? bb - 64 : 0) : z - 1;                             \
        by.advance(take);                                                   \
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

/*
 * The commonest block of real streams: at most DC, zig-zag 1 (row 1, column 0) and zig-zag 2
 * (row 0, column 1).  Column 1 of the first pass is then constant down the rows, so the odd
 * half of every ROW pass is the same eight values D[j], and pixel (r, j) = A[r] + D[j] with
 * A = the column-0 pass.  When |A| + |D| provably stays inside int16 (always, for streams an
 * encoder made), the final butterfly add, DESCALE and the 16..235 clamp run two pixels per
 * instruction: VIADDMNMX.S16x2 (add, lower clamp at 16*8), VIMNMX.S16x2 (upper clamp at
 * 235*8+7), one 32-bit shift by 3 for both halves, one PRMT per four pixels.  Clamping before
 * the shift is exact because >>3 is monotone; the int16 narrowing of DESCALE is the identity
 * inside the bound.  Outside the bound the same sums take the exact 32-bit epilogue.
 */
__device__ __forceinline__ void t2_pixels(int x0, int x1, int q, uint32_t (&px)[16]);

__device__ __forceinline__ void decode_block_t2(const uint8_t *__restrict__ src, const int *__restrict__ iq,
                                                int bt8, uint32_t (&px)[16])
{
    RegBytes<1> by(src);
    const int x0 = wrap16(by.peek_u8() * iq[0]) + 4;            /* +4: DESCALE's rounding term */
    by.advance(true);
    int xs[2];
    int z = 0;
#pragma unroll
    for (int k = 1; k <= 2; k++) {
        const bool take = z == 0;
        const int bb = by.peek_s8();
        const bool run = take && k > bt8 && bb > 63;
        const int v = (take && !run) ? bb : 0;
        z = take ? (run ? bb - 64 : 0) : z - 1;
        by.advance(take);
        xs[k - 1] = wrap16(v * iq[k]);
    }
    t2_pixels(x0, xs[0], xs[1], px);
}

/* the same block from an inline entry (rtj_common.h): coefficients already separated by K1 */
__device__ __forceinline__ void decode_block_inline(uint32_t e, const int *__restrict__ iq, uint32_t (&px)[16])
{
    const int x0 = wrap16((int)(e & 0xFFu) * iq[0]) + 4;
    const int x1 = wrap16((int)(signed char)((e >> 8) & 0xFFu) * iq[1]);
    const int q = wrap16((int)(signed char)((e >> 16) & 0xFFu) * iq[2]);
    t2_pixels(x0, x1, q, px);
}

/* x0 = dequantised DC + 4, x1 = zig-zag 1 (row 1, column 0), q = zig-zag 2 (row 0, column 1) */
__device__ __forceinline__ void t2_pixels(int x0, int x1, int q, uint32_t (&px)[16])
{
    /* pass 1, column 0: inputs (x0, x1, 0, ...): even half = x0, odd half from x1 alone */
    int A[8];
    {
        const int z5 = fxmul(x1, 473);
        const int o6 = z5 - x1;                                  /* o12 = fxmul(0,-669) + z5 = z5 */
        const int o5 = fxmul(x1, 362) - o6;
        const int o4 = fxmul(x1, 277) - z5 + o5;
        A[0] = x0 + x1; A[7] = x0 - x1;
        A[1] = x0 + o6; A[6] = x0 - o6;
        A[2] = x0 + o5; A[5] = x0 - o5;
        A[4] = x0 + o4; A[3] = x0 - o4;
    }
    /* pass 2: column 1 holds q in every row -> one odd half for all rows */
    int D[8];
    {
        const int z5 = fxmul(q, 473);
        const int o6 = z5 - q;
        const int o5 = fxmul(q, 362) - o6;
        const int o4 = fxmul(q, 277) - z5 + o5;
        D[0] = q; D[7] = -q; D[1] = o6; D[6] = -o6; D[2] = o5; D[5] = -o5; D[4] = o4; D[3] = -o4;
    }
    /* |A[r]| <= |x0| + |x1| + 3, |D[j]| <= |q| + 3 (every odd term is below |input| in magnitude) */
    if (abs(x0) + abs(x1) + abs(q) < 32000) {
        const uint32_t d01 = __byte_perm((uint32_t)D[0], (uint32_t)D[1], 0x5410);
        const uint32_t d23 = __byte_perm((uint32_t)D[2], (uint32_t)D[3], 0x5410);
        const uint32_t d45 = __byte_perm((uint32_t)D[4], (uint32_t)D[5], 0x5410);
        const uint32_t d67 = __byte_perm((uint32_t)D[6], (uint32_t)D[7], 0x5410);
        constexpr uint32_t LO = 0x00800080u;     /* 16 * 8 */
        constexpr uint32_t HI = 0x075F075Fu;     /* 235 * 8 + 7 */
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint32_t a2 = __byte_perm((uint32_t)A[r], 0u, 0x1010);
            const uint32_t t0 = __vmins2(__viaddmax_s16x2(a2, d01, LO), HI) >> 3;
            const uint32_t t1 = __vmins2(__viaddmax_s16x2(a2, d23, LO), HI) >> 3;
            const uint32_t t2 = __vmins2(__viaddmax_s16x2(a2, d45, LO), HI) >> 3;
            const uint32_t t3 = __vmins2(__viaddmax_s16x2(a2, d67, LO), HI) >> 3;
            px[2 * r] = __byte_perm(t0, t1, 0x6420);
            px[2 * r + 1] = __byte_perm(t2, t3, 0x6420);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            px[2 * r] = descale_pack4(A[r] + D[0], A[r] + D[1], A[r] + D[2], A[r] + D[3]);
            px[2 * r + 1] = descale_pack4(A[r] + D[4], A[r] + D[5], A[r] + D[6], A[r] + D[7]);
        }
    }
}

/* sparsity classes; the deferred ones are ordered most expensive first so long chunks start early */
enum { CLS_FULLG = 0, CLS_FULL, CLS_T4, CLS_T3, CLS_CARRY, NDEFER, CLS_T2 = NDEFER };

constexpr int IDCT_MAX_MB = 128;     /* macroblocks per CTA strip */
constexpr int IDCT_THREADS = 128;

struct IdctSmemHeader {
    int iq[2][64];
    int cnt[NDEFER];
    int next_chunk;
};

__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, unsigned bytes)
{
    /* TMA 1-D bulk copy shared -> global (UBLKCP) */
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}

/* where block i of the strip lives inside the shared-memory picture strip */
struct TileGeom {
    uint8_t *tileY, *tileU, *tileV;
    int segW, segC;
    __device__ __forceinline__ void store(int i, const uint32_t (&px)[16]) const
    {
        const int mb = i / 6, sub = i - mb * 6;
        uint8_t *dst;
        int pitch;
        if (sub < 4) {
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
};

} // namespace

extern "C" __global__ void __launch_bounds__(IDCT_THREADS, 5)
rtj_idct_kernel(const uint8_t *__restrict__ stream, const rtjgpu_frame_desc *__restrict__ desc,
                const rtj_dev_table *__restrict__ tables, const uint32_t *__restrict__ ent,
                const uint16_t *__restrict__ srcf, int nblk, int w, int h, int seg_mb, int nstrips,
                uint8_t *__restrict__ out, const uint8_t *__restrict__ carry)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARPS = IDCT_THREADS / 32;
    const int f = blockIdx.y;
    const int strip = blockIdx.x % nstrips, my = blockIdx.x / nstrips;
    const int mbw = w >> 4;
    const int mx0 = strip * seg_mb;
    const int mbs = min(seg_mb, mbw - mx0);
    const int nb = mbs * 6;

    TileGeom tile;
    tile.segW = mbs * 16;
    tile.segC = mbs * 8;
    tile.tileY = smem;
    tile.tileU = tile.tileY + 16 * tile.segW;
    tile.tileV = tile.tileU + 8 * tile.segC;
    IdctSmemHeader *hd = reinterpret_cast<IdctSmemHeader *>(tile.tileV + 8 * tile.segC);
    uint32_t *s_ent = reinterpret_cast<uint32_t *>(hd + 1);   /* [NDEFER][nb] entries of deferred blocks */
    uint16_t *s_idx = reinterpret_cast<uint16_t *>(s_ent + NDEFER * nb);   /* their strip index ... */
    uint16_t *s_src = s_idx + NDEFER * nb;                    /* ... and source frame */

    const rtjgpu_frame_desc fd = desc[f];
    const int mytable = fd.table;
    hd->iq[tid >> 6][tid & 63] = tables[mytable].iq[tid >> 6][tid & 63];   /* IDCT_THREADS == 128 entries */
    if (tid < NDEFER) hd->cnt[tid] = 0;
    if (tid == NDEFER) hd->next_chunk = 0;
    const int bt8_l = tables[mytable].bt8[0], bt8_c = tables[mytable].bt8[1];
    __syncthreads();

    /* ---- pass 1, stream order: the sparse majority (<= 3 coded positions) decodes right away,
     *      everything else is deferred into per-class lists ---- */
    const size_t strip_blk0 = (size_t)(my * mbw + mx0) * 6;
    const size_t frame_blk0 = (size_t)f * nblk + strip_blk0;
    const uint8_t *frame_pay = stream + fd.offset + RTJPEG_B200_HEADER_BYTES;
    for (int i0 = 0; i0 < nb; i0 += IDCT_THREADS) {
        const int i = i0 + tid;
        int cls = -1;
        uint32_t e = 0;
        unsigned sf = (unsigned)f;
        if (i < nb) {
            e = ent[frame_blk0 + i];
            if (RTJ_ENT_IS_SKIP(e)) {
                const unsigned s = srcf[frame_blk0 + i];
                if (s != RTJ_SRC_CARRY) {
                    sf = s;
                    e = ent[(size_t)s * nblk + strip_blk0 + i];
                }
            }
            const int eob = RTJ_ENT_IS_INLINE(e) ? 3 : RTJ_ENT_EOB(e);
            if (RTJ_ENT_IS_SKIP(e)) cls = CLS_CARRY;
            else if (sf != (unsigned)f && desc[sf].table != mytable) cls = CLS_FULLG;
            else if (eob <= 3) cls = CLS_T2;
            else if (eob <= 6) cls = CLS_T3;
            else if (eob <= 10) cls = CLS_T4;
            else cls = CLS_FULL;
        }
        if (cls == CLS_T2) {
            const int chroma = (i % 6) >= 4;
            uint32_t px[16];
            if (RTJ_ENT_IS_INLINE(e)) {
                decode_block_inline(e, hd->iq[chroma], px);
            } else {
                const uint8_t *src = (sf == (unsigned)f ? frame_pay : stream + desc[sf].offset + RTJPEG_B200_HEADER_BYTES)
                                     + (e & RTJ_ENT_OFF_MASK);
                decode_block_t2(src, hd->iq[chroma], chroma ? bt8_c : bt8_l, px);
            }
            tile.store(i, px);
        }
        const unsigned deferred = __ballot_sync(FULL, cls >= 0 && cls < NDEFER);
        if (deferred) {                                             /* warp-uniform */
#pragma unroll
            for (int c = 0; c < NDEFER; c++) {
                const unsigned m = __ballot_sync(FULL, cls == c);
                if (m == 0) continue;
                int slot = 0;
                const int leader = __ffs(m) - 1;
                if (lane == leader) slot = atomicAdd(&hd->cnt[c], __popc(m));
                slot = __shfl_sync(FULL, slot, leader);
                if (cls == c) {
                    const int at = c * nb + slot + __popc(m & ((1u << lane) - 1u));
                    s_ent[at] = e;
                    s_idx[at] = (uint16_t)i;
                    s_src[at] = (uint16_t)sf;
                }
            }
        }
    }
    __syncthreads();

    /* ---- pass 2: deferred blocks, one class-homogeneous chunk of 32 per warp step ---- */
    const size_t fsz = (size_t)w * h * 3 / 2;
    int total = 0;
#pragma unroll
    for (int c = 0; c < NDEFER; c++) total += (hd->cnt[c] + 31) >> 5;
    while (total > 0) {
        int ch = 0;
        if (lane == 0) ch = atomicAdd(&hd->next_chunk, 1);
        ch = __shfl_sync(FULL, ch, 0);
        if (ch >= total) break;
        int cls = 0, rel = ch;
#pragma unroll
        for (int c = 0; c < NDEFER; c++) {
            const int nc = (hd->cnt[c] + 31) >> 5;
            if (cls == c) { if (rel >= nc) { rel -= nc; cls = c + 1; } }
        }
        const int idx = rel * 32 + lane;
        if (idx >= hd->cnt[cls]) continue;
        const int at = cls * nb + idx;
        const int i = s_idx[at];
        const int sub = i % 6;
        const int chroma = sub >= 4;
        uint32_t px[16];

        if (cls == CLS_CARRY) {
            if (carry) {
                const int mb = i / 6;
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
            const uint32_t e = s_ent[at];
            const unsigned sf = s_src[at];
            const uint8_t *src = (sf == (unsigned)f ? frame_pay : stream + desc[sf].offset + RTJPEG_B200_HEADER_BYTES)
                                 + (e & RTJ_ENT_OFF_MASK);
            if (cls == CLS_FULLG) {
                const rtj_dev_table *t = &tables[desc[sf].table];
                if (RTJ_ENT_IS_INLINE(e)) {
                    decode_block_inline(e, t->iq[chroma], px);
                } else {
                    MemBytes by(src);
                    decode_block<64>(by, t->iq[chroma], t->bt8[chroma], px);
                }
            } else {
                const int *iq = hd->iq[chroma];
                const int bt8 = chroma ? bt8_c : bt8_l;
                if (cls == CLS_FULL) { MemBytes by(src); decode_block<64>(by, iq, bt8, px); }
                else if (cls == CLS_T4) { RegBytes<3> by(src); decode_block<10>(by, iq, bt8, px); }
                else { RegBytes<2> by(src); decode_block<6>(by, iq, bt8, px); }
            }
        }
        tile.store(i, px);
    }

    /* ---- the strip leaves the SM ---- */
    uint8_t *oy = out + (size_t)f * fsz + (size_t)(my * 16) * w + mx0 * 16;
    const int cw = w >> 1;
    uint8_t *ou = out + (size_t)f * fsz + (size_t)w * h + (size_t)(my * 8) * cw + mx0 * 8;
    uint8_t *ov = ou + (size_t)cw * (h >> 1);
    if (nstrips == 1) {
        /* full-width strip: 16 luma rows and 2 x 8 chroma rows are each one contiguous run in the
         * tight-pitch planes -> three TMA bulk stores issued by one thread */
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            bulk_store(oy, tile.tileY, 16u * (unsigned)tile.segW);
            bulk_store(ou, tile.tileU, 8u * (unsigned)tile.segC);
            bulk_store(ov, tile.tileV, 8u * (unsigned)tile.segC);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else {
        __syncthreads();
        const int vy = tile.segW >> 4;               /* 16-byte vectors per luma row */
        for (int r = warp; r < 16; r += NWARPS)
            for (int c = lane; c < vy; c += 32)
                *reinterpret_cast<uint4 *>(oy + (size_t)r * w + c * 16) =
                    *reinterpret_cast<const uint4 *>(tile.tileY + r * tile.segW + c * 16);
        const int vc = tile.segC >> 3;               /* 8-byte vectors per chroma row */
        for (int r = warp; r < 16; r += NWARPS) {
            const int pl = r >> 3, rr = r & 7;
            for (int c = lane; c < vc; c += 32)
                *reinterpret_cast<uint2 *>((pl ? ov : ou) + (size_t)rr * cw + c * 8) =
                    *reinterpret_cast<const uint2 *>((pl ? tile.tileV : tile.tileU) + rr * tile.segC + c * 8);
        }
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
    s += nb * NDEFER * (4 + 2 + 2);
    return (s + 15) & ~(size_t)15;
}

} // namespace

extern "C" int rtj_kernels_init(void)
{
    int nstrips;
    const size_t worst = idct_smem_bytes(idct_seg_mb(IDCT_MAX_MB, &nstrips));
    cudaError_t e = cudaFuncSetAttribute(rtj_idct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)worst);
    if (e != cudaSuccess) return (int)e;
    return rtj_scan_chunk_init();
}

extern "C" int rtj_launch_scan(const rtj_launch_args *a, void *stream)
{
    const int nblk = (a->w >> 4) * (a->h >> 4) * 6;
    cudaStream_t st = (cudaStream_t)stream;
    int launches = 0, raw_only = 0;
    /* AUTO: the chunk-parallel kernel takes every frame without a raw prefix, the serial kernels
     * the rest.  rtjgpu_set_scan_mode() forces one serial flavour for every frame. */
    if (a->scan_mode == RTJGPU_SCAN_AUTO || a->scan_mode == RTJGPU_SCAN_CHUNK) {
        int e = rtj_launch_scan_chunk(a, stream);
        if (e) return -e;
        launches++;
        raw_only = 1;
    }
    /* many frames: one lane per frame (cheap in issue slots, latency hidden by the batch);
     * few frames: one warp per frame. */
    const bool lane = a->scan_mode == RTJGPU_SCAN_LANE || (a->scan_mode != RTJGPU_SCAN_WARP && a->F >= 512);
    if (lane) {
        rtj_scan_lane_kernel<<<(a->F + 31) / 32, 32, 0, st>>>(
            a->d_stream, a->d_desc, a->d_tables, a->F, nblk, a->d_ent, a->d_frame_skips, a->d_info, raw_only);
    } else {
        const int grid = (a->F + SCAN_WARPS - 1) / SCAN_WARPS;
        rtj_scan_warp_kernel<<<grid, SCAN_WARPS * 32, 0, st>>>(
            a->d_stream, a->d_desc, a->d_tables, a->F, nblk, a->d_ent, a->d_frame_skips, a->d_info, raw_only);
    }
    launches++;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? launches : -(int)e;
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
    dim3 grid((unsigned)(nstrips * mbh), (unsigned)a->F);
    rtj_idct_kernel<<<grid, IDCT_THREADS, idct_smem_bytes(seg_mb), (cudaStream_t)stream>>>(
        a->d_stream, a->d_desc, a->d_tables, a->d_ent, a->d_src, nblk, a->w, a->h, seg_mb, nstrips,
        a->d_out, a->d_carry);
    return (int)cudaGetLastError();
}
