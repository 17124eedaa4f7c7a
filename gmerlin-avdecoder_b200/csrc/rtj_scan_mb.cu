/*
 * rtj_scan_mb.cu -- K1, chunk-parallel flavour for frames whose tables carry a raw 8-bit prefix
 * (lb8 / cb8 != 0: quality > 170 or custom tables, lib/RTjpeg.c:2362-2367, :2389-2394).
 *
 * With a raw prefix the block grammar depends on the block's place in its macroblock (the four
 * luma blocks use lb8, U and V use cb8, lib/RTjpeg.c:2704-2739), so "the block that starts at p"
 * is not one function of p any more.  The scheme of rtj_scan_chunk.cu is kept, one level up:
 *
 *   level 0   TWO length tables, deltaL(p) and deltaC(p): the block that would start at p under
 *             the luma and under the chroma grammar (same SIMD-within-a-register token walk,
 *             shifted by the raw prefix and with the prefix taken off the positions to fill)
 *   compose   deltaMB(p) = length of a whole MACROBLOCK starting at p: L, L, L, L, C, C chained
 *   walk      chunks of MB_C bytes; for every position, where a parse starting a macroblock there leaves
 *             its chunk and how many macroblocks it starts (512 entries per chunk; a macroblock is at
 *             most 6 x 64 bytes long, so a parse can only land in the next chunk)
 *   chain     one thread, chunk to chunk; entries are macroblock boundaries, so no grammar
 *             state travels
 *   emit      every chunk lane walks its macroblocks and writes six entries for each
 *
 * Same outputs, counters and malformed-stream policy as rtj_scan_chunk_kernel.  Blocks coded
 * under a grammar without prefix (chroma at the standard tables) still go inline when short.
 */
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "rtj_common.h"

namespace {

constexpr int MB_THREADS = 128;
constexpr int MB_S = 4096;                          /* segment bytes */
constexpr int MB_C = 512;                           /* chunk bytes: a power of two >= the longest macroblock (384) */
constexpr int MB_NCH = MB_S / MB_C;                 /* 8 DP lanes */
constexpr int MB_MAXLEN = 6 * 64;                   /* longest macroblock */
constexpr int MB_DLA = MB_MAXLEN;                   /* positions behind the segment that still need a block length */
constexpr int MB_LA = MB_DLA + 64 + 64 + 32;        /* payload bytes behind the segment that level 0 may read */
constexpr int MB_RING = MB_C + 2;                   /* u16 per lane: 257 words (bank skew) */
constexpr int MB_STAGE = MB_NCH * MB_RING / 2;      /* staged entries per emit round (aliases the rings) */
constexpr int MB_POS = MB_S + MB_DLA;               /* positions with block lengths */
constexpr int MB_RUN = 36;                          /* positions per thread in level 0 */
constexpr int MB_PFX = 64;                          /* dense path: bytes per thread when the running sums are made */
constexpr int MB_PFX_THREADS = (MB_S + MB_LA + MB_PFX - 1) / MB_PFX;
static_assert(MB_PFX_THREADS <= MB_THREADS, "one thread per 64 bytes of running sums");
static_assert(MB_RUN * MB_THREADS >= MB_POS && (MB_RUN % 4) == 0 && ((MB_RUN / 4) & 1), "level-0 runs cover the positions; odd word stride");

struct MbShared {
    uint32_t pay[(MB_S + MB_LA + 16) / 4];
    uint32_t delL[MB_POS / 4];                      /* one byte per position */
    uint32_t delC[MB_POS / 4];
    uint16_t dmb[MB_S + 2 * MB_NCH];                /* macroblock length per position; chunk rows skewed by one word */
    uint32_t ring[MB_STAGE];                        /* u16 rings during DP/chain, staged u32 entries during emit */
    uint32_t base[MB_NCH + 1];
    uint16_t entq[MB_NCH];
    uint32_t tot[MB_PFX_THREADS + 1];               /* level 0, dense path: what each thread's 64 bytes fill */
    int entry, nb, skips, consumed;
};
/* level 0, dense path: running sums S (16 bit) of what the bytes fill, and T, the inverse of S over half a segment, in the
 * space of dmb and ring */
constexpr int MB_S_ENTRIES = MB_PFX_THREADS * MB_PFX;
constexpr int MB_T_CAP = ((int)(sizeof(uint16_t) * (MB_S + 2 * MB_NCH) + sizeof(uint32_t) * MB_STAGE) - 2 * MB_S_ENTRIES) / 2;
static_assert((offsetof(MbShared, dmb) + 2 * MB_S_ENTRIES) % 16 == 0 && MB_T_CAP % 8 == 0, "T is read and written sixteen bytes at a time");
static_assert(MB_T_CAP >= MB_POS / 2 + 64 + 63 + 256, "the inverse table covers half a segment of mostly one-place bytes");

__device__ __forceinline__ uint32_t swar_runs(uint32_t t) { return t & ~(t >> 1) & 0x40404040u; }
__device__ __forceinline__ uint32_t swar_x(uint32_t t) { return t & ((swar_runs(t) >> 6) * 0x3Fu); }

__device__ __forceinline__ uint32_t lds_u32_unaligned(const uint32_t *w, int byte)
{
    const uint32_t *p = w + (byte >> 2);
    return __funnelshift_r(p[0], p[1], (unsigned)(byte & 3) * 8);
}

/*
 * Level 0 for one grammar (raw prefix b), positions [q_begin, q_end) -- one thread, left to right.
 * The block that would start at q reads its tokens from ws = q + 1 + b on and ends behind the first
 * token that brings the filled positions to 63 - b.  Moving q one byte to the right drops one token
 * from the front of that window and (the fills being positive) can only push its end further right:
 * two pointers, a constant number of steps per position however long the blocks are.
 */
__device__ __forceinline__ void mb_level0_run(const uint8_t *__restrict__ payb, const uint8_t *__restrict__ fills,
                                              uint8_t *__restrict__ del, int q_begin, int q_end, int b)
{
    if (q_begin >= q_end) return;
    if (b >= 63) {                                          /* 63 raw coefficients: no token tail, 64 bytes */
        for (int q = q_begin; q < q_end; q++) del[q] = payb[q] == 0xFFu ? 1 : 64;
        return;
    }
    /* One step per iteration -- either the window grows by a token or a position is finished and the
     * window loses its first token -- so that the lanes of a warp, whose runs need the two kinds of
     * step in different order, never wait for each other.  fills[i] = positions the byte at i fills when read
     * as a token (a run token b > 63, lib/RTjpeg.c:173: b - 63; anything else 1), made once for both grammars. */
    const int need = 63 - b;
    int q = q_begin, ws = q_begin + 1 + b, n = ws, sum = 0;
    while (q < q_end) {
        const bool grow = sum < need;
        const int f = fills[grow ? n : ws];
        if (grow) {
            sum += f;
            n++;
        } else {
            /* skipped block (lib/RTjpeg.c:2704): a byte 0xFF is a block of its own, one byte long */
            del[q] = (uint8_t)(payb[q] == 0xFFu ? 1 : n - q);
            sum -= f;
            ws++;
            q++;
        }
    }
}

} // namespace

/* PHASE 0: the whole frame, segment after segment.  PHASE 1 / 2: one segment (blockIdx.x) of frame
 * blockIdx.y -- its summary for the frame-level chain, resp. its entries (rtj_common.h, rtj_seg_plan). */
/* FMT: the picture format fixes how many blocks a unit holds and how many of them are luma (rtj_common.h) */
template <int PHASE, int FMT>
__global__ void __launch_bounds__(MB_THREADS, 7)
rtj_scan_mb_kernel(const uint8_t *__restrict__ stream, const rtjgpu_frame_desc *__restrict__ desc,
                   const rtj_dev_table *__restrict__ tables, int F, int nblk,
                   uint32_t *__restrict__ ent, uint32_t *__restrict__ frame_skips,
                   rtj_dev_info *__restrict__ info, const rtj_seg_plan sp, int f0, int slice, const uint32_t *__restrict__ redo)
{
    constexpr int unit = RTJ_FMT_UNIT_BLOCKS(FMT), unit_luma = RTJ_FMT_UNIT_LUMA(FMT);
    extern __shared__ __align__(16) uint8_t mb_smem[];
    MbShared &sh = *reinterpret_cast<MbShared *>(mb_smem);
    const int tid = threadIdx.x, lane = tid & 31;
    const int f = (PHASE == 0 ? blockIdx.x : blockIdx.y) + f0;        /* F: one behind the last frame of this launch */
    if (f >= F) return;
    if (PHASE == 0 && redo) {
        /* behind the self-synchronising walks (rtj_scan_sync.cu): only the frames they marked as this kernel's; the marks are
         * final when the grids in front are done */
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (redo[f] != RTJ_REDO_MB) return;
    }
    const rtjgpu_frame_desc d = desc[f];
    const rtj_dev_table &tab = tables[min((int)d.table, RTJ_NUM_TABLES - 1)];          /* descriptors are the caller's memory */
    const int lb8 = tab.bt8[0], cb8 = tab.bt8[1];
    if ((lb8 | cb8) == 0) return;                       /* no raw prefix: rtj_scan_chunk_kernel's frame */

    const uint8_t *pay = stream + d.offset + RTJPEG_B200_HEADER_BYTES;
    const int len = d.length > RTJPEG_B200_HEADER_BYTES ? (int)d.length - RTJPEG_B200_HEADER_BYTES : 0;
    const int mis = (int)(reinterpret_cast<uintptr_t>(pay) & 15);
    const uint8_t *gbase = pay - mis;
    uint32_t *out = ent + (size_t)f * nblk;
    const uint8_t *payb = reinterpret_cast<const uint8_t *>(sh.pay) + mis;
    const uint8_t *dLb = reinterpret_cast<const uint8_t *>(sh.delL);
    const uint8_t *dCb = lb8 == cb8 ? dLb : reinterpret_cast<const uint8_t *>(sh.delC);
    const uint32_t missing = RTJ_ENT(min(len, (int)RTJ_ENT_OFF_MASK), 1);

    const size_t my_seg = (size_t)f * sp.maxseg + blockIdx.x;       /* PHASE 1 / 2 */
    if (PHASE == 2 && sp.base[my_seg] == RTJ_SEG_UNUSED) return;    /* the frame is complete before this segment */
    if (tid == 0) {
        sh.entry = PHASE == 2 ? (int)sp.entry[my_seg] : 0;
        sh.nb = PHASE == 2 ? (int)sp.base[my_seg] : 0;
        sh.skips = 0;
        sh.consumed = 0;
    }
    __syncthreads();

    for (int seg0 = PHASE == 0 ? 0 : (int)blockIdx.x * MB_S; seg0 < len; seg0 += MB_S) {
        const int nb0 = sh.nb;
        if (nb0 >= nblk) break;
        const int lim = len - seg0;
        const int nch = min(MB_NCH, (lim + MB_C - 1) / MB_C);
        const int npos = nch * MB_C;

        /* ---- load; beyond the payload every byte reads 0x7F, a run token that ends any block ---- */
        {
            const uint4 *g4 = reinterpret_cast<const uint4 *>(gbase + seg0);
            uint4 *s4 = reinterpret_cast<uint4 *>(sh.pay);
            const int nvec = (npos + MB_LA + 16) / 16;
            const int vlim = lim + mis;
            for (int v = tid; v < nvec; v += MB_THREADS) {
                const int b0 = v * 16;
                uint4 x = make_uint4(0x7F7F7F7Fu, 0x7F7F7F7Fu, 0x7F7F7F7Fu, 0x7F7F7F7Fu);
                if (b0 < vlim) {
                    x = __ldg(g4 + v);
                    if (b0 + 16 > vlim) {
                        uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int nv = vlim - (b0 + 4 * k);
                            const uint32_t m = nv >= 4 ? 0xFFFFFFFFu : nv <= 0 ? 0u : (1u << (8 * nv)) - 1u;
                            w[k] = (w[k] & m) | (0x7F7F7F7Fu & ~m);
                        }
                        x = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
                s4[v] = x;
            }
        }
        __syncthreads();

        /* ---- level 0: both block-length tables, for the segment and the look-ahead behind it; every
         *      thread owns a run of MB_RUN consecutive positions (17 words: bank-skewed).  The second pass of the
         *      segment-parallel arrangement takes the tables its first pass made, where they were kept. ---- */
        static_assert(2 * MB_POS <= RTJ_SEG_DEL_BYTES && (MB_POS % 16) == 0, "both tables fit the kept copy");
        uint4 *keep = (PHASE != 0 && sp.del) ? reinterpret_cast<uint4 *>(sp.del + my_seg * RTJ_SEG_DEL_BYTES) : nullptr;
        if (PHASE == 2 && keep) {
            const int nv = (lb8 != cb8 ? 2 : 1) * (MB_POS / 16);
            uint4 *d4 = reinterpret_cast<uint4 *>(sh.delL);         /* delL and delC are adjacent */
            for (int v = tid; v < nv; v += MB_THREADS) d4[v] = keep[v];
        } else {
            uint8_t *dL = reinterpret_cast<uint8_t *>(sh.delL), *dC = reinterpret_cast<uint8_t *>(sh.delC);
            const int npos2 = npos + MB_DLA;
            /* DENSE segments -- long blocks, nearly every byte a coefficient that fills one place (the high-quality stress
             * stream: two grammars, 54- and 63-token windows) -- take another way than the two-pointer walk below, which
             * spends some 35 instructions a position there.  With S[i] = places the bytes 0 .. i fill when read as tokens,
             * the block that would start at q ends with the first byte n for which S[n] - S[q + b] >= 63 - b: n is one look-up
             * in T, the inverse of S (T[v] = first i with S[i] >= v).  T has one entry per place filled, so it only fits
             * where runs are short: half a segment at a time, and only if both halves fit. */
            uint16_t *S = reinterpret_cast<uint16_t *>(sh.dmb);
            uint16_t *T = S + MB_S_ENTRIES;
            const int nby = npos + MB_LA;
            const int nthr = (nby + MB_PFX - 1) / MB_PFX;
            /* how many places the segment's bytes fill in all: what tells a dense segment from the usual kind (many times
             * as many places as bytes), before anything is spent on S */
            {
                uint32_t part = 0;
                for (int v = tid; v < (nby + 3) / 4; v += MB_THREADS) {
                    const uint32_t x = swar_x(lds_u32_unaligned(sh.pay, 4 * v + mis)) + 0x01010101u;
                    part += (x & 0x00FF00FFu) + ((x >> 8) & 0x00FF00FFu);              /* two 16-bit sums */
                }
                part = (part & 0xFFFFu) + (part >> 16);
#pragma unroll
                for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
                if (lane == 0) sh.tot[tid >> 5] = part;
            }
            __syncthreads();
            const uint32_t places = sh.tot[0] + sh.tot[1] + sh.tot[2] + sh.tot[3];
            /* rounds the dense path would take (a half or a quarter of the segment's positions each), 0: not dense */
            int rounds = places * 5u <= 8u * (unsigned)MB_T_CAP ? 2 : places * 5u <= 16u * (unsigned)MB_T_CAP ? 4 : 0;
            __syncthreads();                                                         /* tot is used again */
            if (rounds) {
                if (tid < nthr) {
                    uint32_t *S2 = reinterpret_cast<uint32_t *>(S) + tid * (MB_PFX / 2);
                    uint32_t acc = 0;
#pragma unroll 4
                    for (int wv = 0; wv < MB_PFX / 4; wv++) {
                        const uint32_t x = swar_x(lds_u32_unaligned(sh.pay, MB_PFX * tid + 4 * wv + mis)) + 0x01010101u;
                        const uint32_t a0 = acc + (x & 0xFFu), a1 = a0 + ((x >> 8) & 0xFFu), a2 = a1 + ((x >> 16) & 0xFFu);
                        acc = a2 + (x >> 24);
                        S2[2 * wv] = a0 | a1 << 16;
                        S2[2 * wv + 1] = a2 | acc << 16;
                    }
                    sh.tot[tid] = acc;
                }
                __syncthreads();
                if (tid < nthr && tid > 0) {
                    uint32_t off = 0;
                    for (int k = 0; k < tid; k++) off += sh.tot[k];
                    uint32_t *S2 = reinterpret_cast<uint32_t *>(S) + tid * (MB_PFX / 2);
#pragma unroll 4
                    for (int wv = 0; wv < MB_PFX / 2; wv++) S2[wv] += off * 0x00010001u;   /* places < 65536: no carry between the halves */
                }
                __syncthreads();
                /* every round's T must fit: places filled from the round's first position to its last window's end */
                for (;;) {
                    const int R = MB_POS / rounds;
                    bool fits = true;
                    for (int r = 0; r * R < npos2; r++)
                        fits = fits && (int)S[min(min((r + 1) * R, npos2) + 63 + 64, nby - 1)] - (int)S[r * R] < MB_T_CAP - 64;
                    if (fits || rounds == 4) { rounds = fits ? rounds : 0; break; }
                    rounds = 4;
                }
            }
            if (rounds) {
                const int R = MB_POS / rounds;
                for (int r = 0; r * R < npos2; r++) {
                    const int qlo = r * R, qhi = min(qlo + R, npos2);
                    const int v0 = S[qlo];
                    if (r) __syncthreads();                                          /* the round before is done with T */
                    /* T over the places (v0, v0 + cap): byte i covers the places S[i - 1] + 1 .. S[i].  Bytes read as run tokens
                     * cover up to 64 places and every fourth byte of noise reads as one, so T has some nine entries a byte: instead
                     * of a store per entry (a loop whose length differs from lane to lane), every byte marks the FIRST place it
                     * covers with its index, and a running maximum over T -- indices grow with the places -- fills in the rest:
                     * each warp a quarter of T, 256 entries (one 16-byte read a lane) at a time. */
                    const int i2max = min(qhi + 63 + 64, nby - 1);
                    const int nT8 = min(((int)S[i2max] - v0 + 1 + 7) >> 3, MB_T_CAP >> 3);      /* entries in use, in eights */
                    uint4 *T4 = reinterpret_cast<uint4 *>(T);
                    for (int j = tid; j < nT8; j += MB_THREADS) T4[j] = make_uint4(0u, 0u, 0u, 0u);
                    __syncthreads();
                    for (int i2 = qlo + 1 + tid; i2 <= i2max; i2 += MB_THREADS) {
                        const int first = (int)S[i2 - 1] - v0 + 1;
                        if (first < 8 * nT8) T[first] = (uint16_t)i2;
                    }
                    __syncthreads();
                    {
                        const int warp = tid >> 5;
                        const int per = ((nT8 + 4 * 32 - 1) / (4 * 32)) * 32;                  /* eights per warp: whole tiles of 32 */
                        const int j0 = warp * per, j1 = min(j0 + per, nT8);
                        /* what the quarters in front of this warp's hold at most */
                        uint32_t m = 0;
                        for (int j = j0 + lane; j < j1; j += 32) {
                            const uint4 w = T4[j];
                            m = __vmaxu2(__vmaxu2(m, w.x), __vmaxu2(__vmaxu2(w.y, w.z), w.w));
                        }
                        m = max(m & 0xFFFFu, m >> 16);
#pragma unroll
                        for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
                        if (lane == 0) sh.tot[warp] = m;
                        __syncthreads();
                        uint32_t carry = 0;                                                    /* the largest index in front of the tile */
                        for (int k = 0; k < warp; k++) carry = max(carry, sh.tot[k]);
                        for (int jt = j0; jt < j1; jt += 32) {
                            const int j = jt + lane;
                            uint4 w = j < j1 ? T4[j] : make_uint4(0u, 0u, 0u, 0u);
                            /* running maximum inside the lane's eight entries: low half first, then the high half over it */
                            uint32_t run = 0;
#define MB_RUNMAX(word)                                                                       \
                            {                                                                 \
                                const uint32_t lo16 = max(run, (word) & 0xFFFFu);             \
                                run = max(lo16, (word) >> 16);                                \
                                (word) = lo16 | run << 16;                                    \
                            }
                            MB_RUNMAX(w.x) MB_RUNMAX(w.y) MB_RUNMAX(w.z) MB_RUNMAX(w.w)
#undef MB_RUNMAX
                            /* ... over the lanes in front, and the tiles in front */
                            uint32_t incl = run;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                                if (lane >= o) incl = max(incl, up);
                            }
                            uint32_t before = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
                            before = max(lane ? before : 0u, carry);
                            const uint32_t b2 = before * 0x00010001u;
                            if (j < j1) T4[j] = make_uint4(__vmaxu2(w.x, b2), __vmaxu2(w.y, b2), __vmaxu2(w.z, b2), __vmaxu2(w.w, b2));
                            carry = max(carry, __shfl_sync(0xFFFFFFFFu, incl, 31));
                        }
                    }
                    __syncthreads();
                    for (int q = qlo + tid; q < qhi; q += MB_THREADS) {
                        const bool ff = payb[q] == 0xFFu;           /* skipped block (lib/RTjpeg.c:2704): a block of its own, one byte long */
                        /* the block's tokens start behind byte q + b and have 63 - b places to fill */
                        const int nL = lb8 >= 63 ? q + 63 : (int)T[(int)S[q + lb8] + 63 - lb8 - v0];
                        dL[q] = (uint8_t)(ff ? 1 : nL + 1 - q);
                        if (lb8 != cb8) {
                            const int nC = cb8 >= 63 ? q + 63 : (int)T[(int)S[q + cb8] + 63 - cb8 - v0];
                            dC[q] = (uint8_t)(ff ? 1 : nC + 1 - q);
                        }
                    }
                }
            } else {
                __syncthreads();                                                     /* S gives way to the fills */
                /* what every byte fills when read as a token, for both grammars: in the place of the macroblock lengths,
                 * which are made later.  fill = 1 + (b & 63) for a run token b = 64 .. 127, else 1: four bytes per step. */
                uint32_t *fw = reinterpret_cast<uint32_t *>(sh.dmb);
                for (int v = tid; v < (npos + MB_LA + 3) / 4; v += MB_THREADS)
                    fw[v] = swar_x(lds_u32_unaligned(sh.pay, 4 * v + mis)) + 0x01010101u;
                __syncthreads();
                const uint8_t *fills = reinterpret_cast<const uint8_t *>(fw);
                const int q_begin = tid * MB_RUN, q_end = min(q_begin + MB_RUN, npos2);
                mb_level0_run(payb, fills, dL, q_begin, q_end, lb8);
                if (lb8 != cb8) mb_level0_run(payb, fills, dC, q_begin, q_end, cb8);
            }
        }
        __syncthreads();
        if (PHASE == 1 && keep) {
            const int nv = (lb8 != cb8 ? 2 : 1) * (MB_POS / 16);
            const uint4 *d4 = reinterpret_cast<const uint4 *>(sh.delL);
            for (int v = tid; v < nv; v += MB_THREADS) keep[v] = d4[v];
        }

        /* The second pass of the segment-parallel arrangement knows where its segment is entered and how many blocks lie
         * before and in it (rtj_scan_plan_kernel): nothing has to be found out for every possible entry any more -- one
         * thread walks the segment's blocks from the known entry, all threads make the entries. */
        if (PHASE == 2 && tid == 0) {
            const uint32_t nxt = blockIdx.x + 1 < (unsigned)sp.maxseg ? sp.base[my_seg + 1] : RTJ_SEG_UNUSED;
            sh.nb = nxt != RTJ_SEG_UNUSED ? (int)nxt : sp.nbf[f];
            sh.entq[0] = (uint16_t)sh.entry;
            sh.base[0] = (uint32_t)nb0;
        }
        /* ---- compose: length of the macroblock that would start at every position ---- */
        for (int q = tid; q < (PHASE == 2 ? 0 : npos); q += MB_THREADS) {
            int n = q;
#pragma unroll
            for (int k = 0; k < unit_luma; k++) n += dLb[n];
#pragma unroll
            for (int k = unit_luma; k < unit; k++) n += dCb[n];
            sh.dmb[q + ((q / MB_C) << 1)] = (uint16_t)(n - q);
        }
        __syncthreads();

        /* ---- for every position of the segment: where a parse that starts a macroblock there leaves its chunk,
         *      and how many macroblocks it starts on the way.  Ring entry: exit offset (9 bits) | macroblocks << 9.
         *      Every position walks its own chain -- a macroblock is 6 to 384 bytes long, so a chain has few links
         *      (two or three in a dense stream), the walks are independent of each other and all threads take part;
         *      a right-to-left recurrence over the 512 positions of a chunk would keep one lane per chunk busy. ---- */
        {
            uint16_t *rings = reinterpret_cast<uint16_t *>(sh.ring);
            for (int q = tid; q < (PHASE == 2 ? 0 : npos); q += MB_THREADS) {
                const int c = q / MB_C, cend = (c + 1) * MB_C;
                uint32_t v = 0;                                             /* nothing starts behind the payload */
                if (q < lim) {
                    int n = q, cnt = 0;
                    for (;;) {
                        n += (int)sh.dmb[n + 2 * c];
                        cnt++;
                        if (n >= cend) { v = (uint32_t)(n - cend); break; }
                        if (n >= lim) break;                                /* the payload ends inside the chunk: exit 0 */
                    }
                    v |= (uint32_t)cnt << 9;
                }
                rings[c * MB_RING + (q & (MB_C - 1))] = (uint16_t)v;
            }
        }
        __syncthreads();

        if (PHASE == 1) {
            /* ---- summary: for every entry offset, where the parse leaves the segment and what it starts ---- */
            const uint16_t *rings = reinterpret_cast<const uint16_t *>(sh.ring);
            for (int e0 = tid; e0 < 384; e0 += MB_THREADS) {
                int e = e0, units = 0;
                for (int j = 0; j < nch; j++) {
                    const uint32_t v = rings[j * MB_RING + e];
                    e = (int)(v & 511u);
                    units += (int)(v >> 9);
                }
                sp.sum[my_seg * RTJ_SEG_NE + e0] = (uint32_t)e | ((uint32_t)units << 9);
            }
            return;
        }

        /* ---- chain ---- */
        if (PHASE != 2 && tid == 0) {
            const uint16_t *rings = reinterpret_cast<const uint16_t *>(sh.ring);
            int e = sh.entry, nb = nb0;
            for (int j = 0; j < nch; j++) {
                sh.entq[j] = (uint16_t)(j * MB_C + e);
                sh.base[j] = (uint32_t)nb;
                const uint32_t v = rings[j * MB_RING + e];
                e = (int)(v & 511u);
                nb += unit * (int)(v >> 9);
            }
            sh.base[nch] = (uint32_t)nb;
            sh.entry = e;
            sh.nb = nb;
        }
        __syncthreads();

        /* ---- emit, in two steps as in rtj_scan_chunk.cu: the chunk lanes walk their macroblocks and note where every
         *      block starts and which place of its unit it has; all threads then make the entries and store them. ---- */
        const int nb1 = min(sh.nb, nblk);
        const int nwalk = PHASE == 2 ? 1 : nch;                            /* walkers: one per chunk, or one for the segment */
        int q = 0, i = 0, qend = 0, myskips = 0, lastend = -1, k6 = 0;
        if (tid < nwalk) {
            q = sh.entq[tid];
            i = (int)sh.base[tid];
            qend = PHASE == 2 ? npos : (tid + 1) * MB_C;
        }
        __syncthreads();
        uint16_t *starts = reinterpret_cast<uint16_t *>(sh.ring);          /* start (13 bits) | place in the unit << 13; 0xFFFF: missing */
        static_assert(MB_POS < (1 << 13) && RTJ_FMT_UNIT_BLOCKS(0) <= 7, "a block start and its place fit 16 bits");
        for (int r0 = nb0; r0 < nb1; r0 += 2 * MB_STAGE) {
            const int r1 = min(r0 + 2 * MB_STAGE, nb1);
            if (tid < nwalk) {
                /* a macroblock belongs to the chunk it starts in; its later blocks may lie behind the chunk */
                while ((k6 != 0 || q < min(qend, lim)) && i < r1) {
                    if (q >= lim) {
                        starts[i - r0] = 0xFFFFu;                           /* the payload ended inside this macroblock: */
                        lastend = max(lastend, seg0 + lim + 1);             /* a frame that wants more than it has is flagged */
                    } else {
                        starts[i - r0] = (uint16_t)(q | (k6 << 13));
                        q += (k6 < unit_luma ? dLb : dCb)[q];
                        lastend = seg0 + q;
                    }
                    i++;
                    k6 = k6 == unit - 1 ? 0 : k6 + 1;
                }
            }
            __syncthreads();
            for (int k = tid; k < r1 - r0; k += MB_THREADS) {
                const unsigned st = starts[k];
                uint32_t e = missing;
                if (st != 0xFFFFu) {
                    const int qq = (int)(st & 0x1FFFu), kk = (int)(st >> 13);
                    const int bt8 = kk < unit_luma ? lb8 : cb8;
                    const int dl = (kk < unit_luma ? dLb : dCb)[qq];
                    const uint32_t head = lds_u32_unaligned(sh.pay, qq + mis);
                    const uint32_t last = payb[qq + dl - 1];
                    const bool isff = (head & 0xFFu) == 0xFFu;
                    const int eob = bt8 >= 63 ? 64 : ((last - 64u) < 64u ? max(127 - (int)last, dl - 1) : 64);
                    const uint32_t t1 = (head >> 8) & 0xFFu, t2 = (head >> 16) & 0xFFu;
                    const uint32_t c1 = (eob >= 2 && (t1 - 64u) >= 64u) ? t1 : 0u;
                    const uint32_t c2 = (eob >= 3 && (t2 - 64u) >= 64u) ? t2 : 0u;
                    const uint32_t e_inl = RTJ_ENT_INLINE_BIT | (head & 0xFFu) | (c1 << 8) | (c2 << 16);
                    const uint32_t e_gen = RTJ_ENT(seg0 + qq, eob);
                    e = isff ? RTJ_ENT_SKIP : ((bt8 == 0 && eob <= 3) ? e_inl : e_gen);
                    myskips += isff ? 1 : 0;
                }
                out[r0 + k] = e;
            }
            __syncthreads();
        }
        /* the skip counts of all threads, the stream ends of the chunk lanes */
#pragma unroll
        for (int o = 16; o; o >>= 1) myskips += __shfl_xor_sync(0xFFFFFFFFu, myskips, o);
        if (lane == 0 && myskips) atomicAdd(&sh.skips, myskips);
        if (tid < 32) {                                     /* MB_NCH <= 32: the chunk lanes are warp 0 */
#pragma unroll
            for (int o = 16; o; o >>= 1) lastend = max(lastend, __shfl_xor_sync(0xFFFFFFFFu, lastend, o));
            if (lane == 0) sh.consumed = max(sh.consumed, lastend);
        }
        __syncthreads();
        if (PHASE != 0) break;
    }

    if (PHASE == 1) return;                             /* nothing of the payload lies in this segment */
    if (PHASE == 2) {
        /* this segment's share of the frame's counters; the segment holding the frame's last block closes the frame */
        if (tid == 0) {
            const int skips = sh.skips, nbf = sp.nbf[f], nb0 = (int)sp.base[my_seg];
            if (skips) {
                atomicAdd(&frame_skips[f], (uint32_t)skips);
                atomicAdd(&info->skipped_blocks, (unsigned long long)skips);
                atomicAdd(&info->slice_skips[slice], (unsigned)skips);
            }
            if (nb0 < nbf && nbf <= sh.nb) {
                const int consumed = sh.consumed;
                atomicAdd(&info->payload_bytes, (unsigned long long)min(consumed, len));
                if (nbf < nblk || consumed > len || len > (int)RTJGPU_MAX_PAYLOAD_BYTES) {
                    atomicAdd(&info->bad_frames, 1u);
                    atomicMin((unsigned int *)&info->first_bad_frame, (unsigned int)f);
                }
            }
        }
        return;
    }

    /* a frame whose stream ends early or mid-block: flag it, give the missing blocks a harmless entry */
    const int nbf = min(sh.nb, nblk);
    for (int b = nbf + tid; b < nblk; b += MB_THREADS) out[b] = missing;
    if (tid == 0) {
        const int consumed = sh.consumed, skips = sh.skips;
        frame_skips[f] = (uint32_t)skips;
        if (skips) {
            atomicAdd(&info->skipped_blocks, (unsigned long long)skips);
            atomicAdd(&info->slice_skips[slice], (unsigned)skips);
        }
        atomicAdd(&info->payload_bytes, (unsigned long long)min(consumed, len));
        if (nbf < nblk || consumed > len || len > (int)RTJGPU_MAX_PAYLOAD_BYTES) {
            atomicAdd(&info->bad_frames, 1u);
            atomicMin((unsigned int *)&info->first_bad_frame, (unsigned int)f);
        }
    }
}

namespace {

template <int FMT>
cudaError_t scan_mb_launch(const rtj_launch_args *a, int phase, int nblk, const uint32_t *redo, cudaStream_t st)
{
    const int nf = a->f1 - a->f0;
    const dim3 grid = phase == 0 ? dim3((unsigned)nf) : dim3((unsigned)a->seg.maxseg, (unsigned)nf);
    if (phase == 0) {
        /* One CTA per frame, behind the kernel that takes the frames without a raw prefix.  The two share nothing (every frame
         * is one kernel's or the other's), so this grid may start while that one's last CTAs are still at work: programmatic
         * stream serialisation, and no wait for the grid in front -- unless the kernels in front say which frames are left
         * (redo), which is known when they are done. */
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(MB_THREADS);
        cfg.dynamicSmemBytes = sizeof(MbShared);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, rtj_scan_mb_kernel<0, FMT>, a->d_stream, a->d_desc, a->d_tables, a->f1, nblk, a->d_ent,
                                                 a->d_frame_skips, a->d_info, a->seg, a->f0, a->slice, redo);
        if (e != cudaSuccess) return e;
    } else if (phase == 1)
        rtj_scan_mb_kernel<1, FMT><<<grid, MB_THREADS, sizeof(MbShared), st>>>(
            a->d_stream, a->d_desc, a->d_tables, a->f1, nblk, a->d_ent, a->d_frame_skips, a->d_info, a->seg, a->f0, a->slice, nullptr);
    else
        rtj_scan_mb_kernel<2, FMT><<<grid, MB_THREADS, sizeof(MbShared), st>>>(
            a->d_stream, a->d_desc, a->d_tables, a->f1, nblk, a->d_ent, a->d_frame_skips, a->d_info, a->seg, a->f0, a->slice, nullptr);
    return cudaGetLastError();
}

template <int FMT>
cudaError_t scan_mb_attr()
{
    cudaError_t e = cudaFuncSetAttribute(rtj_scan_mb_kernel<0, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MbShared));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rtj_scan_mb_kernel<1, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MbShared));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rtj_scan_mb_kernel<2, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MbShared));
    return e;
}

} // namespace

extern "C" int rtj_launch_scan_mb(const rtj_launch_args *a, int phase, const uint32_t *redo, void *stream)
{
    static_assert(MB_S == RTJ_SEG_BYTES_MB && MB_DLA <= RTJ_SEG_NE, "segment size and entry range are shared with the frame-level chain");
    const int nblk = RTJ_FMT_NBLK(a->fmt, a->w, a->h);
    cudaStream_t st = (cudaStream_t)stream;
    return (int)(a->fmt == 0 ? scan_mb_launch<0>(a, phase, nblk, redo, st) : a->fmt == 1 ? scan_mb_launch<1>(a, phase, nblk, redo, st)
                                                                           : scan_mb_launch<2>(a, phase, nblk, redo, st));
}

extern "C" int rtj_scan_mb_init(void)
{
    cudaError_t e = scan_mb_attr<0>();
    if (e == cudaSuccess) e = scan_mb_attr<1>();
    if (e == cudaSuccess) e = scan_mb_attr<2>();
    return e == cudaSuccess ? 0 : (int)e;
}
