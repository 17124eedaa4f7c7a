/*
 * rtj_scan_chunk.cu -- K1, chunk-parallel flavour: the block-offset scan of one frame
 * spread over a whole CTA.
 *
 * What it replaces: the `sp += RTjpeg_s2b(...)` / `sp++` pointer chase of
 * RTjpeg_decompressYUV420 (lib/RTjpeg.c:2701-2745) with the length rules of RTjpeg_s2b
 * (:157-186).  Where a block starts depends on every token before it, so the reference
 * -- and a one-thread-per-frame GPU scan -- walk the frame serially.  Here the chain is
 * cut into independent pieces:
 *
 *   level 0   for EVERY byte position p of the payload: delta(p) = length of the block
 *             that would start at p.  Purely local (<= 64 bytes ahead), one thread per
 *             four positions, SIMD-within-a-register over four tokens at a time.
 *   DP        the payload is cut into chunks of CS_C bytes, one lane per chunk.  Walking
 *             the chunk right to left, E(p) = E(p + delta(p)) gives for every p the place
 *             where a parse entering at p leaves the chunk and how many blocks it starts
 *             on the way.  Only a 64-entry ring per chunk is live, because delta <= 64.
 *   chain     one thread hops chunk to chunk through the rings: the true entry point
 *             and first block index of every chunk (CS_S / CS_C dependent steps instead
 *             of one per block).
 *   emit      every chunk lane walks its own blocks from its now known entry point and notes
 *             where each starts; the 32-bit entries are then made by all threads, one block
 *             each, and written to the table as coalesced stores.
 *
 * Frames of any size stream through in segments of CS_S bytes; the entry point and the
 * block count carry from segment to segment.
 *
 * Scope: frames whose tables have no raw 8-bit prefix (lb8 == cb8 == 0, i.e. quality
 * <= 170 and most custom tables).  Other frames are left to the serial kernels of
 * rtj_kernels.cu, which run right after this one and skip what is already done.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "rtj_common.h"

namespace {

constexpr int CS_THREADS = 128;
constexpr int CS_S = 8192;                         /* segment bytes */
constexpr int CS_C = 128;                          /* chunk bytes (>= 64, multiple of 64) */
constexpr int CS_NCH = CS_S / CS_C;                /* chunks per segment = DP lanes */
constexpr int CS_LA = 128;                         /* look-ahead bytes behind a segment */
constexpr int CS_RING = 82;                        /* u16 per chunk ring: 64 live + pad to 33 words (bank skew) */
constexpr int CS_STAGE = CS_NCH * CS_RING / 2;     /* staged entries per emit round (aliases the rings) */
constexpr int CS_PAY_WORDS = (CS_S + CS_LA + 16) / 4;
constexpr int CS_DEL_WORDS = CS_NCH * (CS_C / 4 + 1);   /* one pad word per chunk (bank skew) */

static_assert(CS_NCH <= CS_THREADS, "one DP lane per chunk");
static_assert((CS_C % 64) == 0 && CS_C >= 64, "ring indexing assumes chunk starts are multiples of 64");
static_assert((CS_C << 6) + 63 <= 0xFFFF, "ring entries are 16 bit: 6 bits exit offset + block count");

struct CsShared {
    uint32_t pay[CS_PAY_WORDS];        /* payload bytes of the segment (+ look-ahead), shifted by `mis` */
    uint32_t del[CS_DEL_WORDS];        /* delta(p), one byte per position, chunk rows padded by one word */
    uint32_t ring[CS_STAGE];           /* u16 rings during DP/chain, staged u32 entries during emit */
    uint32_t base[CS_NCH + 1];         /* first block index of every chunk */
    uint16_t entq[CS_NCH];             /* entry position of every chunk, relative to the segment */
    int entry;                         /* first block start of the next segment, relative to its first byte */
    int nb;                            /* blocks started so far in this frame */
    int skips;
    int consumed;
};

/* bit 6 of every byte of the form 01xxxxxx: a run token (signed value > 63, lib/RTjpeg.c:173) */
__device__ __forceinline__ uint32_t swar_runs(uint32_t t) { return t & ~(t >> 1) & 0x40404040u; }
/* run length - 1 in run bytes, 0 in coefficient bytes: a token fills 1 + x positions */
__device__ __forceinline__ uint32_t swar_x(uint32_t t) { return t & ((swar_runs(t) >> 6) * 0x3Fu); }

__device__ __forceinline__ uint32_t lds_u32_unaligned(const uint32_t *w, int byte)
{
    const uint32_t *p = w + (byte >> 2);
    return __funnelshift_r(p[0], p[1], (unsigned)(byte & 3) * 8);
}

/* Token count of a block whose first eight tokens (starting at shared byte `tb`) fill
 * `filled` < 63 positions: keep going four tokens at a time.  Ends within 63 tokens
 * because every token fills at least one position. */
__device__ __noinline__ int cs_long_block(const uint32_t *payw, int tb, int filled)
{
    int need = 63 - filled, ntok = 8;
    for (;;) {
        const uint32_t t = lds_u32_unaligned(payw, tb + ntok);
        const uint32_t P = swar_x(t) * 0x01010101u + 0x04030201u;
        const uint32_t c = (P + (uint32_t)(128 - need) * 0x01010101u) & 0x80808080u;
        if (c) return ntok + ((__ffs((int)c) - 1) >> 3) + 1;
        need -= (int)(P >> 24);
        ntok += 4;
    }
}

} // namespace

/* PHASE 0: the whole frame, segment after segment.  PHASE 1 / 2: one segment (blockIdx.x) of frame
 * blockIdx.y -- its summary for the frame-level chain, resp. its entries (rtj_common.h, rtj_seg_plan). */
template <int PHASE>
__global__ void __launch_bounds__(CS_THREADS, 8)
rtj_scan_chunk_kernel(const uint8_t *__restrict__ stream, const rtjgpu_frame_desc *__restrict__ desc,
                      const rtj_dev_table *__restrict__ tables, int F, int nblk,
                      uint32_t *__restrict__ ent, uint32_t *__restrict__ frame_skips,
                      rtj_dev_info *__restrict__ info, const rtj_seg_plan sp, int f0, int slice,
                      const uint32_t *__restrict__ redo)
{
    extern __shared__ __align__(16) uint8_t cs_smem[];
    CsShared &sh = *reinterpret_cast<CsShared *>(cs_smem);
    const int tid = threadIdx.x, lane = tid & 31;
    const int f = (PHASE == 0 ? blockIdx.x : blockIdx.y) + f0;        /* F: one behind the last frame of this launch */
    if (f >= F) return;
    if (PHASE == 0 && redo) {
        /* behind rtj_scan_sync_kernel, launched with programmatic stream serialisation: this grid may be resident before
         * that one is done (and lets the grid behind itself in likewise); its flags are final after the wait */
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (redo[f] != RTJ_REDO_CHUNK) return;                         /* only the frames it left */
    }
    const rtjgpu_frame_desc d = desc[f];
    {
        const rtj_dev_table &tab = tables[min((int)d.table, RTJ_NUM_TABLES - 1)];      /* descriptors are the caller's memory */
        if (tab.bt8[0] | tab.bt8[1]) return;                           /* raw prefix: rtj_scan_mb_kernel's frame */
    }

    const uint8_t *pay = stream + d.offset + RTJPEG_B200_HEADER_BYTES;
    const int len = d.length > RTJPEG_B200_HEADER_BYTES ? (int)d.length - RTJPEG_B200_HEADER_BYTES : 0;
    const int mis = (int)(reinterpret_cast<uintptr_t>(pay) & 15);     /* 0, 4, 8 or 12: packets start 4-byte aligned */
    const uint8_t *gbase = pay - mis;                                   /* 16-byte aligned, never before the packet */
    uint32_t *out = ent + (size_t)f * nblk;
    const uint8_t *payb = reinterpret_cast<const uint8_t *>(sh.pay) + mis;   /* payb[q] = payload byte seg0 + q */
    const uint8_t *delb = reinterpret_cast<const uint8_t *>(sh.del);

    const size_t my_seg = (size_t)f * sp.maxseg + blockIdx.x;       /* PHASE 1 / 2 */
    if (PHASE == 2 && sp.base[my_seg] == RTJ_SEG_UNUSED) return;    /* the frame is complete before this segment */
    if (tid == 0) {
        sh.entry = PHASE == 2 ? (int)sp.entry[my_seg] : 0;
        sh.nb = PHASE == 2 ? (int)sp.base[my_seg] : 0;
        sh.skips = 0;
        sh.consumed = 0;
    }
    __syncthreads();

    for (int seg0 = PHASE == 0 ? 0 : (int)blockIdx.x * CS_S; seg0 < len; seg0 += CS_S) {
        const int nb0 = sh.nb;
        if (nb0 >= nblk) break;                                        /* uniform: everybody reads the same word */
        const int lim = len - seg0;                                    /* payload bytes from here on */
        const int nch = min(CS_NCH, (lim + CS_C - 1) / CS_C);
        const int npos = nch * CS_C;

        /* ---- load: 16-byte vectors; beyond the payload every byte reads 0x7F, a run token that
         *      ends any block (the packet's own bytes are never read past its last 16-byte line).
         *      The next segment's lines are asked into L2 now: they are wanted one segment time later. ---- */
        if (PHASE == 0) {
            const int next = seg0 + CS_S + tid * 128;                  /* one 128-byte line per thread: CS_THREADS * 128 >= CS_S */
            if (tid * 128 < CS_S + CS_LA && next < len)
                asm volatile("prefetch.global.L2 [%0];" :: "l"(gbase + next));
        }
        {
            const uint4 *g4 = reinterpret_cast<const uint4 *>(gbase + seg0);
            uint4 *s4 = reinterpret_cast<uint4 *>(sh.pay);
            const int nvec = (npos + CS_LA + 16) / 16;
            const int vlim = lim + mis;                                /* shared byte index of the payload's end */
            for (int v = tid; v < nvec; v += CS_THREADS) {
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

        /* ---- level 0: delta(p) for every position, four positions per thread.  The second pass of the
         *      segment-parallel arrangement takes the table its first pass made, where it was kept. ---- */
        static_assert(CS_DEL_WORDS * 4 <= RTJ_SEG_DEL_BYTES && (CS_DEL_WORDS % 4) == 0, "the table fits the kept copy");
        uint4 *keep = (PHASE != 0 && sp.del) ? reinterpret_cast<uint4 *>(sp.del + my_seg * RTJ_SEG_DEL_BYTES) : nullptr;
        if (PHASE == 2 && keep) {
            uint4 *d4 = reinterpret_cast<uint4 *>(sh.del);
            for (int v = tid; v < CS_DEL_WORDS / 4; v += CS_THREADS) d4[v] = keep[v];
        }
        for (int q0 = tid * 4; q0 < ((PHASE == 2 && keep) ? 0 : npos); q0 += CS_THREADS * 4) {
            const uint32_t *wp = sh.pay + ((q0 + mis) >> 2);
            const uint32_t W0 = wp[0], W1 = wp[1], W2 = wp[2];
            const uint32_t X0 = swar_x(W0), X1 = swar_x(W1), X2 = swar_x(W2);
            uint32_t packed = 0;
            uint32_t fill8[4];                                          /* positions the first eight tokens fill, for the rare long block */
            unsigned slow = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                /* tokens of a block starting at q0 + i: bytes q0+i+1 .. */
                const uint32_t x0 = i == 3 ? X1 : __funnelshift_r(X0, X1, 8 * (i + 1));
                const uint32_t x1 = i == 3 ? X2 : __funnelshift_r(X1, X2, 8 * (i + 1));
                /* byte k of P = 65 + positions filled by tokens 0..k, so bit 7 <=> >= 63 filled.  Overflows
                 * can only happen behind the first crossing and never disturb it.  The top byte of P0 less
                 * 65 carries the first four tokens' fill into P1: 0x45444342 - 65 * 0x01010101 = 0x04030201. */
                const uint32_t P0 = x0 * 0x01010101u + 0x45444342u;
                const uint32_t P1 = x1 * 0x01010101u + 0x04030201u + (P0 >> 24) * 0x01010101u;
                const uint32_t c0 = P0 & 0x80808080u;
                const uint32_t c1 = P1 & 0x80808080u;
                /* no branch here: a block longer than DC + 8 tokens (both masks empty) is marked and finished below */
                const int bit = c0 ? __ffs((int)c0) - 1 : 31 + __ffs((int)c1);
                packed |= (uint32_t)((bit >> 3) + 2) << (8 * i);        /* DC byte + tokens */
                slow |= (c0 | c1) ? 0u : 1u << i;
                fill8[i] = P1 >> 24;
            }
            /* skipped blocks (lib/RTjpeg.c:2704): a byte 0xFF is a block of its own, one byte long */
            const uint32_t y = ~W0;
            const uint32_t z = ~(((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y | 0x7F7F7F7Fu);          /* bit 7 of every byte of W0 that is 0xFF */
            /* ... and needs no token walk at all: inside a run of skip markers every position sees "coefficients" of -1
             * without end, i.e. the longest possible block -- the slow path, for a length that is thrown away.  (Bit i of
             * the product: byte i of z >> 7.) */
            slow &= ~(((z >> 7) * 0x00204081u) >> 21);
            if (slow) {
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (slow & (1u << i)) {
                        const int dl = 1 + cs_long_block(sh.pay, q0 + mis + i + 1, (int)fill8[i] - 65);
                        packed = (packed & ~(0xFFu << (8 * i))) | (uint32_t)dl << (8 * i);
                    }
            }
            {
                const uint32_t m = (z >> 7) * 0xFFu;
                packed = (packed & ~m) | (0x01010101u & m);
            }
            sh.del[(q0 >> 2) + (q0 / CS_C)] = packed;
        }
        __syncthreads();
        if (PHASE == 1 && keep) {
            const uint4 *d4 = reinterpret_cast<const uint4 *>(sh.del);
            for (int v = tid; v < CS_DEL_WORDS / 4; v += CS_THREADS) keep[v] = d4[v];
        }

        /* ---- DP, one lane per chunk, right to left ---- */
        if (tid < nch) {
            uint16_t *ring = reinterpret_cast<uint16_t *>(sh.ring) + tid * CS_RING;
            const uint32_t *dw = sh.del + tid * (CS_C / 4 + 1);
            const int cq = tid * CS_C;
            /* the 64 positions behind the chunk: a parse that lands there has left the chunk, at that
             * offset, without starting another block.  Their ring slots are exactly their offsets. */
#pragma unroll
            for (int k = 0; k < 64; k += 2) reinterpret_cast<uint32_t *>(ring)[k >> 1] = (uint32_t)k | ((uint32_t)(k + 1) << 16);
            if (cq + CS_C <= lim) {
#pragma unroll 2
                for (int w = CS_C / 4 - 1; w >= 0; --w) {
                    const uint32_t d4 = dw[w];
                    uint16_t *slot = ring + ((4 * w) & 63);
#pragma unroll
                    for (int i = 3; i >= 0; --i) {
                        const int n = 4 * w + i + (int)((d4 >> (8 * i)) & 0xFFu);
                        slot[i] = (uint16_t)(ring[n & 63] + 64u);          /* exit offset | blocks << 6 */
                    }
                }
            } else {                                                       /* the chunk the payload ends in */
                for (int w = CS_C / 4 - 1; w >= 0; --w) {
                    const uint32_t d4 = dw[w];
#pragma unroll
                    for (int i = 3; i >= 0; --i) {
                        const int qq = 4 * w + i;
                        const int n = qq + (int)((d4 >> (8 * i)) & 0xFFu);
                        uint32_t v = ring[n & 63] + 64u;
                        if (cq + qq >= lim) v = 0;                         /* nothing starts behind the payload */
                        ring[qq & 63] = (uint16_t)v;
                    }
                }
            }
        }
        __syncthreads();

        if (PHASE == 1) {
            /* ---- summary: for every entry offset, where the parse leaves the segment and what it starts ---- */
            const uint16_t *rings = reinterpret_cast<const uint16_t *>(sh.ring);
            for (int e0 = tid; e0 < 64; e0 += CS_THREADS) {
                int e = e0, units = 0;
                for (int j = 0; j < nch; j++) {
                    const uint32_t v = rings[j * CS_RING + e];
                    e = (int)(v & 63u);
                    units += (int)(v >> 6);
                }
                sp.sum[my_seg * RTJ_SEG_NE + e0] = (uint32_t)e | ((uint32_t)units << 9);
            }
            return;
        }

        /* ---- chain: entry point and first block index of every chunk ---- */
        if (tid == 0) {
            const uint16_t *rings = reinterpret_cast<const uint16_t *>(sh.ring);
            int e = sh.entry, nb = nb0;
            for (int j = 0; j < nch; j++) {
                sh.entq[j] = (uint16_t)(j * CS_C + e);
                sh.base[j] = (uint32_t)nb;
                const uint32_t v = rings[j * CS_RING + e];
                e = (int)(v & 63u);
                nb += (int)(v >> 6);
            }
            sh.base[nch] = (uint32_t)nb;
            sh.entry = e;
            sh.nb = nb;
        }
        __syncthreads();

        /* ---- emit, in two steps.  The chunk lanes only WALK their blocks (a load and an add per block) and note where
         *      each starts; making the 32-bit entry of a block -- the longer part -- is then shared out evenly over all
         *      threads, which also write the entries straight to the table, coalesced. ---- */
        const int nb1 = min(sh.nb, nblk);
        int q = 0, i = 0, qend = 0, myskips = 0, lastend = -1;
        if (tid < nch) {
            q = sh.entq[tid];
            i = (int)sh.base[tid];
            qend = min((tid + 1) * CS_C, lim);
        }
        __syncthreads();                                   /* the rings are dead: their memory holds the block starts */
        uint16_t *starts = reinterpret_cast<uint16_t *>(sh.ring);
        for (int r0 = nb0; r0 < nb1; r0 += 2 * CS_STAGE) {
            const int r1 = min(r0 + 2 * CS_STAGE, nb1);
            if (tid < nch) {
                while (q < qend && i < r1) {
                    starts[i - r0] = (uint16_t)q;
                    q += delb[q + ((q / CS_C) << 2)];
                    i++;
                    lastend = seg0 + q;
                }
            }
            __syncthreads();
            for (int k = tid; k < r1 - r0; k += CS_THREADS) {
                const int qq = starts[k];
                const int dl = delb[qq + ((qq / CS_C) << 2)];
                const uint32_t head = lds_u32_unaligned(sh.pay, qq + mis);         /* DC, token 1, token 2, token 3 */
                const uint32_t last = payb[qq + dl - 1];
                const bool isff = (head & 0xFFu) == 0xFFu;                         /* skipped block */
                /* positions >= eob are zero: a final run of n zeros ends at 64, so it starts at 64 - n */
                const int eob = (last - 64u) < 64u ? max(127 - (int)last, dl - 1) : 64;
                const uint32_t t1 = (head >> 8) & 0xFFu, t2 = (head >> 16) & 0xFFu;
                const uint32_t c1 = (eob >= 2 && (t1 - 64u) >= 64u) ? t1 : 0u;
                const uint32_t c2 = (eob >= 3 && (t2 - 64u) >= 64u) ? t2 : 0u;
                const uint32_t e_inl = RTJ_ENT_INLINE_BIT | (head & 0xFFu) | (c1 << 8) | (c2 << 16);
                const uint32_t e_gen = RTJ_ENT(seg0 + qq, eob);
                out[r0 + k] = isff ? RTJ_ENT_SKIP : (eob <= 3 ? e_inl : e_gen);
                myskips += isff ? 1 : 0;
            }
            __syncthreads();
        }
        {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                myskips += __shfl_xor_sync(0xFFFFFFFFu, myskips, o);
                lastend = max(lastend, __shfl_xor_sync(0xFFFFFFFFu, lastend, o));
            }
            if (lane == 0) {
                if (myskips) atomicAdd(&sh.skips, myskips);
                if (lastend >= 0) atomicMax(&sh.consumed, lastend);
            }
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
    for (int b = nbf + tid; b < nblk; b += CS_THREADS) out[b] = RTJ_ENT(min(len, (int)RTJ_ENT_OFF_MASK), 1);
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

extern "C" int rtj_scan_chunk_init(void)
{
    cudaError_t e = cudaFuncSetAttribute(rtj_scan_chunk_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CsShared));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rtj_scan_chunk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CsShared));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rtj_scan_chunk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CsShared));
    return e == cudaSuccess ? 0 : (int)e;
}

extern "C" int rtj_launch_scan_chunk(const rtj_launch_args *a, int phase, void *stream)
{
    static_assert(CS_S == RTJ_SEG_BYTES, "segment size is shared with the frame-level chain");
    const int nblk = RTJ_FMT_NBLK(a->fmt, a->w, a->h);
    cudaStream_t st = (cudaStream_t)stream;
    const int nf = a->f1 - a->f0;
    const dim3 grid = phase == 0 ? dim3((unsigned)nf) : dim3((unsigned)a->seg.maxseg, (unsigned)nf);
    if (phase == 0)
        rtj_scan_chunk_kernel<0><<<grid, CS_THREADS, sizeof(CsShared), st>>>(
            a->d_stream, a->d_desc, a->d_tables, a->f1, nblk, a->d_ent, a->d_frame_skips, a->d_info, a->seg, a->f0, a->slice, nullptr);
    else if (phase == 1)
        rtj_scan_chunk_kernel<1><<<grid, CS_THREADS, sizeof(CsShared), st>>>(
            a->d_stream, a->d_desc, a->d_tables, a->f1, nblk, a->d_ent, a->d_frame_skips, a->d_info, a->seg, a->f0, a->slice, nullptr);
    else
        rtj_scan_chunk_kernel<2><<<grid, CS_THREADS, sizeof(CsShared), st>>>(
            a->d_stream, a->d_desc, a->d_tables, a->f1, nblk, a->d_ent, a->d_frame_skips, a->d_info, a->seg, a->f0, a->slice, nullptr);
    return (int)cudaGetLastError();
}

extern "C" int rtj_launch_scan_chunk_redo(const rtj_launch_args *a, const uint32_t *redo, void *stream)
{
    const int nblk = RTJ_FMT_NBLK(a->fmt, a->w, a->h);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(a->f1 - a->f0));
    cfg.blockDim = dim3(CS_THREADS);
    cfg.dynamicSmemBytes = sizeof(CsShared);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, rtj_scan_chunk_kernel<0>, a->d_stream, a->d_desc, a->d_tables, a->f1, nblk, a->d_ent,
                                             a->d_frame_skips, a->d_info, a->seg, a->f0, a->slice, redo);
    return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
}
