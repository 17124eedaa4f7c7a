/*
 * rtj_scan_sync.cu -- K1, self-synchronising flavour: the block-offset scan of one frame by a CTA whose
 * lanes each WALK a piece of the run-length stream, byte by byte, as a small state machine.
 *
 * What it replaces: the `sp += RTjpeg_s2b(...)` / `sp++` pointer chase of RTjpeg_decompressYUV420
 * (lib/RTjpeg.c:2701-2745), block lengths by the rules of RTjpeg_s2b (:157-186).  A parse of the stream is in
 * one of 64 states before every byte: "the next byte starts a block" (a DC byte, or the skip marker 0xFF,
 * :2704), or "r of the block's 63 places behind the DC are still to be filled" (r = 1..63; a byte 64..127 is a run
 * token that fills byte - 63 places, any other byte a coefficient that fills one, :171-183).  Where a block starts
 * depends on every byte before it -- but only through that one state.
 *
 * rtj_scan_chunk.cu makes the scan parallel by working out, for EVERY byte position, the length of the block
 * that would start there (~64 thread instructions per payload byte).  This kernel cuts the payload into one chunk per
 * lane and asks, per chunk, a cheaper question first:
 *
 *   phase A   from the chunk's first byte on, ALL 64 states at once -- the set of states a parse could be in, a 63-bit
 *             mask of the r's plus one flag, advanced by a shift per byte -- until the set has shrunk to ONE state.
 *             From that byte on (the chunk's synchronisation point) the parse is known whatever came before:
 *             run-length streams forget their past quickly (a handful of blocks; measured on the bench stream:
 *             22 bytes in the median, 41 on average).  No guess, no verification: the true state is always one
 *             of the 64.
 *   phase B   every lane walks, one state, ~9 instructions a byte, from its own synchronisation point to the next
 *             lane's and leaves one bit per byte: "a block starts here".  Lane 0 starts from the frame's first
 *             byte (or from the state the previous segment ended in).  A lane that found no synchronisation point
 *             inside its chunk simply has no piece of its own; its left neighbour walks on.
 *   emit      the bit map is counted (prefix sum per 32 positions), turned into a list of block starts, and the
 *             32-bit entries are made by all threads, one block each, exactly as rtj_scan_chunk.cu makes them.
 *
 * Streams that do not forget (every block 64 coefficient bytes long: noise at a high quality) would leave the whole
 * frame to lane 0.  The CTA sees that after phase A -- too many lanes without a synchronisation point -- and hands the
 * frame over (redo[f] = 1) to rtj_scan_chunk_kernel, which is launched behind this kernel and takes the flagged
 * frames only; its cost does not depend on the content.
 *
 * The payload of a frame (<= 39 KB a segment; larger frames in several segments, the state carried from one to the
 * next) arrives in shared memory as ONE bulk copy (cp.async.bulk, completion on an mbarrier).  Lanes read their
 * chunks as 8-byte groups; a chunk is an odd number of groups long, so that the 16 lanes of a half warp hit 16
 * different banks pairs.
 *
 * Scope: frames whose tables have no raw 8-bit prefix (lb8 == cb8 == 0), like rtj_scan_chunk.cu; the others are
 * rtj_scan_mb_kernel's.  Same entries (rtj_common.h), counters and malformed-stream policy as the other flavours.
 */
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "rtj_common.h"

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int SY_THREADS = 64;
constexpr int SY_WARPS = SY_THREADS / 32;
constexpr int SY_GMAX = 79;                                /* 8-byte groups per chunk: odd */
constexpr int SY_LEAD = 32;                                /* groups of the lead-in walk in front of a chunk (256 bytes) */
constexpr int SY_GMIN = SY_LEAD + 1;                       /* short frames use fewer lanes, not shorter chunks */
constexpr int SY_SEG_MAX = SY_THREADS * SY_GMAX * 8;       /* bytes of a frame in shared memory at a time */
constexpr int SY_LA = 80;                                  /* bytes walked behind a segment that is not the frame's last: the block that
                                                            * starts on its last byte (<= 64 bytes) and the start behind it, whole groups */
constexpr int SY_PAY_BYTES = SY_SEG_MAX + SY_LA + 16;      /* + what an unaligned 8-byte read may touch */
constexpr int SY_WORDS = (SY_SEG_MAX + SY_LA + 31) / 32;   /* bit map words */
constexpr int SY_STAGE = 3072;                             /* block starts staged per emit round */
constexpr int SY_KMAX = (SY_WORDS + SY_THREADS - 1) / SY_THREADS;
constexpr int SY_MAX_DIRTY = 8;                            /* more chunks than this entered in the wrong state: not this kernel's stream */

static_assert((SY_GMAX & 1) == 1 && (SY_GMIN & 1) == 1, "odd chunk lengths: conflict-free 8-byte reads");
static_assert(SY_SEG_MAX + SY_LA < 65536, "positions inside a segment are 16 bit");
static_assert((SY_PAY_BYTES % 16) == 0, "bulk copies move 16-byte pieces");

struct SyShared {
    alignas(16) uint32_t pay[SY_PAY_BYTES / 4];      /* the segment's bytes: pay[0] is the 16-byte line the frame's payload starts in */
    uint32_t bits[SY_WORDS + 1];         /* bit p: a block starts at byte p of the segment */
    uint16_t pre[SY_WORDS + 2];          /* blocks that start before word w's 32 positions */
    uint16_t starts[SY_STAGE + 2];       /* one emit round's block starts */
    int8_t   exitst[SY_THREADS];         /* the state every lane's walk left its chunk in */
    int      wsum[SY_WARPS];
    int      carry;                      /* state at the first byte of the next segment */
    int      nb;                         /* blocks started so far in this frame */
    int      skips;
    int      consumed;
    int      total;                      /* block starts in this segment */
    int      sentinel;                   /* first block start behind them */
    alignas(8) unsigned long long mbar;
};
static_assert(offsetof(SyShared, pay) == 0, "bulk copies and 16-byte stores: the payload leads the (16-byte aligned) block");

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, unsigned parity)
{
    unsigned done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, unsigned bytes, uint32_t mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

/* byte k of w, sign-extended (prmt: a selector nibble's top bit replicates the chosen byte's sign) */
template <int K>
__device__ __forceinline__ int sext_byte(uint32_t w)
{
    int d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(0u), "n"(0x8880 + 0x1111 * K));
    return d;
}
template <int K>
__device__ __forceinline__ int zext_byte(uint32_t w)
{
    int d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(0u), "n"(0x4440 + K));
    return d;
}

/* One state over four bytes.  r > 0: places of the block still to fill; r <= 0: the next byte starts a block (a DC byte: 63
 * places behind it; the skip marker 0xFF: a block of its own).  A run token 01xxxxxx fills x + 1 places, any other byte one.
 * bm gets a bit for every byte that starts a block. */
template <bool BITS>
__device__ __forceinline__ void walk4(uint32_t W, int &r, uint32_t &bm, int bit0)
{
    const uint32_t runs = W & ~(W >> 1) & 0x40404040u;                   /* bit 6 of every run token */
    const uint32_t NX = ~(W & (runs - (runs >> 6)));                     /* per byte, as a signed byte: -(what it fills) */
    const uint32_t y = ~W;
    const uint32_t z = ~((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) & W & 0x80808080u;   /* bit 7 of every byte that is 0xFF */
    const uint32_t G = 0x3F3F3F3Fu & ~(z - (z >> 7));                    /* the state behind a block's first byte: 63, or 0 behind a skip marker */
#define SY_STEP(k)                                                     \
    {                                                                  \
        const int nx = sext_byte<k>(NX), g = zext_byte<k>(G);          \
        const bool at = r <= 0;                                        \
        r = at ? g : r + nx;                                           \
        if (BITS && at) bm |= 1u << (bit0 + k);                        \
    }
    SY_STEP(0) SY_STEP(1) SY_STEP(2) SY_STEP(3)
#undef SY_STEP
}

/* groups [g0, g1) from state r; BITS: leave the starts in the bit map */
template <bool BITS>
__device__ __forceinline__ int walk_groups(const uint2 *pay8, uint8_t *bits8, int g0, int g1, int r)
{
    for (int g = g0; g < g1; g++) {
        const uint2 w = pay8[g];
        uint32_t bm = 0;
        walk4<BITS>(w.x, r, bm, 0);
        walk4<BITS>(w.y, r, bm, 4);
        if (BITS) bits8[g] = (uint8_t)bm;
    }
    return r;
}

__device__ __forceinline__ uint32_t lds_u32_unaligned(const uint32_t *w, int byte)
{
    const uint32_t *p = w + (byte >> 2);
    return __funnelshift_r(p[0], p[1], (unsigned)(byte & 3) * 8);
}

/* the starts of word w that count: positions [first, limit) of the segment */
__device__ __forceinline__ uint32_t sy_mask(uint32_t v, int w, int first, int limit)
{
    if (w == 0) v &= ~((1u << first) - 1u);                              /* first <= 12 */
    const int wc = limit >> 5;
    if (w > wc) v = 0u;
    else if (w == wc) v &= (1u << (limit & 31)) - 1u;
    return v;
}

} // namespace

__global__ void __launch_bounds__(SY_THREADS, 4)
rtj_scan_sync_kernel(const uint8_t *__restrict__ stream, const rtjgpu_frame_desc *__restrict__ desc,
                     const rtj_dev_table *__restrict__ tables, int F, int nblk,
                     uint32_t *__restrict__ ent, uint32_t *__restrict__ frame_skips,
                     rtj_dev_info *__restrict__ info, uint32_t *__restrict__ redo, int f0, int slice)
{
    extern __shared__ __align__(16) uint8_t sy_smem[];
    SyShared &sh = *reinterpret_cast<SyShared *>(sy_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.x + f0;
    if (f >= F) return;
    if (redo && tid == 0) redo[f] = 0u;
    const rtjgpu_frame_desc d = desc[f];
    {
        const rtj_dev_table &tab = tables[min((int)d.table, RTJ_NUM_TABLES - 1)];
        if (tab.bt8[0] | tab.bt8[1]) return;                           /* raw prefix: rtj_scan_mb_kernel's frame */
    }

    const uint8_t *pay = stream + d.offset + RTJPEG_B200_HEADER_BYTES;
    const int len = d.length > RTJPEG_B200_HEADER_BYTES ? (int)d.length - RTJPEG_B200_HEADER_BYTES : 0;
    const int mis = (int)(reinterpret_cast<uintptr_t>(pay) & 15);     /* 0, 4, 8 or 12: packets start 4-byte aligned */
    const uint8_t *gbase = pay - mis;                                   /* 16-byte aligned, never before the packet */
    uint32_t *out = ent + (size_t)f * nblk;
    const uint8_t *payb = reinterpret_cast<const uint8_t *>(sh.pay);
    uint8_t *bits8 = reinterpret_cast<uint8_t *>(sh.bits);

    /* positions: byte p of the frame is gbase[p]; the payload is [mis, end).  A frame that fits one segment leaves room for
     * the start behind its last block (end + 1 at most) inside the lanes' chunks. */
    const int end = len > 0 ? mis + len : 0;
    const int G = end + 8 > SY_SEG_MAX ? SY_GMAX : max(SY_GMIN, ((end + 8 + SY_THREADS * 8 - 1) / (SY_THREADS * 8)) | 1);
    const int SEG = SY_THREADS * G * 8;
    const uint32_t mbar = smem_u32(&sh.mbar);

    if (tid == 0) {
        sh.carry = 0;
        sh.nb = 0;
        sh.skips = 0;
        sh.consumed = 0;
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    unsigned parity = 0;
    for (int seg0 = 0; seg0 < end; seg0 += SEG) {
        const int nb0 = sh.nb;
        if (nb0 >= nblk) break;                                        /* uniform: everybody reads the same word */
        const int lim = end - seg0;                                    /* the payload ends at byte lim of this segment */
        const int first = seg0 == 0 ? mis : 0;
        const int climit = min(SEG, lim);                              /* starts before it are this segment's blocks */
        const bool more = lim + 8 > SEG;                               /* the frame goes on behind this segment */
        const int gtot = more ? (SEG + SY_LA) >> 3 : (lim + 8 + 7) >> 3;   /* groups to walk: up to the start behind the last block */
        const int wtot = (gtot * 8 + 31) >> 5;                         /* bit map words they fill (the last one maybe in part) */

        /* ---- load: one bulk copy; what lies behind the payload reads 0x7F, a run token that ends any block (the packet's
         *      own bytes are never read past its last 16-byte line) ---- */
        const int nbuf = gtot * 8 + 16;
        const int nload = min(nbuf & ~15, (lim + 15) & ~15);
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   /* the previous segment's reads are done (barrier below) */
            mbar_expect_tx(mbar, (unsigned)nload);
            bulk_g2s(smem_u32(sh.pay), gbase + seg0, (unsigned)nload, mbar);
        }
        for (int b = nload + tid * 16; b < nbuf; b += SY_THREADS * 16)
            *reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(sh.pay) + b) = make_uint4(0x7F7F7F7Fu, 0x7F7F7F7Fu, 0x7F7F7F7Fu, 0x7F7F7F7Fu);
        mbar_wait(mbar, parity);
        parity ^= 1u;
        if (tid < 16) {
            uint8_t *pb = reinterpret_cast<uint8_t *>(sh.pay);
            if (lim + tid < nload) pb[lim + tid] = 0x7F;               /* at most 15 bytes */
            if (tid < first) pb[tid] = 0xFF;                           /* in front of the payload: "skip markers", whose starts do not count */
        }
        __syncthreads();

        /* ---- lead-in: every lane walks the SY_LEAD groups in front of its chunk from a guessed state ("a block starts
         *      here").  Run-length streams forget their past within a handful of blocks: at the chunk's first byte the
         *      walk is, as a rule, in the true state.  Whether it is, is checked below -- never assumed. ---- */
        const uint2 *pay8 = reinterpret_cast<const uint2 *>(sh.pay);
        const int g0 = tid * G, g1 = min(g0 + G, gtot);                /* this lane's chunk, in groups */
        int r_in = 0;                                                  /* the state the chunk is entered in */
        if (tid == 0) r_in = seg0 == 0 ? 0 : sh.carry;
        else if (g0 < gtot) r_in = walk_groups<false>(pay8, nullptr, g0 - SY_LEAD, g0, 0);

        /* ---- walk: the chunk, from that state; one bit per byte that starts a block ---- */
        int r_out = r_in;
        if (g0 < gtot) r_out = walk_groups<true>(pay8, bits8, g0, g1, r_in);
        sh.exitst[tid] = (int8_t)r_out;
        __syncthreads();

        /* ---- check: a chunk must have been entered in the state its left neighbour ended in.  Where not, it is walked
         *      again from that state; its own exit state may change in turn, so this repeats until nothing changes (one
         *      round as a rule; after round k the first k chunks are final whatever the stream). ---- */
        for (int round = 0;; round++) {
            const int want = tid > 0 ? (int)sh.exitst[tid - 1] : r_in;
            /* states <= 0 all mean "a block starts here" */
            const bool dirty = tid > 0 && g0 < gtot && max(want, 0) != max(r_in, 0);
            const int ndirty = __syncthreads_count(dirty);
            if (ndirty == 0) break;
            if (redo && round == 0 && ndirty > SY_MAX_DIRTY) {
                /* a stream that does not forget its past: leave the frame to the kernel whose cost does not depend on the content */
                if (tid == 0) redo[f] = 1u;
                return;
            }
            if (dirty) {
                r_in = want;
                r_out = walk_groups<true>(pay8, bits8, g0, g1, r_in);
                sh.exitst[tid] = (int8_t)r_out;
            }
            __syncthreads();
        }
        if (more) {
            /* the state the next segment starts in; the look-ahead behind this one */
            if (tid == SY_THREADS - 1) {
                sh.carry = r_out;
                walk_groups<true>(pay8, bits8, SY_THREADS * G, gtot, r_out);
            }
            __syncthreads();
        }

        /* ---- count: blocks that start before every word of the bit map ---- */
        {
            const int K = (wtot + SY_THREADS - 1) / SY_THREADS;
            const int w0 = tid * K;
            int cnt = 0;
#pragma unroll
            for (int i = 0; i < SY_KMAX; i++)
                if (i < K && w0 + i < wtot) cnt += __popc(sy_mask(sh.bits[w0 + i], w0 + i, first, climit));
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += up;
            }
            if (lane == 31) sh.wsum[warp] = incl;
            __syncthreads();
            int base = incl - cnt, total = 0;
#pragma unroll
            for (int k = 0; k < SY_WARPS; k++) {
                const int s = sh.wsum[k];
                if (k < warp) base += s;
                total += s;
            }
#pragma unroll
            for (int i = 0; i < SY_KMAX; i++)
                if (i < K && w0 + i < wtot) {
                    sh.pre[w0 + i] = (uint16_t)base;
                    base += __popc(sy_mask(sh.bits[w0 + i], w0 + i, first, climit));
                }
            if (tid == 0) {
                sh.pre[wtot] = (uint16_t)total;
                sh.total = total;
                /* the first start at or behind climit: where the segment's last block ends */
                int s = climit;
                for (int w = climit >> 5; w < wtot; w++) {
                    uint32_t v = sh.bits[w];
                    if (w == (climit >> 5)) v &= ~((1u << (climit & 31)) - 1u);
                    if (w == wtot - 1 && ((gtot * 8) & 31)) v &= (1u << ((gtot * 8) & 31)) - 1u;
                    if (v) { s = w * 32 + __ffs((int)v) - 1; break; }
                }
                sh.sentinel = s;
            }
        }
        __syncthreads();

        /* ---- emit: the list of starts, round by round, then one block per thread ---- */
        const int total = sh.total;
        const int nemit = min(total, nblk - nb0);
        int myskips = 0, lastend = -1;
        for (int lo = 0; lo < nemit; lo += SY_STAGE) {
            const int n = min(SY_STAGE, nemit - lo);
            {
                /* one start per turn of the loop, words taken in turn: lanes stay busy whatever their words hold */
                int w = tid - SY_THREADS, idx = 0;
                uint32_t v = 0;
                for (;;) {
                    if (v == 0u) {
                        w += SY_THREADS;
                        if (w >= wtot) break;
                        idx = (int)sh.pre[w] - lo;
                        if (idx <= n && (int)sh.pre[w + 1] - lo > 0) v = sy_mask(sh.bits[w], w, first, climit);
                        continue;
                    }
                    const int b = __ffs((int)v) - 1;
                    v &= v - 1u;
                    if ((unsigned)idx <= (unsigned)n) sh.starts[idx] = (uint16_t)(w * 32 + b);
                    idx++;
                }
            }
            if (tid == 0 && lo + n == total) sh.starts[n] = (uint16_t)sh.sentinel;
            __syncthreads();
            for (int k = tid; k < n; k += SY_THREADS) {
                const int qq = sh.starts[k], nx = sh.starts[k + 1];
                const int dl = nx - qq;
                const uint32_t head = lds_u32_unaligned(sh.pay, qq);               /* DC, token 1, token 2, token 3 */
                const uint32_t last = payb[nx - 1];
                const bool isff = (head & 0xFFu) == 0xFFu;                         /* skipped block */
                /* positions >= eob are zero: a final run of n zeros ends at 64, so it starts at 64 - n */
                const int eob = (last - 64u) < 64u ? max(127 - (int)last, dl - 1) : 64;
                const uint32_t t1 = (head >> 8) & 0xFFu, t2 = (head >> 16) & 0xFFu;
                const uint32_t c1 = (eob >= 2 && (t1 - 64u) >= 64u) ? t1 : 0u;
                const uint32_t c2 = (eob >= 3 && (t2 - 64u) >= 64u) ? t2 : 0u;
                const uint32_t e_inl = RTJ_ENT_INLINE_BIT | (head & 0xFFu) | (c1 << 8) | (c2 << 16);
                const uint32_t e_gen = RTJ_ENT(seg0 + qq - mis, eob);
                out[nb0 + lo + k] = isff ? RTJ_ENT_SKIP : (eob <= 3 ? e_inl : e_gen);
                myskips += isff ? 1 : 0;
                lastend = max(lastend, seg0 + nx - mis);
            }
            __syncthreads();
        }
        {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                myskips += __shfl_xor_sync(FULL, myskips, o);
                lastend = max(lastend, __shfl_xor_sync(FULL, lastend, o));
            }
            if (lane == 0) {
                if (myskips) atomicAdd(&sh.skips, myskips);
                if (lastend >= 0) atomicMax(&sh.consumed, lastend);
            }
            if (tid == 0) sh.nb = nb0 + total;
        }
        __syncthreads();
    }

    /* a frame whose stream ends early or mid-block: flag it, give the missing blocks a harmless entry */
    const int nbf = min(sh.nb, nblk);
    for (int b = nbf + tid; b < nblk; b += SY_THREADS) out[b] = RTJ_ENT(min(len, (int)RTJ_ENT_OFF_MASK), 1);
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

extern "C" int rtj_scan_sync_init(void)
{
    cudaError_t e = cudaFuncSetAttribute(rtj_scan_sync_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SyShared));
    return e == cudaSuccess ? 0 : (int)e;
}

/* redo: [F] flags (NULL: never hand a frame over) */
extern "C" int rtj_launch_scan_sync(const rtj_launch_args *a, uint32_t *redo, void *stream)
{
    const int nblk = RTJ_FMT_NBLK(a->fmt, a->w, a->h);
    const int nf = a->f1 - a->f0;
    rtj_scan_sync_kernel<<<(unsigned)nf, SY_THREADS, sizeof(SyShared), (cudaStream_t)stream>>>(
        a->d_stream, a->d_desc, a->d_tables, a->f1, nblk, a->d_ent, a->d_frame_skips, a->d_info, redo, a->f0, a->slice);
    return (int)cudaGetLastError();
}
