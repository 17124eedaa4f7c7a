/*
 * rtj_scan_walk.cu -- K1, walker flavour: ONE LANE walks one frame's run-length stream, block after block.
 *
 * What it replaces is the same `sp += RTjpeg_s2b(...)` / `sp++` pointer chase of RTjpeg_decompressYUV420
 * (lib/RTjpeg.c:2701-2745, block lengths by the rules of RTjpeg_s2b, :157-186) as rtj_scan_chunk.cu -- but where that
 * kernel buys parallelism INSIDE a frame by examining every byte position (about 64 thread instructions per payload
 * byte), this one does the minimum: one step per BLOCK (about 100 instructions), a frame's blocks strictly in turn.  That is
 * a third of the instructions -- and a dependent chain of nblk steps per frame, which is why it is NOT what AUTO picks:
 * measured on a B200, a step takes ~420 cycles of a lone warp, 2.07 ms for a 720x576 frame whatever
 * the batch, against 0.42 ms for 4096 frames with the chunk-parallel scan.  Working the batch through in slices of
 * macroblock rows, walkers ahead of K2 on a second stream, hid K2 behind the walk, not the walk.  It stays as an independent
 * implementation of the grammar that the parity suite runs every case under (rtjgpu_set_scan_mode(RTJGPU_SCAN_WALK)), and as
 * the measured answer to "why not one thread per frame".
 *
 * Every lane keeps a window of its frame's payload in shared memory, a ring of four 128-byte chunks that it fills itself
 * with cp.async (LDGSTS, 16 bytes a piece) three chunks ahead of where it reads -- lane-local copies, lane-local completion
 * (cp.async.wait_group), no barrier -- so that nothing in the chain misses a cache.
 *
 * Scope: frames whose tables have no raw 8-bit prefix (lb8 == cb8 == 0), like rtj_scan_chunk.cu; the others are
 * rtj_scan_mb_kernel's.  Same entries (rtj_common.h), counters and malformed-stream policy as the other flavours.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "rtj_common.h"

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int WK_CHUNK = 128;                        /* bytes per request */
constexpr int WK_RING = 4 * WK_CHUNK;                /* bytes of payload a lane holds */
constexpr int WK_ROW = WK_RING / 4 + 4;              /* words per lane: the ring, then a copy of its first 16 bytes (a block's first
                                                      * twelve bytes are read without wrapping); 528 bytes: 16-byte pieces stay aligned */

__device__ __forceinline__ uint32_t swar_runs(uint32_t t) { return t & ~(t >> 1) & 0x40404040u; }
__device__ __forceinline__ uint32_t swar_x(uint32_t t, uint32_t r) { return t & ((r >> 6) * 0x3Fu); }

__device__ __forceinline__ void cp16(unsigned dst, unsigned long long src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}

/* One 128-byte chunk of the lane's frame into its ring (no commit): the 16-byte pieces inside [lo, hi).  Chunks wholly
 * inside -- all but the first and the last few of a frame -- take eight unconditional copies. */
__device__ __forceinline__ void wk_request(unsigned ring_s, unsigned long long base, int chunk, unsigned long long lo, unsigned long long hi)
{
    const unsigned long long src = base + (unsigned long long)(unsigned)chunk * WK_CHUNK;
    const unsigned dst = ring_s + (unsigned)((chunk & 3) * WK_CHUNK);
    if (src >= lo && src + WK_CHUNK <= hi) {
#pragma unroll
        for (int j = 0; j < WK_CHUNK / 16; j++) cp16(dst + 16u * j, src + 16u * j);
        if ((chunk & 3) == 0) cp16(ring_s + WK_RING, src);                  /* the ring's first piece once more behind its end */
    } else {
#pragma unroll
        for (int j = 0; j < WK_CHUNK / 16; j++) {
            const unsigned long long s = src + 16u * j;
            if (s >= lo && s + 16 <= hi) {
                cp16(dst + 16u * j, s);
                if (j == 0 && (chunk & 3) == 0) cp16(ring_s + WK_RING, s);
            }
        }
    }
}

struct WkFrame {
    unsigned long long base, lo, hi;     /* chunk grid of the frame; the bytes that may be read */
    int skew, len;                       /* ring position of payload byte o: skew + o; payload bytes */
    unsigned ring_s;                     /* the lane's ring in the shared window */
    uint32_t *out;                       /* the frame's entries */
};

/* One block: its entry, and where the next one starts.  `req` is the highest chunk asked for. */
__device__ __forceinline__ void wk_step(const WkFrame &fr, int &o, int &blk, int &skips)
{
    const int r = fr.skew + o;
    /* the block's first byte and its first eight tokens: nine bytes, three words of the ring (no wrap: see WK_ROW) */
    const unsigned at = fr.ring_s + (unsigned)(r & (WK_RING - 1) & ~3);
    uint32_t w0, w1, w2;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(at));
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(at));
    asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(at));
    const unsigned sh = (unsigned)(r & 3) * 8;
    const uint32_t u0 = __funnelshift_r(w0, w1, sh), u1 = __funnelshift_r(w1, w2, sh), u2 = w2 >> sh;
    const uint32_t first = u0 & 0xFFu;
    const uint32_t t0 = __funnelshift_r(u0, u1, 8);               /* bytes o+1 .. o+4 */
    const uint32_t t1 = __funnelshift_r(u1, u2, 8);               /* bytes o+5 .. o+8 */
    const bool isff = first == 0xFFu;                              /* skipped block: one byte, lib/RTjpeg.c:2704 */

    /* eight tokens at once (rtj_kernels.cu, lane_scan_frame): byte k of P = 65 + positions filled by tokens 0..k, so bit 7
     * <=> 63 are filled; 0x45444342 - 65 * 0x01010101 = 0x04030201 carries the first four tokens' fill into the next four */
    const uint32_t r0 = swar_runs(t0), r1 = swar_runs(t1);
    const uint32_t P0 = swar_x(t0, r0) * 0x01010101u + 0x45444342u;
    const uint32_t P1 = swar_x(t1, r1) * 0x01010101u + 0x04030201u + (P0 >> 24) * 0x01010101u;
    const uint32_t c0 = P0 & 0x80808080u, c1 = P1 & 0x80808080u;
    uint32_t c = c0 ? c0 : c1, t = c0 ? t0 : t1, rr = c0 ? r0 : r1;
    int ntok = c0 ? 0 : 4;
    if (!isff && c == 0) {                                         /* long block: keep going four tokens at a time */
        int need = 63 + 65 - (int)(P1 >> 24);
        ntok = 8;
        for (;;) {
            if (o + 1 + ntok >= fr.len + 64) { c = 0x80u; rr = 0; break; }       /* runaway on a truncated frame */
            const int q = r + 1 + ntok;
            const unsigned qa = fr.ring_s + (unsigned)(q & (WK_RING - 1) & ~3), qb = fr.ring_s + (unsigned)((q + 4) & (WK_RING - 1) & ~3);
            uint32_t v0, v1;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v0) : "r"(qa));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v1) : "r"(qb));
            t = __funnelshift_r(v0, v1, (unsigned)(q & 3) * 8);
            rr = swar_runs(t);
            const uint32_t P = swar_x(t, rr) * 0x01010101u + 0x04030201u;
            c = (P + (uint32_t)(128 - need) * 0x01010101u) & 0x80808080u;
            if (c) break;
            need -= (int)(P >> 24);
            ntok += 4;
        }
    }
    const int bit = __ffs((int)c) - 1;                             /* 7, 15, 23 or 31 */
    ntok += (bit >> 3) + 1;
    const uint32_t bk = (t >> (bit - 7)) & 0xFFu;
    /* positions >= eob are zero: a final run of n zeros ends at 64, so it starts at 64 - n (a run that overshoots hides nothing before it) */
    const int eob = ((rr >> (bit - 1)) & 1u) ? max(63 - (int)(bk & 0x3Fu), ntok) : 64;
    /* a block of at most three coefficients travels in its entry (rtj_common.h) */
    const uint32_t a1 = t0 & 0xFFu, a2 = (t0 >> 8) & 0xFFu;
    const uint32_t k1 = (eob >= 2 && (a1 - 64u) >= 64u) ? a1 : 0u;
    const uint32_t k2 = (eob >= 3 && (a2 - 64u) >= 64u) ? a2 : 0u;
    const uint32_t e_inl = RTJ_ENT_INLINE_BIT | first | (k1 << 8) | (k2 << 16);
    const uint32_t e_gen = RTJ_ENT(o, eob);
    fr.out[blk++] = isff ? RTJ_ENT_SKIP : (eob <= 3 ? e_inl : e_gen);
    skips += isff ? 1 : 0;
    o = isff ? o + 1 : o + 1 + ntok;
}

} // namespace

/*
 * One lane per frame.  Blocks [b0, b1) of every frame; where a frame's walk stands between calls is kept in `state`
 * (payload offset of the next block, skip markers so far), so that a batch can be walked in slices of blocks.
 */
extern "C" __global__ void __launch_bounds__(256)
rtj_scan_walk_kernel(const uint8_t *__restrict__ stream, const rtjgpu_frame_desc *__restrict__ desc,
                     const rtj_dev_table *__restrict__ tables, int F, int nblk,
                     uint32_t *__restrict__ ent, uint32_t *__restrict__ frame_skips,
                     rtj_dev_info *__restrict__ info, int2 *__restrict__ state, int b0, int b1, int slice)
{
    extern __shared__ __align__(16) uint32_t ring_mem[];
    const int lane = threadIdx.x & 31;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = f < F;
    rtjgpu_frame_desc d = {0, 0, 0, 0};
    if (live) {
        d = desc[f];
        const rtj_dev_table &tab = tables[min((int)d.table, RTJ_NUM_TABLES - 1)];      /* descriptors are the caller's memory */
        if (tab.bt8[0] | tab.bt8[1]) live = false;                                     /* raw prefix: rtj_scan_mb_kernel's frame */
    }
    WkFrame fr;
    fr.len = d.length > RTJPEG_B200_HEADER_BYTES ? (int)d.length - RTJPEG_B200_HEADER_BYTES : 0;
    const unsigned long long pay = (unsigned long long)(uintptr_t)stream + d.offset + RTJPEG_B200_HEADER_BYTES;
    fr.base = pay & ~(unsigned long long)(WK_CHUNK - 1);
    fr.skew = (int)(pay - fr.base);
    fr.lo = pay & ~15ull;                                                              /* never before the packet (header: 12 bytes) */
    fr.hi = pay + (unsigned long long)fr.len + RTJGPU_STREAM_SLACK_BYTES;            /* the slack the stream buffer guarantees */
    fr.out = ent + (size_t)f * nblk;
    fr.ring_s = (unsigned)__cvta_generic_to_shared(ring_mem + threadIdx.x * WK_ROW);
    const int len = fr.len;

    int o = 0, skips = 0;
    if (live && b0 > 0) { const int2 st = state[f]; o = st.x; skips = st.y; }
    const int skips0 = skips;
    int blk = b0;

    /* the window: chunks c .. c + 3 of where the walk stands, c .. c + 2 landed before the first step */
    int req = 0;                                                   /* highest chunk asked for */
    if (live && o < len) {
        const int c = (fr.skew + o) >> 7;
        wk_request(fr.ring_s, fr.base, c, fr.lo, fr.hi);
        wk_request(fr.ring_s, fr.base, c + 1, fr.lo, fr.hi);
        wk_request(fr.ring_s, fr.base, c + 2, fr.lo, fr.hi);
        asm volatile("cp.async.commit_group;" ::: "memory");
        req = c + 2;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }

    /* Two blocks per round.  A block is at most 64 bytes, so a round moves the walk by at most one chunk: with the chunk
     * three ahead asked for at the top of every round (one cp.async group per round, empty for most), everything up to two
     * chunks ahead -- all a round can read -- has landed once the previous round's group has. */
    while (live && blk < b1 && o < len) {
        const int want = ((fr.skew + o) >> 7) + 3;
        if (want > req) {
            req = want;
            wk_request(fr.ring_s, fr.base, req, fr.lo, fr.hi);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        wk_step(fr, o, blk, skips);
        if (blk < b1 && o < len) wk_step(fr, o, blk, skips);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");           /* nothing of this lane's may still be in flight when the CTA retires */

    if (live) {
        /* a frame whose stream ended early or mid-block: the missing blocks of this slice get a harmless entry */
        const bool ended = blk < b1;
        for (int b = blk; b < b1; b++) fr.out[b] = RTJ_ENT(min(len, (int)RTJ_ENT_OFF_MASK), 1);
        state[f] = make_int2(ended ? max(o, len) : o, skips);
    }
    /* this slice's skip markers (what K3 of the slice decides on), and -- with the frame's last block -- the frame's counters */
    {
        int n = live ? skips - skips0 : 0;
#pragma unroll
        for (int k = 16; k; k >>= 1) n += __shfl_xor_sync(FULL, n, k);
        if (lane == 0 && n) {
            atomicAdd(&info->slice_skips[slice], (unsigned)n);
            atomicAdd(&info->skipped_blocks, (unsigned long long)n);
        }
    }
    if (live && b1 == nblk) {
        const bool bad = blk < nblk || o > len || len > (int)RTJGPU_MAX_PAYLOAD_BYTES;
        frame_skips[f] = (uint32_t)skips;
        atomicAdd(&info->payload_bytes, (unsigned long long)min(o, len));
        if (bad) {
            atomicAdd(&info->bad_frames, 1u);
            atomicMin((unsigned int *)&info->first_bad_frame, (unsigned int)f);
        }
    }
}

namespace {
constexpr int g_walk_threads = 32;                          /* frames per CTA: one warp, so that the walkers spread over the SMs */
constexpr int g_walk_smem = g_walk_threads * WK_ROW * 4;    /* the rings */
}

extern "C" int rtj_scan_walk_init(void)
{
    return (int)cudaFuncSetAttribute(rtj_scan_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_walk_smem);
}

extern "C" int rtj_launch_scan_walk(const rtj_launch_args *a, int b0, int b1, void *stream)
{
    const int nblk = RTJ_FMT_NBLK(a->fmt, a->w, a->h);
    rtj_scan_walk_kernel<<<(a->F + g_walk_threads - 1) / g_walk_threads, g_walk_threads, g_walk_smem, (cudaStream_t)stream>>>(
        a->d_stream, a->d_desc, a->d_tables, a->F, nblk, a->d_ent, a->d_frame_skips, a->d_info,
        reinterpret_cast<int2 *>(a->d_walk), b0, b1, a->slice);
    return (int)cudaGetLastError();
}
