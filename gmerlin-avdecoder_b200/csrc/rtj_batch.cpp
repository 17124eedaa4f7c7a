/*
 * rtj_batch.cpp -- batch context of the B200 RTjpeg decoder (Level 2 of
 * include/rtjpeg_b200.h): device tables, workspaces, the device-resident
 * decode and the pinned-memory host pipeline.
 *
 * Host-side counterpart of what RTjpeg_decompress (lib/RTjpeg.c:3565-3586)
 * does before it starts walking blocks: header parse and lazy size / quality
 * reconfiguration -- done here once per batch on the CPU (rtjgpu_plan), the
 * block work itself runs in rtj_kernels.cu.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "rtj_common.h"

namespace {

struct Workspace {
    uint32_t     *d_ent = nullptr;
    uint16_t     *d_src = nullptr;
    uint32_t     *d_hardq = nullptr;
    uint16_t     *d_chunk_last = nullptr;
    uint32_t     *d_chunk_mask = nullptr;     /* (owns the allocation d_chunk_last points into) */
    size_t        cap_chunk = 0;
    uint16_t     *d_k3_carry = nullptr;  int    k3_carry_cap = 0;    /* [2][nblk]: last writers handed from slice to slice */
    int2         *d_walk = nullptr;      int    walk_cap = 0;        /* [F]: where every frame's walker stands between slices */
    uint32_t     *d_frame_skips = nullptr;
    rtj_dev_info *d_info = nullptr;
    /* segment-parallel scan */
    uint32_t     *d_seg_sum = nullptr;   size_t seg_sum_cap = 0;     /* entries */
    uint32_t     *d_seg_entry = nullptr; size_t seg_cap = 0;         /* entries of entry / base */
    uint32_t     *d_seg_base = nullptr;
    int32_t      *d_seg_nbf = nullptr;   int    seg_frames_cap = 0;
    uint8_t      *d_seg_del = nullptr;   size_t seg_del_cap = 0;     /* segments */
    uint32_t     *d_seg_gsum = nullptr;  size_t seg_gsum_cap = 0;    /* groups */
    uint32_t     *d_seg_gentry = nullptr, *d_seg_gbase = nullptr;
    size_t        cap_entries = 0;
    int           cap_frames = 0;
    /* K2's position table, rebuilt when the geometry changes */
    void         *d_lut = nullptr;       size_t lut_cap = 0;     /* bytes */
    int           lut_fmt = -1, lut_w = 0, lut_h = 0;
};

constexpr int HOST_SLOTS = 3;
constexpr int TIMING_RING = 256;

/* The two streams a large device batch is worked through on (run_kernels): K1 slice by slice on `scan`,
 * K3 / K2 of every slice behind it on `idct`, forked from and joined to the caller's stream with events. */
struct Pipeline {
    cudaStream_t scan = nullptr, idct = nullptr;
    cudaEvent_t  fork = nullptr, join = nullptr;
    cudaEvent_t  scanned[RTJ_MAX_SLICES] = {};
    bool         ready = false;
};

struct HostSlot {
    cudaStream_t       stream = nullptr;
    cudaEvent_t        decoded = nullptr;     /* kernels of the chunk in this slot have finished */
    Workspace          ws;
    uint8_t           *d_in = nullptr;   size_t d_in_cap = 0;
    uint8_t           *d_out = nullptr;  size_t d_out_cap = 0;
    rtjgpu_frame_desc *d_desc = nullptr; int    desc_cap = 0;
    uint8_t           *h_in = nullptr;   size_t h_in_cap = 0;    /* pinned */
    uint8_t           *h_out = nullptr;  size_t h_out_cap = 0;   /* pinned */
    rtjgpu_frame_desc *h_desc = nullptr; int    h_desc_cap = 0;  /* pinned */
    rtj_dev_info      *h_info = nullptr;                         /* pinned read-back of the chunk's counters */
    /* pending drain of h_out into the caller's memory */
    uint8_t           *pending_dst = nullptr;
    size_t             pending_bytes = 0;
    bool               busy = false;
};

} // namespace

struct rtjgpu_ctx {
    int            device = 0;
    int            last_cuda = 0;
    rtj_dev_table *d_tables = nullptr;
    rtj_host_table h_tables[RTJ_NUM_TABLES];
    Workspace      ws;                        /* rtjgpu_decode_device */
    rtj_dev_info  *h_info_reset = nullptr;    /* pinned template {0,0,0,-1} */
    rtj_dev_info  *h_info = nullptr;          /* pinned read-back */
    int            last_F = 0;
    void          *last_stream = nullptr;
    bool           timing = false;
    cudaEvent_t    ev[TIMING_RING][4] = {};   /* stage brackets of the last TIMING_RING device batches */
    uint64_t       timed_calls = 0;
    uint64_t       launches = 0;
    HostSlot       slot[HOST_SLOTS];
    bool           slots_ready = false;
    uint8_t       *d_host_carry = nullptr;    /* carry plane of the host pipeline */
    size_t         d_host_carry_cap = 0;
    int            scan_mode = RTJGPU_SCAN_AUTO;
    int            format = RTJ_YUV420;
    Pipeline       pipe;
    int            pipeline_mode = RTJGPU_PIPELINE_AUTO;
    int            slice_frames = 1184, slice0_frames = 1184;   /* multiples of RTJ_RESOLVE_T */
    bool           scan_priority = false;
    int            frame_run = 0;             /* rtjgpu_set_frame_runs: 0 = by the batch before, 1 = off, n = forced */
    unsigned long long *h_skips_seen = nullptr;    /* pinned: skipped blocks, raw-prefix frames of the last batch whose K3 has run; raw-prefix frames
                                                    * the self-synchronising walk gave up when it was last tried */
    bool           slices_forced = false;     /* rtjgpu_set_pipeline / the environment gave a slice size: it holds in either arrangement */
    /* encoder: configuration, state between calls, workspace */
    int            enc_quality = 0, enc_lb8 = 0, enc_cb8 = 0;
    int            enc_key_rate = 0, enc_key_count = 0, enc_lm = 0, enc_cm = 0;
    bool           enc_clear_old = true;
    int            enc_w = 0, enc_h = 0, enc_fmt = -1;
    int32_t       *d_enc_qt = nullptr;
    int16_t       *d_enc_old = nullptr;       size_t enc_old_cap = 0;      /* blocks */
    uint8_t       *d_enc_slots = nullptr;     size_t enc_blocks_cap = 0;   /* blocks of a batch */
    uint8_t       *d_enc_lens = nullptr;
    uint32_t      *d_enc_boff = nullptr;
    uint32_t      *d_enc_fsize = nullptr;     size_t enc_frames_cap = 0;
    uint64_t      *d_enc_total = nullptr;
    uint64_t      *h_enc_total = nullptr;     /* pinned */
    void          *enc_stream = nullptr;
    uint64_t       host_bad = 0;              /* overrun frames seen by the current rtjgpu_decode_host call */
    uint32_t      *h_host_skips = nullptr;    size_t host_skips_cap = 0;   /* pinned: per-frame skip counts of that call */
    int            host_F = 0;
};

namespace {

#define CK(ctx, call)                                             \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) {                                 \
            (ctx)->last_cuda = (int)e__;                          \
            return RTJGPU_E_CUDA;                                 \
        }                                                         \
    } while (0)

int ws_reserve(rtjgpu_ctx *ctx, Workspace *ws, int F, int nblk)
{
    const size_t need = (size_t)F * (size_t)nblk;
    if (need > ws->cap_entries) {
        if (ws->d_ent) cudaFree(ws->d_ent);
        if (ws->d_src) cudaFree(ws->d_src);
        if (ws->d_hardq) cudaFree(ws->d_hardq);
        ws->d_ent = nullptr; ws->d_src = nullptr; ws->d_hardq = nullptr; ws->cap_entries = 0;
        CK(ctx, cudaMalloc(&ws->d_ent, need * sizeof(uint32_t)));
        CK(ctx, cudaMalloc(&ws->d_src, need * sizeof(uint16_t)));
        CK(ctx, cudaMalloc(&ws->d_hardq, need * 2 * sizeof(uint32_t)));       /* two words an entry: destination and source block */
        ws->cap_entries = need;
    }
    /* K3's per-chunk notes: a 32-bit skip mask and a 16-bit last writer per (chunk of frames, position), the masks first.
     * (Sized on its own: few frames of a large picture need more of these than many frames of a small one.) */
    const size_t nchunk = ((size_t)F / RTJ_RESOLVE_T + 1) * (size_t)nblk;
    if (nchunk > ws->cap_chunk) {
        if (ws->d_chunk_mask) cudaFree(ws->d_chunk_mask);
        ws->d_chunk_mask = nullptr; ws->d_chunk_last = nullptr; ws->cap_chunk = 0;
        CK(ctx, cudaMalloc(&ws->d_chunk_mask, nchunk * (sizeof(uint32_t) + sizeof(uint16_t))));
        ws->cap_chunk = nchunk;
    }
    ws->d_chunk_last = reinterpret_cast<uint16_t *>(ws->d_chunk_mask + ws->cap_chunk);
    if (nblk > ws->k3_carry_cap) {
        if (ws->d_k3_carry) cudaFree(ws->d_k3_carry);
        ws->d_k3_carry = nullptr; ws->k3_carry_cap = 0;
        /* ... and behind them K3's arrival counters, one per 128 positions (zero between launches: the last arrival resets its own) */
        const size_t carry_bytes = ((size_t)2 * nblk * sizeof(uint16_t) + 15) & ~(size_t)15, count_bytes = ((size_t)nblk / 128 + 1) * sizeof(uint32_t);
        CK(ctx, cudaMalloc(&ws->d_k3_carry, carry_bytes + count_bytes));
        CK(ctx, cudaMemset(reinterpret_cast<uint8_t *>(ws->d_k3_carry) + carry_bytes, 0, count_bytes));
        ws->k3_carry_cap = nblk;
    }
    if (F > ws->cap_frames) {
        if (ws->d_frame_skips) cudaFree(ws->d_frame_skips);
        ws->d_frame_skips = nullptr; ws->cap_frames = 0;
        CK(ctx, cudaMalloc(&ws->d_frame_skips, (size_t)2 * F * sizeof(uint32_t)));      /* ... and K1's hand-over flags behind them */
        ws->cap_frames = F;
    }
    if (F > ws->walk_cap) {
        if (ws->d_walk) cudaFree(ws->d_walk);
        ws->d_walk = nullptr; ws->walk_cap = 0;
        CK(ctx, cudaMalloc(&ws->d_walk, (size_t)F * sizeof(int2)));
        ws->walk_cap = F;
    }
    if (!ws->d_info) CK(ctx, cudaMalloc(&ws->d_info, sizeof(rtj_dev_info)));
    return RTJGPU_OK;
}

/* Workspace of the segment-parallel scan, when the batch qualifies: forced by the scan mode, or AUTO
 * with a batch too small to fill the device with one CTA per frame.  Fills *sp (sum == NULL: not used). */
/* AUTO's choice between one CTA per frame and the segment-parallel arrangement.  The host knows the batch's geometry, not
 * its payload (packets and descriptors are device memory): a block of ordinary material is ~4 bytes.  Measured on a B200
 * (DESIGN.md section 3): one CTA per frame takes ~40 us per 40 KB segment of a frame whatever the batch, until the batch
 * fills the device; the segment-parallel passes take ~55 us + 11 ns per KB of batch.  Frames of one segment (up to about
 * 720x576) are never worth cutting up; large frames are, in small batches (a single 1920x1088 frame: 0.06 against 0.18 ms). */
bool auto_wants_segments(const rtjgpu_ctx *ctx, int F, int nblk)
{
    /* Frames with a raw prefix (quality above 170) are another matter: their kernel walks 4 KB segments at ~25 us each, and
     * a frame is many of them.  Whether a batch holds such frames only the device knows; the batch before is the guide (its
     * count arrives in pinned memory, like the skip count that K2's arrangement goes by). */
    /* (The self-synchronising walk takes such frames at a few us per 40 KB when their streams forget their past -- ordinary
     * material does; noise at a high quality does not, and is what the segment-parallel passes remain for.) */
    const volatile unsigned long long *seen = ctx->h_skips_seen;
    if (seen[1]) {
        if (seen[2]) return F <= 256;
        /* measured at a quality of 255 (~9 bytes a block): the walk ~21 us per 40 KB segment of a frame whatever the batch
         * (720x576: 0.067 ms for 1 .. 96 frames; 1920x1088: 0.22 - 0.25 ms for 4 .. 32), the segment-parallel passes
         * ~90 us + 20 ns per KB of batch (720x576: 0.097 / 0.133 / 0.239 ms for 1 / 32 / 96 frames; 1920x1088: 0.123 / 0.370 ms
         * for 4 / 32): segments for a handful of large frames only */
        const double raw_kb = (double)nblk * 9.0 / 1024.0;
        const int raw_nseg = (int)((raw_kb + 39.99) / 40.0);
        return 0.090 + 2.0e-5 * (double)F * raw_kb < 0.021 * (double)raw_nseg;
    }
    const double frame_kb = (double)nblk * 4.0 / 1024.0;
    const int nseg = (int)((frame_kb + 39.99) / 40.0);
    return nseg > 1 && 0.055 + 1.1e-5 * (double)F * frame_kb < 0.040 * (double)nseg;
}
constexpr size_t SEG_MAX_SUM_BYTES = (size_t)1 << 30;
constexpr size_t SEG_MAX_DEL_BYTES = (size_t)2 << 30;

int seg_reserve(rtjgpu_ctx *ctx, Workspace *ws, int F, int nblk, int scan_mode, rtj_seg_plan *sp)
{
    memset(sp, 0, sizeof(*sp));
    if (scan_mode != RTJGPU_SCAN_SEGMENT && !(scan_mode == RTJGPU_SCAN_AUTO && auto_wants_segments(ctx, F, nblk))) return RTJGPU_OK;
    /* a frame needs at most 64 bytes per block */
    const size_t maxseg = ((size_t)nblk * 64 + RTJ_SEG_BYTES_MB - 1) / RTJ_SEG_BYTES_MB + 1;   /* the smaller segment size rules */
    const size_t nseg = (size_t)F * maxseg, nsum = nseg * RTJ_SEG_NE;
    if (nsum * sizeof(uint32_t) > SEG_MAX_SUM_BYTES) return RTJGPU_OK;          /* too big: one CTA per frame instead */
    if (nsum > ws->seg_sum_cap) {
        if (ws->d_seg_sum) cudaFree(ws->d_seg_sum);
        ws->d_seg_sum = nullptr; ws->seg_sum_cap = 0;
        CK(ctx, cudaMalloc(&ws->d_seg_sum, nsum * sizeof(uint32_t)));
        ws->seg_sum_cap = nsum;
    }
    if (nseg > ws->seg_cap) {
        if (ws->d_seg_entry) cudaFree(ws->d_seg_entry);
        if (ws->d_seg_base) cudaFree(ws->d_seg_base);
        ws->d_seg_entry = ws->d_seg_base = nullptr; ws->seg_cap = 0;
        CK(ctx, cudaMalloc(&ws->d_seg_entry, nseg * sizeof(uint32_t)));
        CK(ctx, cudaMalloc(&ws->d_seg_base, nseg * sizeof(uint32_t)));
        ws->seg_cap = nseg;
    }
    if (F > ws->seg_frames_cap) {
        if (ws->d_seg_nbf) cudaFree(ws->d_seg_nbf);
        ws->d_seg_nbf = nullptr; ws->seg_frames_cap = 0;
        CK(ctx, cudaMalloc(&ws->d_seg_nbf, (size_t)F * sizeof(int32_t)));
        ws->seg_frames_cap = F;
    }
    /* the block lengths of the first pass are kept for the second when that fits (it saves the larger half of the second) */
    if (nseg * RTJ_SEG_DEL_BYTES <= SEG_MAX_DEL_BYTES) {
        if (nseg > ws->seg_del_cap) {
            if (ws->d_seg_del) cudaFree(ws->d_seg_del);
            ws->d_seg_del = nullptr; ws->seg_del_cap = 0;
            if (cudaMalloc(&ws->d_seg_del, nseg * RTJ_SEG_DEL_BYTES) == cudaSuccess) ws->seg_del_cap = nseg;
            else { ws->d_seg_del = nullptr; cudaGetLastError(); }         /* no room: the second pass works them out again */
        }
        sp->del = ws->seg_del_cap >= nseg ? ws->d_seg_del : nullptr;
    }
    if (maxseg >= RTJ_SEG_GROUP_MIN_SEGS) {
        const size_t ngroups = (maxseg + RTJ_SEG_GROUP - 1) / RTJ_SEG_GROUP, ng = (size_t)F * ngroups;
        if (ng > ws->seg_gsum_cap) {
            if (ws->d_seg_gsum) cudaFree(ws->d_seg_gsum);
            if (ws->d_seg_gentry) cudaFree(ws->d_seg_gentry);
            if (ws->d_seg_gbase) cudaFree(ws->d_seg_gbase);
            ws->d_seg_gsum = ws->d_seg_gentry = ws->d_seg_gbase = nullptr; ws->seg_gsum_cap = 0;
            CK(ctx, cudaMalloc(&ws->d_seg_gsum, ng * RTJ_SEG_NE * sizeof(uint32_t)));
            CK(ctx, cudaMalloc(&ws->d_seg_gentry, ng * sizeof(uint32_t)));
            CK(ctx, cudaMalloc(&ws->d_seg_gbase, ng * sizeof(uint32_t)));
            ws->seg_gsum_cap = ng;
        }
        sp->gsum = ws->d_seg_gsum; sp->gentry = ws->d_seg_gentry; sp->gbase = ws->d_seg_gbase; sp->ngroups = (int)ngroups;
    }
    sp->sum = ws->d_seg_sum; sp->entry = ws->d_seg_entry; sp->base = ws->d_seg_base; sp->nbf = ws->d_seg_nbf;
    sp->maxseg = (int)maxseg;
    return RTJGPU_OK;
}

void ws_release(Workspace *ws)
{
    if (ws->d_ent) cudaFree(ws->d_ent);
    if (ws->d_src) cudaFree(ws->d_src);
    if (ws->d_hardq) cudaFree(ws->d_hardq);
    if (ws->d_chunk_mask) cudaFree(ws->d_chunk_mask);
    if (ws->d_k3_carry) cudaFree(ws->d_k3_carry);
    if (ws->d_walk) cudaFree(ws->d_walk);
    if (ws->d_frame_skips) cudaFree(ws->d_frame_skips);
    if (ws->d_info) cudaFree(ws->d_info);
    if (ws->d_seg_sum) cudaFree(ws->d_seg_sum);
    if (ws->d_seg_entry) cudaFree(ws->d_seg_entry);
    if (ws->d_seg_base) cudaFree(ws->d_seg_base);
    if (ws->d_seg_nbf) cudaFree(ws->d_seg_nbf);
    if (ws->d_seg_del) cudaFree(ws->d_seg_del);
    if (ws->d_seg_gsum) cudaFree(ws->d_seg_gsum);
    if (ws->d_seg_gentry) cudaFree(ws->d_seg_gentry);
    if (ws->d_seg_gbase) cudaFree(ws->d_seg_gbase);
    if (ws->d_lut) cudaFree(ws->d_lut);
    *ws = Workspace();
}

int pipeline_init(rtjgpu_ctx *ctx)
{
    Pipeline &p = ctx->pipe;
    if (p.ready) return RTJGPU_OK;
    int least = 0, greatest = 0;
    CK(ctx, cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CK(ctx, cudaStreamCreateWithPriority(&p.scan, cudaStreamNonBlocking, ctx->scan_priority ? greatest : least));
    CK(ctx, cudaStreamCreateWithPriority(&p.idct, cudaStreamNonBlocking, least));
    CK(ctx, cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming));
    CK(ctx, cudaEventCreateWithFlags(&p.join, cudaEventDisableTiming));
    for (int i = 0; i < RTJ_MAX_SLICES; i++) CK(ctx, cudaEventCreateWithFlags(&p.scanned[i], cudaEventDisableTiming));
    p.ready = true;
    return RTJGPU_OK;
}

void pipeline_release(Pipeline *p)
{
    if (p->scan) cudaStreamDestroy(p->scan);
    if (p->idct) cudaStreamDestroy(p->idct);
    if (p->fork) cudaEventDestroy(p->fork);
    if (p->join) cudaEventDestroy(p->join);
    for (int i = 0; i < RTJ_MAX_SLICES; i++) if (p->scanned[i]) cudaEventDestroy(p->scanned[i]);
    *p = Pipeline();
}

/* The slices of a batch: [first[i], first[i + 1]), whole chunks of RTJ_RESOLVE_T frames, at most RTJ_MAX_SLICES. */
int plan_slices(const rtjgpu_ctx *ctx, int F, int *first)
{
    auto up = [](int v) { return (v + RTJ_RESOLVE_T - 1) / RTJ_RESOLVE_T * RTJ_RESOLVE_T; };
    int s0 = up(std::max(ctx->slice0_frames, 1)), sl = up(std::max(ctx->slice_frames, 1));
    /* stage after stage, slices only serve to bound K3's look-back: few and large (two launches each) */
    if (ctx->pipeline_mode != RTJGPU_PIPELINE_SLICED && !ctx->slices_forced) s0 = sl = std::max(sl, 4096);
    if (F > s0 && (F - s0 + sl - 1) / sl + 1 > RTJ_MAX_SLICES) sl = up((F - s0 + RTJ_MAX_SLICES - 2) / (RTJ_MAX_SLICES - 1));
    int n = 0;
    first[0] = 0;
    for (int at = 0; at < F; n++) {
        at = std::min(F, at + (n == 0 ? s0 : sl));
        first[n + 1] = at;
    }
    return n;
}

/*
 * K1 -> K3 -> K2 -> K2b.  ev != NULL brackets the stages with events.
 *
 * The batch is worked through in slices of frames.  Serial arrangement (small batches, the segment-parallel and the
 * serial scans, RTJGPU_PIPELINE_SERIAL): everything on the caller's stream -- K1 over the batch, K3 slice by slice
 * (its look-back never leaves a slice), K2 over the batch.  Pipelined arrangement (allow_pipe, two slices or more):
 * K1 of slice s + 1 runs on a second stream next to K3 / K2 of slice s.  K1 is bound by the ALU pipe and by its
 * barriers, K2 by the FMA pipe: side by side on an SM they fill each other's idle issue slots.  K1's stream has the
 * higher priority, so the CTAs of its next slice (one wave, sized by the slice) take their share of every SM as
 * K2's short-lived CTAs retire.  The last writer of a skipped block lies in an earlier frame, i.e. in a slice
 * that is already scanned, so nothing else has to be ordered.
 */
/* where K2's pixels go when they leave as packed RGB (rtjgpu_decode_device_rgb) */
struct RgbOut {
    uint8_t *d_rgb = nullptr;
    size_t   row_pitch = 0, frame_pitch = 0;
    int      kind = 0;
    unsigned alpha = 0;
    uint8_t *d_last_yuv = nullptr;
};

int run_kernels(rtjgpu_ctx *ctx, Workspace *ws, const uint8_t *d_stream, const rtjgpu_frame_desc *d_desc,
                int F, int w, int h, uint8_t *d_out, const uint8_t *d_carry, cudaStream_t st, cudaEvent_t *ev,
                bool allow_pipe, const RgbOut *rgb = nullptr)
{
    rtj_launch_args a;
    a.d_stream = d_stream; a.d_desc = d_desc; a.d_tables = ctx->d_tables;
    a.F = F; a.w = w; a.h = h;
    a.f0 = 0; a.f1 = F; a.slice = 0;
    a.fmt = ctx->format;
    a.row0 = 0; a.row1 = RTJ_FMT_UNITS_Y(ctx->format, h);
    /* K2 works through runs of frames with the strip staying on chip when the stream skips blocks -- which only the device
     * knows for this batch.  The batch before is the guide (its count arrives in pinned memory; a count that is not there yet
     * is the one before it): streams do not change their nature from batch to batch, and either arrangement is exact. */
    a.k2_run = ctx->frame_run ? ctx->frame_run : (*(volatile unsigned long long *)ctx->h_skips_seen ? RTJ_K2_RUN_FRAMES : 1);
    a.h_skips_seen = ctx->h_skips_seen;
    a.raw_expected = ((volatile unsigned long long *)ctx->h_skips_seen)[1] != 0;
    a.d_walk = ws->d_walk;
    a.d_redo = ws->d_frame_skips + ws->cap_frames;
    a.d_ent = ws->d_ent; a.d_src = ws->d_src; a.d_frame_skips = ws->d_frame_skips; a.d_info = ws->d_info;
    a.d_hardq = ws->d_hardq; a.d_chunk_last = ws->d_chunk_last; a.d_chunk_mask = ws->d_chunk_mask;
    a.d_k3_in = nullptr; a.d_k3_out = ws->d_k3_carry;
    a.d_k3_count = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(ws->d_k3_carry) + (((size_t)2 * ws->k3_carry_cap * sizeof(uint16_t) + 15) & ~(size_t)15));
    a.d_out = d_out; a.d_carry = d_carry;
    a.d_rgb = rgb ? rgb->d_rgb : nullptr;
    a.rgb_row_pitch = rgb ? rgb->row_pitch : 0; a.rgb_frame_pitch = rgb ? rgb->frame_pitch : 0;
    a.rgb_kind = rgb ? rgb->kind : 0; a.rgb_alpha = rgb ? rgb->alpha : 0; a.d_last_yuv = rgb ? rgb->d_last_yuv : nullptr;
    a.scan_mode = ctx->scan_mode;
    const int nblk = RTJ_FMT_NBLK(ctx->format, w, h);
    {
        const int rc = seg_reserve(ctx, ws, F, nblk, ctx->scan_mode, &a.seg);
        if (rc) return rc;
    }

    a.d_lut = nullptr;
    {
        if (ws->lut_fmt != ctx->format || ws->lut_w != w || ws->lut_h != h) {
            const size_t need = rtj_lut_bytes(ctx->format, w, h);
            if (need > ws->lut_cap) {
                if (ws->d_lut) cudaFree(ws->d_lut);
                ws->d_lut = nullptr; ws->lut_cap = 0; ws->lut_fmt = -1;
                CK(ctx, cudaMalloc(&ws->d_lut, need));
                ws->lut_cap = need;
            }
            const int e0 = rtj_launch_build_lut(ctx->format, w, h, ws->d_lut, st);
            if (e0) { ctx->last_cuda = e0; return RTJGPU_E_CUDA; }
            ws->lut_fmt = ctx->format; ws->lut_w = w; ws->lut_h = h;
            ctx->launches += 1;
        }
        a.d_lut = ws->d_lut;
    }

    int first[RTJ_MAX_SLICES + 1];
    const int nslices = plan_slices(ctx, F, first);
    const bool chunk_scan = !a.seg.sum && (ctx->scan_mode == RTJGPU_SCAN_AUTO || ctx->scan_mode == RTJGPU_SCAN_CHUNK);
    const bool piped = allow_pipe && chunk_scan && nslices >= 2 && ctx->pipeline_mode == RTJGPU_PIPELINE_SLICED;
    auto slice_args = [&](int i) {
        a.f0 = first[i]; a.f1 = first[i + 1]; a.slice = i;
        a.d_k3_in = i == 0 ? nullptr : ws->d_k3_carry + (size_t)(i & 1) * nblk;
        a.d_k3_out = ws->d_k3_carry + (size_t)((i + 1) & 1) * nblk;
    };
#define LAUNCHED(call, count)                                                  \
    do {                                                                       \
        const int e__ = (call);                                                \
        if (e__ < 0 || ((count) && e__ > 0)) { ctx->last_cuda = e__ < 0 ? -e__ : e__; return RTJGPU_E_CUDA; } \
        ctx->launches += (count) ? (uint64_t)(count) : (uint64_t)e__;          \
    } while (0)

    CK(ctx, cudaMemcpyAsync(ws->d_info, ctx->h_info_reset, sizeof(rtj_dev_info), cudaMemcpyHostToDevice, st));
    if (ev) CK(ctx, cudaEventRecord(ev[0], st));
    if (!piped) {
        LAUNCHED(rtj_launch_scan(&a, st), 0);                  /* returns its launches */
        if (ev) CK(ctx, cudaEventRecord(ev[1], st));
        for (int i = 0; i < nslices; i++) {
            slice_args(i);                                      /* (the one scan over the batch counted its skips as slice 0) */
            LAUNCHED(rtj_launch_resolve(&a, st), 2);
        }
        a.f0 = 0; a.f1 = F;
        if (ev) CK(ctx, cudaEventRecord(ev[2], st));
        LAUNCHED(rtj_launch_idct(&a, st), 1);
        if (!rgb) LAUNCHED(rtj_launch_idct_hard(&a, st), 2);
        if (ev) CK(ctx, cudaEventRecord(ev[3], st));
        return RTJGPU_OK;
    }

    {
        const int rc = pipeline_init(ctx);
        if (rc) return rc;
    }
    Pipeline &p = ctx->pipe;
    CK(ctx, cudaEventRecord(p.fork, st));
    CK(ctx, cudaStreamWaitEvent(p.scan, p.fork, 0));
    CK(ctx, cudaStreamWaitEvent(p.idct, p.fork, 0));
    for (int i = 0; i < nslices; i++) {
        slice_args(i);
        LAUNCHED(rtj_launch_scan(&a, p.scan), 0);
        CK(ctx, cudaEventRecord(p.scanned[i], p.scan));
        CK(ctx, cudaStreamWaitEvent(p.idct, p.scanned[i], 0));
        LAUNCHED(rtj_launch_resolve(&a, p.idct), 2);
        LAUNCHED(rtj_launch_idct(&a, p.idct), 1);
    }
    if (ev) {                                                   /* stages overlap: scan = until the last slice is scanned, idct = the rest */
        CK(ctx, cudaEventRecord(ev[1], p.scan));
        CK(ctx, cudaEventRecord(ev[2], p.scan));
    }
    a.f0 = 0; a.f1 = F;
    if (!rgb) LAUNCHED(rtj_launch_idct_hard(&a, p.idct), 2);
    CK(ctx, cudaEventRecord(p.join, p.idct));
    CK(ctx, cudaStreamWaitEvent(st, p.join, 0));
    if (ev) CK(ctx, cudaEventRecord(ev[3], st));
#undef LAUNCHED
    return RTJGPU_OK;
}

inline uint32_t rd_u32le(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
inline uint16_t rd_u16le(const uint8_t *p) { return (uint16_t)(p[0] | p[1] << 8); }

template <typename T>
int grow_device(rtjgpu_ctx *ctx, T **p, size_t *cap, size_t need)
{
    if (need <= *cap) return RTJGPU_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    CK(ctx, cudaMalloc(p, need * sizeof(T)));
    *cap = need;
    return RTJGPU_OK;
}

template <typename T>
int grow_pinned(rtjgpu_ctx *ctx, T **p, size_t *cap, size_t need)
{
    if (need <= *cap) return RTJGPU_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr; *cap = 0;
    CK(ctx, cudaMallocHost(p, need * sizeof(T)));
    *cap = need;
    return RTJGPU_OK;
}

} // namespace

extern "C" {

const char *rtjgpu_strerror(int code)
{
    switch (code) {
    case RTJGPU_OK:        return "ok";
    case RTJGPU_E_CUDA:    return "CUDA call failed";
    case RTJGPU_E_ARG:     return "bad argument";
    case RTJGPU_E_HEADER:  return "packet shorter than its header or framesize";
    case RTJGPU_E_SIZE:    return "width/height zero, not a multiple of 16, or changing inside a batch";
    case RTJGPU_E_FORMAT:  return "unknown picture format or converter";
    case RTJGPU_E_OVERRUN: return "block stream runs past the end of its packet";
    case RTJGPU_E_TOOBIG:  return "batch exceeds RTJGPU_MAX_FRAMES_PER_BATCH / RTJGPU_MAX_PAYLOAD_BYTES";
    case RTJGPU_E_NOMEM:   return "out of memory";
    default:               return "unknown error";
    }
}

int rtjgpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int rtjgpu_create(int device, rtjgpu_ctx **out)
{
    if (!out) return RTJGPU_E_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || device < 0 || device >= n) return RTJGPU_E_CUDA;   /* no CPU fallback */
    rtjgpu_ctx *ctx = new (std::nothrow) rtjgpu_ctx();
    if (!ctx) return RTJGPU_E_NOMEM;
    ctx->device = device;
    /* development knobs (tools/): slice sizes of the pipelined arrangement, K1's stream priority */
    if (const char *v = getenv("RTJPEG_B200_SLICE")) { ctx->slice_frames = ctx->slice0_frames = std::max(32, atoi(v)); ctx->slices_forced = true; }
    if (const char *v = getenv("RTJPEG_B200_SLICE0")) ctx->slice0_frames = std::max(32, atoi(v));
    if (const char *v = getenv("RTJPEG_B200_SCAN_PRIO")) ctx->scan_priority = atoi(v) != 0;
    if (const char *v = getenv("RTJPEG_B200_SCAN")) ctx->scan_mode = std::min(std::max(atoi(v), 0), (int)RTJGPU_SCAN_SYNC);
    if (const char *v = getenv("RTJPEG_B200_K2_RUN")) ctx->frame_run = std::min(std::max(atoi(v), 0), 64);
    if (const char *v = getenv("RTJPEG_B200_PIPELINE")) ctx->pipeline_mode = std::min(std::max(atoi(v), 0), (int)RTJGPU_PIPELINE_SLICED);
    int rc = RTJGPU_OK;
    do {
        if ((e = cudaSetDevice(device)) != cudaSuccess) { rc = RTJGPU_E_CUDA; break; }
        if (int k = rtj_kernels_init()) { e = (cudaError_t)k; rc = RTJGPU_E_CUDA; break; }
        /* table 0: never-configured instance (all zero, lb8 = cb8 = 0); 1..255: quality; 256: custom */
        memset(ctx->h_tables, 0, sizeof(ctx->h_tables));
        std::vector<rtj_dev_table> dev(RTJ_NUM_TABLES);
        for (int q = 1; q <= 255; q++) rtj_table_from_quality(q, &ctx->h_tables[q]);
        for (int t = 0; t < RTJ_NUM_TABLES; t++) rtj_table_to_device_layout(&ctx->h_tables[t], &dev[t]);
        if ((e = cudaMalloc(&ctx->d_tables, sizeof(rtj_dev_table) * RTJ_NUM_TABLES)) != cudaSuccess) { rc = RTJGPU_E_CUDA; break; }
        if ((e = cudaMemcpy(ctx->d_tables, dev.data(), sizeof(rtj_dev_table) * RTJ_NUM_TABLES, cudaMemcpyHostToDevice)) != cudaSuccess) { rc = RTJGPU_E_CUDA; break; }
        if ((e = cudaMallocHost(&ctx->h_info_reset, sizeof(rtj_dev_info))) != cudaSuccess) { rc = RTJGPU_E_CUDA; break; }
        if ((e = cudaMallocHost(&ctx->h_info, sizeof(rtj_dev_info))) != cudaSuccess) { rc = RTJGPU_E_CUDA; break; }
        if ((e = cudaMallocHost(&ctx->h_skips_seen, 3 * sizeof(unsigned long long))) != cudaSuccess) { rc = RTJGPU_E_CUDA; break; }
        ctx->h_skips_seen[0] = ctx->h_skips_seen[1] = ctx->h_skips_seen[2] = 0;
        memset(ctx->h_info_reset, 0, sizeof(rtj_dev_info));
        ctx->h_info_reset->first_bad_frame = -1;
        *ctx->h_info = *ctx->h_info_reset;
        for (int i = 0; i < TIMING_RING * 4 && rc == RTJGPU_OK; i++)
            if ((e = cudaEventCreate(&ctx->ev[i / 4][i % 4])) != cudaSuccess) rc = RTJGPU_E_CUDA;
    } while (0);
    if (rc != RTJGPU_OK) {
        ctx->last_cuda = (int)e;
        rtjgpu_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return RTJGPU_OK;
}

void rtjgpu_destroy(rtjgpu_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    ws_release(&ctx->ws);
    pipeline_release(&ctx->pipe);
    for (int i = 0; i < HOST_SLOTS; i++) {
        HostSlot &s = ctx->slot[i];
        ws_release(&s.ws);
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        if (s.d_desc) cudaFree(s.d_desc);
        if (s.h_in) cudaFreeHost(s.h_in);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.h_desc) cudaFreeHost(s.h_desc);
        if (s.h_info) cudaFreeHost(s.h_info);
        if (s.decoded) cudaEventDestroy(s.decoded);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    if (ctx->d_host_carry) cudaFree(ctx->d_host_carry);
    if (ctx->h_host_skips) cudaFreeHost(ctx->h_host_skips);
    if (ctx->d_enc_qt) cudaFree(ctx->d_enc_qt);
    if (ctx->d_enc_old) cudaFree(ctx->d_enc_old);
    if (ctx->d_enc_slots) cudaFree(ctx->d_enc_slots);
    if (ctx->d_enc_lens) cudaFree(ctx->d_enc_lens);
    if (ctx->d_enc_boff) cudaFree(ctx->d_enc_boff);
    if (ctx->d_enc_fsize) cudaFree(ctx->d_enc_fsize);
    if (ctx->d_enc_total) cudaFree(ctx->d_enc_total);
    if (ctx->h_enc_total) cudaFreeHost(ctx->h_enc_total);
    if (ctx->d_tables) cudaFree(ctx->d_tables);
    if (ctx->h_info_reset) cudaFreeHost(ctx->h_info_reset);
    if (ctx->h_info) cudaFreeHost(ctx->h_info);
    if (ctx->h_skips_seen) cudaFreeHost(ctx->h_skips_seen);
    for (int i = 0; i < TIMING_RING * 4; i++) if (ctx->ev[i / 4][i % 4]) cudaEventDestroy(ctx->ev[i / 4][i % 4]);
    delete ctx;
}

int rtjgpu_last_cuda_error(const rtjgpu_ctx *ctx) { return ctx ? ctx->last_cuda : 0; }
uint64_t rtjgpu_launch_count(const rtjgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }
void rtjgpu_enable_timing(rtjgpu_ctx *ctx, int on) { if (ctx) ctx->timing = on != 0; }

int rtjgpu_set_scan_mode(rtjgpu_ctx *ctx, int mode)
{
    if (!ctx || mode < RTJGPU_SCAN_AUTO || mode > RTJGPU_SCAN_SYNC) return RTJGPU_E_ARG;
    ctx->scan_mode = mode;
    return RTJGPU_OK;
}

int rtjgpu_set_pipeline(rtjgpu_ctx *ctx, int mode, int slice_frames)
{
    if (!ctx || mode < RTJGPU_PIPELINE_AUTO || mode > RTJGPU_PIPELINE_SLICED || slice_frames < 0) return RTJGPU_E_ARG;
    ctx->pipeline_mode = mode;
    if (slice_frames) {
        ctx->slice_frames = ctx->slice0_frames = (slice_frames + RTJ_RESOLVE_T - 1) / RTJ_RESOLVE_T * RTJ_RESOLVE_T;
        ctx->slices_forced = true;
    }
    return RTJGPU_OK;
}

int rtjgpu_set_frame_runs(rtjgpu_ctx *ctx, int frames)
{
    if (!ctx || frames < 0 || frames > 64) return RTJGPU_E_ARG;
    ctx->frame_run = frames;
    return RTJGPU_OK;
}

int rtjgpu_set_format(rtjgpu_ctx *ctx, int format)
{
    if (!ctx || format < RTJ_YUV420 || format > RTJ_RGB8) return RTJGPU_E_ARG;
    ctx->format = format;
    return RTJGPU_OK;
}

/* ---- encoder ---------------------------------------------------------------------------------------- */

int rtjgpu_encoder_set_quality(rtjgpu_ctx *ctx, int quality)
{
    if (!ctx) return RTJGPU_E_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    if (quality < 1) quality = 1;
    if (quality > 255) quality = 255;
    int32_t qt[128];
    rtj_encoder_table_from_quality(quality, qt, &ctx->enc_lb8, &ctx->enc_cb8);
    if (!ctx->d_enc_qt) CK(ctx, cudaMalloc(&ctx->d_enc_qt, sizeof(qt)));
    if (ctx->enc_stream) CK(ctx, cudaStreamSynchronize((cudaStream_t)ctx->enc_stream));   /* a batch may still read the old tables */
    CK(ctx, cudaMemcpy(ctx->d_enc_qt, qt, sizeof(qt), cudaMemcpyHostToDevice));
    ctx->enc_quality = quality;
    return RTJGPU_OK;
}

int rtjgpu_encoder_set_intra(rtjgpu_ctx *ctx, int key_rate, int lm, int cm)
{
    if (!ctx) return RTJGPU_E_ARG;
    ctx->enc_key_rate = key_rate < 0 ? 0 : key_rate > 255 ? 255 : key_rate;
    ctx->enc_lm = lm < 0 ? 0 : lm > 16 ? 16 : lm;
    ctx->enc_cm = cm < 0 ? 0 : cm > 16 ? 16 : cm;
    ctx->enc_clear_old = true;           /* lib/RTjpeg.c:2488: the stored blocks are cleared; the key counter is not touched */
    return RTJGPU_OK;
}

int rtjgpu_encoder_reset(rtjgpu_ctx *ctx)
{
    if (!ctx) return RTJGPU_E_ARG;
    ctx->enc_key_count = 0;
    ctx->enc_clear_old = true;
    return RTJGPU_OK;
}

int rtjgpu_encode_device(rtjgpu_ctx *ctx, const uint8_t *d_frames, int F, int w, int h,
                         uint8_t *d_stream, size_t capacity, uint64_t *d_offsets, void *cuda_stream)
{
    if (!ctx || F < 0) return RTJGPU_E_ARG;
    if (ctx->format != RTJ_YUV420 && ctx->format != RTJ_YUV422) return RTJGPU_E_FORMAT;
    if (w <= 0 || h <= 0 || (w & 15) || (h & 15) || w > 65535 || h > 65535) return RTJGPU_E_SIZE;
    if (F > RTJGPU_MAX_FRAMES_PER_BATCH) return RTJGPU_E_TOOBIG;
    if (!d_offsets || (F && (!d_frames || !d_stream)) || ((uintptr_t)d_frames & 7) || ((uintptr_t)d_stream & 3)) return RTJGPU_E_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t nblk = (size_t)RTJ_FMT_NBLK(ctx->format, w, h);
    if (!ctx->d_enc_qt) {                /* never configured: the all-zero tables of a fresh RTjpeg_t (lib/RTjpeg.c:2499) */
        CK(ctx, cudaMalloc(&ctx->d_enc_qt, 128 * sizeof(int32_t)));
        CK(ctx, cudaMemset(ctx->d_enc_qt, 0, 128 * sizeof(int32_t)));
    }
    if (!ctx->d_enc_total) {
        CK(ctx, cudaMalloc(&ctx->d_enc_total, 2 * sizeof(uint64_t)));
        CK(ctx, cudaMallocHost(&ctx->h_enc_total, 2 * sizeof(uint64_t)));
    }
    if (ctx->enc_w != w || ctx->enc_h != h || ctx->enc_fmt != ctx->format) {
        ctx->enc_w = w; ctx->enc_h = h; ctx->enc_fmt = ctx->format;
        ctx->enc_clear_old = true;
    }
    int rc = grow_device(ctx, &ctx->d_enc_old, &ctx->enc_old_cap, nblk * 64);
    if (rc) return rc;
    if (ctx->enc_clear_old) {
        CK(ctx, cudaMemsetAsync(ctx->d_enc_old, 0, nblk * 64 * sizeof(int16_t), st));
        ctx->enc_clear_old = false;
    }
    ctx->enc_stream = cuda_stream;
    if (F == 0) {
        CK(ctx, cudaMemsetAsync(d_offsets, 0, sizeof(uint64_t), st));
        CK(ctx, cudaMemsetAsync(ctx->d_enc_total, 0, 2 * sizeof(uint64_t), st));
        return RTJGPU_OK;
    }
    const size_t nb = nblk * (size_t)F;
    if (nb > ctx->enc_blocks_cap) {
        if (ctx->d_enc_slots) cudaFree(ctx->d_enc_slots);
        if (ctx->d_enc_lens) cudaFree(ctx->d_enc_lens);
        if (ctx->d_enc_boff) cudaFree(ctx->d_enc_boff);
        ctx->d_enc_slots = nullptr; ctx->d_enc_lens = nullptr; ctx->d_enc_boff = nullptr; ctx->enc_blocks_cap = 0;
        CK(ctx, cudaMalloc(&ctx->d_enc_slots, nb * 64));
        CK(ctx, cudaMalloc(&ctx->d_enc_lens, nb));
        CK(ctx, cudaMalloc(&ctx->d_enc_boff, nb * sizeof(uint32_t)));
        ctx->enc_blocks_cap = nb;
    }
    if ((rc = grow_device(ctx, &ctx->d_enc_fsize, &ctx->enc_frames_cap, (size_t)F))) return rc;
    rtj_encode_args a;
    a.d_frames = d_frames; a.F = F; a.w = w; a.h = h; a.fmt = ctx->format;
    a.d_qt = ctx->d_enc_qt; a.lb8 = ctx->enc_lb8; a.cb8 = ctx->enc_cb8; a.quality = ctx->enc_quality;
    a.key_rate = ctx->enc_key_rate; a.key_count0 = ctx->enc_key_count; a.lmask = ctx->enc_lm; a.cmask = ctx->enc_cm;
    a.d_old = ctx->d_enc_old; a.d_slots = ctx->d_enc_slots; a.d_lens = ctx->d_enc_lens; a.d_boff = ctx->d_enc_boff;
    a.d_fsize = ctx->d_enc_fsize; a.d_stream = d_stream; a.capacity = capacity; a.d_offsets = d_offsets;
    a.d_total = ctx->d_enc_total;
    const int e = rtj_launch_encode(&a, cuda_stream);
    if (e < 0) { ctx->last_cuda = -e; return RTJGPU_E_CUDA; }
    ctx->launches += (uint64_t)e;
    if (ctx->enc_key_rate) ctx->enc_key_count = (ctx->enc_key_count + F) % (ctx->enc_key_rate + 1);
    return RTJGPU_OK;
}

int rtjgpu_get_encode_info(rtjgpu_ctx *ctx, uint64_t *bytes, int *overflow)
{
    if (!ctx || !ctx->d_enc_total) return RTJGPU_E_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)ctx->enc_stream;
    CK(ctx, cudaMemcpyAsync(ctx->h_enc_total, ctx->d_enc_total, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    if (bytes) *bytes = ctx->h_enc_total[0];
    if (overflow) *overflow = ctx->h_enc_total[1] != 0;
    return RTJGPU_OK;
}

int rtjgpu_convert_bpp(int kind) { return rtj_convert_bpp(kind); }

int rtjgpu_convert_device(rtjgpu_ctx *ctx, int kind, const uint8_t *d_frames, size_t src_frame_bytes,
                          int F, int w, int h, uint8_t *d_out, size_t row_pitch, size_t frame_pitch,
                          int alpha, void *cuda_stream)
{
    if (!ctx || F < 0) return RTJGPU_E_ARG;
    const int bpp = rtj_convert_bpp(kind);
    if (!bpp) return RTJGPU_E_FORMAT;
    if (F == 0) return RTJGPU_OK;
    if (!d_frames || !d_out) return RTJGPU_E_ARG;
    if (w <= 0 || h <= 0 || (w & 15) || (h & 15) || w > 65535 || h > 65535) return RTJGPU_E_SIZE;
    if (F > 65535) return RTJGPU_E_TOOBIG;                                   /* frames are the grid's second dimension */
    if ((row_pitch & 15) || row_pitch < (size_t)w * bpp || frame_pitch < row_pitch * (size_t)h
        || ((uintptr_t)d_out & 15) || ((uintptr_t)d_frames & 7) || (src_frame_bytes & 7) || (frame_pitch & 15))
        return RTJGPU_E_ARG;
    const size_t ysz = (size_t)w * h;
    const size_t need = kind == RTJ_CONV_RGB8 ? ysz : kind == RTJ_CONV_YUV422_RGB24 ? 2 * ysz : ysz + ysz / 2;
    if (src_frame_bytes < need) return RTJGPU_E_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    const int e = rtj_launch_convert(kind, d_frames, src_frame_bytes, F, w, h, d_out, row_pitch, frame_pitch,
                                     (unsigned)alpha, cuda_stream);
    if (e) { ctx->last_cuda = e; return RTJGPU_E_CUDA; }
    ctx->launches += 1;
    return RTJGPU_OK;
}

int rtjgpu_set_custom_tables(rtjgpu_ctx *ctx, const uint32_t raw[128])
{
    if (!ctx || !raw) return RTJGPU_E_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    rtj_table_from_raw(raw, &ctx->h_tables[RTJGPU_TABLE_CUSTOM]);
    rtj_dev_table dev;
    rtj_table_to_device_layout(&ctx->h_tables[RTJGPU_TABLE_CUSTOM], &dev);
    /* ordered after the batches this context has queued (its own streams and the caller's last one); other contexts'
     * work on the device is none of its business */
    if (cudaStreamSynchronize((cudaStream_t)ctx->last_stream) != cudaSuccess) {    /* the caller's stream may be gone by now */
        cudaGetLastError();
        CK(ctx, cudaDeviceSynchronize());
    }
    if (ctx->pipe.ready) { CK(ctx, cudaStreamSynchronize(ctx->pipe.scan)); CK(ctx, cudaStreamSynchronize(ctx->pipe.idct)); }
    if (ctx->slots_ready) for (int i = 0; i < HOST_SLOTS; i++) CK(ctx, cudaStreamSynchronize(ctx->slot[i].stream));
    CK(ctx, cudaMemcpy(ctx->d_tables + RTJGPU_TABLE_CUSTOM, &dev, sizeof(dev), cudaMemcpyHostToDevice));
    return RTJGPU_OK;
}

/* internal: host tables in the reference's raster order (for RTjpeg_get_tables) */
const rtj_host_table *rtj_ctx_host_table(const rtjgpu_ctx *ctx, int table)
{
    if (!ctx || table < 0 || table >= RTJ_NUM_TABLES) return nullptr;
    return &ctx->h_tables[table];
}

int rtjgpu_plan(const uint8_t *stream, const uint64_t *offsets, int F, rtjgpu_state *state, rtjgpu_frame_desc *desc)
{
    return rtjgpu_plan_n(stream, offsets, nullptr, F, state, desc);
}

int rtjgpu_plan_n(const uint8_t *stream, const uint64_t *offsets, const uint32_t *lengths, int F, rtjgpu_state *state,
                  rtjgpu_frame_desc *desc)
{
    if (!stream || !offsets || !state || (F > 0 && !desc) || F < 0) return RTJGPU_E_ARG;
    if (F > RTJGPU_MAX_FRAMES_PER_BATCH) return RTJGPU_E_TOOBIG;
    rtjgpu_state st = *state;
    int bw = 0, bh = 0;
    for (int f = 0; f < F; f++) {
        if (offsets[f + 1] < offsets[f]) return RTJGPU_E_ARG;
        uint64_t avail = offsets[f + 1] - offsets[f];
        if (lengths) {
            if (lengths[f] > avail) return RTJGPU_E_ARG;
            avail = lengths[f];
        }
        if (avail < RTJPEG_B200_HEADER_BYTES) return RTJGPU_E_HEADER;
        if (offsets[f] & 3u) return RTJGPU_E_ARG;
        const uint8_t *p = stream + offsets[f];
        /* packed little-endian header, include/RTjpeg.h:100-109 */
        const uint32_t framesize = rd_u32le(p);
        const int w = rd_u16le(p + 6), h = rd_u16le(p + 8), q = p[10];
        /* the reference never reads framesize/headersize (lib/RTjpeg.c:3565-3586).  With the packets' true lengths
         * given, neither does this; without them the slot [offsets[f], offsets[f + 1]) may end in alignment padding,
         * and the smaller of framesize and the slot bounds every read */
        uint64_t len = avail;
        if (!lengths && framesize >= RTJPEG_B200_HEADER_BYTES && framesize < len) len = framesize;
        if (len - RTJPEG_B200_HEADER_BYTES > RTJGPU_MAX_PAYLOAD_BYTES) return RTJGPU_E_TOOBIG;
        if (w != st.width || h != st.height) { st.width = w; st.height = h; }   /* :3568-3574 */
        if (w == 0 || h == 0 || (w & 15) || (h & 15)) return RTJGPU_E_SIZE;     /* the row loop of :2701 needs /16 */
        if (f == 0) { bw = w; bh = h; }
        else if (w != bw || h != bh) return RTJGPU_E_SIZE;
        if (q != st.quality) {                                                   /* :3575-3579 */
            const int qc = q < 1 ? 1 : q;                                        /* set_quality clamps, :2410-2412 */
            st.quality = qc;
            st.table = qc;
        }
        desc[f].offset = offsets[f];
        desc[f].length = (uint32_t)len;
        desc[f].table = (uint16_t)st.table;
        desc[f].flags = 0;
    }
    *state = st;
    return RTJGPU_OK;
}

int rtjgpu_decode_device(rtjgpu_ctx *ctx, const uint8_t *d_stream, const rtjgpu_frame_desc *d_desc, int F,
                         int w, int h, uint8_t *d_out, const uint8_t *d_carry, void *cuda_stream)
{
    if (!ctx || F < 0) return RTJGPU_E_ARG;
    if (F == 0) { ctx->last_F = 0; return RTJGPU_OK; }
    if (!d_stream || !d_desc || !d_out) return RTJGPU_E_ARG;
    /* K1 reads the stream as 32-bit words, K2 leaves its strips as 16-byte bulk stores and reads the carry as 8-byte rows */
    if (((uintptr_t)d_stream & 3) || ((uintptr_t)d_desc & 7) || ((uintptr_t)d_out & 15) || ((uintptr_t)d_carry & 7)) return RTJGPU_E_ARG;
    if (w <= 0 || h <= 0 || (w & 15) || (h & 15) || w > 65535 || h > 65535) return RTJGPU_E_SIZE;
    if (F > RTJGPU_MAX_FRAMES_PER_BATCH) return RTJGPU_E_TOOBIG;
    CK(ctx, cudaSetDevice(ctx->device));
    const int nblk = RTJ_FMT_NBLK(ctx->format, w, h);
    if ((uint64_t)F * (uint64_t)nblk >= (1ull << 32)) return RTJGPU_E_TOOBIG;   /* block indices are 32 bit */
    if ((uint64_t)RTJ_FMT_FRAME_BYTES(ctx->format, w, h) >= (1ull << 32)) return RTJGPU_E_TOOBIG;   /* and so are offsets inside a frame */
    int rc = ws_reserve(ctx, &ctx->ws, F, nblk);
    if (rc) return rc;
    cudaEvent_t *ev = ctx->timing ? ctx->ev[ctx->timed_calls % TIMING_RING] : nullptr;
    rc = run_kernels(ctx, &ctx->ws, d_stream, d_desc, F, w, h, d_out, d_carry, (cudaStream_t)cuda_stream, ev, true);
    if (ev && rc == RTJGPU_OK) ctx->timed_calls++;
    ctx->last_F = F;
    ctx->last_stream = cuda_stream;
    return rc;
}

int rtjgpu_decode_device_rgb(rtjgpu_ctx *ctx, const uint8_t *d_stream, const rtjgpu_frame_desc *d_desc, int F,
                             int w, int h, int kind, uint8_t *d_rgb, size_t row_pitch, size_t frame_pitch, int alpha,
                             const uint8_t *d_carry, uint8_t *d_last_yuv, void *cuda_stream)
{
    if (!ctx || F < 0) return RTJGPU_E_ARG;
    if (ctx->format != RTJ_YUV420) return RTJGPU_E_FORMAT;                     /* the reference's converters are yuv420 ones */
    if (kind != RTJ_CONV_RGB32 && kind != RTJ_CONV_BGR32 && kind != RTJ_CONV_RGB24 && kind != RTJ_CONV_BGR24 && kind != RTJ_CONV_RGB16)
        return RTJGPU_E_FORMAT;
    if (F == 0) { ctx->last_F = 0; return RTJGPU_OK; }
    if (!d_stream || !d_desc || !d_rgb) return RTJGPU_E_ARG;
    if (w <= 0 || h <= 0 || (w & 15) || (h & 15) || w > 65535 || h > 65535) return RTJGPU_E_SIZE;
    if (F > RTJGPU_MAX_FRAMES_PER_BATCH) return RTJGPU_E_TOOBIG;
    const int bpp = rtj_convert_bpp(kind);
    if (((uintptr_t)d_stream & 3) || ((uintptr_t)d_desc & 7) || ((uintptr_t)d_rgb & 15) || ((uintptr_t)d_carry & 7)
        || ((uintptr_t)d_last_yuv & 15) || (row_pitch & 15) || (frame_pitch & 15) || row_pitch < (size_t)w * bpp
        || frame_pitch < row_pitch * (size_t)h)
        return RTJGPU_E_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    const int nblk = RTJ_FMT_NBLK(ctx->format, w, h);
    if ((uint64_t)F * (uint64_t)nblk >= (1ull << 32)) return RTJGPU_E_TOOBIG;
    int rc = ws_reserve(ctx, &ctx->ws, F, nblk);
    if (rc) return rc;
    RgbOut o;
    o.d_rgb = d_rgb; o.row_pitch = row_pitch; o.frame_pitch = frame_pitch; o.kind = kind; o.alpha = (unsigned)alpha & 0xFFu;
    o.d_last_yuv = d_last_yuv;
    cudaEvent_t *ev = ctx->timing ? ctx->ev[ctx->timed_calls % TIMING_RING] : nullptr;
    rc = run_kernels(ctx, &ctx->ws, d_stream, d_desc, F, w, h, nullptr, d_carry, (cudaStream_t)cuda_stream, ev, true, &o);
    if (ev && rc == RTJGPU_OK) ctx->timed_calls++;
    ctx->last_F = F;
    ctx->last_stream = cuda_stream;
    return rc;
}

int rtjgpu_sync(rtjgpu_ctx *ctx)
{
    if (!ctx) return RTJGPU_E_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaDeviceSynchronize());
    return RTJGPU_OK;
}

int rtjgpu_get_timing_at(rtjgpu_ctx *ctx, int calls_ago, rtjgpu_timing *out)
{
    if (!ctx || !out || calls_ago < 0 || calls_ago >= TIMING_RING) return RTJGPU_E_ARG;
    memset(out, 0, sizeof(*out));
    if ((uint64_t)calls_ago >= ctx->timed_calls) return RTJGPU_E_ARG;
    cudaEvent_t *ev = ctx->ev[(ctx->timed_calls - 1 - (uint64_t)calls_ago) % TIMING_RING];
    CK(ctx, cudaEventSynchronize(ev[3]));
    CK(ctx, cudaEventElapsedTime(&out->scan_ms, ev[0], ev[1]));
    CK(ctx, cudaEventElapsedTime(&out->resolve_ms, ev[1], ev[2]));
    CK(ctx, cudaEventElapsedTime(&out->idct_ms, ev[2], ev[3]));
    CK(ctx, cudaEventElapsedTime(&out->total_ms, ev[0], ev[3]));
    return RTJGPU_OK;
}

int rtjgpu_get_timing(rtjgpu_ctx *ctx, rtjgpu_timing *out)
{
    if (!ctx || !out) return RTJGPU_E_ARG;
    memset(out, 0, sizeof(*out));
    if (!ctx->timed_calls) return RTJGPU_OK;
    return rtjgpu_get_timing_at(ctx, 0, out);
}

int rtjgpu_get_batch_info(rtjgpu_ctx *ctx, rtjgpu_batch_info *out)
{
    if (!ctx || !out) return RTJGPU_E_ARG;
    memset(out, 0, sizeof(*out));
    out->first_bad_frame = -1;
    if (ctx->last_F == 0 || !ctx->ws.d_info) return RTJGPU_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaDeviceSynchronize());
    CK(ctx, cudaMemcpy(ctx->h_info, ctx->ws.d_info, sizeof(rtj_dev_info), cudaMemcpyDeviceToHost));
    out->skipped_blocks = ctx->h_info->skipped_blocks;
    out->payload_bytes = ctx->h_info->payload_bytes;
    out->bad_frames = ctx->h_info->bad_frames;
    out->first_bad_frame = ctx->h_info->first_bad_frame;
    return RTJGPU_OK;
}

int rtjgpu_get_skip_counts(rtjgpu_ctx *ctx, uint32_t *counts, int F)
{
    if (!ctx || !counts || F < 0 || F > ctx->last_F) return RTJGPU_E_ARG;
    if (F == 0) return RTJGPU_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaDeviceSynchronize());
    CK(ctx, cudaMemcpy(counts, ctx->ws.d_frame_skips, (size_t)F * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return RTJGPU_OK;
}

int rtjgpu_get_entries(rtjgpu_ctx *ctx, uint32_t *entries, size_t n)
{
    if (!ctx || !entries || n > ctx->ws.cap_entries) return RTJGPU_E_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaDeviceSynchronize());
    CK(ctx, cudaMemcpy(entries, ctx->ws.d_ent, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return RTJGPU_OK;
}

/* nearest clean frame to `ideal` inside [lo, hi], the later one on a tie; -1: none */
static int nearest_clean(const uint8_t *clean, int ideal, int lo, int hi)
{
    for (int d = 0; ideal - d >= lo || ideal + d <= hi; d++) {
        if (ideal + d >= lo && ideal + d <= hi && clean[ideal + d]) return ideal + d;
        if (ideal - d >= lo && ideal - d <= hi && clean[ideal - d]) return ideal - d;
    }
    return -1;
}

int rtjgpu_split_shards(const uint8_t *clean, int F, int n, int *first)
{
    if (!clean || !first || n < 1 || F < 0) return RTJGPU_E_ARG;
    first[0] = 0;
    int empty = 0;
    for (int i = 1; i < n; i++) {
        /* the clean frame nearest to the ideal cut, behind the previous cut and leaving a frame for every later shard
         * where the clean frames allow it; without any, the shard comes out empty -- and is counted */
        const int ideal = (int)((long long)F * i / n);
        const int lo = first[i - 1] + 1, hi = F - 1;
        int cut = lo <= hi ? nearest_clean(clean, std::min(std::max(ideal, lo), hi), lo, hi) : -1;
        if (cut < 0) cut = F;
        first[i] = cut;
    }
    first[n] = F;
    for (int i = 0; i < n; i++) empty += first[i + 1] == first[i];
    return F == 0 ? 0 : empty;
}

int rtjgpu_split_shards_lead(const uint8_t *clean, int F, int n, int *first, int *lead)
{
    if (!clean || !first || !lead || n < 1 || F < 0) return RTJGPU_E_ARG;
    first[0] = 0;
    lead[0] = 0;                                            /* frame 0 starts from the caller's picture */
    const int win = std::max(1, F / (2 * n));              /* how far a cut may move to find a clean frame: half a shard */
    for (int i = 1; i < n; i++) {
        const int ideal = (int)((long long)F * i / n);
        const int lo = std::min(first[i - 1] + 1, F);
        const int hi = std::max(lo, F - (n - i));                          /* a frame for every shard, while F >= n */
        const int at = std::min(std::max(ideal, lo), hi);
        if (at >= F) { first[i] = F; lead[i] = 0; continue; }              /* fewer frames than shards: an empty one */
        const int c = nearest_clean(clean, at, std::max(lo, at - win), std::min(hi, at + win));
        if (c >= 0) { first[i] = c; lead[i] = 0; continue; }
        /* no clean frame near: cut at the ideal place and decode again from the last clean frame before it --
         * or from frame 0 and the caller's picture when there is none */
        int back = at;
        while (back > 0 && !clean[back]) back--;
        first[i] = at;
        lead[i] = at - back;
    }
    first[n] = F;
    return RTJGPU_OK;
}

int rtjgpu_scan_device(rtjgpu_ctx *ctx, const uint8_t *d_stream, const rtjgpu_frame_desc *d_desc, int F, int w, int h,
                       void *cuda_stream)
{
    if (!ctx || F < 0) return RTJGPU_E_ARG;
    if (F == 0) { ctx->last_F = 0; return RTJGPU_OK; }
    if (!d_stream || !d_desc || ((uintptr_t)d_stream & 3)) return RTJGPU_E_ARG;
    if (w <= 0 || h <= 0 || (w & 15) || (h & 15) || w > 65535 || h > 65535) return RTJGPU_E_SIZE;
    if (F > RTJGPU_MAX_FRAMES_PER_BATCH) return RTJGPU_E_TOOBIG;
    CK(ctx, cudaSetDevice(ctx->device));
    const int nblk = RTJ_FMT_NBLK(ctx->format, w, h);
    if ((uint64_t)F * (uint64_t)nblk >= (1ull << 32)) return RTJGPU_E_TOOBIG;
    int rc = ws_reserve(ctx, &ctx->ws, F, nblk);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    rtj_launch_args a;
    memset(&a, 0, sizeof(a));
    a.d_stream = d_stream; a.d_desc = d_desc; a.d_tables = ctx->d_tables;
    a.F = F; a.w = w; a.h = h; a.f0 = 0; a.f1 = F; a.slice = 0;
    a.fmt = ctx->format;
    a.d_ent = ctx->ws.d_ent; a.d_frame_skips = ctx->ws.d_frame_skips; a.d_info = ctx->ws.d_info;
    a.d_walk = ctx->ws.d_walk;
    a.d_redo = ctx->ws.d_frame_skips + ctx->ws.cap_frames;
    a.raw_expected = ((volatile unsigned long long *)ctx->h_skips_seen)[1] != 0;
    a.row1 = RTJ_FMT_UNITS_Y(ctx->format, h);
    a.scan_mode = ctx->scan_mode;
    if ((rc = seg_reserve(ctx, &ctx->ws, F, nblk, ctx->scan_mode, &a.seg))) return rc;
    CK(ctx, cudaMemcpyAsync(ctx->ws.d_info, ctx->h_info_reset, sizeof(rtj_dev_info), cudaMemcpyHostToDevice, st));
    const int e = rtj_launch_scan(&a, st);
    if (e < 0) { ctx->last_cuda = -e; return RTJGPU_E_CUDA; }
    ctx->launches += (uint64_t)e;
    ctx->last_F = F;
    ctx->last_stream = cuda_stream;
    return RTJGPU_OK;
}

static void export_table(const rtj_host_table &t, uint32_t scaled[128], int *lb8, int *cb8)
{
    for (int i = 0; i < 64; i++) {
        scaled[i] = (uint32_t)t.liqt[i];
        scaled[64 + i] = (uint32_t)t.ciqt[i];
    }
    if (lb8) *lb8 = t.lb8;
    if (cb8) *cb8 = t.cb8;
}

void rtjgpu_tables_for_quality(int Q, uint32_t scaled[128], int *lb8, int *cb8)
{
    rtj_host_table t;
    rtj_table_from_quality(Q, &t);
    export_table(t, scaled, lb8, cb8);
}

void rtjgpu_tables_from_raw(const uint32_t raw[128], uint32_t scaled[128], int *lb8, int *cb8)
{
    rtj_host_table t;
    rtj_table_from_raw(raw, &t);
    export_table(t, scaled, lb8, cb8);
}

void *rtjgpu_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}

void rtjgpu_host_free(void *p) { if (p) cudaFreeHost(p); }

/* ------------------------------------------------------------------------ */
/* host pipeline                                                              */
/* ------------------------------------------------------------------------ */

static int slots_init(rtjgpu_ctx *ctx)
{
    if (ctx->slots_ready) return RTJGPU_OK;
    for (int i = 0; i < HOST_SLOTS; i++) {
        CK(ctx, cudaStreamCreateWithFlags(&ctx->slot[i].stream, cudaStreamNonBlocking));
        CK(ctx, cudaEventCreateWithFlags(&ctx->slot[i].decoded, cudaEventDisableTiming));
        CK(ctx, cudaMallocHost(&ctx->slot[i].h_info, sizeof(rtj_dev_info)));
    }
    ctx->slots_ready = true;
    return RTJGPU_OK;
}

static int slot_drain(rtjgpu_ctx *ctx, HostSlot &s)
{
    if (!s.busy) return RTJGPU_OK;
    const cudaError_t e = cudaStreamSynchronize(s.stream);
    if (e == cudaSuccess) {
        ctx->host_bad += s.h_info->bad_frames;
        if (s.pending_dst) memcpy(s.pending_dst, s.h_out, s.pending_bytes);
    }
    s.pending_dst = nullptr;              /* whatever happened: never again touch the memory of the call that queued this */
    s.pending_bytes = 0;
    s.busy = false;
    if (e != cudaSuccess) { ctx->last_cuda = (int)e; return RTJGPU_E_CUDA; }
    return RTJGPU_OK;
}

int rtjgpu_decode_host(rtjgpu_ctx *ctx, const uint8_t *h_stream, const uint64_t *offsets, int F,
                       rtjgpu_state *state, uint8_t *h_out, uint8_t *h_carry_inout, int flags)
{
    return rtjgpu_decode_host_n(ctx, h_stream, offsets, nullptr, F, state, h_out, h_carry_inout, flags);
}

int rtjgpu_decode_host_n(rtjgpu_ctx *ctx, const uint8_t *h_stream, const uint64_t *offsets, const uint32_t *lengths, int F,
                         rtjgpu_state *state, uint8_t *h_out, uint8_t *h_carry_inout, int flags)
{
    if (!ctx || !state || F < 0) return RTJGPU_E_ARG;
    if (F == 0) return RTJGPU_OK;
    if (!h_stream || !offsets || !h_out) return RTJGPU_E_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    int rc = slots_init(ctx);
    if (rc) return rc;

    /* geometry from the first header; rtjgpu_plan re-checks every frame */
    if ((lengths ? (uint64_t)lengths[0] : offsets[1] - offsets[0]) < RTJPEG_B200_HEADER_BYTES) return RTJGPU_E_HEADER;
    const int w = rd_u16le(h_stream + offsets[0] + 6), h = rd_u16le(h_stream + offsets[0] + 8);
    if (w == 0 || h == 0 || (w & 15) || (h & 15)) return RTJGPU_E_SIZE;
    const size_t fsz = RTJ_FMT_FRAME_BYTES(ctx->format, w, h);
    const int nblk = RTJ_FMT_NBLK(ctx->format, w, h);

    /* chunk size: ~64 MB of output per chunk keeps three slots in flight without
     * hoarding memory; at least 1, at most the batch limit */
    int chunk = (int)std::max<size_t>(1, (64u << 20) / fsz);
    chunk = std::min(chunk, RTJGPU_MAX_FRAMES_PER_BATCH);
    chunk = std::min(chunk, F);

    rc = grow_device(ctx, &ctx->d_host_carry, &ctx->d_host_carry_cap, fsz);
    if (rc) return rc;
    if ((rc = grow_pinned(ctx, &ctx->h_host_skips, &ctx->host_skips_cap, (size_t)F))) return rc;
    ctx->host_F = 0;
    const bool have_carry = h_carry_inout != nullptr;
    if (have_carry)
        CK(ctx, cudaMemcpy(ctx->d_host_carry, h_carry_inout, fsz, cudaMemcpyHostToDevice));

    const uint8_t *d_prev = have_carry ? ctx->d_host_carry : nullptr;
    cudaEvent_t prev_decoded = nullptr;
    rtjgpu_state st = *state;
    int result = RTJGPU_OK;
    ctx->host_bad = 0;
    std::vector<uint64_t> rel;

    /* inside the loop a failing CUDA call must not return: the slots already in flight hold pointers into this call's
     * h_out and have to be drained first */
#define CKB(call)                                                                          \
    { const cudaError_t e__ = (call); if (e__ != cudaSuccess) { ctx->last_cuda = (int)e__; result = RTJGPU_E_CUDA; break; } }
    for (int c0 = 0, ci = 0; c0 < F; c0 += chunk, ci++) {
        const int n = std::min(chunk, F - c0);
        HostSlot &s = ctx->slot[ci % HOST_SLOTS];
        if ((rc = slot_drain(ctx, s))) { result = rc; break; }

        const uint64_t b0 = offsets[c0], b1 = offsets[c0 + n];
        const size_t in_bytes = (size_t)(b1 - b0);
        if ((rc = grow_device(ctx, &s.d_in, &s.d_in_cap, in_bytes + RTJGPU_STREAM_SLACK_BYTES))) { result = rc; break; }
        if ((rc = grow_device(ctx, &s.d_out, &s.d_out_cap, fsz * (size_t)n))) { result = rc; break; }
        size_t dcap = (size_t)s.desc_cap;
        if ((rc = grow_device(ctx, &s.d_desc, &dcap, (size_t)n))) { result = rc; break; }
        s.desc_cap = (int)dcap;
        size_t hcap = (size_t)s.h_desc_cap;
        if ((rc = grow_pinned(ctx, &s.h_desc, &hcap, (size_t)n))) { result = rc; break; }
        s.h_desc_cap = (int)hcap;
        if ((rc = ws_reserve(ctx, &s.ws, n, nblk))) { result = rc; break; }

        /* descriptors relative to the chunk's own device buffer */
        rel.resize((size_t)n + 1);
        for (int i = 0; i <= n; i++) rel[(size_t)i] = offsets[c0 + i] - b0;
        if ((rc = rtjgpu_plan_n(h_stream + b0, rel.data(), lengths ? lengths + c0 : nullptr, n, &st, s.h_desc))) { result = rc; break; }
        if (st.width != w || st.height != h) { result = RTJGPU_E_SIZE; break; }

        const uint8_t *src = h_stream + b0;
        if (!(flags & RTJGPU_HOST_IN_PINNED)) {
            if ((rc = grow_pinned(ctx, &s.h_in, &s.h_in_cap, in_bytes))) { result = rc; break; }
            memcpy(s.h_in, src, in_bytes);
            src = s.h_in;
        }
        CKB(cudaMemcpyAsync(s.d_in, src, in_bytes, cudaMemcpyHostToDevice, s.stream));
        CKB(cudaMemsetAsync(s.d_in + in_bytes, 0x7F, RTJGPU_STREAM_SLACK_BYTES, s.stream));
        CKB(cudaMemcpyAsync(s.d_desc, s.h_desc, sizeof(rtjgpu_frame_desc) * (size_t)n, cudaMemcpyHostToDevice, s.stream));
        if (prev_decoded) CKB(cudaStreamWaitEvent(s.stream, prev_decoded, 0));   /* carry comes from the previous chunk */
        if ((rc = run_kernels(ctx, &s.ws, s.d_in, s.d_desc, n, w, h, s.d_out, d_prev, s.stream, nullptr, false))) { result = rc; break; }
        CKB(cudaEventRecord(s.decoded, s.stream));
        CKB(cudaMemcpyAsync(s.h_info, s.ws.d_info, sizeof(rtj_dev_info), cudaMemcpyDeviceToHost, s.stream));
        CKB(cudaMemcpyAsync(ctx->h_host_skips + c0, s.ws.d_frame_skips, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, s.stream));
        prev_decoded = s.decoded;
        d_prev = s.d_out + fsz * (size_t)(n - 1);

        uint8_t *dst = h_out + fsz * (size_t)c0;
        if (flags & RTJGPU_HOST_OUT_PINNED) {
            CKB(cudaMemcpyAsync(dst, s.d_out, fsz * (size_t)n, cudaMemcpyDeviceToHost, s.stream));
            s.pending_dst = nullptr;
        } else {
            if ((rc = grow_pinned(ctx, &s.h_out, &s.h_out_cap, fsz * (size_t)n))) { result = rc; break; }
            CKB(cudaMemcpyAsync(s.h_out, s.d_out, fsz * (size_t)n, cudaMemcpyDeviceToHost, s.stream));
            s.pending_dst = dst;
            s.pending_bytes = fsz * (size_t)n;
        }
        s.busy = true;
    }
#undef CKB
    /* drain in issue order so that later chunks' carries are complete */
    for (int i = 0; i < HOST_SLOTS; i++) {
        int rc2 = slot_drain(ctx, ctx->slot[i]);
        if (rc2 && !result) result = rc2;
    }
    if (result == RTJGPU_OK && ctx->host_bad) result = RTJGPU_E_OVERRUN;
    if (result == RTJGPU_OK || result == RTJGPU_E_OVERRUN) {
        if (have_carry) memcpy(h_carry_inout, h_out + fsz * (size_t)(F - 1), fsz);
        *state = st;
        ctx->host_F = F;
    }
    return result;
}

int rtjgpu_get_host_skip_counts(rtjgpu_ctx *ctx, uint32_t *counts, int F)
{
    if (!ctx || !counts || F < 0 || F > ctx->host_F) return RTJGPU_E_ARG;
    if (F) memcpy(counts, ctx->h_host_skips, sizeof(uint32_t) * (size_t)F);
    return RTJGPU_OK;
}

} // extern "C"
