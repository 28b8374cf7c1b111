// mog_pipe.cuh -- the steady-state form of the fused kernel (mog_fused.cuh has the semantics and
// the generic fallback): the same per-pixel MOG2 + setTo(0) + BGR->HSV + inRange arithmetic, fed by
// an asynchronous bulk-copy pipeline instead of per-thread global loads, and RESIDENT across frames:
// one launch works through a queue of frame descriptors (a whole clip, frames of several independent
// streams interleaved, or a single frame), so CTA start-up, pipeline fill and drain are paid once
// per launch instead of once per frame and the host does nothing per frame.
//
//   * persistent grid (CTAs/SM x SM count); work items are (frame, tile) pairs, tile = PIPE_TILE = 1024
//     consecutive padded pixels (4 per thread), drawn from ONE global counter in frame-major order;
//   * the hot GMM planes -- mode 0 {weight, variance, mean b/g/r} and the mode-count bytes, i.e.
//     everything a single-mode pixel needs -- and the tile's BGR bytes are staged global -> shared with
//     cp.async.bulk (TMA engine, SASS UBLKCP) into a PIPE_STAGES-deep ring tracked by mbarriers,
//     updated IN PLACE in shared memory and written back with cp.async.bulk shared -> global; no warp
//     ever waits on a global load of state in the common case, and the bytes in flight per SM are set
//     by the ring depth, not by registers/occupancy;
//   * frame t+1 of a model needs frame t's state only tile by tile: every finished tile is PUBLISHED
//     (release) in tile_seq[] and acquired by whoever loads that tile for the next frame -- inside one
//     launch and across consecutive launches alike, so neither a frame boundary nor a launch boundary
//     drains the machine;
//   * modes 1..4 (pixels whose model currently has more than one live mode, ~1-2 % on the
//     benchmark stream) are read/written with direct 128-bit global accesses in a slow path that
//     runs the literal mog2_pixel(); a thread whose 4 pixels all have one mode that fits takes a
//     straight-line fast path (identical arithmetic, no sort/insert/shadow/HSV code).
//
// Parity: the fast path evaluates exactly the expressions mog2_pixel() evaluates for n == 1 with a
// fitting sample (same operations, same order, no FMA contraction), so state and masks stay
// bit-identical to the oracle; any other case falls through to mog2_pixel() itself.
#pragma once
#include "mog_fused.cuh"

namespace oat {

#ifndef PIPE_CTHREADS_CFG
#define PIPE_CTHREADS_CFG 256
#endif
#ifndef PIPE_STAGES_CFG
#define PIPE_STAGES_CFG 4
#endif
// Launches that run next to a detect tail (the resident tail server, or per-frame tail kernels) leave this many CTA
// slots of the persistent grid unused: the SMs that then hold one fused CTA instead of two are where the tail's
// CTAs find room (two fused CTAs fill an SM's shared memory and registers).
#ifndef PIPE_RESERVED_CTAS_CFG
#define PIPE_RESERVED_CTAS_CFG 24
#endif
#ifndef PIPE_STORERS_CFG
#define PIPE_STORERS_CFG 2   // storer lanes per CTA (each in a warp of its own), taking the CTA's tiles in turn
#endif
#ifndef PIPE_TAIL_GRID_CFG
#define PIPE_TAIL_GRID_CFG 24   // CTAs of the resident tail server: one beside the single fused CTA of a reserved SM
#endif
#ifndef PIPE_TAIL_SMEM_KB_CFG
#define PIPE_TAIL_SMEM_KB_CFG 110  // ... with the rest of that SM's shared memory for its labelling tables
#endif
#ifndef PIPE_MINBLOCKS_CFG
#define PIPE_MINBLOCKS_CFG 2
#endif
constexpr int PIPE_CTHREADS = PIPE_CTHREADS_CFG;    // compute threads (8 warps), 4 pixels each
constexpr int PIPE_STORERS = PIPE_STORERS_CFG;
constexpr int PIPE_THREADS = PIPE_CTHREADS + 32 + 32 * PIPE_STORERS;  // + a loader warp (bulk loads) and the storer warps (write-back + publication)
constexpr int PIPE_TILE = PIPE_CTHREADS * 4;        // pixels per tile
constexpr int PIPE_STAGES = PIPE_STAGES_CFG;
static_assert(PIPE_STORERS >= 1 && PIPE_STORERS <= PIPE_STAGES_CFG, "every storer lane needs a stage to work on");
#ifndef PIPE_CTAS_PER_SM_CFG
#define PIPE_CTAS_PER_SM_CFG PIPE_MINBLOCKS_CFG
#endif
constexpr int PIPE_CTAS_PER_SM = PIPE_CTAS_PER_SM_CFG;  // persistent grid = this x SM count
constexpr int PIPE_OFF_NM = 5 * PIPE_TILE * 4;       // stage layout: 5 fp32 planes | mode counts | BGR | flag
constexpr int PIPE_OFF_BGR = PIPE_OFF_NM + PIPE_TILE;
constexpr int PIPE_OFF_FLAG = PIPE_OFF_BGR + 3 * PIPE_TILE;   // stage header, 128 B (u32 words, see HDR_*)
constexpr int PIPE_OFF_BITS = PIPE_OFF_FLAG + 128;            // threshold bits of the tile (PIPE_TILE / 8 bytes)
constexpr int PIPE_OFF_QUEUE = PIPE_OFF_BITS + PIPE_TILE / 8;  // slow-pixel queue (u16 pixel-in-tile, 0xffff = empty)
constexpr int PIPE_STAGE_BYTES = PIPE_OFF_QUEUE + 2 * PIPE_TILE;  // 26880 B
constexpr int PIPE_SMEM_BYTES = PIPE_STAGES * PIPE_STAGE_BYTES;
// stage header words: what the producer lane tells the compute warps (tile, state) and its own store side
// (everything else) about the tile a stage holds
enum {
    HDR_DIRTY = 0,    // TRACK: some state word of the tile changed
    HDR_TILE = 1,     // tile number inside its frame, -1 = no more work
    HDR_QCNT = 2,     // slow-pixel queue: entries pushed
    HDR_QHEAD = 3,    //                   entries claimed
    HDR_FRAME = 4,    // frame index inside the launch
    HDR_SEQ_OUT = 5,  // value to publish in tile_seq
    HDR_STATE = 6,    // (64-bit) GMM planes of the frame's model
    HDR_NMODES = 8,
    HDR_THR = 10,
    HDR_TSEQ = 12,
    HDR_SLOW = 14,
    HDR_DONE = 16,
};

// ---- PTX wrappers: mbarrier + bulk async copies (sm_90+; SASS UBLKCP / SYNCS) ----------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const unsigned int *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
// one bounded attempt: true once the phase has completed, false after roughly `ns` nanoseconds without it
__device__ __forceinline__ bool mbar_try_wait_ns(uint64_t *b, uint32_t parity, uint32_t ns)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// same, with an L2 eviction-priority hint (createpolicy): the input frame is read exactly once, so it
// must not displace the GMM planes, which the next frame of the stream re-reads from L2
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all()  // writes of all but the N latest groups are COMPLETE (not just read)
{
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// release side of the tile hand-off: everything this thread has observed (its own completed bulk stores, and
// -- through the CTA-scope mbarrier it waited on -- the compute warps' direct stores) becomes visible at GPU
// scope before the flag stores that follow
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// acquire side: the async proxy (bulk copies) must not read global memory ahead of the generic-proxy acquire
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void st_relaxed_gpu(unsigned int *p, unsigned int v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_relaxed_gpu(unsigned int *p, unsigned int v)
{
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One frame of work.  A launch reads its descriptors from device memory (a long queue) or, up to
// PIPE_INLINE_DESCS frames, from the kernel parameters themselves (no copy to enqueue, nothing to wait for).  All frames of a launch share geometry, pitches, GMM
// constants and the inRange band (StreamArgs::f); what differs per frame is below.
struct FrameDesc {
    const uint8_t *bgr;        // input frame
    float *state;              // GMM planes of the frame's model
    uint8_t *nmodes;
    uint32_t *thr_bits;        // threshold bits out (or NULL)
    unsigned int *tile_seq;    // per tile of the model: sequence number of the last frame that finished it
    unsigned int *slow_count;  // or NULL: += 4-pixel groups that left the fast path in this frame
    unsigned int *done_count;  // or NULL: += 1 per published tile (the tail server starts the frame at ntiles)
    unsigned int seq_expect;   // tile_seq value that means "the model's previous frame has finished this tile"
    unsigned int flags;        // FD_CHAIN
};
constexpr int PIPE_INLINE_DESCS = 32;
enum { FD_CHAIN = 1u };        // order every tile behind tile_seq (else: the launch is ordered behind the stream)
static_assert(sizeof(FrameDesc) == 64, "FrameDesc is 64 bytes");

struct StreamArgs {
    FusedArgs f;                   // geometry, pitches, constants, band, egress (bgr/state/nmodes/thr_bits: see FrameDesc)
    int ntiles;                    // tiles per frame
    int nframes;                   // frames in this launch
    const FrameDesc *descs;        // [nframes] in device memory, or NULL: `inl` (short queues travel with the launch)
    FrameDesc inl[PIPE_INLINE_DESCS];
    unsigned long long div_magic;  // ceil(2^64 / pitch_px): y = umul64hi(pidx, magic), exact for every pidx < 2^32
    int zero_in;                   // HSV (0,0,0) lies inside the inRange band
    int wait_grid;                 // 1: order the whole launch behind the previous grid on the stream (griddepcontrol.wait)
    // Work scheduler: every (frame, tile) item is drawn from this counter (0 when the launch starts; the last CTA
    // to leave re-arms it and the ticket, then tells the host).  The host hands the slots of a ring out to
    // consecutive launches and re-uses a slot only after its previous user has said so (api.cu), so no two live
    // launches ever share one.
    unsigned int *work_counter;
    unsigned int *exit_ticket;
    unsigned int *done_flag;       // pinned host word of the slot: the last CTA to leave stores launch_id there
    unsigned int launch_id;
    int relaxed_publish;           // measurement switch ONLY (OAT_B200_RELAXED_PUBLISH): never fence, to price the fences
    int fence_always;              // measurement switch ONLY (OAT_B200_FENCE_ALWAYS): one fence per tile instead of per batch
};

// One pixel with one or two live modes whose sample fits mode 0 (the heavier one): the m = 0
// iteration of mog2_pixel(), the weight decay / pruning of mode 1 (the only thing the m = 1 iteration
// does once a fit was found) and the normalisation -- the same operations in the same order.  Valid
// when the sample fits mode 0, is background and mode 0 is not pruned (returns false otherwise:
// nothing may be stored).  W1 is mode 1's weight (ignored when n == 1).
template <bool TRACK>
__device__ __forceinline__ bool fast_px(const float x0, const float x1, const float x2, float &W, float &V, float &A,
                                        float &B, float &C, float &W1, uint32_t &n, const MogConsts &c, bool &dirty)
{
    const float w0 = fadd(fmul(c.a1, W), c.prune);
    const float d0 = fsub(A, x0), d1 = fsub(B, x1), d2 = fsub(C, x2);
    const float dist2 = fadd(fadd(fmul(d0, d0), fmul(d1, d1)), fmul(d2, d2));
    const bool bg = (0.f < c.TB) && (dist2 < fmul(c.Tb, V));
    const bool fits = dist2 < fmul(c.Tg, V);
    const float w = fadd(w0, c.aT);
    const float k = fdiv(c.aT, w);
    const float nA = fsub(A, fmul(k, d0)), nB = fsub(B, fmul(k, d1)), nC = fsub(C, fmul(k, d2));
    float vn = fadd(V, fmul(k, fsub(dist2, V)));
    vn = (vn < c.varMin) ? c.varMin : vn;
    vn = (vn > c.varMax) ? c.varMax : vn;
    const bool pruned = w < -c.prune;
    float tot = fadd(0.f, w);
    float w1 = 0.f;
    uint32_t nn = n;
    if (n == 2u) {
        w1 = fadd(fmul(c.a1, W1), c.prune);
        if (w1 < -c.prune) {
            w1 = 0.f;
            nn = 1u;
        }
        tot = fadd(tot, w1);
    }
    float inv = 0.f;
    if (fabsf(tot) > FLT_EPSILON) inv = fdiv(1.f, tot);
    const float nW = fmul(w, inv);
    const float nW1 = (nn == 2u) ? fmul(w1, inv) : w1;
    const bool ok = bg && fits && !pruned;
    if (ok) {
        upd<TRACK>(W, nW, dirty);
        upd<TRACK>(V, vn, dirty);
        upd<TRACK>(A, nA, dirty);
        upd<TRACK>(B, nB, dirty);
        upd<TRACK>(C, nC, dirty);
        if (n == 2u) upd<TRACK>(W1, nW1, dirty);
        if (TRACK) dirty |= (nn != n);
        n = nn;
    }
    return ok;
}

// Frozen model (learning rate 0 -- Oat's default -a 0 -- where nothing is ever inserted), one live mode:
// the whole of mog2_pixel() for that case inline, classification included, so that foreground pixels do
// not leave the fast path: returns the mask value {0, shadow, 255}, or 0xffffffff if the mode would be
// pruned (left to the slow path).  Same operations in the same order as mog2_pixel().
__device__ __forceinline__ uint32_t frozen_px(const float x0, const float x1, const float x2, float &W, float &V, float &A,
                                              float &B, float &C, const MogConsts &c, bool &dirty)
{
    float w = fadd(fmul(c.a1, W), c.prune);
    const float d0 = fsub(A, x0), d1 = fsub(B, x1), d2 = fsub(C, x2);
    const float dist2 = fadd(fadd(fmul(d0, d0), fmul(d1, d1)), fmul(d2, d2));
    const bool bg = (0.f < c.TB) && (dist2 < fmul(c.Tb, V));
    float nA = A, nB = B, nC = C, vn = V;
    if (dist2 < fmul(c.Tg, V)) {
        w = fadd(w, c.aT);
        // (this function only runs with aT == 0: 0 / w is exactly +0 for every w > 0 -- no division needed)
        const float k = (c.aT == 0.f && w > 0.f) ? 0.f : fdiv(c.aT, w);
        nA = fsub(A, fmul(k, d0));
        nB = fsub(B, fmul(k, d1));
        nC = fsub(C, fmul(k, d2));
        vn = fadd(V, fmul(k, fsub(dist2, V)));
        vn = (vn < c.varMin) ? c.varMin : vn;
        vn = (vn > c.varMax) ? c.varMax : vn;
    }
    if (w < -c.prune) return 0xffffffffu;
    const float tot = fadd(0.f, w);
    float inv = 0.f;
    if (tot == 1.f) inv = 1.f;  // the frozen one-mode model's weight: 1 / 1 is exact
    else if (fabsf(tot) > FLT_EPSILON) inv = fdiv(1.f, tot);
    const float nW = fmul(w, inv);
    upd<true>(W, nW, dirty);
    upd<true>(V, vn, dirty);
    upd<true>(A, nA, dirty);
    upd<true>(B, nB, dirty);
    upd<true>(C, nC, dirty);
    if (bg) return 0u;
    if (c.detect_shadows) {  // detectShadowGMM over the single mode
        const float num = fadd(fadd(fadd(0.f, fmul(x0, nA)), fmul(x1, nB)), fmul(x2, nC));
        const float den = fadd(fadd(fadd(0.f, fmul(nA, nA)), fmul(nB, nB)), fmul(nC, nC));
        if (den != 0.f && num <= den && num >= fmul(c.tau, den)) {
            const float a = fdiv(num, den);
            const float e0 = fsub(fmul(a, nA), x0), e1 = fsub(fmul(a, nB), x1), e2 = fsub(fmul(a, nC), x2);
            const float dist2a = fadd(fadd(fadd(0.f, fmul(e0, e0)), fmul(e1, e1)), fmul(e2, e2));
            if (dist2a < fmul(fmul(fmul(c.Tb, vn), a), a)) return (uint32_t)c.shadow_value;
        }
    }
    return 255u;
}

__device__ __forceinline__ float sel4(const float4 &v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
__device__ __forceinline__ void set4(float4 &v, int i, float x)
{
    if (i == 0) v.x = x;
    if (i == 1) v.y = x;
    if (i == 2) v.z = x;
    if (i == 3) v.w = x;
}

// Slow path, ONE pixel (rare): more than two live modes, no fit on mode 0, foreground / shadow, or a pruned
// mode.  Pixels that leave the fast path are queued per tile and drained by whichever lanes of the CTA
// are free (a warp whose 128 pixels all sit on the blob would otherwise walk them 4 deep while the
// other seven warps -- and the stage refill behind them -- wait), so this takes any pixel `e` of the
// tile, not the caller's own.  Runs the rolled mog2_pixel_rolled(): cold code, small footprint.
// Mode 0, the counts and the BGR bytes are in the shared-memory stage `st`; modes >= 1 in global memory.
template <int K, bool TRACK>
__device__ __forceinline__ void slow_pixel(const StreamArgs &pa, float *state, uint8_t *st, const int e, const size_t pidx)
{
    const FusedArgs &a = pa.f;
    float *sm0 = reinterpret_cast<float *>(st) + e;
    uint8_t *smn = st + PIPE_OFF_NM + e;
    const uint8_t *smb = st + PIPE_OFF_BGR + 3 * e;
    int n = *smn;
    const int n_old = n;
    float W[K], V[K], A[K], B[K], C[K];
    W[0] = sm0[0];
    V[0] = sm0[PIPE_TILE];
    A[0] = sm0[2 * PIPE_TILE];
    B[0] = sm0[3 * PIPE_TILE];
    C[0] = sm0[4 * PIPE_TILE];
#pragma unroll 1
    for (int m = 1; m < n; ++m) {
        const float *g = state + (size_t)(m * 5) * a.plane + pidx;
        W[m] = ld_state_f1(g);
        V[m] = ld_state_f1(g + a.plane);
        A[m] = ld_state_f1(g + 2 * a.plane);
        B[m] = ld_state_f1(g + 3 * a.plane);
        C[m] = ld_state_f1(g + 4 * a.plane);
    }
    float oW[K], oV[K], oA[K], oB[K], oC[K];
    if (TRACK) {
#pragma unroll 1
        for (int m = 0; m < n; ++m) {
            oW[m] = W[m];
            oV[m] = V[m];
            oA[m] = A[m];
            oB[m] = B[m];
            oC[m] = C[m];
        }
    }
    const int b = smb[0], g_ = smb[1], r = smb[2];
    const uint32_t mk = mog2_pixel_rolled<K>((float)b, (float)g_, (float)r, n, W, V, A, B, C, a.c);
    const int nw = max(n, n_old);
    bool d = true;
    if (TRACK) {
        d = (n != n_old);
#pragma unroll 1
        for (int m = 0; m < n_old && m < n; ++m)
            d |= (__float_as_uint(oW[m]) != __float_as_uint(W[m])) | (__float_as_uint(oV[m]) != __float_as_uint(V[m])) |
                 (__float_as_uint(oA[m]) != __float_as_uint(A[m])) | (__float_as_uint(oB[m]) != __float_as_uint(B[m])) |
                 (__float_as_uint(oC[m]) != __float_as_uint(C[m]));
    }
    if (d) {
        if (TRACK) *reinterpret_cast<volatile uint32_t *>(st + PIPE_OFF_FLAG) = 1u;
        sm0[0] = W[0];
        sm0[PIPE_TILE] = V[0];
        sm0[2 * PIPE_TILE] = A[0];
        sm0[3 * PIPE_TILE] = B[0];
        sm0[4 * PIPE_TILE] = C[0];
        *smn = (uint8_t)n;
#pragma unroll 1
        for (int m = 1; m < nw; ++m) {
            float *g = state + (size_t)(m * 5) * a.plane + pidx;
            st_state_f1(g, W[m]);
            st_state_f1(g + a.plane, V[m]);
            st_state_f1(g + 2 * a.plane, A[m]);
            st_state_f1(g + 3 * a.plane, B[m]);
            st_state_f1(g + 4 * a.plane, C[m]);
        }
    }
    const int ob = mk ? b : 0, og = mk ? g_ : 0, orr = mk ? r : 0;
    int h = 0, sa = 0, v = 0;
    if (a.do_hsv) {
        if (mk) bgr2hsv_px_div(ob, og, orr, h, sa, v);
        const bool in = (a.lo[0] <= h) & (h <= a.hi[0]) & (a.lo[1] <= sa) & (sa <= a.hi[1]) & (a.lo[2] <= v) & (v <= a.hi[2]);
        if (in) atomicOr(reinterpret_cast<uint32_t *>(st + PIPE_OFF_BITS) + (e >> 5), 1u << (e & 31));
    }
    if (a.bgr_out || a.hsv_out || a.fg_out) {
        const int y = (int)__umul64hi((unsigned long long)pidx, pa.div_magic);
        const int x = (int)(pidx - (size_t)y * ((size_t)a.wpr * 32));
        if (a.bgr_out) {
            uint8_t *dst = a.bgr_out + (size_t)y * a.bgr_out_pitch + 3 * x;
            dst[0] = (uint8_t)ob;
            dst[1] = (uint8_t)og;
            dst[2] = (uint8_t)orr;
        }
        if (a.hsv_out) {
            uint8_t *dst = a.hsv_out + (size_t)y * a.hsv_pitch + 3 * x;
            dst[0] = (uint8_t)h;
            dst[1] = (uint8_t)sa;
            dst[2] = (uint8_t)v;
        }
        if (a.fg_out) a.fg_out[(size_t)y * a.fg_pitch + x] = (uint8_t)mk;
    }
}


// ---------------------------------------------------------------------------------------------------
// The resident fused kernel.
//
// Roles inside a CTA: PIPE_CTHREADS compute threads, a LOADER lane (draws work, waits for the tile's previous
// frame, issues the bulk loads) and PIPE_STORERS STORER lanes (write finished tiles back with bulk stores and publish
// them, taking the CTA's tiles in turn), each in a warp of its own, around a ring of PIPE_STAGES stages:
//
//     loader --full[s]--> compute warps --done[s]--> storer --freed[s]--> loader
//
// Ordering of a tile between the frame that wrote it (any CTA, this launch or the previous one) and the frame
// that reads it next -- a release/acquire pair at GPU scope, correct under the PTX memory model:
//   writer   compute warps' direct stores (modes >= 1)  --mbarrier done[s] (release.cta / acquire.cta)-->  storer lane;
//            the tile's bulk stores are COMPLETE (cp.async.bulk.wait_group, not .read);
//            fence.acq_rel.gpu  (cumulative: covers what the storer lane observed through the mbarrier);
//            st.relaxed.gpu tile_seq[tile]  (+ red.add done_count for the tail server)
//   reader   loader lane: ld.acquire.gpu tile_seq[tile] == seq_expect;  fence.proxy.async.global;  bulk loads;
//            the compute warps' own ld.global.cg follow the mbarrier full[s] the loads complete on.
// The release fence is a MEMBAR.GPU: a round trip through a memory system that this kernel keeps saturated.  It
// lives in the storer lanes, where it delays nothing but that lane's next write-back -- and with two lanes the other
// lane's write-back goes on beside it.  (One lane doing loads and stores, fence per tile: 0.71 of the HBM peak; one
// storer lane: 0.83 per tile, 0.95 with one fence per 2 tiles; NO fence -- flags behind the mere completion of the
// bulk stores -- reaches 0.97-0.99 but is wrong: see the storer lanes.  As built: profiles/.)
//
// LINEAR: cols % 32 == 0 and every image pitch is tight, so a pixel's byte offsets are plain multiples of its
// padded index (one bulk copy brings a tile's BGR bytes).  Otherwise the loader lane issues one bulk copy
// per row segment of the tile (rows start 16-byte aligned: the host checks pointer and pitch), landing the
// bytes at the same place, 3 * (pixel in tile); padding pixels are never evaluated.
// ---------------------------------------------------------------------------------------------------
#ifdef PIPE_MAXNREG_CFG
#define PIPE_KERNEL_BOUNDS __maxnreg__(PIPE_MAXNREG_CFG)
#else
#define PIPE_KERNEL_BOUNDS __launch_bounds__(PIPE_THREADS, PIPE_MINBLOCKS_CFG)
#endif
template <int K, bool TRACK, bool LINEAR>
__global__ void PIPE_KERNEL_BOUNDS mog_stream_kernel(const __grid_constant__ StreamArgs pa)
{
    extern __shared__ __align__(128) uint8_t stage_mem[];
    __shared__ __align__(8) uint64_t full[PIPE_STAGES];   // stage loaded (expect_tx of the loader lane) or end marker
    __shared__ __align__(8) uint64_t done[PIPE_STAGES];   // stage updated in place by all compute threads
    __shared__ __align__(8) uint64_t freed[PIPE_STAGES];  // stage written back (its bulk stores have read it)
    const FusedArgs &a = pa.f;
    const int tid = threadIdx.x;
    const bool is_loader = (tid == PIPE_CTHREADS);
    const bool is_storer = (tid >= PIPE_CTHREADS + 32) && ((tid & 31) == 0);  // lane 0 of every storer warp
    const int storer_x = (tid - PIPE_CTHREADS - 32) >> 5;                       // which one: it takes the stage uses x, x + PIPE_STORERS, ...
    // let the next launch on this stream (programmatic dependent launch) become resident as CTAs of
    // this one retire; it either parks in griddepcontrol.wait below or orders itself tile by tile
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // the scheduler's first answer (PIPE_STAGES consecutive items) travels while the CTA initialises
    uint32_t raw0 = 0;
    if (is_loader) raw0 = atomicAdd(pa.work_counter, (unsigned)PIPE_STAGES);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < PIPE_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&done[s], PIPE_CTHREADS);
            mbar_init(&freed[s], 1);
        }
        fence_mbar_init();
    }
    for (int s = 0; s < PIPE_STAGES; ++s) {
        uint8_t *st = stage_mem + (size_t)s * PIPE_STAGE_BYTES;
        if (tid < 4) reinterpret_cast<uint32_t *>(st + PIPE_OFF_FLAG)[tid] = 0u;  // dirty, tile, queue count, queue head
        for (int i = tid; i < PIPE_TILE; i += PIPE_THREADS) reinterpret_cast<uint16_t *>(st + PIPE_OFF_QUEUE)[i] = 0xffffu;
    }
    __syncthreads();
    // everything above touched shared memory and the scheduler counter only; frames, GMM state and flags are
    // ordered behind the previous grid on the stream from here on (FD_CHAIN frames order themselves tile by tile)
    if (pa.wait_grid) asm volatile("griddepcontrol.wait;" ::: "memory");

    auto hdr_of = [&](int s) -> volatile uint32_t * {
        return reinterpret_cast<volatile uint32_t *>(stage_mem + (size_t)s * PIPE_STAGE_BYTES + PIPE_OFF_FLAG);
    };
    auto hdr_ptr = [&](volatile uint32_t *h, int word) -> void *volatile * {
        return reinterpret_cast<void *volatile *>(const_cast<uint32_t *>(h + word));
    };
    auto tile_span = [&](int tile, size_t &p0, uint32_t &npx) {
        p0 = (size_t)tile * PIPE_TILE;
        const size_t rem = a.plane - p0;
        npx = rem < (size_t)PIPE_TILE ? (uint32_t)rem : (uint32_t)PIPE_TILE;
    };

    if (tid >= PIPE_CTHREADS) {
        if (is_loader) {
            // ---- loader lane -----------------------------------------------------------------------
            const uint64_t pol_first = l2_policy_evict_first();
            const uint32_t total = (uint32_t)pa.nframes * (uint32_t)pa.ntiles;
            const unsigned pitch_px = (unsigned)a.wpr * 32u;
            // a drawn work item, resolved against its frame's descriptor
            struct Item {
                int tile;  // -1: the queue is exhausted
                int frame;
                const uint8_t *bgr;
                float *state;
                uint8_t *nmodes;
                uint32_t *thr;
                unsigned int *tseq, *slow, *donec;
                uint32_t seq, flags, seen;
            };
            int dframe = -1;  // descriptor cache (draws are monotonic: the frame index never goes back)
            FrameDesc dcur;
            uint32_t fbase = 0;  // first work number of frame fpos
            int fpos = 0;
            auto resolve = [&](uint32_t g, Item &it) {
                if (g >= total) {
                    it.tile = -1;
                    return;
                }
                while (g >= fbase + (uint32_t)pa.ntiles) {
                    fbase += (uint32_t)pa.ntiles;
                    ++fpos;
                }
                if (fpos != dframe) {
                    const FrameDesc *dp = pa.descs ? pa.descs + fpos : &pa.inl[fpos];
                    const uint4 *q = reinterpret_cast<const uint4 *>(dp);
                    uint4 *w = reinterpret_cast<uint4 *>(&dcur);
                    w[0] = q[0];
                    w[1] = q[1];
                    w[2] = q[2];
                    w[3] = q[3];
                    dframe = fpos;
                }
                it.tile = (int)(g - fbase);
                it.frame = fpos;
                it.bgr = dcur.bgr;
                it.state = dcur.state;
                it.nmodes = dcur.nmodes;
                it.thr = dcur.thr_bits;
                it.tseq = dcur.tile_seq;
                it.slow = dcur.slow_count;
                it.donec = dcur.done_count;
                it.seq = dcur.seq_expect;
                it.flags = dcur.flags;
                // a look at the tile's flag well before the loads are due (an acquire: if it already shows the
                // expected value nothing more is needed when the tile is loaded)
                it.seen = it.seq + 1u;
                if (it.flags & FD_CHAIN) it.seen = ld_acquire_gpu(it.tseq + it.tile);
            };
            auto issue_load = [&](int s, const Item &it) {
                uint8_t *st = stage_mem + (size_t)s * PIPE_STAGE_BYTES;
                volatile uint32_t *h = hdr_of(s);
                h[HDR_TILE] = (uint32_t)it.tile;
                h[HDR_FRAME] = (uint32_t)it.frame;
                h[HDR_SEQ_OUT] = it.seq + 1u;
                *hdr_ptr(h, HDR_STATE) = it.state;
                *hdr_ptr(h, HDR_NMODES) = it.nmodes;
                *hdr_ptr(h, HDR_THR) = it.thr;
                *hdr_ptr(h, HDR_TSEQ) = it.tseq;
                *hdr_ptr(h, HDR_SLOW) = it.slow;
                *hdr_ptr(h, HDR_DONE) = it.donec;
                // (the caller has seen tile_seq == seq_expect with an acquire load)
                if (it.flags & FD_CHAIN) fence_proxy_async_global();
                size_t p0;
                uint32_t npx;
                tile_span(it.tile, p0, npx);
                if (LINEAR) {
                    mbar_expect_tx(&full[s], npx * 24u);
                    bulk_g2s_hint(st + PIPE_OFF_BGR, it.bgr + 3 * p0, npx * 3u, &full[s], pol_first);
                } else {
                    // BGR: one bulk copy per row segment of the tile (16-byte aligned on both sides, see the header)
                    const uint32_t y0 = (uint32_t)__umul64hi((unsigned long long)p0, pa.div_magic);
                    uint32_t bytes = 0;
                    {
                        uint32_t xa = (uint32_t)(p0 - (size_t)y0 * pitch_px), left = npx;
                        while (left) {
                            const uint32_t xb = min(pitch_px, xa + left), xv = min(xb, (uint32_t)a.cols);
                            if (xv > xa) bytes += (3u * (xv - xa) + 15u) & ~15u;
                            left -= xb - xa;
                            xa = 0;
                        }
                    }
                    mbar_expect_tx(&full[s], npx * 21u + bytes);
                    uint32_t y = y0, xa = (uint32_t)(p0 - (size_t)y0 * pitch_px), left = npx, e0 = 0;
                    while (left) {
                        const uint32_t xb = min(pitch_px, xa + left), xv = min(xb, (uint32_t)a.cols);
                        if (xv > xa)
                            bulk_g2s_hint(st + PIPE_OFF_BGR + 3u * e0, it.bgr + (size_t)y * a.in_pitch + 3u * xa,
                                          (3u * (xv - xa) + 15u) & ~15u, &full[s], pol_first);
                        left -= xb - xa;
                        e0 += xb - xa;
                        xa = 0;
                        ++y;
                    }
                }
#pragma unroll
                for (int cc = 0; cc < 5; ++cc)
                    bulk_g2s(st + cc * (PIPE_TILE * 4), it.state + (size_t)cc * a.plane + p0, npx * 4u, &full[s]);
                bulk_g2s(st + PIPE_OFF_NM, it.nmodes + p0, npx, &full[s]);
            };

            Item nxt, cur;
            uint32_t batch_next = raw0 + 1u, batch_left = (uint32_t)PIPE_STAGES - 1u;  // the first PIPE_STAGES items are consecutive
            resolve(raw0, nxt);
            for (int li = 0;; ++li) {
                const int s = li % PIPE_STAGES;
                // the stage's previous tile has been written back
                if (li >= PIPE_STAGES) mbar_wait(&freed[s], (uint32_t)(li / PIPE_STAGES - 1) & 1u);
                if (nxt.tile < 0) {
                    // End marker: one for every storer lane, in the stage uses li .. li + PIPE_STORERS - 1 (each lane owns
                    // every PIPE_STORERS-th use).  All of them are written before the compute warps are released, which
                    // pass them on together.
#pragma unroll
                    for (int k = 1; k < PIPE_STORERS; ++k)
                        if (li + k >= PIPE_STAGES) mbar_wait(&freed[(li + k) % PIPE_STAGES], (uint32_t)((li + k) / PIPE_STAGES - 1) & 1u);
#pragma unroll
                    for (int k = 0; k < PIPE_STORERS; ++k) hdr_of((li + k) % PIPE_STAGES)[HDR_TILE] = 0xffffffffu;
                    mbar_arrive(&full[s]);  // completes the phase: the compute warps read the marker, pass it on and stop
                    break;
                }
                if ((nxt.flags & FD_CHAIN) && nxt.seen != nxt.seq) {
                    // The tile's previous frame is still in flight (in another CTA, in the previous launch, or -- frames
                    // smaller than the ring -- in this CTA's own stages: the storer lane retires and publishes those on
                    // its own).  Bounded: a writer that died (launch error) must surface as an error, not hang the GPU.
                    unsigned spins = 0;
                    while ((nxt.seen = ld_acquire_gpu(nxt.tseq + nxt.tile)) != nxt.seq) {
                        __nanosleep(32);
                        if (++spins > (1u << 23)) __trap();
                    }
                }
                cur = nxt;
                uint32_t rawn;
                if (batch_left) {
                    rawn = batch_next++;
                    --batch_left;
                } else {
                    rawn = atomicAdd(pa.work_counter, 1u);  // issued now, consumed after this tile's loads are on their way
                }
                issue_load(s, cur);
                resolve(rawn, nxt);
            }
            return;
        }
        if (!is_storer) return;
        // ---- storer lanes --------------------------------------------------------------------------
        // PIPE_STORERS lanes take the CTA's finished tiles in turn.  Each writes its tile back, hands the stage to the
        // loader as soon as the bulk stores have READ it, and then publishes: waits until the stores are COMPLETE,
        // fence.acq_rel.gpu, flag.  A MEMBAR.GPU is a round trip through a memory system this kernel keeps saturated
        // (~1.5 us, most of a tile period): in ONE lane, paid per tile, it makes the lane the bottleneck (0.83 instead of
        // 0.95 of the HBM peak at 1080p) and has to be shared by a batch of tiles (NB = 2: 0.95; the second tile of a
        // batch is withheld from its waiter for a tile period); with two lanes every tile is published as soon as its
        // own stores have landed, and the fence of one lane runs beside the write-back of the other.  It cannot be
        // dropped -- without it a reader that has seen the flag can still see a stale plane of the tile (observed on
        // 640x480 frames, where consecutive frames of a stream are in flight together).
        constexpr int NB = PIPE_STORERS == 1 ? 2 : 1;
        unsigned int *pend_tseq[NB], *pend_done[NB];
        uint32_t pend_tile[NB], pend_seq[NB];
        int npend = 0;
        auto publish_pending = [&]() {
            if (npend == 0) return;
            bulk_wait_all<0>();  // their bulk stores are complete: performed, not merely read out of shared memory
            // ... and, with the compute warps' direct stores of those tiles (observed by this lane through done[s]),
            // visible at GPU scope before the flags
            if (!pa.relaxed_publish) fence_acq_rel_gpu();
#pragma unroll
            for (int k = 0; k < NB; ++k)
                if (k < npend) {
                    st_relaxed_gpu(pend_tseq[k] + pend_tile[k], pend_seq[k]);
                    if (pend_done[k]) red_add_relaxed_gpu(pend_done[k], 1u);
                }
            npend = 0;
        };
        for (int si = storer_x;; si += PIPE_STORERS) {
            const int s = si % PIPE_STAGES;
            const uint32_t par = (uint32_t)(si / PIPE_STAGES) & 1u;
            volatile uint32_t *h = hdr_of(s);
            uint8_t *st = stage_mem + (size_t)s * PIPE_STAGE_BYTES;
            // (one lane) the next tile normally arrives within a tile period; if it does not, somebody may be waiting
            // for the tile this lane still owes: publish it, then wait for as long as it takes
            if (NB > 1 && !mbar_try_wait_ns(&done[s], par, 4000u)) publish_pending();
            mbar_wait(&done[s], par);
            const int tile = (int)h[HDR_TILE];
            if (tile < 0) break;
            // 1. a full batch of retired tiles (the youngest has had a whole tile period for its bulk stores): publish it
            if (npend == NB || pa.fence_always) publish_pending();
            // 2. write this tile back
            float *state = reinterpret_cast<float *>(*hdr_ptr(h, HDR_STATE));
            uint8_t *nmodes = reinterpret_cast<uint8_t *>(*hdr_ptr(h, HDR_NMODES));
            uint32_t *thr = reinterpret_cast<uint32_t *>(*hdr_ptr(h, HDR_THR));
            bool store = true;
            if (TRACK) {
                store = (h[HDR_DIRTY] != 0u);
                h[HDR_DIRTY] = 0u;
            }
            size_t p0;
            uint32_t npx;
            tile_span(tile, p0, npx);
            if (store) {
#pragma unroll
                for (int cc = 0; cc < 5; ++cc)
                    bulk_s2g(state + (size_t)cc * a.plane + p0, st + cc * (PIPE_TILE * 4), npx * 4u);
                bulk_s2g(nmodes + p0, st + PIPE_OFF_NM, npx);
            }
            if (thr) {  // the tile's threshold bits: 128 B, one bulk store (word loop for a ragged last tile)
                if (npx == (uint32_t)PIPE_TILE) {
                    bulk_s2g(thr + (p0 >> 5), st + PIPE_OFF_BITS, PIPE_TILE / 8);
                } else {
                    const volatile uint32_t *bw = reinterpret_cast<const volatile uint32_t *>(st + PIPE_OFF_BITS);
                    for (uint32_t w = 0; w < npx / 32u; ++w) thr[(p0 >> 5) + w] = bw[w];
                }
            }
            bulk_commit();
            // slow-path census of the tile (the queue count is 4 x the number of 4-pixel groups that left the fast path)
            {
                const uint32_t q = h[HDR_QCNT];
                unsigned int *slow = reinterpret_cast<unsigned int *>(*hdr_ptr(h, HDR_SLOW));
                if (q && slow) red_add_relaxed_gpu(slow, q >> 2);
            }
            h[HDR_QCNT] = 0u;  // re-arm the slow-pixel queue
            h[HDR_QHEAD] = 0u;
#pragma unroll
            for (int k = 0; k < NB; ++k)
                if (k == npend) {
                    pend_tseq[k] = reinterpret_cast<unsigned int *>(*hdr_ptr(h, HDR_TSEQ));
                    pend_done[k] = reinterpret_cast<unsigned int *>(*hdr_ptr(h, HDR_DONE));
                    pend_tile[k] = (uint32_t)tile;
                    pend_seq[k] = h[HDR_SEQ_OUT];
                }
            ++npend;
            // 3. hand the stage back as soon as its stores have READ it
            bulk_wait_read<0>();
            mbar_arrive(&freed[s]);
            // 4. (several lanes) publish the tile as soon as its stores have landed: the other lane has the next one
            if (NB == 1) publish_pending();
        }
        publish_pending();
        bulk_wait_all<0>();  // shared memory must outlive the last bulk stores
        // (the loader lane has drawn its last number before it passed the end markers on)
        // the last storer lane of the last CTA to leave re-arms the scheduler slot for its next user (a launch the host
        // starts only after this one has said so) and tells the host
        if (atomicAdd(pa.exit_ticket, 1u) == gridDim.x * (unsigned)PIPE_STORERS - 1u) {
            *pa.work_counter = 0u;
            *pa.exit_ticket = 0u;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned int *>(pa.done_flag) = pa.launch_id;
        }
        return;
    }

    // ---- compute warps ---------------------------------------------------------------------------
    const unsigned pitch_px = (unsigned)a.wpr * 32u;
    auto locate = [&](int tile, size_t &pidx, int &y, int &x) -> bool {
        pidx = (size_t)tile * PIPE_TILE + (size_t)tid * 4;
        if (LINEAR) {
            y = 0;
            x = 0;
            return pidx < a.plane;
        }
        y = (int)__umul64hi((unsigned long long)pidx, pa.div_magic);
        x = (int)(pidx - (size_t)y * pitch_px);
        return (pidx < a.plane) && (x < a.cols);
    };
    const uint32_t zero_nib = pa.zero_in ? 0xFu : 0u;

    for (int i = 0;; ++i) {
        const int s = i % PIPE_STAGES;
        uint8_t *st = stage_mem + (size_t)s * PIPE_STAGE_BYTES;
        float *sm0 = reinterpret_cast<float *>(st) + tid * 4;
        volatile uint32_t *h = hdr_of(s);

        mbar_wait(&full[s], (uint32_t)(i / PIPE_STAGES) & 1u);

        const int tile = (int)h[HDR_TILE];
        if (tile < 0) {
            // every storer lane reads a marker of its own: this stage's and the next stages' (written by the loader
            // lane before it released this one; those stages' previous tiles are long done)
#pragma unroll
            for (int k = 0; k < PIPE_STORERS; ++k) mbar_arrive(&done[(s + k) % PIPE_STAGES]);
            break;
        }
        size_t pidx;
        int y, x;
        const bool active = locate(tile, pidx, y, x);

        bool dirty = !TRACK, slow = false;
        uint32_t nib = 0;
        if (active) {
            const uint32_t nm = *(reinterpret_cast<const uint32_t *>(st + PIPE_OFF_NM) + tid);
            bool fast = false;
            const uint32_t two = nm - 0x01010101u;  // per pixel: 0 = one mode, 1 = two modes
            if (TRACK && two == 0u && a.c.aT == 0.f && !a.bgr_out && !a.hsv_out && !a.fg_out) {
                // frozen one-mode model: classify in place, foreground included
                const uint32_t *smb = reinterpret_cast<const uint32_t *>(st + PIPE_OFF_BGR) + tid * 3;
                const uint32_t w0 = smb[0], w1 = smb[1], w2 = smb[2];
                float4 W = *reinterpret_cast<const float4 *>(sm0);
                float4 V = *reinterpret_cast<const float4 *>(sm0 + PIPE_TILE);
                float4 A = *reinterpret_cast<const float4 *>(sm0 + 2 * PIPE_TILE);
                float4 B = *reinterpret_cast<const float4 *>(sm0 + 3 * PIPE_TILE);
                float4 C = *reinterpret_cast<const float4 *>(sm0 + 4 * PIPE_TILE);
                const uint32_t b0 = w0 & 255u, g0 = (w0 >> 8) & 255u, r0 = (w0 >> 16) & 255u, b1 = w0 >> 24, g1 = w1 & 255u,
                               r1 = (w1 >> 8) & 255u, b2 = (w1 >> 16) & 255u, g2 = w1 >> 24, r2 = w2 & 255u, b3 = (w2 >> 8) & 255u,
                               g3 = (w2 >> 16) & 255u, r3 = w2 >> 24;
                bool d2 = false;
                const uint32_t m0 = frozen_px((float)b0, (float)g0, (float)r0, W.x, V.x, A.x, B.x, C.x, a.c, d2);
                const uint32_t m1 = frozen_px((float)b1, (float)g1, (float)r1, W.y, V.y, A.y, B.y, C.y, a.c, d2);
                const uint32_t m2 = frozen_px((float)b2, (float)g2, (float)r2, W.z, V.z, A.z, B.z, C.z, a.c, d2);
                const uint32_t m3 = frozen_px((float)b3, (float)g3, (float)r3, W.w, V.w, A.w, B.w, C.w, a.c, d2);
                if ((m0 | m1 | m2 | m3) != 0xffffffffu) {
                    fast = true;
                    if (d2) {
                        *reinterpret_cast<float4 *>(sm0) = W;
                        *reinterpret_cast<float4 *>(sm0 + PIPE_TILE) = V;
                        *reinterpret_cast<float4 *>(sm0 + 2 * PIPE_TILE) = A;
                        *reinterpret_cast<float4 *>(sm0 + 3 * PIPE_TILE) = B;
                        *reinterpret_cast<float4 *>(sm0 + 4 * PIPE_TILE) = C;
                        dirty = true;
                    }
                    nib = zero_nib;
                    if ((m0 | m1 | m2 | m3) != 0u && a.do_hsv) {  // some pixel keeps its colour: threshold it
                        auto bit = [&](uint32_t mk, uint32_t b, uint32_t g, uint32_t r) -> uint32_t {
                            if (!mk) return zero_nib & 1u;
                            int hh, sa, v;
                            bgr2hsv_px_div((int)b, (int)g, (int)r, hh, sa, v);
                            return ((a.lo[0] <= hh) & (hh <= a.hi[0]) & (a.lo[1] <= sa) & (sa <= a.hi[1]) & (a.lo[2] <= v) & (v <= a.hi[2])) ? 1u : 0u;
                        };
                        nib = bit(m0, b0, g0, r0) | (bit(m1, b1, g1, r1) << 1) | (bit(m2, b2, g2, r2) << 2) | (bit(m3, b3, g3, r3) << 3);
                    }
                }
            } else if ((two & ~0x01010101u) == 0u && (K >= 2 || two == 0u)) {
                // mode 1's weights first: the only global read of this path, in flight during the mode-0 maths
                float4 W1 = make_float4(0.f, 0.f, 0.f, 0.f);
                float *w1p = nullptr;
                if (two != 0u) {
                    float *state = reinterpret_cast<float *>(*hdr_ptr(h, HDR_STATE));
                    w1p = state + (size_t)5 * a.plane + pidx;
                    W1 = ld_state_f4(w1p);
                }
                const uint32_t *smb = reinterpret_cast<const uint32_t *>(st + PIPE_OFF_BGR) + tid * 3;
                const uint32_t w0 = smb[0], w1 = smb[1], w2 = smb[2];
                float4 W = *reinterpret_cast<const float4 *>(sm0);
                float4 V = *reinterpret_cast<const float4 *>(sm0 + PIPE_TILE);
                float4 A = *reinterpret_cast<const float4 *>(sm0 + 2 * PIPE_TILE);
                float4 B = *reinterpret_cast<const float4 *>(sm0 + 3 * PIPE_TILE);
                float4 C = *reinterpret_cast<const float4 *>(sm0 + 4 * PIPE_TILE);
                uint32_t n0 = nm & 255u, n1 = (nm >> 8) & 255u, n2 = (nm >> 16) & 255u, n3 = nm >> 24;
                bool d2 = false;
                bool ok = fast_px<TRACK>((float)(w0 & 255u), (float)((w0 >> 8) & 255u), (float)((w0 >> 16) & 255u), W.x,
                                         V.x, A.x, B.x, C.x, W1.x, n0, a.c, d2);
                ok &= fast_px<TRACK>((float)(w0 >> 24), (float)(w1 & 255u), (float)((w1 >> 8) & 255u), W.y, V.y, A.y, B.y,
                                     C.y, W1.y, n1, a.c, d2);
                ok &= fast_px<TRACK>((float)((w1 >> 16) & 255u), (float)(w1 >> 24), (float)(w2 & 255u), W.z, V.z, A.z, B.z,
                                     C.z, W1.z, n2, a.c, d2);
                ok &= fast_px<TRACK>((float)((w2 >> 8) & 255u), (float)((w2 >> 16) & 255u), (float)(w2 >> 24), W.w, V.w,
                                     A.w, B.w, C.w, W1.w, n3, a.c, d2);
                if (ok) {
                    fast = true;
                    if (!TRACK || d2) {
                        *reinterpret_cast<float4 *>(sm0) = W;
                        *reinterpret_cast<float4 *>(sm0 + PIPE_TILE) = V;
                        *reinterpret_cast<float4 *>(sm0 + 2 * PIPE_TILE) = A;
                        *reinterpret_cast<float4 *>(sm0 + 3 * PIPE_TILE) = B;
                        *reinterpret_cast<float4 *>(sm0 + 4 * PIPE_TILE) = C;
                        if (two != 0u) {
                            st_state_f4(w1p, W1);
                            const uint32_t nnew = n0 | (n1 << 8) | (n2 << 16) | (n3 << 24);
                            if (nnew != nm) *(reinterpret_cast<uint32_t *>(st + PIPE_OFF_NM) + tid) = nnew;
                        }
                        dirty = true;
                    }
                    nib = zero_nib;
                    // all four pixels are background: the published frame / HSV / mask are zero here
                    if (a.bgr_out) {
                        uint8_t *d = LINEAR ? a.bgr_out + 3 * pidx : a.bgr_out + (size_t)y * a.bgr_out_pitch + 3 * x;
                        st_stream_u32(d, 0u);
                        st_stream_u32(d + 4, 0u);
                        st_stream_u32(d + 8, 0u);
                    }
                    if (a.hsv_out) {
                        uint8_t *d = LINEAR ? a.hsv_out + 3 * pidx : a.hsv_out + (size_t)y * a.hsv_pitch + 3 * x;
                        st_stream_u32(d, 0u);
                        st_stream_u32(d + 4, 0u);
                        st_stream_u32(d + 8, 0u);
                    }
                    if (a.fg_out) st_stream_u32(LINEAR ? a.fg_out + pidx : a.fg_out + (size_t)y * a.fg_pitch + x, 0u);
                }
            }
            slow = !fast;
        }
        // threshold mask, 1 bit/pixel: 8 lanes x 4 px -> one word of the stage's bit block
        uint32_t *sbits = reinterpret_cast<uint32_t *>(st + PIPE_OFF_BITS);
        {
            const unsigned lane = tid & 31u;
            uint32_t wv = nib << (4 * (lane & 7u));
            wv |= __shfl_xor_sync(0xffffffffu, wv, 1);
            wv |= __shfl_xor_sync(0xffffffffu, wv, 2);
            wv |= __shfl_xor_sync(0xffffffffu, wv, 4);
            if ((lane & 7u) == 0) sbits[tid >> 3] = wv;
        }
        __syncwarp();  // the words are in place before any of this warp's pixels can be claimed
        // queue the pixels that left the fast path, then help drain the tile's queue
        volatile uint32_t *qcnt = h + HDR_QCNT;
        volatile uint32_t *qhead = h + HDR_QHEAD;
        volatile uint16_t *queue = reinterpret_cast<volatile uint16_t *>(st + PIPE_OFF_QUEUE);
        if (slow) {
            const uint32_t base = atomicAdd(const_cast<uint32_t *>(qcnt), 4u);
            __threadfence_block();  // release: this warp's threshold words are visible before its pixels can be claimed
#pragma unroll
            for (int q = 0; q < 4; ++q) queue[base + q] = (uint16_t)(tid * 4 + q);  // relaxed (volatile) flag-style publish
        }
        __syncwarp();  // this warp's entries are published before any of its lanes starts claiming
        for (;;) {     // warp-level claiming: lane 0 takes up to 32 queue slots, one pixel per lane
            uint32_t h0 = 0, take = 0;
            if ((tid & 31) == 0) {
                const uint32_t hd = *qhead, c = *qcnt;
                if (hd < c) {
                    take = min(32u, c - hd);
                    if (atomicCAS(const_cast<uint32_t *>(qhead), hd, hd + take) == hd)
                        h0 = hd;
                    else
                        take = 0xffffffffu;  // lost the race: look again
                }
            }
            h0 = __shfl_sync(0xffffffffu, h0, 0);
            take = __shfl_sync(0xffffffffu, take, 0);
            if (take == 0u) break;
            if (take == 0xffffffffu) continue;
            if ((uint32_t)(tid & 31) < take) {
                const uint32_t hq = h0 + (uint32_t)(tid & 31);
                uint32_t e;
                while ((e = queue[hq]) == 0xffffu) {}  // reserved by a pusher of another warp that is about to fill it
                __threadfence_block();                 // acquire: pairs with the pusher's fence
                queue[hq] = 0xffffu;
                float *state = reinterpret_cast<float *>(*hdr_ptr(h, HDR_STATE));
                slow_pixel<K, TRACK>(pa, state, st, (int)e, (size_t)tile * PIPE_TILE + e);
            }
            __syncwarp();
        }
        if (TRACK && dirty) h[HDR_DIRTY] = 1u;
        // hand the stage to the producer: generic-proxy writes -> async proxy, then arrive
        fence_proxy_async();
        mbar_arrive(&done[s]);
    }
}

}  // namespace oat
