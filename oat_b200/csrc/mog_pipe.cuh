// mog_pipe.cuh -- the steady-state form of the fused kernel (mog_fused.cuh has the semantics and
// the generic fallback): the same per-pixel MOG2 + setTo(0) + BGR->HSV + inRange arithmetic, fed by
// an asynchronous bulk-copy pipeline instead of per-thread global loads.
//
//   * persistent grid (CTAs/SM x SM count), each CTA walks tiles of PIPE_TILE = 1024 consecutive
//     padded pixels (4 per thread);
//   * the hot GMM planes -- mode 0 {weight, variance, mean b/g/r} and the mode-count bytes, i.e.
//     everything a single-mode pixel needs -- are staged global -> shared with cp.async.bulk
//     (TMA engine, SASS UBLKCP) into a PIPE_STAGES-deep ring tracked by mbarriers, updated IN PLACE
//     in shared memory and written back with cp.async.bulk shared -> global; no warp ever waits on
//     a global load of state in the common case, and the bytes in flight per SM are set by the ring
//     depth, not by registers/occupancy;
//   * BGR input (12 B/thread) is register-prefetched one tile ahead;
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
#define PIPE_STAGES_CFG 3
#endif
#ifndef PIPE_MINBLOCKS_CFG
#define PIPE_MINBLOCKS_CFG 2
#endif
constexpr int PIPE_CTHREADS = PIPE_CTHREADS_CFG;    // compute threads (8 warps), 4 pixels each
constexpr int PIPE_THREADS = PIPE_CTHREADS + 32;     // + one producer warp (bulk loads / stores)
constexpr int PIPE_TILE = PIPE_CTHREADS * 4;        // pixels per tile
constexpr int PIPE_STAGES = PIPE_STAGES_CFG;
#ifndef PIPE_CTAS_PER_SM_CFG
#define PIPE_CTAS_PER_SM_CFG PIPE_MINBLOCKS_CFG
#endif
constexpr int PIPE_CTAS_PER_SM = PIPE_CTAS_PER_SM_CFG;  // persistent grid = this x SM count
constexpr int PIPE_OFF_NM = 5 * PIPE_TILE * 4;       // stage layout: 5 fp32 planes | mode counts | BGR | flag
constexpr int PIPE_OFF_BGR = PIPE_OFF_NM + PIPE_TILE;
constexpr int PIPE_OFF_FLAG = PIPE_OFF_BGR + 3 * PIPE_TILE;   // +0 dirty flag, +4 tile number, +8 queue count, +12 queue head
constexpr int PIPE_OFF_BITS = PIPE_OFF_FLAG + 128;            // threshold bits of the tile (PIPE_TILE / 8 bytes)
constexpr int PIPE_OFF_QUEUE = PIPE_OFF_BITS + PIPE_TILE / 8;  // slow-pixel queue (u16 pixel-in-tile, 0xffff = empty)
constexpr int PIPE_STAGE_BYTES = PIPE_OFF_QUEUE + 2 * PIPE_TILE;  // 26880 B
constexpr int PIPE_SMEM_BYTES = PIPE_STAGES * PIPE_STAGE_BYTES;

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
// Ampere-style 4-byte async copy (SASS LDGSTS) whose completion is counted by an mbarrier.
__device__ __forceinline__ void cp_async4(void *dst_smem, const void *src_gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t *b)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}

struct PipeArgs {
    FusedArgs f;
    int ntiles;
    unsigned long long div_magic;  // ceil(2^64 / pitch_px): y = umul64hi(pidx, magic), exact for every pidx < 2^32
    int zero_in;                   // HSV (0,0,0) lies inside the inRange band
    int grid_tiles;                // tile stride between consecutive tiles of one CTA (= gridDim.x)
    unsigned int *tile_counter;    // dynamic tile scheduler (LINEAR frames): monotonic draw counter of THIS launch's slot
    unsigned int counter_base;     // value of *tile_counter before this launch's first draw (the host knows every launch's draw count)
    // Tile-granular ordering between consecutive launches on one model ("chain"): tile_seq[i] holds the
    // sequence number of the last launch that has finished (stored + made visible) tile i.  A chained
    // launch is NOT ordered behind its predecessor grid (no griddepcontrol.wait): its producer loads
    // tile i once tile_seq[i] == seq_expect, so its first tiles stream in while the predecessor's
    // last tiles are still being computed.  Every launch publishes seq_expect + 1.
    unsigned int *tile_seq;
    unsigned int seq_expect;
    int chain;
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
__device__ __forceinline__ void slow_pixel(const PipeArgs &pa, uint8_t *st, const int e, const size_t pidx)
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
        const float *g = a.state + (size_t)(m * 5) * a.plane + pidx;
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
            float *g = a.state + (size_t)(m * 5) * a.plane + pidx;
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

// LINEAR: cols % 32 == 0 and every image pitch is tight, so a pixel's byte offsets are plain
// multiples of its padded index (no row/column split anywhere in the steady-state loop).
template <int K, bool TRACK, bool LINEAR>
__global__ void __launch_bounds__(PIPE_THREADS, PIPE_MINBLOCKS_CFG) mog_pipe_kernel(const __grid_constant__ PipeArgs pa)
{
    extern __shared__ __align__(128) uint8_t stage_mem[];
    __shared__ __align__(8) uint64_t full[PIPE_STAGES];  // stage loaded: 1 expect_tx arrive + 256 cp.async arrives
    __shared__ __align__(8) uint64_t done[PIPE_STAGES];  // stage updated in place by all compute threads
    const FusedArgs &a = pa.f;
    const int tid = threadIdx.x;
    // let the next launch on this stream (programmatic dependent launch) become resident as CTAs of
    // this one retire; it parks in griddepcontrol.wait below until this grid has completed and flushed
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < PIPE_STAGES; ++s) {
            mbar_init(&full[s], LINEAR ? 1 : 1 + PIPE_CTHREADS);
            mbar_init(&done[s], PIPE_CTHREADS);
        }
        fence_mbar_init();
    }
    for (int s = 0; s < PIPE_STAGES; ++s) {
        uint8_t *st = stage_mem + (size_t)s * PIPE_STAGE_BYTES;
        if (tid < 4) reinterpret_cast<uint32_t *>(st + PIPE_OFF_FLAG)[tid] = 0u;  // dirty, tile, queue count, queue head
        for (int i = tid; i < PIPE_TILE; i += PIPE_THREADS) reinterpret_cast<uint16_t *>(st + PIPE_OFF_QUEUE)[i] = 0xffffu;
    }
    __syncthreads();
    // everything above touched shared memory only; global memory (GMM state, scheduler counters) is
    // ordered behind the previous grid on the stream from here on
    // (a chained launch orders itself tile by tile through pa.tile_seq instead)
    if (!pa.chain) asm volatile("griddepcontrol.wait;" ::: "memory");

    // Tile order.  LINEAR frames use a dynamic scheduler: the producer draws tile numbers from a
    // global counter (tiles that hit the multi-mode slow path take several times longer than the
    // rest, so a static split leaves the kernel waiting for its unluckiest CTA) and also brings the
    // BGR bytes in with one more bulk copy.  Otherwise tile i of this CTA is blockIdx.x + i*gridDim.x
    // and every compute thread copies its own 12 input bytes with cp.async.
    constexpr bool DYN = LINEAR;
    const int first = blockIdx.x, stride = pa.grid_tiles;
    const int my_n = DYN ? 0x7fffffff : (first < pa.ntiles ? (pa.ntiles - first + stride - 1) / stride : 0);
    auto stage_tile = [&](int s) -> volatile int * {
        return reinterpret_cast<volatile int *>(stage_mem + (size_t)s * PIPE_STAGE_BYTES + PIPE_OFF_FLAG + 4);
    };

    if (tid >= PIPE_CTHREADS) {
        // ---- producer warp: one lane drives the bulk-copy engine -------------------------------
        if (tid != PIPE_CTHREADS) return;
        const uint64_t pol_first = l2_policy_evict_first();
        auto tile_span = [&](int tile, size_t &p0, uint32_t &npx) {
            p0 = (size_t)tile * PIPE_TILE;
            const size_t rem = a.plane - p0;
            npx = rem < (size_t)PIPE_TILE ? (uint32_t)rem : (uint32_t)PIPE_TILE;
        };
        auto issue_load = [&](int s, int tile, uint32_t seen) {
            size_t p0;
            uint32_t npx;
            tile_span(tile, p0, npx);
            uint8_t *st = stage_mem + (size_t)s * PIPE_STAGE_BYTES;
            // chained launch: the predecessor must have published this tile.  `seen` is a relaxed read of
            // the flag taken one refill earlier (hides the L2 round trip in the steady state, where the
            // predecessor is long past this tile); state bytes reach this SM only through bulk copies and
            // L1::no_allocate loads, i.e. from L2, where the publisher's release ordered them before the flag.
            if (pa.chain && seen != pa.seq_expect) {
                // bounded: a predecessor that died (launch error) must surface as an error, not hang the GPU
                unsigned spins = 0;
                while (ld_acquire_gpu(pa.tile_seq + tile) != pa.seq_expect) {
                    __nanosleep(20);
                    if (++spins > (1u << 23)) __trap();
                }
            }
            mbar_expect_tx(&full[s], npx * (DYN ? 24u : 21u));
#pragma unroll
            for (int cc = 0; cc < 5; ++cc)
                bulk_g2s(st + cc * (PIPE_TILE * 4), a.state + (size_t)cc * a.plane + p0, npx * 4u, &full[s]);
            bulk_g2s(st + PIPE_OFF_NM, a.nmodes + p0, npx, &full[s]);
            if (DYN) bulk_g2s_hint(st + PIPE_OFF_BGR, a.bgr + 3 * p0, npx * 3u, &full[s], pol_first);
        };
        // next tile of this CTA's sequence, or -1 when the frame is exhausted (DYN: every CTA draws
        // exactly one number >= ntiles, so the host knows a launch's draw count and the counters never
        // need re-arming: each launch is told where its numbers start, pa.counter_base)
        // DYN: the first PIPE_STAGES tiles of a CTA are fixed (no atomic on the start-up path); later
        // ones are drawn from the global counter one refill AHEAD of their use, so the L2 round trip
        // of the atomic hides behind the wait for the stage.
        int seq = 0, ahead = 0;
        bool ended = false;
        // publish a finished tile (its bulk stores are complete, the compute warps' direct stores
        // were ordered by done[s]) to the next launch on this model
        const uint32_t seq_out = pa.seq_expect + 1u;
        // Steady-state publishes are RELAXED stores: a release would cost a MEMBAR.GPU that also waits for
        // the bulk loads just issued by the refill (measured: +0.6 us per tile, 4K frame 62 -> 80 us).  What
        // orders the data before the flag instead: the tile's bulk stores are complete
        // (cp.async.bulk.wait_group, non-.read) and the compute warps' few direct stores (modes >= 1,
        // write-through) were issued before those bulk stores, i.e. at least a store round trip earlier (a whole
        // tile time, ~2 us, for all but a CTA's last tile); the consumer is a full frame behind except at
        // the frame boundary, and reads state only from L2 (bulk copies, ld.cg).
        auto publish = [&](int tile) {
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(pa.tile_seq + tile), "r"(seq_out) : "memory");
        };
        int unpublished = -1;
        uint32_t ahead_seen = pa.seq_expect + 1u, seen = pa.seq_expect + 1u;  // "not seen yet"
        auto peek = [&]() {  // relaxed look at the flag of the tile drawn for the NEXT refill
            if (pa.chain && ahead < pa.ntiles) ahead_seen = *reinterpret_cast<const volatile unsigned int *>(pa.tile_seq + ahead);
        };
        uint32_t raw = 0;
        bool drew = false;
        auto next_tile = [&]() -> int {
            int t;
            seen = pa.seq_expect + 1u;
            if (DYN && seq >= PIPE_STAGES) {
                t = ahead;
                seen = ahead_seen;
                if (t < pa.ntiles) {
                    raw = atomicAdd(pa.tile_counter, 1u);  // issued now, consumed after this tile's loads are on their way
                    drew = true;
                }
            } else {
                t = first + seq * stride;
                if (DYN && seq == PIPE_STAGES - 1) {
                    raw = atomicAdd(pa.tile_counter, 1u);
                    drew = true;
                }
            }
            ++seq;
            if (t >= pa.ntiles) {
                ended = true;
                return -1;
            }
            return t;
        };
        auto refill = [&](int s) {  // give stage s its next tile, or the end marker
            const int t = next_tile();
            *stage_tile(s) = t;
            if (t >= 0)
                issue_load(s, t, seen);
            else if (DYN)
                mbar_arrive(&full[s]);  // completes the phase: the compute warps read the marker and stop
            if (drew) {  // the scheduler's answer (an L2 round trip) and the look at that tile's flag: off the load's path
                const uint32_t d = raw - pa.counter_base;
                ahead = d < 0x40000000u ? (int)d + PIPE_STAGES * (int)gridDim.x : 0x7fffffff;
                peek();
                drew = false;
            }
        };
        for (int k = 0; k < PIPE_STAGES && !ended; ++k) refill(k);  // every stage starts loaded
        for (int i = 0;; ++i) {
            const int s = i % PIPE_STAGES;
            const int tile = *stage_tile(s);
            if (tile < 0) break;
            uint8_t *st = stage_mem + (size_t)s * PIPE_STAGE_BYTES;
            mbar_wait(&done[s], (uint32_t)(i / PIPE_STAGES) & 1u);
            bool store = true;
            if (TRACK) {
                volatile uint32_t *flag = reinterpret_cast<volatile uint32_t *>(st + PIPE_OFF_FLAG);
                store = (*flag != 0u);
                *flag = 0u;
            }
            if (store) {
                size_t p0;
                uint32_t npx;
                tile_span(tile, p0, npx);
#pragma unroll
                for (int cc = 0; cc < 5; ++cc)
                    bulk_s2g(a.state + (size_t)cc * a.plane + p0, st + cc * (PIPE_TILE * 4), npx * 4u);
                bulk_s2g(a.nmodes + p0, st + PIPE_OFF_NM, npx);
            }
            if (a.thr_bits) {  // the tile's threshold bits: 128 B, one bulk store (word loop for a ragged last tile)
                size_t p0;
                uint32_t npx;
                tile_span(tile, p0, npx);
                if (npx == (uint32_t)PIPE_TILE) {
                    bulk_s2g(a.thr_bits + (p0 >> 5), st + PIPE_OFF_BITS, PIPE_TILE / 8);
                } else {
                    const volatile uint32_t *bw = reinterpret_cast<const volatile uint32_t *>(st + PIPE_OFF_BITS);
                    for (uint32_t w = 0; w < npx / 32u; ++w) a.thr_bits[(p0 >> 5) + w] = bw[w];
                }
            }
            reinterpret_cast<volatile uint32_t *>(st + PIPE_OFF_FLAG)[2] = 0u;  // re-arm the slow-pixel queue
            reinterpret_cast<volatile uint32_t *>(st + PIPE_OFF_FLAG)[3] = 0u;
            bulk_commit();
            // refill THIS stage as soon as its stores have read it (not one tile later): its next tile is
            // then in flight for PIPE_STAGES-1 tile times instead of one
            if (!ended) {
                bulk_wait_read<0>();
                refill(s);
            } else if (!DYN) {
                *stage_tile(s) = -1;
            }
            // off the refill's critical path: the PREVIOUS tile's stores (committed one tile time ago) are
            // complete by now -- publish it to the next launch on this model
            if (unpublished >= 0) {
                bulk_wait_all<1>();
                publish(unpublished);
            }
            unpublished = tile;
        }
        bulk_wait_all<0>();  // shared memory must outlive the last bulk stores; and they must be complete
        // the last tile too: its bulk stores are complete, and the compute warps' direct stores were issued before
        // those (a release here is a MEMBAR.GPU on every CTA's exit path, ~0.7 us that the successor's CTA waits for)
        if (unpublished >= 0) publish(unpublished);
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
    auto bgr_issue = [&](int i) {  // (static order only) this thread's 12 input bytes of its i-th tile
        const int s = i % PIPE_STAGES;
        size_t pidx;
        int y, x;
        if (locate(first + i * stride, pidx, y, x)) {
            const uint8_t *src = a.bgr + (size_t)y * a.in_pitch + 3 * x;
            uint8_t *dst = stage_mem + (size_t)s * PIPE_STAGE_BYTES + PIPE_OFF_BGR + tid * 12;
            cp_async4(dst, src);
            cp_async4(dst + 4, src + 4);
            cp_async4(dst + 8, src + 8);
        }
        cp_async_arrive_noinc(&full[s]);
    };
    if (!DYN) {
        const int pre = my_n < PIPE_STAGES - 1 ? my_n : PIPE_STAGES - 1;
        for (int i = 0; i < pre; ++i) bgr_issue(i);
    }
    const uint32_t zero_nib = pa.zero_in ? 0xFu : 0u;
    unsigned nslow = 0;

    for (int i = 0; i < my_n; ++i) {
        const int s = i % PIPE_STAGES;
        if (!DYN && i + PIPE_STAGES - 1 < my_n) bgr_issue(i + PIPE_STAGES - 1);
        uint8_t *st = stage_mem + (size_t)s * PIPE_STAGE_BYTES;
        float *sm0 = reinterpret_cast<float *>(st) + tid * 4;

        mbar_wait(&full[s], (uint32_t)(i / PIPE_STAGES) & 1u);

        const int tile = DYN ? *stage_tile(s) : first + i * stride;
        if (DYN && tile < 0) break;
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
                            int h, sa, v;
                            bgr2hsv_px_div((int)b, (int)g, (int)r, h, sa, v);
                            return ((a.lo[0] <= h) & (h <= a.hi[0]) & (a.lo[1] <= sa) & (sa <= a.hi[1]) & (a.lo[2] <= v) & (v <= a.hi[2])) ? 1u : 0u;
                        };
                        nib = bit(m0, b0, g0, r0) | (bit(m1, b1, g1, r1) << 1) | (bit(m2, b2, g2, r2) << 2) | (bit(m3, b3, g3, r3) << 3);
                    }
                }
            } else if ((two & ~0x01010101u) == 0u && (K >= 2 || two == 0u)) {
                // mode 1's weights first: the only global read of this path, in flight during the mode-0 maths
                float4 W1 = make_float4(0.f, 0.f, 0.f, 0.f);
                float *w1p = a.state + (size_t)5 * a.plane + pidx;
                if (two != 0u) W1 = ld_state_f4(w1p);
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
        volatile uint32_t *qcnt = reinterpret_cast<volatile uint32_t *>(st + PIPE_OFF_FLAG) + 2;
        volatile uint32_t *qhead = qcnt + 1;
        volatile uint16_t *queue = reinterpret_cast<volatile uint16_t *>(st + PIPE_OFF_QUEUE);
        if (slow) {
            ++nslow;
            const uint32_t base = atomicAdd(const_cast<uint32_t *>(qcnt), 4u);
            __threadfence_block();  // release: this warp's threshold words are visible before its pixels can be claimed
#pragma unroll
            for (int i = 0; i < 4; ++i) queue[base + i] = (uint16_t)(tid * 4 + i);  // relaxed (volatile) flag-style publish
        }
        __syncwarp();  // this warp's entries are published before any of its lanes starts claiming
        for (;;) {     // warp-level claiming: lane 0 takes up to 32 queue slots, one pixel per lane
            uint32_t h0 = 0, take = 0;
            if ((tid & 31) == 0) {
                const uint32_t h = *qhead, c = *qcnt;
                if (h < c) {
                    take = min(32u, c - h);
                    if (atomicCAS(const_cast<uint32_t *>(qhead), h, h + take) == h)
                        h0 = h;
                    else
                        take = 0xffffffffu;  // lost the race: look again
                }
            }
            h0 = __shfl_sync(0xffffffffu, h0, 0);
            take = __shfl_sync(0xffffffffu, take, 0);
            if (take == 0u) break;
            if (take == 0xffffffffu) continue;
            if ((uint32_t)(tid & 31) < take) {
                const uint32_t h = h0 + (uint32_t)(tid & 31);
                uint32_t e;
                while ((e = queue[h]) == 0xffffu) {}  // reserved by a pusher of another warp that is about to fill it
                __threadfence_block();                // acquire: pairs with the pusher's fence
                queue[h] = 0xffffu;
                slow_pixel<K, TRACK>(pa, st, (int)e, (size_t)tile * PIPE_TILE + e);
            }
            __syncwarp();
        }
        if (TRACK && dirty) *reinterpret_cast<volatile uint32_t *>(st + PIPE_OFF_FLAG) = 1u;
        // hand the stage to the producer: generic-proxy writes -> async proxy, then arrive
        fence_proxy_async();
        mbar_arrive(&done[s]);
    }
    // slow-path census (drives the host's choice between this kernel and the generic one)
    nslow = __reduce_add_sync(0xffffffffu, nslow);
    if ((tid & 31) == 0 && nslow && a.slow_count) atomicAdd(a.slow_count, nslow);
}

}  // namespace oat
