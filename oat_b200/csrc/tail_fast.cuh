// tail_fast.cuh -- the whole detect tail of `posidet hsv` in ONE launch for the common case
// (a handful of blobs): morphology + row extents by all CTAs, then the LAST CTA to finish labels
// the mask in shared memory and produces the detection.
//
//   cv::erode / cv::dilate, MORPH_RECT k x k      (src/positiondetector/HSVDetector.cpp:152-156, :253-273)
//   siftContours()                                  (src/positiondetector/DetectorFunc.cpp:31-66)
//
// Same contour semantics as tail.cuh (external contours = 8-connected foreground + its 4-connected
// holes; exact integer 2x2-cell moments; reverse-raster tie-break), different machinery: the
// union-find nodes are whole row RUNS (foreground runs and "candidate" background runs between
// the first and last foreground pixel of a row), extracted from the bit mask with warp scans and
// kept -- together with the parent array and the per-contour accumulators -- in shared memory,
// so every dependent pointer hop costs ~30 cycles instead of an L2 round trip, and nothing is
// proportional to the frame size except one streaming pass over the mask.
//
// The run table is bounded by the shared-memory budget.  A mask that does not fit (more runs or
// contours than the table holds: the first frame's whole-image blob, heavy noise) makes the
// per-frame kernel report TAIL_OVERFLOW; the host then replays the frame through the unbounded
// global-memory path of tail.cuh.  Results are identical either way.
//
// tail_stream_kernel is the RESIDENT form: one launch serves the tails of a whole queue of frames as the
// fused kernel (mog_pipe.cuh) publishes their masks.  For busy masks its band CTAs pre-label their own rows
// (band_prelabel) and the labelling CTA only merges across band borders; a table that outgrows shared memory
// is labelled in a global-memory scratch area on the device.
#pragma once
#include "tail.cuh"

namespace oat {

enum { TAIL_OK = 0, TAIL_OVERFLOW = 1 };

struct TailResult {
    oat_detection det;
    int32_t status;      // TAIL_OK / TAIL_OVERFLOW
    uint32_t nodes;      // run-table entries the mask needed (diagnostic / sizing)
    uint32_t slow_groups;  // fused kernel's census: 4-pixel groups that left its fast path in this frame
    uint32_t pad;        // 1: the run table was copied from the bands' pre-labelled pool (diagnostic)
    uint32_t cyc[8];     // SM-clock stamps of the labelling CTA (diagnostic): start, ticket, extents staged, runs counted, mask staged + run table filled, merged, holes filled, end
};

struct FastArgs;
__device__ __forceinline__ void store_result(const FastArgs &a, const TailResult &r);

struct FastArgs {
    const uint32_t *in;  // threshold bits before morphology
    uint32_t *out;       // post-morphology bits (what thresh egress publishes)
    int ke, kd, R;
    BitGeom g;
    int2 *rowext;
    int2 *rowcnt;            // per row: foreground runs, candidate-background runs (counted by the CTA that made the row)
    int *bbox;               // ymin, ymax (reset by the last CTA for the next launch)
    unsigned int *ticket;    // CTA completion counter (reset likewise)
    double min_area, max_area;
    TailResult *res;
    TailResult *res_host;    // or NULL: pinned-host mirror written by the kernel itself (no D2H copy to enqueue)
    int smem_bytes;          // dynamic shared memory per CTA
    int max_comps;
    unsigned int *slow_in;   // or NULL: the fused kernel's slow-path census of this frame (read into the result, re-armed)
    int in_place_ok;         // nobody reads `out` after the labelling (no thresh egress): a mask too large to be staged beside
                             // its run table may be read -- and have its holes filled -- where it is
    // Band pre-labelling (resident tail server; NULL elsewhere): every band CTA leaves the runs of its rows -- already
    // merged inside the band, with their moment sums -- in the frame's pool, so the labelling CTA of a busy mask only
    // has to merge across band borders (see band_prelabel / tail_label_phase).
    uint2 *pool_runs;               // [pool_cap] x = start | end << 16, y = band-local root | band-local row << 16
    uint4 *pool_sums;               // [pool_cap] foreground runs: 2*m00, 6*m10, 6*m01 of the cells the run owns (holes not filled)
    uint4 *pool_agg;                // [pool_cap] at band-local foreground roots: the same, summed over the root's runs in the band
    uint4 *band_hdr;                // [bands] base, foreground runs, candidate runs, ok
    unsigned int *pool_alloc;       // entries handed out for this frame (re-armed by the CTA that finishes the frame)
    int pool_cap;
    int force_pre;                  // test switch (OAT_B200_FORCE_PRELABEL): use the pool even for a mask that could be staged
};

// the frame's result: device copy (epilogues, replays read it) and, when asked for, the pinned-host mirror the
// collecting thread reads after the frame's `done` event -- written over PCIe by the kernel, which is shorter
// than a separate D2H copy on the stream (one API call and one DMA start-up less per frame)
__device__ __forceinline__ void store_result(const FastArgs &a, const TailResult &r)
{
    *a.res = r;
    if (a.res_host) {
        *a.res_host = r;
        __threadfence_system();
    }
}


// ---- shared-memory union-find (parents only ever decrease) ----------------------------------
__device__ __forceinline__ uint32_t suf_find(volatile uint32_t *P, uint32_t x)
{
    uint32_t p = P[x];
    while (p != x) {
        const uint32_t gp = P[p];
        if (gp != p) atomicMin(const_cast<uint32_t *>(P) + x, gp);
        x = p;
        p = gp;
    }
    return x;
}
__device__ __forceinline__ void suf_union(volatile uint32_t *P, uint32_t a, uint32_t b)
{
    for (;;) {
        a = suf_find(P, a);
        b = suf_find(P, b);
        if (a == b) return;
        if (a < b) {
            const uint32_t t = a;
            a = b;
            b = t;
        }
        const uint32_t old = atomicMin(const_cast<uint32_t *>(P) + a, b);
        if (old == a) return;
        a = old;
    }
}

// union of run `id` with a run `k` of the row above it (k < id).  The first neighbour above is linked directly -- a
// run that nobody has merged into anything yet is its own root, and one atomicMin makes k its parent; no chain is
// walked, which is what a union costs for the 30-odd rows of a blob inside a band -- unless somebody was faster:
// then the displaced parent is united with k the long way.
__device__ __forceinline__ void suf_link_up(volatile uint32_t *P, uint32_t id, uint32_t k, bool first)
{
    if (!first) {
        suf_union(P, id, k);
        return;
    }
    const uint32_t old = atomicMin(const_cast<uint32_t *>(P) + id, k);
    if (old != id) suf_union(P, old, k);
}

// Region view of a bit image staged in shared memory: rows [y0, y1], words [j0, j1]; 0 elsewhere.
struct RegionView {
    const uint32_t *w;
    int y0, y1, j0, j1, wd;
    bool vol;  // the words live in global memory and are updated with atomics (holes): never read them through L1
    __device__ __forceinline__ uint32_t at(int y, int j) const
    {
        if (y < y0 || y > y1 || j < j0 || j > j1) return 0u;
        const uint32_t *p = w + (y - y0) * wd + (j - j0);
        return vol ? *reinterpret_cast<const volatile uint32_t *>(p) : *p;
    }
};
// candidate background of an inner row: background between the row's first and last foreground pixel
__device__ __forceinline__ uint32_t fast_candw(const RegionView &m, const BitGeom &g, int y, int j, int2 e)
{
    if (j < m.j0 || j > m.j1) return 0u;
    return ~m.at(y, j) & g.valid_mask(j) & range_mask(j, e.x, e.y);
}
// any background pixel in columns [s, e] of row y ([s, e] lies inside the region's columns)
__device__ __forceinline__ bool fast_any_bg(const RegionView &m, const BitGeom &g, int y, int s, int e)
{
    if (y < m.y0 || y > m.y1) return true;
    for (int j = s >> 5; j <= (e >> 5); ++j)
        if (~m.at(y, j) & g.valid_mask(j) & range_mask(j, s, e)) return true;
    return false;
}

// Exact 2x2-cell sums (2*m00, 6*m10, 6*m01) of the cells owned by pixels [s, e] of row y -- a cell is owned by its
// top-left pixel, or by its top-right one when the top-left is background -- over the words jfirst, jfirst + jstep, ...
// of the run.  at(yy, j) is word j of row yy of the (hole-filled) mask, 0 outside the image.
template <class At>
__device__ __forceinline__ void run_cell_sums(At at, const int y, const int s, const int e, const int jfirst, const int jstep,
                                              unsigned long long &t00, unsigned long long &t10, unsigned long long &t01)
{
    for (int j = jfirst; j <= (e >> 5); j += jstep) {
        const uint32_t T = at(y, j), Bw = at(y + 1, j);
        const uint64_t T64 = (uint64_t)(at(y, j - 1) >> 31) | ((uint64_t)T << 1) | ((uint64_t)(at(y, j + 1) & 1u) << 33);
        const uint64_t B64 = (uint64_t)(at(y + 1, j - 1) >> 31) | ((uint64_t)Bw << 1) | ((uint64_t)(at(y + 1, j + 1) & 1u) << 33);
        const uint64_t tl = T64, tr = T64 >> 1, bl = B64, br = B64 >> 1;
        const uint32_t sm = range_mask(j, s, e);
        const uint32_t F = (uint32_t)((tl & tr & bl & br) >> 1) & sm;
        const uint32_t ma = (uint32_t)((tl & ~tr & bl & br) >> 1) & sm;
        const uint32_t mb = (uint32_t)((tl & tr & ~bl & br) >> 1) & sm;
        const uint32_t mc = (uint32_t)((tl & tr & bl & ~br) >> 1) & sm;
        const uint32_t md = (uint32_t)(~tl & tr & bl & br) & sm;  // owner = top-right pixel, cell x = owner - 1
        if (!(F | ma | mb | mc | md)) continue;
        const uint32_t xb = 32u * (uint32_t)j;
        const uint32_t nFc = __popc(F), na = __popc(ma), nb = __popc(mb), nc = __popc(mc), nd = __popc(md);
        t00 += 2u * nFc + na + nb + nc + nd;
        t10 += (unsigned long long)(6u * (sum_pos(F) + nFc * xb) + 3u * nFc + 3u * (sum_pos(ma) + na * xb) + na +
                                    3u * (sum_pos(mb) + nb * xb) + 2u * nb + 3u * (sum_pos(mc) + nc * xb) + nc +
                                    3u * (sum_pos(md) + nd * xb) - nd);
        const uint32_t yy = (uint32_t)y;
        t01 += (unsigned long long)(nFc * (6u * yy + 3u) + (na + nd) * (3u * yy + 2u) + (nb + nc) * (3u * yy + 1u));
    }
}

// The labelling phase, run by one CTA after every CTA has published its rows.  Everything it
// touches repeatedly -- the mask's bounding region, row extents, run table, parents, accumulators
// -- is staged in shared memory first (one coalesced pass over L2), so no step chases pointers
// through global memory.
// `work` is the phase's working memory (a.smem_bytes of it): the CTA's shared memory, or -- second attempt of the
// resident tail server for a mask whose tables do not fit there -- a per-CTA scratch area in global memory (same
// code, every hop an L2 round trip instead of ~30 cycles).  Returns false WITHOUT publishing a result if the tables
// do not fit and this is not the final attempt; with `final` it publishes TAIL_OVERFLOW and the host replays the frame.
template <bool PRE>
__device__ bool tail_label_phase(const FastArgs &a, uint8_t *smem, const int ymin, const int ymax, const uint32_t a_t0,
                                 const uint32_t slow_groups, const bool final)
{
    const BitGeom g = a.g;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
    __shared__ uint32_t s_nF, s_nB, s_ncomp, s_fail;
    __shared__ int s_xmin, s_xmax;
    __shared__ unsigned long long s_best;
    TailResult r;
    r.det.position_valid = 0;
    r.det.n_components = 0;
    r.det.x = r.det.y = r.det.area = 0.0;
    r.status = TAIL_OK;
    r.nodes = 0;
    r.slow_groups = slow_groups;
    r.pad = 0;
    for (int i = 0; i < 8; ++i) r.cyc[i] = 0;
    r.cyc[0] = a_t0;
    r.cyc[1] = (uint32_t)clock64();
    if (ymax < ymin) {  // empty mask
        if (tid == 0) store_result(a, r);
        return true;
    }
    const int H = ymax - ymin + 1;
    const int C = a.max_comps;
    // ---- 0. stage row extents, then the bounding region of the mask ---------------------------
    // layout: ext[H] (int2) | offF[H+1] offB[H+1] (u32) | acc[3*C] (u64) | [M[H*Wd] (u32): the staged mask] | node arrays
    size_t used = 0;
    int2 *ext = reinterpret_cast<int2 *>(smem);
    used += (size_t)H * 8;
    uint32_t *offF = reinterpret_cast<uint32_t *>(smem + used);
    uint32_t *offB = offF + (H + 1);
    used += (size_t)2 * (H + 1) * 4;
    used = (used + 7) & ~(size_t)7;
    unsigned long long *acc = reinterpret_cast<unsigned long long *>(smem + used);
    used += (size_t)3 * C * 8;
    if (used + 1024 > (size_t)a.smem_bytes) {
        if (final && tid == 0) {
            r.status = TAIL_OVERFLOW;
            store_result(a, r);
        }
        return false;
    }
    if (tid == 0) {
        s_ncomp = 0;
        s_fail = 0;
        s_best = 0ull;
        s_xmin = INT_MAX;
        s_xmax = -1;
    }
    __syncthreads();
    {
        int xmin = INT_MAX, xmax = -1;
        for (int yb = tid; yb < H; yb += 4 * NT) {  // (four rows per thread in flight: a tall region is a few L2 round trips, not H / NT)
            int2 e[4], cnt[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int y = yb + u * NT;
                e[u] = y < H ? __ldcg(a.rowext + ymin + y) : make_int2(INT_MAX, -1);
                cnt[u] = y < H ? __ldcg(a.rowcnt + ymin + y) : make_int2(0, 0);  // run counts, made by the CTA that produced the row
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int y = yb + u * NT;
                if (y >= H) continue;
                ext[y] = e[u];
                offF[y] = (uint32_t)cnt[u].x;
                offB[y] = (uint32_t)cnt[u].y;
                if (e[u].y >= 0) {
                    xmin = min(xmin, e[u].x);
                    xmax = max(xmax, e[u].y);
                }
            }
        }
        xmin = __reduce_min_sync(0xffffffffu, xmin);
        xmax = __reduce_max_sync(0xffffffffu, xmax);
        if (lane == 0) {
            atomicMin(&s_xmin, xmin);
            atomicMax(&s_xmax, xmax);
        }
    }
    __syncthreads();
    const int jmin = s_xmin >> 5, jmax = s_xmax >> 5, Wd = jmax - jmin + 1;
    const size_t RW = (size_t)H * Wd;
    r.cyc[2] = (uint32_t)clock64();
    // ---- 1. exclusive scan over rows: the run counts size the table.  Chunks of 32 rows, a warp each; then the chunk
    //         totals (one warp); then every row adds its chunk's base.  (More than 128 chunks: one warp walks them.) ----
    __shared__ uint32_t s_ctF[128], s_ctB[128];
    const int nchunks = (H + 31) >> 5;
    if (nchunks <= 128) {
        for (int ck = warp; ck < nchunks; ck += nwarps) {
            const int y = 32 * ck + lane;
            const uint32_t vf = y < H ? offF[y] : 0u, vb = y < H ? offB[y] : 0u;
            uint32_t sf = vf, sb = vb;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t tf = __shfl_up_sync(0xffffffffu, sf, d), tb = __shfl_up_sync(0xffffffffu, sb, d);
                if (lane >= d) {
                    sf += tf;
                    sb += tb;
                }
            }
            if (y < H) {
                offF[y] = sf - vf;
                offB[y] = sb - vb;
            }
            if (lane == 31) {
                s_ctF[ck] = sf;
                s_ctB[ck] = sb;
            }
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t baseF = 0, baseB = 0;
            for (int c0 = 0; c0 < nchunks; c0 += 32) {
                const int ck = c0 + lane;
                const uint32_t vf = ck < nchunks ? s_ctF[ck] : 0u, vb = ck < nchunks ? s_ctB[ck] : 0u;
                uint32_t sf = vf, sb = vb;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t tf = __shfl_up_sync(0xffffffffu, sf, d), tb = __shfl_up_sync(0xffffffffu, sb, d);
                    if (lane >= d) {
                        sf += tf;
                        sb += tb;
                    }
                }
                if (ck < nchunks) {
                    s_ctF[ck] = baseF + sf - vf;
                    s_ctB[ck] = baseB + sb - vb;
                }
                baseF += __shfl_sync(0xffffffffu, sf, 31);
                baseB += __shfl_sync(0xffffffffu, sb, 31);
            }
            if (lane == 0) {
                offF[H] = baseF;
                offB[H] = baseB;
                s_nF = baseF;
                s_nB = baseB;
            }
        }
        __syncthreads();
        for (int y = tid; y < H; y += NT) {
            offF[y] += s_ctF[y >> 5];
            offB[y] += s_ctB[y >> 5];
        }
    } else if (warp == 0) {
        uint32_t baseF = 0, baseB = 0;
        for (int y0 = 0; y0 < H; y0 += 32) {
            const int y = y0 + lane;
            const uint32_t vf = y < H ? offF[y] : 0u, vb = y < H ? offB[y] : 0u;
            uint32_t sf = vf, sb = vb;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t tf = __shfl_up_sync(0xffffffffu, sf, d), tb = __shfl_up_sync(0xffffffffu, sb, d);
                if (lane >= d) {
                    sf += tf;
                    sb += tb;
                }
            }
            if (y < H) {
                offF[y] = baseF + sf - vf;
                offB[y] = baseB + sb - vb;
            }
            baseF += __shfl_sync(0xffffffffu, sf, 31);
            baseB += __shfl_sync(0xffffffffu, sb, 31);
        }
        if (lane == 0) {
            offF[H] = baseF;
            offB[H] = baseB;
            s_nF = baseF;
            s_nB = baseB;
        }
    }
    __syncthreads();
    const uint32_t nF = s_nF, nB = s_nB;
    r.nodes = nF + nB + 1;
    r.cyc[3] = (uint32_t)clock64();
    // ---- 2. working memory: the run table (start,end,row,comp: u16 x4 + parent u32 = 12 B per run) and, if it fits
    //         beside it, a staged copy of the mask's bounding region.  A region too large for that (many blobs spread
    //         over the frame: the region is the frame) is read where the bands left it -- streaming reads through L2,
    //         while everything the union-find chases stays in this memory
    const size_t need_nodes = ((size_t)nF + nB + 1) * 12 + 16;
    bool staged = used + RW * 4 + need_nodes <= (size_t)a.smem_bytes;
    // Bands that left their runs pre-labelled in the frame's pool (band_prelabel): used for the masks that cannot be
    // staged -- the table is then COPIED (coalesced, independent loads) instead of being extracted from the in-place mask
    // one dependent L2 round trip per row, only the rows on band borders are merged here, and the moments come from
    // the bands' sums except where a filled hole changes them.
    const int Rb = a.R, b0 = ymin / Rb, b1 = ymax / Rb, nbd = b1 - b0 + 1;
    auto ri = [&](int b) -> int { return min(max(b * Rb - ymin, 0), H); };  // index into offF / offB of band b's first row
    uint32_t *bbase = nullptr, *dirty = nullptr, *lroot = nullptr, *holeb = nullptr;
    const uint32_t nFw = (nF + 32u) / 32u, nBw = (nB + 32u) / 32u;
    bool pre = false;
    // (a table that lives in global memory -- the second attempt -- prefers the copy even if a staged mask would fit)
    if (PRE && (!staged || __isGlobal(smem) != 0 || a.force_pre) && a.pool_runs != nullptr && a.in_place_ok) {
        const size_t extra = (((size_t)nbd + 2 * (size_t)nFw + nBw) * 4 + 15) & ~(size_t)15;
        if (used + extra + need_nodes <= (size_t)a.smem_bytes) {
            bbase = reinterpret_cast<uint32_t *>(smem + used);
            dirty = bbase + nbd;
            lroot = dirty + nFw;
            holeb = lroot + nFw;
            int ok = 1;
            for (int b = b0 + tid; b <= b1; b += NT) {
                const uint32_t f = offF[ri(b + 1)] - offF[ri(b)], cnd = offB[ri(b + 1)] - offB[ri(b)];
                uint32_t base = 0;
                if (f + cnd) {
                    const uint4 hd = __ldcg(a.band_hdr + b);
                    if (!hd.w || hd.y != f || hd.z != cnd) ok = 0;
                    base = hd.x;
                }
                bbase[b - b0] = base;
            }
            for (uint32_t i = tid; i < 2 * nFw + nBw; i += NT) dirty[i] = 0u;
            pre = __syncthreads_and(ok) != 0;
            if (pre) {
                used += extra;
                staged = false;
                r.pad = 1u;
            }
        }
    }
    if (!staged && !(a.in_place_ok && used + need_nodes <= (size_t)a.smem_bytes)) {
        if (final && tid == 0) {
            r.status = TAIL_OVERFLOW;
            store_result(a, r);
        }
        return false;
    }
    // pool index of run k (0-based: foreground runs 0..nF-1, then the candidates), and its band
    auto pool_index = [&](uint32_t k, int &b, uint32_t &li, uint32_t &nFb) -> uint32_t {
        const bool isB = k >= nF;
        const uint32_t kk = isB ? k - nF : k;
        const uint32_t *off = isB ? offB : offF;
        int lo = b0, hi = b1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (off[ri(mid)] <= kk) lo = mid; else hi = mid - 1;
        }
        b = lo;
        nFb = offF[ri(b + 1)] - offF[ri(b)];
        li = isB ? nFb + (kk - offB[ri(b)]) : kk - offF[ri(b)];
        return bbase[b - b0] + li;
    };
    uint32_t *Ms = staged ? reinterpret_cast<uint32_t *>(smem + used) : a.out + (size_t)ymin * g.wpr + jmin;
    uint32_t *Gs = Ms;  // holes are OR-ed into the same words once the merges (the last readers of the bare mask) are done
    const int Ws = staged ? Wd : g.wpr;  // row stride of Ms / Gs in words
    if (staged) used += RW * 4;
    const int N = (int)(((size_t)a.smem_bytes - used) / 12);
    uint32_t *parent = reinterpret_cast<uint32_t *>(smem + used);
    uint16_t *nstart = reinterpret_cast<uint16_t *>(parent + N);
    uint16_t *nend = nstart + N;
    uint16_t *nrow = nend + N;
    uint16_t *ncomp = nrow + N;
    if (staged)
        for (size_t t = tid; t < RW; t += NT) {
            const int y = (int)(t / Wd), j = jmin + (int)(t % Wd);
            Ms[t] = __ldcg(a.out + (size_t)(ymin + y) * g.wpr + j);
        }
    for (int c = tid; c < 3 * C; c += NT) acc[c] = 0ull;
    __syncthreads();
    const bool vol = !staged || __isGlobal(smem) != 0;
    RegionView M{Ms, ymin, ymax, jmin, jmax, Ws, vol};
    RegionView G{Gs, ymin, ymax, jmin, jmax, Ws, vol};
    // node ids: 0 = EXT, 1..nF foreground runs (raster order), nF+1..nF+nB candidate background runs
    // ---- 3. fill the run table: copied from the bands' pool if they pre-labelled it ...
    if (pre) {
        constexpr int U = 4;  // runs per thread and round: their loads are in flight together
        for (uint32_t k0 = tid; k0 < nF + nB; k0 += U * NT) {
            uint2 ent[U];
            int bb[U];
            uint32_t li[U], nFb[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t k = k0 + u * NT;
                bb[u] = b0;
                li[u] = nFb[u] = 0;
                ent[u] = make_uint2(0u, 0u);
                if (k < nF + nB) ent[u] = __ldcg(a.pool_runs + pool_index(k, bb[u], li[u], nFb[u]));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t k = k0 + u * NT;
                if (k >= nF + nB) continue;
                const uint32_t id = 1 + k, lr = ent[u].y & 0xffffu;
                nstart[id] = (uint16_t)(ent[u].x & 0xffffu);
                nend[id] = (uint16_t)(ent[u].x >> 16);
                nrow[id] = (uint16_t)(bb[u] * Rb + (int)(ent[u].y >> 16));
                // band-local ids (0 = exterior, foreground runs, candidates) -> ids of this table; order is preserved,
                // so parents still only point downwards
                const uint32_t fb = offF[ri(bb[u])], cb = offB[ri(bb[u])];
                parent[id] = lr == 0u ? 0u : (lr <= nFb[u] ? fb + lr : nF + cb + (lr - nFb[u]));
                if (k < nF && lr == li[u] + 1u) atomicOr(lroot + (k >> 5), 1u << (k & 31u));  // a band-local root: carries the band's sums
            }
        }
    } else {
    // ----    ... else extracted from the mask: G lanes per row (G = region width in words rounded up to a power of two,
    //         so a narrow blob puts 32/G rows in flight per warp), one word per lane, segmented warp scans ----
        int G = 1;
        while (G < Wd && G < 32) G <<= 1;
        const int rpw = 32 / G, gl = lane & (G - 1), gi = lane / G;
        for (int yb = ymin + warp * rpw; yb <= ymax; yb += nwarps * rpw) {
            const int y = yb + gi;
            const bool row_ok = y <= ymax;
            const int2 e = row_ok ? ext[y - ymin] : make_int2(INT_MAX, -1);
            const bool live = row_ok && e.y >= 0;
            const bool inner = (y > 0) && (y < g.rows - 1);
            uint32_t bsF = 0, beF = 0, bsB = 0, beB = 0;
            if (live) {
                bsF = beF = 1 + offF[y - ymin];
                bsB = beB = 1 + nF + offB[y - ymin];
            }
            // G == 32: one row per warp, as many 32-word chunks as the row needs (warp-uniform trip count);
            // G < 32: every row of the region fits one chunk
            const int nchunk = (G == 32) ? (live ? ((e.y >> 5) - (e.x >> 5)) / 32 + 1 : 0) : 1;
            for (int ch = 0; ch < nchunk; ++ch) {
                const int j = (live ? (e.x >> 5) : 0) + ch * 32 + gl;
                uint32_t sF = 0, eF = 0, sB = 0, eB = 0;
                if (live && j <= (e.y >> 5)) {
                    const uint32_t w = M.at(y, j), pw = M.at(y, j - 1), nw = M.at(y, j + 1);
                    sF = w & ~((w << 1) | (pw >> 31));
                    eF = w & ~((w >> 1) | (nw << 31));
                    if (inner) {
                        const uint32_t c = fast_candw(M, g, y, j, e), pc = fast_candw(M, g, y, j - 1, e),
                                       nc = fast_candw(M, g, y, j + 1, e);
                        sB = c & ~((c << 1) | (pc >> 31));
                        eB = c & ~((c >> 1) | (nc << 31));
                    }
                }
                // exclusive scans of the four counts inside each group of G lanes (packed 2 x 16 bit)
                const uint32_t cnt1 = (uint32_t)__popc(sF) | ((uint32_t)__popc(eF) << 16);
                const uint32_t cnt2 = (uint32_t)__popc(sB) | ((uint32_t)__popc(eB) << 16);
                uint32_t x1 = cnt1, x2 = cnt2;
                for (int d = 1; d < G; d <<= 1) {
                    const uint32_t t1 = __shfl_up_sync(0xffffffffu, x1, d, G), t2 = __shfl_up_sync(0xffffffffu, x2, d, G);
                    if (gl >= d) {
                        x1 += t1;
                        x2 += t2;
                    }
                }
                const uint32_t tot1 = __shfl_sync(0xffffffffu, x1, G - 1, G), tot2 = __shfl_sync(0xffffffffu, x2, G - 1, G);
                x1 -= cnt1;
                x2 -= cnt2;
                uint32_t k;
                k = bsF + (x1 & 0xffffu);
                for (uint32_t m = sF; m; m &= m - 1, ++k) {
                    nstart[k] = (uint16_t)(32 * j + __ffs(m) - 1);
                    nrow[k] = (uint16_t)y;
                    parent[k] = k;
                }
                k = beF + (x1 >> 16);
                for (uint32_t m = eF; m; m &= m - 1, ++k) nend[k] = (uint16_t)(32 * j + __ffs(m) - 1);
                k = bsB + (x2 & 0xffffu);
                for (uint32_t m = sB; m; m &= m - 1, ++k) {
                    nstart[k] = (uint16_t)(32 * j + __ffs(m) - 1);
                    nrow[k] = (uint16_t)y;
                    parent[k] = k;
                }
                k = beB + (x2 >> 16);
                for (uint32_t m = eB; m; m &= m - 1, ++k) nend[k] = (uint16_t)(32 * j + __ffs(m) - 1);
                bsF += tot1 & 0xffffu;
                beF += tot1 >> 16;
                bsB += tot2 & 0xffffu;
                beB += tot2 >> 16;
            }
        }
    }
    if (tid == 0) parent[0] = 0;
    __syncthreads();
    volatile uint32_t *P = parent;
    r.cyc[4] = (uint32_t)clock64();
    // ---- 4. vertical merges (thread per run) --------------------------------------------------
    auto merge_run = [&](const uint32_t id) {
        const int y = nrow[id], s = nstart[id], e = nend[id];
        const bool fg = id <= nF;
        if (fg) {
            if (y > ymin) {
                const uint32_t lo0 = 1 + offF[y - 1 - ymin], hi0 = 1 + offF[y - ymin];
                uint32_t lo = lo0, hi = hi0;  // first run of the previous row with end >= s - 1
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if ((int)nend[mid] < s - 1) lo = mid + 1; else hi = mid;
                }
                for (uint32_t k = lo; k < hi0 && (int)nstart[k] <= e + 1; ++k) suf_link_up(P, id, k, k == lo);
            }
        } else {
            // 4-connected to candidate background of the previous row
            if (y > ymin) {
                const uint32_t lo0 = 1 + nF + offB[y - 1 - ymin], hi0 = 1 + nF + offB[y - ymin];
                uint32_t lo = lo0, hi = hi0;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if ((int)nend[mid] < s) lo = mid + 1; else hi = mid;
                }
                for (uint32_t k = lo; k < hi0 && (int)nstart[k] <= e; ++k) suf_link_up(P, id, k, k == lo);
            }
            // exterior if it touches row-exterior background above or below: in rows 0 and
            // rows-1 every background pixel is exterior; elsewhere the background outside the
            // row's foreground extent is.
            bool isext = false;
#pragma unroll
            for (int dy = -1; dy <= 1; dy += 2) {
                const int yy = y + dy;  // candidate rows are inner rows, so 0 <= yy <= rows-1
                if (yy == 0 || yy == g.rows - 1) {
                    isext |= fast_any_bg(M, g, yy, s, e);
                } else if (yy < ymin || yy > ymax) {
                    isext = true;
                } else {
                    const int2 ee = ext[yy - ymin];
                    isext |= (ee.y < 0) || (s < ee.x) || (e > ee.y);
                }
            }
            if (isext) suf_union(P, id, 0u);
        }
    };
    if (pre) {
        // rows inside a band were merged by the band: only its first and its last row have a neighbour it did not see
        // (a warp per such row: the work is where the rows are, not spread over every run of the table)
        for (int q = tid >> 3; q < 2 * nbd; q += NT >> 3) {  // (eight lanes per row: a row of a busy mask has about that many runs)
            const int b = b0 + (q >> 1);
            const int y = (q & 1) ? min(b * Rb + Rb - 1, g.rows - 1) : b * Rb;
            if (y < ymin || y > ymax || ((q & 1) && Rb == 1)) continue;
            const uint32_t f0 = offF[y - ymin], nf = offF[y - ymin + 1] - f0, c0 = offB[y - ymin], nc = offB[y - ymin + 1] - c0;
            for (uint32_t i = (uint32_t)(tid & 7); i < nf + nc; i += 8) merge_run(i < nf ? 1 + f0 + i : 1 + nF + c0 + (i - nf));
        }
    } else {
        for (uint32_t id = 1 + tid; id <= nF + nB; id += NT) merge_run(id);
    }
    __syncthreads();
    r.cyc[5] = (uint32_t)clock64();
    // ---- 5. holes join the foreground beside them; G = foreground + holes -----------------------
    for (uint32_t id = 1 + nF + tid; id <= nF + nB; id += NT) {
        if (suf_find(P, id) == 0u) continue;  // exterior
        const int y = nrow[id], s = nstart[id], e = nend[id];
        // a candidate run lies strictly inside the row's foreground extent: runs end at s-1 and start at e+1
        const uint32_t lo0 = 1 + offF[y - ymin], hi0 = 1 + offF[y - ymin + 1];
        uint32_t lo = lo0, hi = hi0;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if ((int)nend[mid] < s - 1) lo = mid + 1; else hi = mid;
        }
        suf_union(P, id, lo);  // nend[lo] == s - 1
        if (lo + 1 < hi0) suf_union(P, id, lo + 1);  // nstart[lo + 1] == e + 1
        if (pre) {
            atomicOr(holeb + ((id - 1 - nF) >> 5), 1u << ((id - 1 - nF) & 31u));
            // the filled pixels change the cells of the runs beside them (this row) and above them (cells of row y - 1
            // reach into row y): those runs' sums are recomputed from the filled mask
            atomicOr(dirty + ((lo - 1) >> 5), 1u << ((lo - 1) & 31u));
            if (lo + 1 < hi0) atomicOr(dirty + (lo >> 5), 1u << (lo & 31u));
            if (y > ymin) {
                const uint32_t plo0 = 1 + offF[y - 1 - ymin], phi0 = 1 + offF[y - ymin];
                uint32_t pl = plo0, ph = phi0;
                while (pl < ph) {
                    const uint32_t mid = (pl + ph) >> 1;
                    if ((int)nend[mid] < s - 1) pl = mid + 1; else ph = mid;
                }
                for (uint32_t k = pl; k < phi0 && (int)nstart[k] <= e + 1; ++k) atomicOr(dirty + ((k - 1) >> 5), 1u << ((k - 1) & 31u));
            }
        }
        for (int j = s >> 5; j <= (e >> 5); ++j) atomicOr(Gs + (size_t)(y - ymin) * Ws + (j - jmin), range_mask(j, s, e));
    }
    __syncthreads();
    r.cyc[6] = (uint32_t)clock64();
    // ---- 6. contours = roots among the foreground runs: give them compact accumulator slots ----
    for (uint32_t id = 1 + tid; id <= nF; id += NT) {
        if (P[id] == id) {
            const uint32_t c = atomicAdd(&s_ncomp, 1u);
            ncomp[id] = (uint16_t)c;
            if (c >= (uint32_t)C) s_fail = 1u;
        }
    }
    __syncthreads();
    if (s_fail) {
        if (final && tid == 0) {
            r.status = TAIL_OVERFLOW;
            store_result(a, r);
        }
        return false;
    }
    // ---- 7. exact 2x2-cell moments (cells are owned by their top row).  Four lanes share a run, each
    //         taking every fourth word of it; a warp whose lanes all feed the same contour (the usual
    //         single-blob mask) folds its sums with shuffles and issues ONE set of atomics ------------
    auto G_at = [&](int yy, int j) { return G.at(yy, j); };
    if (!pre) {
        const uint32_t nwork = (nF + nB) * 4u;
        const uint32_t nround = (nwork + NT - 1) / NT;
        for (uint32_t rd = 0; rd < nround; ++rd) {
            const uint32_t wk = rd * NT + tid;
            uint32_t root = 0u;
            unsigned long long t00 = 0, t10 = 0, t01 = 0;
            if (wk < nwork) {
                const uint32_t id = 1u + (wk >> 2);
                root = suf_find(P, id);
                const int y = nrow[id], s = nstart[id], e = nend[id];
                if (root != 0u && y < g.rows - 1)  // exterior background owns nothing; the last row owns no cells
                    run_cell_sums(G_at, y, s, e, (s >> 5) + (int)(wk & 3u), 4, t00, t10, t01);
            }
            const bool have = (t00 | t10 | t01) != 0ull;
            // one contour for the whole warp?  (lanes with nothing to add do not count)
            const unsigned contrib = __ballot_sync(0xffffffffu, have);
            if (contrib == 0u) continue;
            const uint32_t root0 = __shfl_sync(0xffffffffu, root, __ffs(contrib) - 1);
            if (__all_sync(0xffffffffu, !have || root == root0)) {
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    t00 += __shfl_xor_sync(0xffffffffu, t00, d);
                    t10 += __shfl_xor_sync(0xffffffffu, t10, d);
                    t01 += __shfl_xor_sync(0xffffffffu, t01, d);
                }
                if (lane == 0) {
                    const uint32_t c = ncomp[root0];
                    atomicAdd(acc + c, t00);
                    atomicAdd(acc + C + c, t10);
                    atomicAdd(acc + 2 * C + c, t01);
                }
            } else if (have) {
                const uint32_t c = ncomp[root];
                atomicAdd(acc + c, t00);
                atomicAdd(acc + C + c, t10);
                atomicAdd(acc + 2 * C + c, t01);
            }
        }
    } else {
        // Pre-labelled: the foreground runs arrive with their sums, one set per band-local root (c).  Computed here,
        // from the filled mask, is only what a band could not know: (a) the last row of every band -- its cells reach
        // into the next band --, (b) the filled holes themselves and the runs beside or above them, whose band sums
        // are REPLACED (subtracted modulo 2^64, the accumulators being plain sums).  Each list is walked densely:
        // the L2 round trips of a mask that is read in place are paid by the runs that need them, not by every round
        // of a loop over the whole table.
        auto add_to = [&](const uint32_t id, const unsigned long long t00, const unsigned long long t10, const unsigned long long t01) {
            if ((t00 | t10 | t01) == 0ull) return;
            const uint32_t c = ncomp[suf_find(P, id)];
            atomicAdd(acc + c, t00);
            atomicAdd(acc + C + c, t10);
            atomicAdd(acc + 2 * C + c, t01);
        };
        // (a) eight lanes per band-last row, four lanes per run, folded with shuffles (warp-uniform trip counts)
        for (int bq = b0 + (warp << 2); bq <= b1; bq += nwarps << 2) {
            const int b = bq + (lane >> 3);
            const int y = b * Rb + Rb - 1;
            const bool rowok = b <= b1 && y >= ymin && y <= ymax && y < g.rows - 1;
            const uint32_t lo = rowok ? 1 + offF[y - ymin] : 0u, hi = rowok ? 1 + offF[y - ymin + 1] : 0u;
            const uint32_t trips = __reduce_max_sync(0xffffffffu, (hi - lo + 1u) >> 1);
            for (uint32_t it = 0; it < trips; ++it) {
                const uint32_t id = lo + 2u * it + (uint32_t)((lane & 7) >> 2);
                unsigned long long t00 = 0, t10 = 0, t01 = 0;
                if (id < hi) run_cell_sums(G_at, y, (int)nstart[id], (int)nend[id], (nstart[id] >> 5) + (lane & 3), 4, t00, t10, t01);
#pragma unroll
                for (int d = 1; d <= 2; d <<= 1) {
                    t00 += __shfl_xor_sync(0xffffffffu, t00, d);
                    t10 += __shfl_xor_sync(0xffffffffu, t10, d);
                    t01 += __shfl_xor_sync(0xffffffffu, t01, d);
                }
                if (id < hi && (lane & 3) == 0) add_to(id, t00, t10, t01);
            }
        }
        // (b) filled holes, and the foreground runs whose cells they changed (not those of (a): nothing was summed for them)
        for (uint32_t w = tid; w < nBw + nFw; w += NT) {
            const bool hole = w < nBw;
            for (uint32_t m = hole ? holeb[w] : dirty[w - nBw]; m; m &= m - 1) {
                const uint32_t k = 32u * (hole ? w : w - nBw) + (uint32_t)__ffs(m) - 1u;
                const uint32_t id = 1u + (hole ? nF + k : k);
                const int y = nrow[id], s = nstart[id], e = nend[id];
                if (y >= g.rows - 1 || (!hole && (y % Rb == Rb - 1))) continue;
                unsigned long long t00 = 0, t10 = 0, t01 = 0;
                run_cell_sums(G_at, y, s, e, s >> 5, 1, t00, t10, t01);
                if (!hole) {
                    int b;
                    uint32_t li, nFb;
                    const uint4 old = __ldcg(a.pool_sums + pool_index(k, b, li, nFb));
                    t00 -= old.x;
                    t10 -= old.y;
                    t01 -= old.z;
                }
                add_to(id, t00, t10, t01);
            }
        }
        // (c) the bands' sums
        for (uint32_t w = tid; w < nFw; w += NT) {
            for (uint32_t m = lroot[w]; m; m &= m - 1) {
                const uint32_t k = 32u * w + (uint32_t)__ffs(m) - 1u;
                int b;
                uint32_t li, nFb;
                const uint4 v = __ldcg(a.pool_agg + pool_index(k, b, li, nFb));
                add_to(1u + k, (unsigned long long)v.x, (unsigned long long)v.y, (unsigned long long)v.z);
            }
        }
    }
    __syncthreads();
    // ---- 8. select: largest area in [min, max); ties go to the raster-last contour ---------------
    for (uint32_t id = 1 + tid; id <= nF; id += NT) {
        if (P[id] != id) continue;
        const unsigned long long s00 = acc[ncomp[id]];
        const double area = 0.5 * (double)s00;
        if (area >= a.min_area && area < a.max_area && s00 > 0ull) atomicMax(&s_best, (s00 << 32) | (unsigned long long)id);
    }
    __syncthreads();
    if (tid == 0) {
        r.det.n_components = (int32_t)s_ncomp;
        r.cyc[7] = (uint32_t)clock64();
        if (s_best) {
            const uint32_t id = (uint32_t)(s_best & 0xffffffffull);
            const uint32_t c = ncomp[id];
            const double m00 = (double)acc[c] * 0.5;
            const double m10 = (double)acc[C + c] * 0.16666666666666666666666666666667;
            const double m01 = (double)acc[2 * C + c] * 0.16666666666666666666666666666667;
            r.det.position_valid = 1;
            r.det.x = m10 / m00;
            r.det.y = m01 / m00;
            r.det.area = m00;
        }
        store_result(a, r);
    }
    return true;
}

// Band pre-labelling: what the labelling CTA of a busy mask would otherwise do alone, one dependent L2 round trip at
// a time (the mask does not fit its shared memory), done here by every band CTA for its own rows while they are still
// in shared memory: the run table of the band's rows (same node definition as tail_label_phase: foreground runs, and
// candidate background runs of inner rows), the vertical merges between rows of the band, the exterior test against
// neighbour rows inside the band, and the moment sums of every foreground run whose cells lie inside the band (all
// rows but the band's last), per run and summed per band-local root.  What crosses a band border is left to the
// labelling CTA.  `rowsA` = the band's post-morphology rows (row y at rowsA + (y - i0) * wpr, padding bits clear).
__device__ __forceinline__ void band_prelabel(const FastArgs &a, const int band, const uint32_t *rowsA, const int i0, const int y0,
                                              const int y1, const int2 *sb_ext, const int2 *sb_cnt, uint8_t *work, const size_t work_bytes)
{
    const BitGeom g = a.g;
    const int wpr = g.wpr, rows = g.rows, nr = y1 - y0 + 1;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
    __shared__ uint32_t sb_offF[33], sb_offB[33], sb_base, sb_ok;
    if (warp == 0) {  // exclusive scan of the rows' run counts (nr <= 32)
        const uint32_t vf = lane < nr ? (uint32_t)sb_cnt[lane].x : 0u, vb = lane < nr ? (uint32_t)sb_cnt[lane].y : 0u;
        uint32_t sf = vf, sb = vb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t tf = __shfl_up_sync(0xffffffffu, sf, d), tb = __shfl_up_sync(0xffffffffu, sb, d);
            if (lane >= d) {
                sf += tf;
                sb += tb;
            }
        }
        sb_offF[lane] = sf - vf;
        sb_offB[lane] = sb - vb;
        if (lane == 31) {
            sb_offF[32] = sf;
            sb_offB[32] = sb;
        }
    }
    __syncthreads();
    const uint32_t nF = sb_offF[32], nB = sb_offB[32], n = nF + nB;
    if (n == 0) return;  // (the labelling CTA only asks bands that have runs)
    // table: parent u32 (x2) | start, end u16 | row u8 | band-local sums of the roots 3 x u32 (native shared-memory atomics;
    // a band-local contour of a frame of up to 4096 x 4096 stays below 2^32: 6 * 32 rows * sum of x, cols * 32 * (6 y + 3))
    const size_t need = (size_t)(n + 1) * (4 + 4 + 2 + 2 + 1) + 16 + (size_t)(nF + 1) * 12;
    if (need > work_bytes || g.cols > 4096 || g.rows > 4096 || n >= 65535u) {  // (16-bit band-local ids)
        if (tid == 0) a.band_hdr[band] = make_uint4(0u, nF, nB, 0u);
        return;
    }
    uint32_t *agg = reinterpret_cast<uint32_t *>(work);  // [3][nF + 1]
    uint32_t *par = agg + (size_t)3 * (nF + 1);
    uint32_t *par2 = par + n + 1;  // second copy for the pointer jumping
    uint16_t *st = reinterpret_cast<uint16_t *>(par2 + n + 1);
    uint16_t *en = st + n + 1;
    uint8_t *rw = reinterpret_cast<uint8_t *>(en + n + 1);
    for (uint32_t i = tid; i < 3u * (nF + 1u); i += NT) agg[i] = 0u;
    auto word = [&](int y, int j) -> uint32_t { return (j < 0 || j >= wpr) ? 0u : rowsA[(size_t)(y - i0) * wpr + j]; };
    // ---- run table: one warp per row, 32 words per chunk, segmented warp scans (tail_label_phase step 3) ----
    for (int r = warp; r < nr; r += nwarps) {
        const int y = y0 + r;
        const int2 e = sb_ext[r];
        if (e.y < 0) continue;
        const bool inner = (y > 0) && (y < rows - 1);
        uint32_t bsF = 1 + sb_offF[r], beF = bsF, bsB = 1 + nF + sb_offB[r], beB = bsB;
        auto cand = [&](int j) -> uint32_t { return (j < 0 || j >= wpr) ? 0u : (~word(y, j) & g.valid_mask(j) & range_mask(j, e.x, e.y)); };
        const int nchunk = ((e.y >> 5) - (e.x >> 5)) / 32 + 1;
        for (int ch = 0; ch < nchunk; ++ch) {
            const int j = (e.x >> 5) + ch * 32 + lane;
            uint32_t sF = 0, eF = 0, sB = 0, eB = 0;
            if (j <= (e.y >> 5)) {
                const uint32_t w = word(y, j), pw = word(y, j - 1), nw = word(y, j + 1);
                sF = w & ~((w << 1) | (pw >> 31));
                eF = w & ~((w >> 1) | (nw << 31));
                if (inner) {
                    const uint32_t c = cand(j), pc = cand(j - 1), nc = cand(j + 1);
                    sB = c & ~((c << 1) | (pc >> 31));
                    eB = c & ~((c >> 1) | (nc << 31));
                }
            }
            const uint32_t cnt1 = (uint32_t)__popc(sF) | ((uint32_t)__popc(eF) << 16);
            const uint32_t cnt2 = (uint32_t)__popc(sB) | ((uint32_t)__popc(eB) << 16);
            uint32_t x1 = cnt1, x2 = cnt2;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t1 = __shfl_up_sync(0xffffffffu, x1, d), t2 = __shfl_up_sync(0xffffffffu, x2, d);
                if (lane >= d) {
                    x1 += t1;
                    x2 += t2;
                }
            }
            const uint32_t tot1 = __shfl_sync(0xffffffffu, x1, 31), tot2 = __shfl_sync(0xffffffffu, x2, 31);
            x1 -= cnt1;
            x2 -= cnt2;
            uint32_t k;
            k = bsF + (x1 & 0xffffu);
            for (uint32_t m = sF; m; m &= m - 1, ++k) {
                st[k] = (uint16_t)(32 * j + __ffs(m) - 1);
                rw[k] = (uint8_t)r;
                par[k] = k;
            }
            k = beF + (x1 >> 16);
            for (uint32_t m = eF; m; m &= m - 1, ++k) en[k] = (uint16_t)(32 * j + __ffs(m) - 1);
            k = bsB + (x2 & 0xffffu);
            for (uint32_t m = sB; m; m &= m - 1, ++k) {
                st[k] = (uint16_t)(32 * j + __ffs(m) - 1);
                rw[k] = (uint8_t)r;
                par[k] = k;
            }
            k = beB + (x2 >> 16);
            for (uint32_t m = eB; m; m &= m - 1, ++k) en[k] = (uint16_t)(32 * j + __ffs(m) - 1);
            bsF += tot1 & 0xffffu;
            beF += tot1 >> 16;
            bsB += tot2 & 0xffffu;
            beB += tot2 >> 16;
        }
    }
    if (tid == 0) {
        par[0] = 0;
        const unsigned int base = atomicAdd(a.pool_alloc, n);
        sb_base = base;
        sb_ok = (base + n <= (unsigned int)a.pool_cap) ? 1u : 0u;
        a.band_hdr[band] = make_uint4(base, nF, nB, sb_ok);
    }
    __syncthreads();
    if (!sb_ok) return;
    volatile uint32_t *P = par;
    // ---- vertical merges between rows of the band; exterior test against neighbour rows inside the band ----
    for (uint32_t id = 1 + tid; id <= n; id += NT) {
        const int r = rw[id], y = y0 + r, s = st[id], e = en[id];
        if (id <= nF) {
            if (r > 0) {
                const uint32_t lo0 = 1 + sb_offF[r - 1], hi0 = 1 + sb_offF[r];
                uint32_t lo = lo0, hi = hi0;  // first run of the previous row with end >= s - 1
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if ((int)en[mid] < s - 1) lo = mid + 1; else hi = mid;
                }
                for (uint32_t k = lo; k < hi0 && (int)st[k] <= e + 1; ++k) suf_link_up(P, id, k, k == lo);
            }
        } else {
            if (r > 0) {
                const uint32_t lo0 = 1 + nF + sb_offB[r - 1], hi0 = 1 + nF + sb_offB[r];
                uint32_t lo = lo0, hi = hi0;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if ((int)en[mid] < s) lo = mid + 1; else hi = mid;
                }
                for (uint32_t k = lo; k < hi0 && (int)st[k] <= e; ++k) suf_link_up(P, id, k, k == lo);
            }
            bool isext = false;
#pragma unroll
            for (int dy = -1; dy <= 1; dy += 2) {
                const int rr = r + dy, yy = y + dy;
                if (rr < 0 || rr >= nr) continue;  // a neighbour row in another band: the labelling CTA's business
                if (yy == 0 || yy == rows - 1) {
                    for (int j = s >> 5; j <= (e >> 5); ++j)
                        if (~word(yy, j) & g.valid_mask(j) & range_mask(j, s, e)) isext = true;
                } else {
                    const int2 ee = sb_ext[rr];
                    isext |= (ee.y < 0) || (s < ee.x) || (e > ee.y);
                }
            }
            if (isext) suf_union(P, id, 0u);
        }
    }
    __syncthreads();
    // ---- publish: runs with their band-local root, per-run moment sums, per-root sums ----
    // the direct links left chains (a run -> the run above it -> ...): rounds of pointer jumping bring a depth of
    // 32 (64) down to 1; whatever unions of several branches left deeper is walked by the plain finds below
    // (between two copies of the parent array: no round reads what it writes; six rounds, so the result is back in `par`)
    for (int it = 0; it < 6; ++it) {
        const uint32_t *src = (it & 1) ? par2 : par;
        uint32_t *dst = (it & 1) ? par : par2;
        for (uint32_t id = tid; id <= n; id += NT) dst[id] = src[src[id]];
        __syncthreads();
    }
    const unsigned int base = sb_base;
    for (uint32_t id = 1 + tid; id <= n; id += NT) {
        uint32_t root = id;
        for (uint32_t pp = P[root]; pp != root; pp = P[root]) root = pp;
        const int r = rw[id], y = y0 + r, s = st[id], e = en[id];
        a.pool_runs[base + id - 1] = make_uint2((uint32_t)s | ((uint32_t)e << 16), root | ((uint32_t)r << 16));
        if (id <= nF) {
            unsigned long long t00 = 0, t10 = 0, t01 = 0;
            if (r < nr - 1) {  // (the cells of the band's last row need the next band's first row)
                run_cell_sums(word, y, s, e, s >> 5, 1, t00, t10, t01);
                // (a foreground run's root is a foreground run: ids of foreground runs are the smaller ones)
                atomicAdd(agg + root, (uint32_t)t00);
                atomicAdd(agg + (nF + 1) + root, (uint32_t)t10);
                atomicAdd(agg + 2 * (size_t)(nF + 1) + root, (uint32_t)t01);
            }
            a.pool_sums[base + id - 1] = make_uint4((uint32_t)t00, (uint32_t)t10, (uint32_t)t01, 0u);
        }
    }
    __syncthreads();
    for (uint32_t id = 1 + tid; id <= nF; id += NT)
        if (par[id] == id) a.pool_agg[base + id - 1] = make_uint4(agg[id], agg[(nF + 1) + id], agg[2 * (size_t)(nF + 1) + id], 0u);
}

// Horizontal pass of a k-wide rectangle (k <= 32) on word `cur` of a row, `prev` / `next` being its neighbours (0 outside
// the row): bit x of the result = OR over the window [x - k/2, x - k/2 + k - 1] (hpass_word<true>), by doubling on the
// 96 bits prev:cur:next instead of one funnel shift per window position.  Erosion is the same on the complement (the
// caller passes ~row & valid, out-of-row words 0, and complements the result): samples outside the image are ignored.
__device__ __forceinline__ uint32_t hwin_or(const uint32_t prev, const uint32_t cur, const uint32_t next, const int k)
{
    unsigned long long lo = (unsigned long long)prev | ((unsigned long long)cur << 32), hi = next;  // bits 0..63, 64..95
    auto shr_or = [&](int sft) {  // R |= R >> sft, 0 < sft < 64
        const unsigned long long nlo = (lo >> sft) | (hi << (64 - sft)), nhi = hi >> sft;
        lo |= nlo;
        hi |= nhi;
    };
    int w = 1;
    while (2 * w <= k) {
        shr_or(w);
        w <<= 1;
    }
    if (w < k) shr_or(k - w);
    // R bit p = OR of bits p .. p + k - 1; result bit x = R bit (32 + x - k/2)
    const int sft = 32 - k / 2;
    return (uint32_t)((lo >> sft) | (hi << (64 - sft)));
}

// One band of R rows: [erode] -> [dilate] -> publish the rows (mask, extents, run counts, vertical bounding range).
// Dynamic shared memory: 2 x (R + ke - 1 + kd - 1) rows of the mask.
template <bool PRE>
__device__ __forceinline__ void tail_band(const FastArgs &a, const int band, uint32_t *sm)
{
    const BitGeom g = a.g;
    const int wpr = g.wpr, rows = g.rows;
    const int y0 = band * a.R, y1 = min(y0 + a.R, rows) - 1;
    const int ae = a.ke / 2, ad = a.kd / 2;
    const int e0 = max(a.kd > 0 ? y0 - ad : y0, 0), e1 = min(a.kd > 0 ? y1 - ad + a.kd - 1 : y1, rows - 1);
    const int i0 = max(a.ke > 0 ? e0 - ae : e0, 0), i1 = min(a.ke > 0 ? e1 - ae + a.ke - 1 : e1, rows - 1);
    const int nin = i1 - i0 + 1;
    uint32_t *A = sm, *B = sm + (size_t)nin * wpr;
    uint32_t any = 0;
    // Rows that are a whole number of 16-byte vectors and kernels of at most 32 pixels (every practical case) take
    // 128-bit loads issued together, the doubling window of hwin_or and 128-bit vertical passes; index arithmetic by
    // multiplication (a band has a few hundred vectors).  Anything else: word by word, hpass_word.
    const bool vec = ((wpr & 3) == 0) && a.ke <= 32 && a.kd <= 32;
    const int wpr4 = wpr >> 2;
    // t / wpr4 == umulhi(t, magic4) for the t of a band (t * wpr4 < 2^32); a row of one vector needs no division
    const uint32_t magic4 = (vec && wpr4 > 1) ? (uint32_t)((((unsigned long long)1 << 32) + (uint32_t)wpr4 - 1u) / (uint32_t)wpr4) : 0u;
    auto div4 = [&](int t) -> int { return wpr4 > 1 ? (int)__umulhi((uint32_t)t, magic4) : t; };
    const uint32_t lastmask = g.valid_mask(wpr - 1);
    if (vec) {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.in + (size_t)i0 * wpr);
        uint4 *A4 = reinterpret_cast<uint4 *>(A);
        const int n4 = nin * wpr4;
        constexpr int LU = PRE ? 4 : 2;  // vectors in flight per thread (the per-frame kernel runs six CTAs per SM on 40 registers)
        for (int t0 = threadIdx.x; t0 < n4; t0 += LU * blockDim.x) {
            uint4 v[LU];
#pragma unroll
            for (int u = 0; u < LU; ++u) {
                const int t = t0 + u * blockDim.x;
                v[u] = t < n4 ? __ldcg(src + t) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < LU; ++u) {
                const int t = t0 + u * blockDim.x;
                if (t >= n4) continue;
                const int r = div4(t);
                if (t - r * wpr4 == wpr4 - 1) v[u].w &= lastmask;  // (padding bits of the row's last word)
                A4[t] = v[u];
                any |= v[u].x | v[u].y | v[u].z | v[u].w;
            }
        }
    } else {
        for (int t = threadIdx.x; t < nin * wpr; t += blockDim.x) {
            const int r = t / wpr, j = t % wpr;
            const uint32_t w = __ldcg(a.in + (size_t)(i0 + r) * wpr + j) & g.valid_mask(j);
            A[t] = w;
            any |= w;
        }
    }
    // Most bands of a tracking mask are empty: nothing to erode or dilate, no extents, no runs.  (An empty input
    // band gives an empty output band for dilation, and for erosion a fortiori.)
    const bool band_empty = __syncthreads_or(any != 0u) == 0;
    if (band_empty) {
        for (int t = threadIdx.x; t < (y1 - y0 + 1) * wpr; t += blockDim.x) a.out[(size_t)y0 * wpr + t] = 0u;
        for (int y = y0 + (int)threadIdx.x; y <= y1; y += blockDim.x) {
            a.rowext[y] = make_int2(INT_MAX, -1);
            a.rowcnt[y] = make_int2(0, 0);
        }
        return;
    }
    // horizontal pass over rows [ra, rb] (indices into the staged rows) of src -> dst, 4 words per thread
    auto hpass_vec = [&](const uint32_t *srcw, uint32_t *dstw, const int ra, const int rb, const int k, const bool dilate) {
        const int n4 = (rb - ra + 1) * wpr4;
        for (int t = threadIdx.x; t < n4; t += blockDim.x) {
            const int rr = div4(t), q = t - rr * wpr4;
            const uint32_t *row = srcw + (size_t)(ra + rr) * wpr;
            uint4 c = *reinterpret_cast<const uint4 *>(row + 4 * q);
            uint32_t p = q > 0 ? row[4 * q - 1] : 0u, nx = q < wpr4 - 1 ? row[4 * q + 4] : 0u;
            if (!dilate) {  // complement inside the image; outside stays 0 (ignored)
                c.x = ~c.x;
                c.y = ~c.y;
                c.z = ~c.z;
                c.w = ~c.w;
                if (q == wpr4 - 1) c.w &= lastmask;
                p = q > 0 ? ~p : 0u;
                nx = q < wpr4 - 1 ? (~nx & (4 * q + 4 == wpr - 1 ? lastmask : 0xffffffffu)) : 0u;
            }
            uint4 o;
            o.x = hwin_or(p, c.x, c.y, k);
            o.y = hwin_or(c.x, c.y, c.z, k);
            o.z = hwin_or(c.y, c.z, c.w, k);
            o.w = hwin_or(c.z, c.w, nx, k);
            if (!dilate) {
                o.x = ~o.x;
                o.y = ~o.y;
                o.z = ~o.z;
                o.w = ~o.w;
            }
            if (q == wpr4 - 1) o.w &= lastmask;
            *reinterpret_cast<uint4 *>(dstw + (size_t)(ra + rr) * wpr + 4 * q) = o;
        }
    };
    // vertical pass: rows [ya, yb] of dst = AND / OR of the window rows of src (rows outside the image are ignored)
    auto vpass_vec = [&](const uint32_t *srcw, uint32_t *dstw, const int ya, const int yb, const int k, const bool dilate) {
        const int n4 = (yb - ya + 1) * wpr4, an = k / 2;
        for (int t = threadIdx.x; t < n4; t += blockDim.x) {
            const int rr = div4(t), q = t - rr * wpr4, y = ya + rr;
            const int v0 = max(y - an, 0), v1 = min(y - an + k - 1, rows - 1);
            const uint4 *col = reinterpret_cast<const uint4 *>(srcw + (size_t)(v0 - i0) * wpr) + q;
            uint4 acc = *col;
            for (int yy = v0 + 1; yy <= v1; ++yy) {
                col += wpr4;
                const uint4 v = *col;
                if (dilate) {
                    acc.x |= v.x;
                    acc.y |= v.y;
                    acc.z |= v.z;
                    acc.w |= v.w;
                } else {
                    acc.x &= v.x;
                    acc.y &= v.y;
                    acc.z &= v.z;
                    acc.w &= v.w;
                }
            }
            *reinterpret_cast<uint4 *>(dstw + (size_t)(y - i0) * wpr + 4 * q) = acc;
        }
    };
    if (a.ke > 0) {
        if (vec) {
            hpass_vec(A, B, 0, nin - 1, a.ke, false);
            __syncthreads();
            vpass_vec(B, A, e0, e1, a.ke, false);
        } else {
            for (int t = threadIdx.x; t < nin * wpr; t += blockDim.x) B[t] = hpass_word<false>(A + (t / wpr) * wpr, t % wpr, g, a.ke);
            __syncthreads();
            const int ne = e1 - e0 + 1;
            for (int t = threadIdx.x; t < ne * wpr; t += blockDim.x) {
                const int y = e0 + t / wpr, j = t % wpr;
                const int v0 = max(y - ae, 0), v1 = min(y - ae + a.ke - 1, rows - 1);
                uint32_t accw = 0xffffffffu;
                for (int yy = v0; yy <= v1; ++yy) accw &= B[(yy - i0) * wpr + j];
                A[(y - i0) * wpr + j] = accw & g.valid_mask(j);
            }
        }
        __syncthreads();
    }
    if (a.kd > 0) {
        if (vec) {
            hpass_vec(A, B, e0 - i0, e1 - i0, a.kd, true);
            __syncthreads();
            vpass_vec(B, A, y0, y1, a.kd, true);
        } else {
            const int ne = e1 - e0 + 1;
            for (int t = threadIdx.x; t < ne * wpr; t += blockDim.x) {
                const int r = e0 - i0 + t / wpr;
                B[r * wpr + t % wpr] = hpass_word<true>(A + r * wpr, t % wpr, g, a.kd);
            }
            __syncthreads();
            const int no = y1 - y0 + 1;
            for (int t = threadIdx.x; t < no * wpr; t += blockDim.x) {
                const int y = y0 + t / wpr, j = t % wpr;
                const int v0 = max(y - ad, 0), v1 = min(y - ad + a.kd - 1, rows - 1);
                uint32_t accw = 0u;
                for (int yy = v0; yy <= v1; ++yy) accw |= B[(yy - i0) * wpr + j];
                A[(y - i0) * wpr + j] = accw;
            }
        }
        __syncthreads();
    }
    // publish rows: mask, hole-fill seed, extents, vertical bounding range
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    __shared__ int2 sb_ext[32], sb_cnt[32];  // the band's rows again, for band_prelabel
    const bool prelabel = PRE && a.pool_runs != nullptr && a.R <= 32;
    int bymin = INT_MAX, bymax = -1;
    for (int y = y0 + warp; y <= y1; y += nwarps) {
        const uint32_t *row = A + (size_t)(y - i0) * wpr;
        int xmin = INT_MAX, xmax = -1;
        for (int j = lane; j < wpr; j += 32) {
            const uint32_t w = row[j];
            a.out[(size_t)y * wpr + j] = w;
            if (w) {
                xmin = min(xmin, 32 * j + __ffs(w) - 1);
                xmax = max(xmax, 32 * j + 31 - __clz(w));
            }
        }
        xmin = __reduce_min_sync(0xffffffffu, xmin);
        xmax = __reduce_max_sync(0xffffffffu, xmax);
        // run counts of the row (the labelling CTA only has to scan them): foreground runs, and candidate
        // background runs = background between the row's first and last foreground pixel (inner rows only)
        int cf = 0, cb = 0;
        if (xmax >= 0) {
            const bool inner = (y > 0) && (y < rows - 1);
            for (int j = (xmin >> 5) + lane; j <= (xmax >> 5); j += 32) {
                const uint32_t w = row[j], pw = j > 0 ? row[j - 1] : 0u;
                cf += __popc(w & ~((w << 1) | (pw >> 31)));
                if (inner) {
                    const uint32_t c = ~w & g.valid_mask(j) & range_mask(j, xmin, xmax);
                    const uint32_t pc = j > 0 ? (~row[j - 1] & g.valid_mask(j - 1) & range_mask(j - 1, xmin, xmax)) : 0u;
                    cb += __popc(c & ~((c << 1) | (pc >> 31)));
                }
            }
        }
        cf = __reduce_add_sync(0xffffffffu, cf);
        cb = __reduce_add_sync(0xffffffffu, cb);
        if (lane == 0) {
            a.rowext[y] = make_int2(xmin, xmax);
            a.rowcnt[y] = make_int2(cf, cb);
            if (prelabel) {
                sb_ext[y - y0] = make_int2(xmin, xmax);
                sb_cnt[y - y0] = make_int2(cf, cb);
            }
        }
        if (xmax >= 0) {
            bymin = min(bymin, y);
            bymax = max(bymax, y);
        }
    }
    if (lane == 0 && bymax >= 0) {
        atomicMin(a.bbox, bymin);
        atomicMax(a.bbox + 1, bymax);
    }
    if (prelabel) {
        __syncthreads();
        const size_t stage = ((size_t)2 * nin * wpr * sizeof(uint32_t) + 15) & ~(size_t)15;
        if ((size_t)a.smem_bytes > stage)
            band_prelabel(a, band, A, i0, y0, y1, sb_ext, sb_cnt, reinterpret_cast<uint8_t *>(sm) + stage, (size_t)a.smem_bytes - stage);
        else if (threadIdx.x == 0)
            a.band_hdr[band] = make_uint4(0u, 0u, 0u, 0u);
    }
}

// One launch: [erode] -> [dilate] -> row extents on bands of R rows (all CTAs), then the last CTA
// labels.  Dynamic shared memory: max(morphology staging, a.smem_bytes).
__global__ void __launch_bounds__(256, 6) tail_fast_kernel(const FastArgs a)
{
    extern __shared__ __align__(16) uint32_t sm[];
    const uint32_t t_start = (uint32_t)clock64();
    tail_band<false>(a, (int)blockIdx.x, sm);
    __shared__ bool s_last;
    __shared__ int s_ymin, s_ymax;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
        if (s_last) {
            __threadfence();
            s_ymin = atomicExch(a.bbox, INT_MAX);   // read + reset for the next launch
            s_ymax = atomicExch(a.bbox + 1, -1);
            *a.ticket = 0u;
        }
    }
    __syncthreads();
    if (!s_last) return;
    __shared__ uint32_t s_slow;
    if (threadIdx.x == 0) s_slow = a.slow_in ? atomicExch(a.slow_in, 0u) : 0u;  // the fused kernel's census of this frame, re-armed
    __syncthreads();
    tail_label_phase<false>(a, reinterpret_cast<uint8_t *>(sm), s_ymin, s_ymax, t_start, s_slow, true);
}

// ---------------------------------------------------------------------------------------------------
// The resident tail server: ONE launch serves the detect tails of a whole queue of frames (the clip the
// resident fused kernel, mog_pipe.cuh, is working through).  Per frame: wait until the fused kernel has
// published every tile of the frame's threshold mask (done_count, release/acquire at GPU scope), then the
// CTAs claim the frame's bands dynamically; the CTA that finishes the frame's last band labels it (~15 us)
// while the others are already on the next frame.  Nothing is re-armed inside the launch -- a CTA that
// comes back late from labelling finds a frame's band counter exhausted and moves on -- the per-launch
// counters are zeroed by a memset the host puts in front of the launch.
// ---------------------------------------------------------------------------------------------------
struct TailFrame {
    FastArgs a;                       // a.ticket is not used here
    const unsigned int *done_count;   // or NULL: the frame's mask is ready when the launch starts
    unsigned int done_target;         // the frame is complete when (int)(*done_count - done_target) >= 0
    unsigned int pad;
};

// Up to TAIL_INLINE_DESCS frames travel in the kernel parameters (no copy to enqueue in front of the launch).
constexpr int TAIL_INLINE_DESCS = 32;
struct TailQueue {
    TailFrame inl[TAIL_INLINE_DESCS];
};
static_assert(sizeof(TailQueue) + 128 <= 32764, "kernel parameters are limited to 32764 bytes (CUDA 12.1+, sm_70+)");
static_assert(sizeof(TailFrame) % 4 == 0, "TailFrame is copied word by word");

#ifndef TAIL_STREAM_MINBLOCKS_CFG
#define TAIL_STREAM_MINBLOCKS_CFG 2  // its CTAs sit beside ONE fused CTA on a reserved SM: registers are not what limits them
#endif
__global__ void __launch_bounds__(256, TAIL_STREAM_MINBLOCKS_CFG) tail_stream_kernel(const TailFrame *frames /* or NULL: q.inl */,
                                                             const __grid_constant__ TailQueue q, const int nframes,
                                                             unsigned int *band_ctr /* [nframes], zero when the launch starts */,
                                                             unsigned int *band_done /* [nframes], likewise */,
                                                             unsigned int *exit_ticket /* likewise: the last CTA to leave re-arms all three */,
                                                             uint8_t *scratch /* or NULL: gridDim.x areas of scratch_bytes */,
                                                             const int scratch_bytes, const int scratch_comps)
{
    extern __shared__ __align__(16) uint32_t sm[];
    __shared__ TailFrame s_tf;
    __shared__ int s_band;
    __shared__ bool s_last;
    __shared__ int s_ymin, s_ymax;
    for (int f = 0; f < nframes; ++f) {
        __syncthreads();  // the previous frame's use of s_tf / s_band is over
        {
            const TailFrame *src = frames ? frames + f : &q.inl[f];
            for (int i = threadIdx.x; i < (int)(sizeof(TailFrame) / 4); i += blockDim.x)
                reinterpret_cast<uint32_t *>(&s_tf)[i] = reinterpret_cast<const uint32_t *>(src)[i];
        }
        __syncthreads();
        const int nbands = (s_tf.a.g.rows + s_tf.a.R - 1) / s_tf.a.R;
        if (threadIdx.x == 0) {
            int b = (int)atomicAdd(band_ctr + f, 1u);
            if (b < nbands && s_tf.done_count) {
                // bounded: a fused kernel that died must surface as an error, not hang the GPU
                unsigned spins = 0;
                unsigned int v;
                for (;;) {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(s_tf.done_count) : "memory");
                    if ((int)(v - s_tf.done_target) >= 0) break;
                    __nanosleep(256);
                    if (++spins > (1u << 24)) __trap();
                }
            }
            s_band = b;
        }
        __syncthreads();
        while (s_band < nbands) {
            const uint32_t t_start = (uint32_t)clock64();
            tail_band<true>(s_tf.a, s_band, sm);
            __threadfence();
            __syncthreads();
            // two independent round trips, two threads: this band is done / the next band to work on.  (If this CTA turns
            // out to be the frame's finisher every band had been claimed already: its draw is >= nbands, as it must be.)
            if (threadIdx.x == 0) {
                s_last = (atomicAdd(band_done + f, 1u) == (unsigned)nbands - 1u);
                if (s_last) __threadfence();
            }
            if (threadIdx.x == 32) s_band = (int)atomicAdd(band_ctr + f, 1u);
            __syncthreads();
            if (s_last) {
                __shared__ uint32_t s_slow;
                // the frame's bounding rows and the fused kernel's census, read and re-armed for the slot's next frame
                // (likewise independent: one thread each)
                if (threadIdx.x == 0) s_ymin = atomicExch(s_tf.a.bbox, INT_MAX);
                if (threadIdx.x == 32) s_ymax = atomicExch(s_tf.a.bbox + 1, -1);
                if (threadIdx.x == 64) s_slow = s_tf.a.slow_in ? atomicExch(s_tf.a.slow_in, 0u) : 0u;
                if (threadIdx.x == 96 && s_tf.a.pool_alloc) *s_tf.a.pool_alloc = 0u;  // (every band of the frame has drawn its share)
                __syncthreads();
                // label in shared memory; a mask whose tables do not fit (many blobs, noise) is labelled again by this
                // CTA in its global-memory scratch area -- slower, but on the device and beside the other CTAs' frames,
                // instead of a round trip through the host
                if (!tail_label_phase<true>(s_tf.a, reinterpret_cast<uint8_t *>(sm), s_ymin, s_ymax, t_start, s_slow, scratch == nullptr)) {
                    __syncthreads();
                    FastArgs big = s_tf.a;
                    big.smem_bytes = scratch_bytes;
                    big.max_comps = scratch_comps;
                    tail_label_phase<true>(big, scratch + (size_t)blockIdx.x * (size_t)scratch_bytes, s_ymin, s_ymax, t_start, s_slow, true);
                }
                __syncthreads();
                if (threadIdx.x == 0) s_last = false;
            }
        }
    }
    // the last CTA to leave re-arms the launch's counters for the next user of this half of the ring (a launch the host
    // starts only after this one has completed), so no memset has to be enqueued in front of a launch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(exit_ticket, 1u) == gridDim.x - 1u);
    }
    __syncthreads();
    if (s_last) {
        for (int i = threadIdx.x; i < nframes; i += blockDim.x) {
            band_ctr[i] = 0u;
            band_done[i] = 0u;
        }
        if (threadIdx.x == 0) *exit_ticket = 0u;
    }
}

}  // namespace oat
