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
// kernel report TAIL_OVERFLOW; the host then replays the frame through the unbounded
// global-memory path of tail.cuh.  Results are identical either way.
#pragma once
#include "tail.cuh"

namespace oat {

enum { TAIL_OK = 0, TAIL_OVERFLOW = 1 };

struct TailResult {
    oat_detection det;
    int32_t status;      // TAIL_OK / TAIL_OVERFLOW
    uint32_t nodes;      // run-table entries the mask needed (diagnostic / sizing)
    uint32_t slow_groups;  // fused kernel's census: 4-pixel groups that left its fast path in this frame
    uint32_t pad;
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

// The labelling phase, run by one CTA after every CTA has published its rows.  Everything it
// touches repeatedly -- the mask's bounding region, row extents, run table, parents, accumulators
// -- is staged in shared memory first (one coalesced pass over L2), so no step chases pointers
// through global memory.
// `work` is the phase's working memory (a.smem_bytes of it): the CTA's shared memory, or -- second attempt of the
// resident tail server for a mask whose tables do not fit there -- a per-CTA scratch area in global memory (same
// code, every hop an L2 round trip instead of ~30 cycles).  Returns false WITHOUT publishing a result if the tables
// do not fit and this is not the final attempt; with `final` it publishes TAIL_OVERFLOW and the host replays the frame.
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
        for (int y = tid; y < H; y += NT) {
            const int2 e = __ldcg(a.rowext + ymin + y);
            ext[y] = e;
            const int2 cnt = __ldcg(a.rowcnt + ymin + y);  // run counts, made by the CTA that produced the row
            offF[y] = (uint32_t)cnt.x;
            offB[y] = (uint32_t)cnt.y;
            if (e.y >= 0) {
                xmin = min(xmin, e.x);
                xmax = max(xmax, e.y);
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
    // ---- 1. exclusive scan over rows (one warp, chunked): the run counts size the table --------
    if (warp == 0) {
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
    const bool staged = used + RW * 4 + need_nodes <= (size_t)a.smem_bytes;
    if (!staged && !(a.in_place_ok && used + need_nodes <= (size_t)a.smem_bytes)) {
        if (final && tid == 0) {
            r.status = TAIL_OVERFLOW;
            store_result(a, r);
        }
        return false;
    }
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
    // ---- 3. fill the run table: G lanes per row (G = region width in words rounded up to a power of two,
    //         so a narrow blob puts 32/G rows in flight per warp), one word per lane, segmented warp scans ----
    {
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
    for (uint32_t id = 1 + tid; id <= nF + nB; id += NT) {
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
                for (uint32_t k = lo; k < hi0 && (int)nstart[k] <= e + 1; ++k) suf_union(P, id, k);
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
                for (uint32_t k = lo; k < hi0 && (int)nstart[k] <= e; ++k) suf_union(P, id, k);
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
    {
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
                if (root != 0u && y < g.rows - 1) {  // exterior background owns nothing; the last row owns no cells
                    for (int j = (s >> 5) + (int)(wk & 3u); j <= (e >> 5); j += 4) {
                        const uint32_t T = G.at(y, j), Bw = G.at(y + 1, j);
                        const uint64_t T64 = (uint64_t)(G.at(y, j - 1) >> 31) | ((uint64_t)T << 1) | ((uint64_t)(G.at(y, j + 1) & 1u) << 33);
                        const uint64_t B64 = (uint64_t)(G.at(y + 1, j - 1) >> 31) | ((uint64_t)Bw << 1) |
                                             ((uint64_t)(G.at(y + 1, j + 1) & 1u) << 33);
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

// One band of R rows: [erode] -> [dilate] -> publish the rows (mask, extents, run counts, vertical bounding range).
// Dynamic shared memory: 2 x (R + ke - 1 + kd - 1) rows of the mask.
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
    for (int t = threadIdx.x; t < nin * wpr; t += blockDim.x) {
        const int r = t / wpr, j = t % wpr;
        const uint32_t w = __ldcg(a.in + (size_t)(i0 + r) * wpr + j) & g.valid_mask(j);
        A[t] = w;
        any |= w;
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
    if (a.ke > 0) {
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
        __syncthreads();
    }
    if (a.kd > 0) {
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
        __syncthreads();
    }
    // publish rows: mask, hole-fill seed, extents, vertical bounding range
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
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
}

// One launch: [erode] -> [dilate] -> row extents on bands of R rows (all CTAs), then the last CTA
// labels.  Dynamic shared memory: max(morphology staging, a.smem_bytes).
__global__ void __launch_bounds__(256, 6) tail_fast_kernel(const FastArgs a)
{
    extern __shared__ __align__(16) uint32_t sm[];
    const uint32_t t_start = (uint32_t)clock64();
    tail_band(a, (int)blockIdx.x, sm);
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
    tail_label_phase(a, reinterpret_cast<uint8_t *>(sm), s_ymin, s_ymax, t_start, s_slow, true);
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

#ifndef TAIL_STREAM_MINBLOCKS_CFG
#define TAIL_STREAM_MINBLOCKS_CFG 2  // its CTAs sit beside ONE fused CTA on a reserved SM: registers are not what limits them
#endif
__global__ void __launch_bounds__(256, TAIL_STREAM_MINBLOCKS_CFG) tail_stream_kernel(const TailFrame *frames, const int nframes,
                                                             unsigned int *band_ctr /* [nframes], zeroed */,
                                                             unsigned int *band_done /* [nframes], zeroed */,
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
        for (int i = threadIdx.x; i < (int)(sizeof(TailFrame) / 4); i += blockDim.x)
            reinterpret_cast<uint32_t *>(&s_tf)[i] = reinterpret_cast<const uint32_t *>(frames + f)[i];
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
            tail_band(s_tf.a, s_band, sm);
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) {
                s_last = (atomicAdd(band_done + f, 1u) == (unsigned)nbands - 1u);
                if (s_last) {
                    __threadfence();
                    s_ymin = atomicExch(s_tf.a.bbox, INT_MAX);  // read + reset for the slot's next frame
                    s_ymax = atomicExch(s_tf.a.bbox + 1, -1);
                }
                s_band = s_last ? nbands : (int)atomicAdd(band_ctr + f, 1u);
            }
            __syncthreads();
            if (s_last) {
                __shared__ uint32_t s_slow;
                if (threadIdx.x == 0) s_slow = s_tf.a.slow_in ? atomicExch(s_tf.a.slow_in, 0u) : 0u;  // the frame's census, re-armed
                __syncthreads();
                // label in shared memory; a mask whose tables do not fit (many blobs, noise) is labelled again by this
                // CTA in its global-memory scratch area -- slower, but on the device and beside the other CTAs' frames,
                // instead of a round trip through the host
                if (!tail_label_phase(s_tf.a, reinterpret_cast<uint8_t *>(sm), s_ymin, s_ymax, t_start, s_slow, scratch == nullptr)) {
                    __syncthreads();
                    FastArgs big = s_tf.a;
                    big.smem_bytes = scratch_bytes;
                    big.max_comps = scratch_comps;
                    tail_label_phase(big, scratch + (size_t)blockIdx.x * (size_t)scratch_bytes, s_ymin, s_ymax, t_start, s_slow, true);
                }
                __syncthreads();
                if (threadIdx.x == 0) s_last = false;
            }
        }
    }
}

}  // namespace oat
