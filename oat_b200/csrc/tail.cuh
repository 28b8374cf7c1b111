// tail.cuh -- the detect tail of `posidet hsv` on a 1-bit/pixel mask:
//
//   cv::erode / cv::dilate, MORPH_RECT k x k         (src/positiondetector/HSVDetector.cpp:152-156, :253-273)
//   siftContours: findContours(RETR_EXTERNAL) -> cv::moments(contour) -> largest area in [min,max)
//                                                     (src/positiondetector/DetectorFunc.cpp:31-66)
//
// Nothing here follows OpenCV's serial border tracing.  The external contours of a mask and
// their Green's-theorem polygon moments are obtained from a union-find labelling instead
// (SURVEY.md 8(a) row a8, Appendix A10-A13):
//
//   phase A  label foreground (8-connected) and background (4-connected) word-segments in one
//            union-find forest; background touching the image border is linked to node 0 (EXT);
//   phase B  every background set that is NOT exterior is a hole: union it with the foreground
//            next to it.  Sets are now  G = component + its holes (+ islands inside the holes),
//            one per external contour, rooted at the contour's raster-first pixel;
//   moments  every 2x2 block of pixel centres with 4 pixels in G adds area 1 at its centre, with
//            exactly 3 adds area 1/2 at the mean of the three -- summed as exact integers
//            (2*m00, 6*m10, 6*m01) per root with warp-aggregated atomics;
//   select   arg-max of area over roots with min <= area < max; an area tie goes to the contour
//            whose first pixel is LAST in raster order (cv2 lists contours in reverse raster
//            order and DetectorFunc.cpp:56 compares with a strict '>').
//
// Union-find nodes are "word segments": maximal runs of equal bits inside one 32-bit word, named
// by 1 + padded linear index of their first pixel.  Everything is bit arithmetic on L2-resident
// words (259 KB per 1080p mask); the tail is latency-bound, not bandwidth-bound.
#pragma once
#include "common.cuh"

namespace oat {

struct TailBuffers {
    BitGeom g;
    int *parent;                 // [rows*wpr*32 + 1]
    unsigned long long *acc;     // 3 planes of [rows*wpr*32 + 1]: 2*m00, 6*m10, 6*m01 (valid at roots)
    size_t nnodes;               // rows*wpr*32 + 1
    uint32_t *G;                 // [rows][wpr] hole-filled mask
    int2 *rowext;                // [rows] first / last foreground x of the row (INT_MAX / -1 if none)
    unsigned long long *best;    // arg-max key  (s00 << 32 | id)
    unsigned int *count;         // number of external contours
    unsigned int *ticket;        // last-block election of the select kernel
};

// Background pixels left of the first / right of the last foreground pixel of their row, and
// the first and last image rows, reach the image border along their own row (or are on it):
// they are exterior by construction and never become union-find nodes.  Only background
// BETWEEN foreground pixels of a row ("candidates") has to be resolved by labelling.
__device__ __forceinline__ uint32_t ones_below(int n) { return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u)); }
__device__ __forceinline__ uint32_t range_mask(int j, int xmin, int xmax)
{
    return ones_below(xmax - 32 * j + 1) & ~ones_below(xmin - 32 * j);
}

// ---- bit helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t seg_mask(int s, int len) { return (len >= 32) ? 0xffffffffu : (((1u << len) - 1u) << s); }
// length of the run of ones in w starting at bit s (bit s must be set)
__device__ __forceinline__ int run_len(uint32_t w, int s)
{
    const uint32_t rest = ~(w >> s);
    const int l = rest ? (__ffs(rest) - 1) : 32;
    return min(l, 32 - s);
}
// first bit of the run of ones in w that contains bit b (bit b must be set)
__device__ __forceinline__ int run_start(uint32_t w, int b)
{
    const uint32_t x = (~w) << (31 - b);
    const int l = min(__clz(x), b + 1);
    return b - l + 1;
}
__device__ __forceinline__ int node_id(const BitGeom &g, int y, int j, int b) { return 1 + ((y * g.wpr + j) << 5) + b; }

__device__ __forceinline__ int uf_find(int *P, int x)
{
    int p = __ldcg(P + x);
    while (p != x) {
        x = p;
        p = __ldcg(P + x);
    }
    return x;
}
// find with path halving (monotone: parents only ever decrease, so atomicMin is safe)
__device__ __forceinline__ int uf_find_halve(int *P, int x)
{
    int p = __ldcg(P + x);
    while (p != x) {
        const int gp = __ldcg(P + p);
        if (gp != p) atomicMin(P + x, gp);
        x = p;
        p = gp;
    }
    return x;
}
__device__ __forceinline__ int uf_find_compress(int *P, int x)
{
    const int r = uf_find(P, x);
    if (r != x) atomicMin(P + x, r);
    return r;
}
__device__ __forceinline__ void uf_union(int *P, int a, int b)
{
    for (;;) {
        a = uf_find_halve(P, a);
        b = uf_find_halve(P, b);
        if (a == b) return;
        if (a < b) {
            const int t = a;
            a = b;
            b = t;
        }
        const int old = atomicMin(P + a, b);
        if (old == a) return;
        a = old;
    }
}

__device__ __forceinline__ uint32_t load_word(const uint32_t *bits, const BitGeom &g, int y, int j)
{
    return (y < 0 || y >= g.rows || j < 0 || j >= g.wpr) ? 0u : bits[(size_t)y * g.wpr + j];
}
// polarity view: fg = bits, bg = ~bits inside the image; 0 outside
__device__ __forceinline__ uint32_t pol_word(const uint32_t *bits, const BitGeom &g, int y, int j, bool fg)
{
    if (y < 0 || y >= g.rows || j < 0 || j >= g.wpr) return 0u;
    const uint32_t v = bits[(size_t)y * g.wpr + j];
    const uint32_t vm = g.valid_mask(j);
    return fg ? (v & vm) : (~v & vm);
}

// ---- morphology ------------------------------------------------------------------------------
// MORPH_RECT k x k with the default anchor (k/2, k/2): both ops sample the input window
// [x - k/2, x - k/2 + k - 1] (same in y); out-of-image samples are ignored.
template <bool DILATE>
__global__ void morph_h_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, BitGeom g, int k)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.rows * g.wpr) return;
    const int y = t / g.wpr, j = t % g.wpr;
    const uint32_t pad = DILATE ? 0u : 0xffffffffu;
    auto word = [&](int jj) -> uint32_t {
        if (jj < 0 || jj >= g.wpr) return pad;
        const uint32_t vm = g.valid_mask(jj);
        return (in[(size_t)y * g.wpr + jj] & vm) | (pad & ~vm);
    };
    const int a = k / 2;
    uint32_t acc = pad;
    int curq = INT_MIN;
    uint32_t lo = 0, hi = 0;
    for (int d = -a; d <= k - 1 - a; ++d) {
        const int q = (d >= 0) ? (d >> 5) : -((-d + 31) >> 5);
        const int r = d - 32 * q;
        if (q != curq) {
            lo = word(j + q);
            hi = word(j + q + 1);
            curq = q;
        }
        const uint32_t v = __funnelshift_r(lo, hi, r);  // bit b <- pixel 32j + b + d
        acc = DILATE ? (acc | v) : (acc & v);
    }
    out[t] = acc & g.valid_mask(j);
}

template <bool DILATE>
__global__ void morph_v_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, BitGeom g, int k)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.rows * g.wpr) return;
    const int y = t / g.wpr, j = t % g.wpr;
    const int a = k / 2;
    const int y0 = max(y - a, 0), y1 = min(y - a + k - 1, g.rows - 1);
    uint32_t acc = DILATE ? 0u : 0xffffffffu;
    for (int yy = y0; yy <= y1; ++yy) {
        const uint32_t v = in[(size_t)yy * g.wpr + j];
        acc = DILATE ? (acc | v) : (acc & v);
    }
    out[t] = acc & g.valid_mask(j);
}

// ---- union-find labelling -------------------------------------------------------------------
__device__ __forceinline__ uint32_t cand_word(const uint32_t *bits, const int2 *rowext, const BitGeom &g, int y, int j)
{
    if (y <= 0 || y >= g.rows - 1 || j < 0 || j >= g.wpr) return 0u;
    const int2 e = rowext[y];
    return ~bits[(size_t)y * g.wpr + j] & g.valid_mask(j) & range_mask(j, e.x, e.y);
}
// row-exterior background of word (y, j), rows 0 <= y < rows
__device__ __forceinline__ uint32_t rowext_bg_word(const uint32_t *bits, const int2 *rowext, const BitGeom &g, int y, int j)
{
    const uint32_t bgw = ~bits[(size_t)y * g.wpr + j] & g.valid_mask(j);
    if (y == 0 || y == g.rows - 1) return bgw;
    const int2 e = rowext[y];
    return bgw & ~range_mask(j, e.x, e.y);
}

// prep: [erode] -> [dilate] on a band of rows staged in shared memory, then per output row the
// foreground extent and the union-find initialisation (every foreground / candidate-background
// segment start becomes its own root, foreground accumulators are zeroed).
struct PrepArgs {
    const uint32_t *in;
    uint32_t *out;
    int ke, kd;  // erode / dilate kernel sizes (0 = off)
    int R;       // output rows per CTA
    TailBuffers tb;
};

template <bool DILATE>
__device__ __forceinline__ uint32_t hpass_word(const uint32_t *row, int j, const BitGeom &g, int k)
{
    const uint32_t pad = DILATE ? 0u : 0xffffffffu;
    auto word = [&](int jj) -> uint32_t {
        if (jj < 0 || jj >= g.wpr) return pad;
        return DILATE ? row[jj] : (row[jj] | ~g.valid_mask(jj));
    };
    const int a = k / 2;
    uint32_t acc = pad;
    int curq = INT_MIN;
    uint32_t lo = 0, hi = 0;
    for (int d = -a; d <= k - 1 - a; ++d) {
        const int q = (d >= 0) ? (d >> 5) : -((-d + 31) >> 5);
        const int r = d - 32 * q;
        if (q != curq) {
            lo = word(j + q);
            hi = word(j + q + 1);
            curq = q;
        }
        const uint32_t v = __funnelshift_r(lo, hi, r);
        acc = DILATE ? (acc | v) : (acc & v);
    }
    return acc & g.valid_mask(j);
}

__global__ void tail_prep_kernel(const PrepArgs a)
{
    extern __shared__ uint32_t sm[];
    const BitGeom g = a.tb.g;
    const int wpr = g.wpr, rows = g.rows;
    const int y0 = blockIdx.x * a.R, y1 = min(y0 + a.R, rows) - 1;
    const int ae = a.ke / 2, ad = a.kd / 2;
    // rows of the eroded image the dilate needs, rows of the input the erode needs
    const int e0 = max(a.kd > 0 ? y0 - ad : y0, 0), e1 = min(a.kd > 0 ? y1 - ad + a.kd - 1 : y1, rows - 1);
    const int i0 = max(a.ke > 0 ? e0 - ae : e0, 0), i1 = min(a.ke > 0 ? e1 - ae + a.ke - 1 : e1, rows - 1);
    const int nin = i1 - i0 + 1;
    uint32_t *A = sm, *B = sm + (size_t)nin * wpr;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.tb.parent[0] = 0;
        *a.tb.best = 0ull;
        *a.tb.count = 0u;
        *a.tb.ticket = 0u;
    }
    for (int t = threadIdx.x; t < nin * wpr; t += blockDim.x) {
        const int r = t / wpr, j = t % wpr;
        A[t] = a.in[(size_t)(i0 + r) * wpr + j] & g.valid_mask(j);
    }
    __syncthreads();
    if (a.ke > 0) {
        for (int t = threadIdx.x; t < nin * wpr; t += blockDim.x) B[t] = hpass_word<false>(A + (t / wpr) * wpr, t % wpr, g, a.ke);
        __syncthreads();
        const int ne = e1 - e0 + 1;
        for (int t = threadIdx.x; t < ne * wpr; t += blockDim.x) {
            const int y = e0 + t / wpr, j = t % wpr;
            const int v0 = max(y - ae, 0), v1 = min(y - ae + a.ke - 1, rows - 1);
            uint32_t acc = 0xffffffffu;
            for (int yy = v0; yy <= v1; ++yy) acc &= B[(yy - i0) * wpr + j];
            A[(y - i0) * wpr + j] = acc & g.valid_mask(j);
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
            uint32_t acc = 0u;
            for (int yy = v0; yy <= v1; ++yy) acc |= B[(yy - i0) * wpr + j];
            A[(y - i0) * wpr + j] = acc;
        }
        __syncthreads();
    }
    // epilogue: one warp per output row
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
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
        if (lane == 0) a.tb.rowext[y] = make_int2(xmin, xmax);
        if (xmax < 0) continue;  // no foreground in this row: nothing to initialise
        const bool inner = (y > 0) && (y < rows - 1);
        for (int j = lane; j < wpr; j += 32) {
            const uint32_t w = row[j];
            for (uint32_t st = w & ~(w << 1); st; st &= st - 1) {
                const int id = node_id(g, y, j, __ffs(st) - 1);
                a.tb.parent[id] = id;
                a.tb.acc[id] = 0ull;
                a.tb.acc[a.tb.nnodes + id] = 0ull;
                a.tb.acc[2 * a.tb.nnodes + id] = 0ull;
            }
            if (inner) {
                const uint32_t c = ~w & g.valid_mask(j) & range_mask(j, xmin, xmax);
                for (uint32_t st = c & ~(c << 1); st; st &= st - 1) {
                    const int id = node_id(g, y, j, __ffs(st) - 1);
                    a.tb.parent[id] = id;
                }
            }
        }
    }
}

// phase A: foreground 8-connected; candidate background 4-connected, linked to EXT (node 0)
// wherever it touches row-exterior background above or below.
__global__ void ccl_merge_kernel(const uint32_t *__restrict__ bits, TailBuffers tb)
{
    const BitGeom g = tb.g;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.rows * g.wpr) return;
    const int y = t / g.wpr, j = t % g.wpr;
    const int2 ext = tb.rowext[y];
    if (ext.y < 0 || 32 * j > ext.y || 32 * j + 31 < ext.x) return;  // nothing but exterior background here
    int *P = tb.parent;
    {  // foreground
        const uint32_t w = pol_word(bits, g, y, j, true);
        if (w) {
            const uint32_t up = pol_word(bits, g, y - 1, j, true);
            const uint32_t left = pol_word(bits, g, y, j - 1, true);
            for (uint32_t st = w & ~(w << 1); st; st &= st - 1) {
                const int s = __ffs(st) - 1;
                const int len = run_len(w, s);
                const int e = s + len - 1;
                const uint32_t sm = seg_mask(s, len);
                const int id = node_id(g, y, j, s);
                if (s == 0 && (left >> 31)) uf_union(P, id, node_id(g, y, j - 1, run_start(left, 31)));
                uint32_t m = up & (sm | (sm << 1) | (sm >> 1));
                while (m) {
                    const int b = __ffs(m) - 1;
                    const int us = run_start(up, b);
                    m &= ~seg_mask(us, run_len(up, us));
                    uf_union(P, id, node_id(g, y - 1, j, us));
                }
                if (s == 0) {
                    const uint32_t ul = pol_word(bits, g, y - 1, j - 1, true);
                    if (ul >> 31) uf_union(P, id, node_id(g, y - 1, j - 1, run_start(ul, 31)));
                }
                if (e == 31) {
                    const uint32_t ur = pol_word(bits, g, y - 1, j + 1, true);
                    if (ur & 1u) uf_union(P, id, node_id(g, y - 1, j + 1, 0));
                }
            }
        }
    }
    {  // candidate background
        const uint32_t w = cand_word(bits, tb.rowext, g, y, j);
        if (w) {
            const uint32_t up = cand_word(bits, tb.rowext, g, y - 1, j);
            const uint32_t left = cand_word(bits, tb.rowext, g, y, j - 1);
            const uint32_t ext_up = rowext_bg_word(bits, tb.rowext, g, y - 1, j);
            const uint32_t ext_dn = rowext_bg_word(bits, tb.rowext, g, y + 1, j);
            for (uint32_t st = w & ~(w << 1); st; st &= st - 1) {
                const int s = __ffs(st) - 1;
                const int len = run_len(w, s);
                const uint32_t sm = seg_mask(s, len);
                const int id = node_id(g, y, j, s);
                if ((ext_up | ext_dn) & sm) uf_union(P, id, 0);
                if (s == 0 && (left >> 31)) uf_union(P, id, node_id(g, y, j - 1, run_start(left, 31)));
                uint32_t m = up & sm;
                while (m) {
                    const int b = __ffs(m) - 1;
                    const int us = run_start(up, b);
                    m &= ~seg_mask(us, run_len(up, us));
                    uf_union(P, id, node_id(g, y - 1, j, us));
                }
            }
        }
    }
}

// Optional egress: label of every foreground pixel = unpadded linear index of its
// 8-connected component's raster-first pixel; -1 for background.  Runs after phase A.
__global__ void ccl_labels_kernel(const uint32_t *__restrict__ bits, TailBuffers tb, int32_t *__restrict__ labels)
{
    const BitGeom g = tb.g;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.rows * g.wpr) return;
    const int y = t / g.wpr, j = t % g.wpr;
    const uint32_t w = pol_word(bits, g, y, j, true);
    const int nb = min(32, g.cols - 32 * j);
    int32_t *row = labels + (size_t)y * g.cols + 32 * j;
    for (int b = 0; b < nb; ++b) row[b] = -1;
    for (uint32_t st = w & ~(w << 1); st; st &= st - 1) {
        const int s = __ffs(st) - 1;
        const int len = run_len(w, s);
        const int r = uf_find(tb.parent, node_id(g, y, j, s)) - 1;
        const int ry = r / g.pitch_px(), rx = r % g.pitch_px();
        const int32_t lab = ry * g.cols + rx;
        for (int b = s; b < s + len; ++b) row[b] = lab;
    }
}

// phase B: holes (candidate-background sets not linked to EXT) join the foreground beside them;
// writes G = foreground + holes.
__global__ void ccl_fill_kernel(const uint32_t *__restrict__ bits, TailBuffers tb)
{
    const BitGeom g = tb.g;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.rows * g.wpr) return;
    const int y = t / g.wpr, j = t % g.wpr;
    int *P = tb.parent;
    const uint32_t f = bits[t] & g.valid_mask(j);
    uint32_t G = f;
    const uint32_t c = cand_word(bits, tb.rowext, g, y, j);
    for (uint32_t st = c & ~(c << 1); st; st &= st - 1) {
        const int s = __ffs(st) - 1;
        const int len = run_len(c, s);
        const int id = node_id(g, y, j, s);
        if (uf_find_compress(P, id) == 0) continue;  // exterior
        G |= seg_mask(s, len);
        // the pixel left of a hole segment is foreground or the same hole (previous word)
        if (s > 0) {
            uf_union(P, id, node_id(g, y, j, run_start(f, s - 1)));
        } else {
            const uint32_t lf = pol_word(bits, g, y, j - 1, true);
            if (lf >> 31) uf_union(P, id, node_id(g, y, j - 1, run_start(lf, 31)));
        }
        // ... and the pixel right of it (covers islands whose left neighbour is the hole)
        const int e = s + len;
        if (e < 32) {
            if ((f >> e) & 1u) uf_union(P, id, node_id(g, y, j, e));
        } else {
            const uint32_t rf = pol_word(bits, g, y, j + 1, true);
            if (rf & 1u) uf_union(P, id, node_id(g, y, j + 1, 0));
        }
    }
    tb.G[t] = G;
}

__device__ __forceinline__ uint32_t sum_pos(uint32_t m)
{
    return __popc(m & 0xAAAAAAAAu) + 2u * __popc(m & 0xCCCCCCCCu) + 4u * __popc(m & 0xF0F0F0F0u) +
           8u * __popc(m & 0xFF00FF00u) + 16u * __popc(m & 0xFFFF0000u);
}

// moments: exact integer 2x2-cell sums per root.
__global__ void ccl_moments_kernel(TailBuffers tb)
{
    const BitGeom g = tb.g;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (g.rows - 1) * g.wpr) return;
    const int y = t / g.wpr, j = t % g.wpr;
    const uint32_t T = tb.G[t];
    if (!T) return;
    const uint32_t Bw = tb.G[t + g.wpr];
    const uint64_t T64 = (uint64_t)(load_word(tb.G, g, y, j - 1) >> 31) | ((uint64_t)T << 1) |
                         ((uint64_t)(load_word(tb.G, g, y, j + 1) & 1u) << 33);
    const uint64_t B64 = (uint64_t)(load_word(tb.G, g, y + 1, j - 1) >> 31) | ((uint64_t)Bw << 1) |
                         ((uint64_t)(load_word(tb.G, g, y + 1, j + 1) & 1u) << 33);
    const uint64_t tl = T64, tr = T64 >> 1, bl = B64, br = B64 >> 1;
    // indexed by the OWNER bit (the top pixel inside this word that names the cell's segment)
    const uint32_t full = (uint32_t)((tl & tr & bl & br) >> 1);
    const uint32_t mTR = (uint32_t)((tl & ~tr & bl & br) >> 1);
    const uint32_t mBL = (uint32_t)((tl & tr & ~bl & br) >> 1);
    const uint32_t mBR = (uint32_t)((tl & tr & bl & ~br) >> 1);
    const uint32_t mTL = (uint32_t)(~tl & tr & bl & br);  // owner = top-right pixel, cell x = owner - 1
    if (!(full | mTR | mBL | mBR | mTL)) return;
    const uint32_t xb = 32u * (uint32_t)j;
    for (uint32_t st = T & ~(T << 1); st; st &= st - 1) {
        const int s = __ffs(st) - 1;
        const int len = run_len(T, s);
        const uint32_t sm = seg_mask(s, len);
        const uint32_t F = full & sm, a = mTR & sm, b = mBL & sm, c = mBR & sm, d = mTL & sm;
        if (!(F | a | b | c | d)) continue;
        const uint32_t nF = __popc(F), na = __popc(a), nb = __popc(b), nc = __popc(c), nd = __popc(d);
        const uint32_t s00 = 2u * nF + na + nb + nc + nd;
        // x of a cell = xb + owner bit (A cells) or xb + owner bit - 1 (TL-missing cells)
        const uint32_t s10 = 6u * (sum_pos(F) + nF * xb) + 3u * nF + 3u * (sum_pos(a) + na * xb) + na +
                             3u * (sum_pos(b) + nb * xb) + 2u * nb + 3u * (sum_pos(c) + nc * xb) + nc +
                             3u * (sum_pos(d) + nd * xb) - nd;
        const uint32_t yy = (uint32_t)y;
        const uint32_t s01 = nF * (6u * yy + 3u) + (na + nd) * (3u * yy + 2u) + (nb + nc) * (3u * yy + 1u);
        const int root = uf_find_compress(tb.parent, node_id(g, y, j, s));
        // warp-aggregated: lanes holding the same root add once
        const unsigned act = __activemask();
        const unsigned peers = __match_any_sync(act, root);
        const uint32_t r00 = __reduce_add_sync(peers, s00);
        const uint32_t r10 = __reduce_add_sync(peers, s10);
        const uint32_t r01 = __reduce_add_sync(peers, s01);
        if ((int)(threadIdx.x & 31u) == __ffs(peers) - 1) {
            atomicAdd(tb.acc + root, (unsigned long long)r00);
            atomicAdd(tb.acc + tb.nnodes + root, (unsigned long long)r10);
            atomicAdd(tb.acc + 2 * tb.nnodes + root, (unsigned long long)r01);
        }
    }
}

// select: arg-max over roots; the last CTA to finish turns the winner into the detection
// (cv::moments' contourMoments scaling: m00 = a00/2, m10 = a10/6, m01 = a01/6 in doubles, then
// DetectorFunc.cpp:58-59 x = m10/m00, y = m01/m00).
__global__ void ccl_select_kernel(const uint32_t *__restrict__ bits, TailBuffers tb, double min_area,
                                  double max_area, oat_detection *out)
{
    const BitGeom g = tb.g;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned nroots = 0;
    unsigned long long key = 0ull;
    if (t < g.rows * g.wpr) {
        const int y = t / g.wpr, j = t % g.wpr;
        const uint32_t w = pol_word(bits, g, y, j, true);
        for (uint32_t st = w & ~(w << 1); st; st &= st - 1) {
            const int s = __ffs(st) - 1;
            const int id = node_id(g, y, j, s);
            if (__ldcg(tb.parent + id) != id) continue;
            ++nroots;
            const unsigned long long s00 = __ldcg(tb.acc + id);
            const double area = 0.5 * (double)s00;
            if (area >= min_area && area < max_area && s00 > 0ull) {
                const unsigned long long k = (s00 << 32) | (unsigned long long)(unsigned)id;
                key = k > key ? k : key;
            }
        }
    }
    const unsigned tot = __reduce_add_sync(0xffffffffu, nroots);
    if (tot && (threadIdx.x & 31u) == 0) atomicAdd(tb.count, tot);
    if (key) atomicMax(tb.best, key);
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = (atomicAdd(tb.ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        const unsigned long long best = __ldcg(tb.best);
        oat_detection d;
        d.position_valid = 0;
        d.n_components = (int32_t)__ldcg(tb.count);
        d.x = d.y = d.area = 0.0;
        if (best) {
            const int id = (int)(best & 0xffffffffull);
            const double m00 = (double)__ldcg(tb.acc + id) * 0.5;
            const double m10 = (double)__ldcg(tb.acc + tb.nnodes + id) * 0.16666666666666666666666666666667;
            const double m01 = (double)__ldcg(tb.acc + 2 * tb.nnodes + id) * 0.16666666666666666666666666666667;
            d.position_valid = 1;
            d.x = m10 / m00;
            d.y = m01 / m00;
            d.area = m00;
        }
        *out = d;
    }
}

}  // namespace oat
