// mog_fused.cuh -- the dominant (HBM-bound) kernel of the path:
//
//   framefilt mog     cv::BackgroundSubtractorMOG2::apply + frame.setTo(0, mask == 0)
//                     (reference: src/framefilter/BackgroundSubtractorMOG.cpp:83, :124-125)
//   framefilt col     cv::cvtColor(COLOR_BGR2HSV), 8 bit      (src/framefilter/ColorConvert.cpp:101-107)
//   posidet hsv (1)   cv::inRange on the HSV frame             (src/positiondetector/HSVDetector.cpp:146-149)
//
// fused into ONE pass over HBM: per pixel it reads 3 B of BGR, the mode count and the live GMM
// modes, writes the updated modes back, and emits the post-threshold mask as 1 bit/pixel for
// the detect tail (morph.cu / ccl.cu).  Frame/mask/HSV egress (what the reference components
// publish to their SINKs) is optional and only costs bytes when a consumer asked for it.
//
// GMM state layout in HBM ("mode-major planes"):
//     state[(m*5 + c) * plane + y*pitch_px + x],  c = {weight, variance, mean_b, mean_g, mean_r}
//     nmodes[y*pitch_px + x]                      live mode count (cv's bgmodelUsedModes)
// pitch_px = 32*wpr, so one warp-wide float4 load covers 128 consecutive pixels of one plane
// (512 B, fully coalesced) and planes of modes no pixel of the thread uses are never fetched.
//
// Arithmetic is the literal MOG2Invoker sequence in fp32 with NO fused multiply-add
// (__fmul_rn/__fadd_rn are never contracted), so masks and state are bit-identical to the
// SSE-baseline OpenCV build the reference links.
#pragma once
#include <float.h>

#include "common.cuh"

namespace oat {

struct MogConsts {
    float aT, a1, prune;  // per frame: alphaT, 1-alphaT, float(-lr*CT)
    float Tb, TB, Tg, varInit, varMin, varMax, tau;
    int detect_shadows, shadow_value;
};

struct FusedArgs {
    const uint8_t *bgr;
    size_t in_pitch;
    int rows, cols, wpr;
    float *state;
    size_t plane;  // floats per plane = rows * wpr * 32
    uint8_t *nmodes;
    int reset;  // 1: model is empty (first frame / learning rate >= 1)
    MogConsts c;
    int do_hsv;        // HSV needed (threshold bits and/or HSV egress)
    int lo[3], hi[3];  // inRange band on (H,S,V), inclusive
    uint32_t *thr_bits;  // [rows][wpr] or NULL
    uint8_t *bgr_out;    // or NULL
    size_t bgr_out_pitch;
    uint8_t *fg_out;  // or NULL
    size_t fg_pitch;
    uint8_t *hsv_out;  // or NULL
    size_t hsv_pitch;
    const int *hsv_lut;  // sdiv[256] then hdiv[256]
    unsigned int *slow_count;  // or NULL: += number of 4-pixel groups that are not "one mode, fits, background"
};

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

template <bool TRACK>
__device__ __forceinline__ void upd(float &dst, float v, bool &dirty)
{
    if (TRACK) dirty |= (__float_as_uint(dst) != __float_as_uint(v));
    dst = v;
}
__device__ __forceinline__ void fswap(float &a, float &b)
{
    float t = a;
    a = b;
    b = t;
}

// One pixel of cv::MOG2Invoker (modules/video/src/bgfg_gaussmix2.cpp); NA mode slots in
// registers (n <= NA on entry and after a possible insertion), K = nmixtures (capacity).
// TRACK: compare every state write with the old value so that an unchanged model (frozen
// learning rate) is not written back.  W/V/A/B/C = weight, variance, mean b/g/r, sorted by weight descending.
// Returns the mask value {0, shadow_value, 255}; `dirty` is set if any state bit changed.
template <int K, int NA, bool TRACK>
__device__ __forceinline__ uint32_t mog2_pixel(const float x0, const float x1, const float x2, int &n,
                                               float (&W)[NA], float (&V)[NA], float (&A)[NA], float (&B)[NA],
                                               float (&C)[NA], const MogConsts &c, bool &dirty)
{
    bool bg = false, fits = false;
    float tot = 0.f;
#pragma unroll
    for (int m = 0; m < NA; ++m) {
        if (m < n) {  // n shrinks inside the loop when a mode is pruned, exactly like the reference
            float w = fadd(fmul(c.a1, W[m]), c.prune);
            int sc = 0;
            if (!fits) {
                const float var = V[m];
                const float d0 = fsub(A[m], x0), d1 = fsub(B[m], x1), d2 = fsub(C[m], x2);
                const float dist2 = fadd(fadd(fmul(d0, d0), fmul(d1, d1)), fmul(d2, d2));
                if (tot < c.TB && dist2 < fmul(c.Tb, var)) bg = true;
                if (dist2 < fmul(c.Tg, var)) {
                    fits = true;
                    w = fadd(w, c.aT);
                    const float k = fdiv(c.aT, w);
                    upd<TRACK>(A[m], fsub(A[m], fmul(k, d0)), dirty);
                    upd<TRACK>(B[m], fsub(B[m], fmul(k, d1)), dirty);
                    upd<TRACK>(C[m], fsub(C[m], fmul(k, d2)), dirty);
                    float vn = fadd(var, fmul(k, fsub(dist2, var)));
                    vn = (vn < c.varMin) ? c.varMin : vn;  // MAX(varnew, varMin)
                    vn = (vn > c.varMax) ? c.varMax : vn;  // MIN(varnew, varMax)
                    upd<TRACK>(V[m], vn, dirty);
                    bool go = true;
#pragma unroll
                    for (int i = m; i > 0; --i) {
                        if (go) {
                            if (w < W[i - 1]) {
                                go = false;
                            } else {
                                ++sc;
                                fswap(W[i], W[i - 1]);
                                fswap(V[i], V[i - 1]);
                                fswap(A[i], A[i - 1]);
                                fswap(B[i], B[i - 1]);
                                fswap(C[i], C[i - 1]);
                                dirty = true;
                            }
                        }
                    }
                }
            }
            if (w < -c.prune) {
                w = 0.f;
                --n;
                dirty = true;
            }
#pragma unroll
            for (int j = 0; j <= m; ++j)
                if (j == m - sc) upd<TRACK>(W[j], w, dirty);
            tot = fadd(tot, w);
        }
    }
    float inv = 0.f;
    if (fabsf(tot) > FLT_EPSILON) inv = fdiv(1.f, tot);
#pragma unroll
    for (int m = 0; m < NA; ++m)
        if (m < n) upd<TRACK>(W[m], fmul(W[m], inv), dirty);

    if (!fits && c.aT > 0.f) {
        const int mode = (n == K) ? K - 1 : n++;
        float wn;
        if (n == 1) {
            wn = 1.f;
        } else {
            wn = c.aT;
#pragma unroll
            for (int i = 0; i < NA - 1; ++i)
                if (i < n - 1) W[i] = fmul(W[i], c.a1);
        }
#pragma unroll
        for (int j = 0; j < NA; ++j)
            if (j == mode) {
                W[j] = wn;
                V[j] = c.varInit;
                A[j] = x0;
                B[j] = x1;
                C[j] = x2;
            }
        bool go = true;
#pragma unroll
        for (int i = NA - 1; i > 0; --i) {
            if (go && i <= n - 1) {
                if (c.aT < W[i - 1]) {
                    go = false;
                } else {
                    fswap(W[i], W[i - 1]);
                    fswap(V[i], V[i - 1]);
                    fswap(A[i], A[i - 1]);
                    fswap(B[i], B[i - 1]);
                    fswap(C[i], C[i - 1]);
                }
            }
        }
        dirty = true;
    }

    if (bg) return 0u;
    if (c.detect_shadows) {  // detectShadowGMM
        float tw = 0.f;
        bool done = false, shadow = false;
#pragma unroll
        for (int m = 0; m < NA; ++m) {
            if (!done && m < n) {
                const float num = fadd(fadd(fadd(0.f, fmul(x0, A[m])), fmul(x1, B[m])), fmul(x2, C[m]));
                const float den =
                    fadd(fadd(fadd(0.f, fmul(A[m], A[m])), fmul(B[m], B[m])), fmul(C[m], C[m]));
                if (den == 0.f) {
                    done = true;
                } else {
                    if (num <= den && num >= fmul(c.tau, den)) {
                        const float a = fdiv(num, den);
                        const float e0 = fsub(fmul(a, A[m]), x0), e1 = fsub(fmul(a, B[m]), x1),
                                    e2 = fsub(fmul(a, C[m]), x2);
                        const float dist2a =
                            fadd(fadd(fadd(0.f, fmul(e0, e0)), fmul(e1, e1)), fmul(e2, e2));
                        if (dist2a < fmul(fmul(fmul(c.Tb, V[m]), a), a)) {
                            shadow = true;
                            done = true;
                        }
                    }
                    if (!done) {
                        tw = fadd(tw, W[m]);
                        if (tw > c.TB) done = true;
                    }
                }
            }
        }
        if (shadow) return (uint32_t)c.shadow_value;
    }
    return 255u;
}

// The same algorithm as mog2_pixel(), written as ROLLED loops over mode arrays (which therefore live in
// local memory).  Several times slower per pixel, but a fraction of the instruction footprint: the
// pipelined kernel's slow path runs on a fraction of a percent of the pixels and its code is usually
// cold, so what it costs is instruction fetch, not arithmetic.  Identical operation order.
template <int K>
__device__ __noinline__ uint32_t mog2_pixel_rolled(const float x0, const float x1, const float x2, int &n_io, float *W,
                                                   float *V, float *A, float *B, float *C, const MogConsts &c)
{
    bool bg = false, fits = false;
    float tot = 0.f;
    int n = n_io;
    const int n_in = n;
#pragma unroll 1
    for (int m = 0; m < n_in; ++m) {
        if (m >= n) break;  // n shrinks inside the loop when a mode is pruned, exactly like the reference
        float w = fadd(fmul(c.a1, W[m]), c.prune);
        int sc = 0;
        if (!fits) {
            const float var = V[m];
            const float d0 = fsub(A[m], x0), d1 = fsub(B[m], x1), d2 = fsub(C[m], x2);
            const float dist2 = fadd(fadd(fmul(d0, d0), fmul(d1, d1)), fmul(d2, d2));
            if (tot < c.TB && dist2 < fmul(c.Tb, var)) bg = true;
            if (dist2 < fmul(c.Tg, var)) {
                fits = true;
                w = fadd(w, c.aT);
                const float k = fdiv(c.aT, w);
                A[m] = fsub(A[m], fmul(k, d0));
                B[m] = fsub(B[m], fmul(k, d1));
                C[m] = fsub(C[m], fmul(k, d2));
                float vn = fadd(var, fmul(k, fsub(dist2, var)));
                vn = (vn < c.varMin) ? c.varMin : vn;
                vn = (vn > c.varMax) ? c.varMax : vn;
                V[m] = vn;
#pragma unroll 1
                for (int i = m; i > 0; --i) {
                    if (w < W[i - 1]) break;
                    ++sc;
                    fswap(W[i], W[i - 1]);
                    fswap(V[i], V[i - 1]);
                    fswap(A[i], A[i - 1]);
                    fswap(B[i], B[i - 1]);
                    fswap(C[i], C[i - 1]);
                }
            }
        }
        if (w < -c.prune) {
            w = 0.f;
            --n;
        }
        W[m - sc] = w;
        tot = fadd(tot, w);
    }
    float inv = 0.f;
    if (fabsf(tot) > FLT_EPSILON) inv = fdiv(1.f, tot);
#pragma unroll 1
    for (int m = 0; m < n; ++m) W[m] = fmul(W[m], inv);
    if (!fits && c.aT > 0.f) {
        const int mode = (n == K) ? K - 1 : n++;
        if (n == 1) {
            W[mode] = 1.f;
        } else {
            W[mode] = c.aT;
#pragma unroll 1
            for (int i = 0; i < n - 1; ++i) W[i] = fmul(W[i], c.a1);
        }
        V[mode] = c.varInit;
        A[mode] = x0;
        B[mode] = x1;
        C[mode] = x2;
#pragma unroll 1
        for (int i = n - 1; i > 0; --i) {
            if (c.aT < W[i - 1]) break;
            fswap(W[i], W[i - 1]);
            fswap(V[i], V[i - 1]);
            fswap(A[i], A[i - 1]);
            fswap(B[i], B[i - 1]);
            fswap(C[i], C[i - 1]);
        }
    }
    n_io = n;
    if (bg) return 0u;
    if (c.detect_shadows) {
        float tw = 0.f;
#pragma unroll 1
        for (int m = 0; m < n; ++m) {
            const float num = fadd(fadd(fadd(0.f, fmul(x0, A[m])), fmul(x1, B[m])), fmul(x2, C[m]));
            const float den = fadd(fadd(fadd(0.f, fmul(A[m], A[m])), fmul(B[m], B[m])), fmul(C[m], C[m]));
            if (den == 0.f) break;
            if (num <= den && num >= fmul(c.tau, den)) {
                const float a = fdiv(num, den);
                const float e0 = fsub(fmul(a, A[m]), x0), e1 = fsub(fmul(a, B[m]), x1), e2 = fsub(fmul(a, C[m]), x2);
                const float dist2a = fadd(fadd(fadd(0.f, fmul(e0, e0)), fmul(e1, e1)), fmul(e2, e2));
                if (dist2a < fmul(fmul(fmul(c.Tb, V[m]), a), a)) return (uint32_t)c.shadow_value;
            }
            tw = fadd(tw, W[m]);
            if (tw > c.TB) break;
        }
    }
    return 255u;
}

// 8-bit BGR -> HSV, OpenCV's integer RGB2HSV_b (hsv_shift 12, hrange 180); lut = sdiv|hdiv.
__device__ __forceinline__ void bgr2hsv_px(int b, int g, int r, const int *lut, int &h, int &s, int &v)
{
    v = max(b, max(g, r));
    const int vmin = min(b, min(g, r));
    const int diff = v - vmin;
    s = (diff * lut[v] + (1 << 11)) >> 12;
    int hh = (v == r) ? (g - b) : ((v == g) ? (b - r + 2 * diff) : (r - g + 4 * diff));
    hh = (hh * lut[256 + diff] + (1 << 11)) >> 12;
    h = hh < 0 ? hh + 180 : hh;
}

// Same conversion without the table: sdiv[i] = round((255 << 12) / i) and hdiv[i] = round((180 << 12) / (6 i))
// never hit a rounding tie (2N/i is never an odd integer for i <= 255), so round-half-up integer division
// reproduces cvRound exactly.  Used where a table lookup would be a cold global load.
__device__ __forceinline__ void bgr2hsv_px_div(int b, int g, int r, int &h, int &s, int &v)
{
    v = max(b, max(g, r));
    const int vmin = min(b, min(g, r));
    const int diff = v - vmin;
    const int sdiv = v ? (2 * (255 << 12) + v) / (2 * v) : 0;
    const int hdiv = diff ? (2 * 122880 + diff) / (2 * diff) : 0;
    s = (diff * sdiv + (1 << 11)) >> 12;
    int hh = (v == r) ? (g - b) : ((v == g) ? (b - r + 2 * diff) : (r - g + 4 * diff));
    hh = (hh * hdiv + (1 << 11)) >> 12;
    h = hh < 0 ? hh + 180 : hh;
}

// Loads NL modes, runs the PX pixels of this thread, stores what changed.  NL (= the largest
// live-mode count among the thread's pixels) is a template parameter so that the common case
// (one or two live modes) is a compact, register-light code path; NA = modes held in registers
// (one spare slot for a mode inserted this frame).
template <int K, int NL, int PX, bool TRACK>
__device__ __forceinline__ uint32_t mog_body(const FusedArgs &a, const int *lut, const int y, const int x,
                                             const uint32_t (&px)[3 * PX], int (&n)[PX])
{
    constexpr int NA = (NL + 1 < K) ? NL + 1 : K;
    const size_t pidx = (size_t)y * (a.wpr * 32) + x;
    float S[NA][5][PX];
#pragma unroll
    for (int m = 0; m < NA; ++m) {
#pragma unroll
        for (int cc = 0; cc < 5; ++cc) {
            if (m < NL && !a.reset) {
                const float *p = a.state + (size_t)(m * 5 + cc) * a.plane + pidx;
                if (PX == 4) {
                    const float4 v = ld_state_f4(p);
                    S[m][cc][0] = v.x;
                    S[m][cc][PX > 1 ? 1 : 0] = v.y;
                    S[m][cc][PX > 2 ? 2 : 0] = v.z;
                    S[m][cc][PX > 3 ? 3 : 0] = v.w;
                } else {
                    S[m][cc][0] = ld_state_f1(p);
                }
            } else {
#pragma unroll
                for (int i = 0; i < PX; ++i) S[m][cc][i] = 0.f;
            }
        }
    }

    bool dirty = !TRACK || (a.reset != 0);
    uint32_t nib = 0;
    uint32_t fgm[PX], outb[3 * PX], outh[3 * PX];
    int nnew_max = 0;
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        float W[NA], V[NA], A[NA], B[NA], C[NA];
#pragma unroll
        for (int m = 0; m < NA; ++m) {
            W[m] = S[m][0][i];
            V[m] = S[m][1][i];
            A[m] = S[m][2][i];
            B[m] = S[m][3][i];
            C[m] = S[m][4][i];
        }
        const int b = (int)px[3 * i], gch = (int)px[3 * i + 1], r = (int)px[3 * i + 2];
        int ni = n[i];
        const uint32_t mk = mog2_pixel<K, NA, TRACK>((float)b, (float)gch, (float)r, ni, W, V, A, B, C, a.c, dirty);
        n[i] = ni;
#pragma unroll
        for (int m = 0; m < NA; ++m) {
            S[m][0][i] = W[m];
            S[m][1][i] = V[m];
            S[m][2][i] = A[m];
            S[m][3][i] = B[m];
            S[m][4][i] = C[m];
        }
        nnew_max = max(nnew_max, ni);
        fgm[i] = mk;
        const int ob = mk ? b : 0, og = mk ? gch : 0, orr = mk ? r : 0;
        outb[3 * i] = ob;
        outb[3 * i + 1] = og;
        outb[3 * i + 2] = orr;
        if (a.do_hsv) {
            int h = 0, s = 0, v = 0;
            if (mk) bgr2hsv_px(ob, og, orr, lut, h, s, v);
            outh[3 * i] = h;
            outh[3 * i + 1] = s;
            outh[3 * i + 2] = v;
            const bool in = (a.lo[0] <= h) & (h <= a.hi[0]) & (a.lo[1] <= s) & (s <= a.hi[1]) & (a.lo[2] <= v) &
                            (v <= a.hi[2]);
            nib |= (in ? 1u : 0u) << i;
        }
    }

    if (dirty) {
#pragma unroll
        for (int m = 0; m < NA; ++m) {
            if (m < NL || m < nnew_max) {
#pragma unroll
                for (int cc = 0; cc < 5; ++cc) {
                    float *p = a.state + (size_t)(m * 5 + cc) * a.plane + pidx;
                    if (PX == 4)
                        st_state_f4(p, make_float4(S[m][cc][0], S[m][cc][PX > 1 ? 1 : 0], S[m][cc][PX > 2 ? 2 : 0],
                                                   S[m][cc][PX > 3 ? 3 : 0]));
                    else
                        st_state_f1(p, S[m][cc][0]);
                }
            }
        }
        if (PX == 4)
            *reinterpret_cast<uint32_t *>(a.nmodes + pidx) = (uint32_t)n[0] | ((uint32_t)n[PX > 1 ? 1 : 0] << 8) |
                                                             ((uint32_t)n[PX > 2 ? 2 : 0] << 16) |
                                                             ((uint32_t)n[PX > 3 ? 3 : 0] << 24);
        else
            a.nmodes[pidx] = (uint8_t)n[0];
    }
    if (PX == 4) {
        if (a.bgr_out) {
            uint8_t *d = a.bgr_out + (size_t)y * a.bgr_out_pitch + 3 * x;
#pragma unroll
            for (int wd = 0; wd < 3; ++wd)
                st_stream_u32(d + 4 * wd, outb[(4 * wd) % (3 * PX)] | (outb[(4 * wd + 1) % (3 * PX)] << 8) |
                                              (outb[(4 * wd + 2) % (3 * PX)] << 16) |
                                              (outb[(4 * wd + 3) % (3 * PX)] << 24));
        }
        if (a.hsv_out) {
            uint8_t *d = a.hsv_out + (size_t)y * a.hsv_pitch + 3 * x;
#pragma unroll
            for (int wd = 0; wd < 3; ++wd)
                st_stream_u32(d + 4 * wd, outh[(4 * wd) % (3 * PX)] | (outh[(4 * wd + 1) % (3 * PX)] << 8) |
                                              (outh[(4 * wd + 2) % (3 * PX)] << 16) |
                                              (outh[(4 * wd + 3) % (3 * PX)] << 24));
        }
        if (a.fg_out)
            st_stream_u32(a.fg_out + (size_t)y * a.fg_pitch + x,
                          fgm[0] | (fgm[PX > 1 ? 1 : 0] << 8) | (fgm[PX > 2 ? 2 : 0] << 16) |
                              (fgm[PX > 3 ? 3 : 0] << 24));
    } else {
        if (a.bgr_out) {
            uint8_t *d = a.bgr_out + (size_t)y * a.bgr_out_pitch + 3 * x;
            d[0] = (uint8_t)outb[0];
            d[1] = (uint8_t)outb[1];
            d[2] = (uint8_t)outb[2];
        }
        if (a.hsv_out) {
            uint8_t *d = a.hsv_out + (size_t)y * a.hsv_pitch + 3 * x;
            d[0] = (uint8_t)outh[0];
            d[1] = (uint8_t)outh[1];
            d[2] = (uint8_t)outh[2];
        }
        if (a.fg_out) a.fg_out[(size_t)y * a.fg_pitch + x] = (uint8_t)fgm[0];
    }
    return nib;
}

#ifndef OAT_FUSED_MIN_BLOCKS
#define OAT_FUSED_MIN_BLOCKS 4
#endif

template <int K, int PX, bool TRACK>
__global__ void __launch_bounds__(128, OAT_FUSED_MIN_BLOCKS) mog_fused_kernel(const FusedArgs a)
{
    __shared__ int lut[512];
    if (a.do_hsv) {
        for (int i = threadIdx.x; i < 512; i += blockDim.x) lut[i] = a.hsv_lut[i];
        __syncthreads();
    }
    const int gpr = a.wpr * (32 / PX);
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int y = (int)(g / gpr);
    const int x = (int)(g % gpr) * PX;
    const bool active = (y < a.rows) && (x < a.cols);  // PX==4 requires cols % 4 == 0
    uint32_t nib = 0;
    bool multi = false;  // census: this thread's pixels carry more than one live mode

    if (active) {
        const size_t pidx = (size_t)y * (a.wpr * 32) + x;
        uint32_t px[3 * PX];
        int n[PX];
        const uint8_t *src = a.bgr + (size_t)y * a.in_pitch + 3 * x;
        if (PX == 4) {
            const uint32_t w0 = ld_stream_u32(src), w1 = ld_stream_u32(src + 4), w2 = ld_stream_u32(src + 8);
            const uint32_t nm = a.reset ? 0u : *reinterpret_cast<const uint32_t *>(a.nmodes + pidx);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                px[i % (3 * PX)] = (w0 >> (8 * i)) & 255u;
                px[(4 + i) % (3 * PX)] = (w1 >> (8 * i)) & 255u;
                px[(8 + i) % (3 * PX)] = (w2 >> (8 * i)) & 255u;
                n[i % PX] = (int)((nm >> (8 * i)) & 255u);
            }
        } else {
            px[0] = ld_stream_u8(src);
            px[1] = ld_stream_u8(src + 1);
            px[2] = ld_stream_u8(src + 2);
            n[0] = a.reset ? 0 : (int)a.nmodes[pidx];
        }
        int nmax = n[0];
#pragma unroll
        for (int i = 1; i < PX; ++i) nmax = max(nmax, n[i]);
        multi = nmax >= 2;
        // thread-level dispatch on the number of live modes (spatially coherent in practice)
        if (nmax <= 1 || K == 1)
            nib = mog_body<K, 1, PX, TRACK>(a, lut, y, x, px, n);
        else if (nmax == 2 || K == 2)
            nib = mog_body<K, (K < 2 ? K : 2), PX, TRACK>(a, lut, y, x, px, n);
        else if (nmax == 3 || K == 3)
            nib = mog_body<K, (K < 3 ? K : 3), PX, TRACK>(a, lut, y, x, px, n);
        else if (nmax == 4 || K == 4)
            nib = mog_body<K, (K < 4 ? K : 4), PX, TRACK>(a, lut, y, x, px, n);
        else
            nib = mog_body<K, K, PX, TRACK>(a, lut, y, x, px, n);
    }

    if (a.slow_count) {
        const unsigned nm_ = __popc(__ballot_sync(0xffffffffu, multi));
        if ((threadIdx.x & 31u) == 0 && nm_) atomicAdd(a.slow_count, (PX == 4 ? 1u : 0u) * nm_ + (PX == 4 ? 0u : (nm_ + 3u) / 4u));
    }
    // ---- threshold mask, 1 bit/pixel: 8 lanes x 4 px (or 32 lanes x 1 px) -> one word --------
    if (a.thr_bits) {  // uniform
        const unsigned lane = threadIdx.x & 31u;
        if (PX == 4) {
            uint32_t wv = nib << (4 * (lane & 7u));
            wv |= __shfl_xor_sync(0xffffffffu, wv, 1);
            wv |= __shfl_xor_sync(0xffffffffu, wv, 2);
            wv |= __shfl_xor_sync(0xffffffffu, wv, 4);
            if ((lane & 7u) == 0 && y < a.rows) a.thr_bits[(size_t)y * a.wpr + (x >> 5)] = wv;
        } else {
            const uint32_t wv = __ballot_sync(0xffffffffu, nib & 1u);
            if (lane == 0 && y < a.rows) a.thr_bits[(size_t)y * a.wpr + (x >> 5)] = wv;
        }
    }
}

}  // namespace oat
