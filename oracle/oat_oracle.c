/*
 * oat_oracle.c -- CPU restatement of the arithmetic on Oat's tracking hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker.  The product path (oat_b200/, liboatgpu.so) never calls it.
 *
 * Where the algorithm lives.  Oat delegates every image operation on this path to OpenCV
 * (find_package(OpenCV REQUIRED), /root/reference/CMakeLists.txt:88; no version pinned, Travis
 * used libopencv3-dev, .travis.yml:28), which is NOT vendored under /root/reference.  Each
 * function below therefore restates the published OpenCV algorithm behind one reference call
 * site (cited per function) and is PINNED against the real library -- the `cv2` 4.13.0 wheel
 * in this image, which exports those same C++ functions -- by tests/test_oracle_vs_cv2.py and
 * the committed fixtures under tests/golden/ (made by tests/golden/make_golden.py).  The
 * reference's own tests hold no golden vectors for this path.
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off: MOG2 parity needs plain mul/add, the cv2
 * build does not contract to FMA).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------ */
/* Synthetic stream; identical integer arithmetic in oracle/synth.py and */
/* in the CUDA generator (oat_b200/csrc/synth.cuh).                                      */
/* ------------------------------------------------------------------------------------ */
static inline uint32_t fmix32(uint32_t h)
{
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

ORC_API void orc_synth_frame(uint8_t *dst, size_t pitch, int rows, int cols, uint32_t seed,
                             uint32_t t)
{
    const uint32_t kbg = fmix32(seed ^ 0x9e3779b9u);
    const uint32_t knz = fmix32(seed + 0x7f4a7c15u * (t + 1u));
    const int r = rows / 20;
    const int cx = cols / 4 + (int)((7u * t) % (uint32_t)(cols / 2));
    const int cy = rows / 3 + (int)((4u * t) % (uint32_t)(rows / 3));
    static const uint8_t disc[3] = {40, 220, 60};
    for (int y = 0; y < rows; ++y) {
        uint8_t *row = dst + (size_t)y * pitch;
        for (int x = 0; x < cols; ++x) {
            int dx = x - cx, dy = y - cy;
            int in_disc = (t != 0) && (dx * dx + dy * dy <= r * r);
            for (int c = 0; c < 3; ++c) {
                uint32_t idx = ((uint32_t)y * (uint32_t)cols + (uint32_t)x) * 3u + (uint32_t)c;
                int bg = 40 + (int)(fmix32(idx ^ kbg) % 81u);
                int nz = (int)(fmix32(idx ^ knz) % 7u) - 3;
                row[3 * x + c] = in_disc ? disc[c] : (uint8_t)(bg + nz);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* framefilt mog: cv::BackgroundSubtractorMOG2::apply                                    */
/* call site: /root/reference/src/framefilter/BackgroundSubtractorMOG.cpp:83 (create,     */
/* all defaults) and :124 (apply(frame, mask, learning_coeff_)).                          */
/* Algorithm: Zivkovic 2004/2006 as implemented by OpenCV's MOG2Invoker                  */
/* (modules/video/src/bgfg_gaussmix2.cpp); state layout = OpenCV's (AoS per pixel).      */
/* ------------------------------------------------------------------------------------ */
typedef struct orc_mog_params {
    int history;
    int nmixtures;
    float var_threshold;
    float var_threshold_gen;
    float background_ratio;
    float var_init, var_min, var_max;
    float ct;
    int detect_shadows;
    int shadow_value;
    float shadow_threshold;
} orc_mog_params;

typedef struct orc_mog {
    int rows, cols, K;
    orc_mog_params p;
    int nframes;
    uint8_t *modes_used; /* [rows*cols] */
    float *weight;       /* [rows*cols*K] */
    float *variance;     /* [rows*cols*K] */
    float *mean;         /* [rows*cols*K*3] */
} orc_mog;

ORC_API void orc_mog_default_params(orc_mog_params *p)
{
    p->history = 500;
    p->nmixtures = 5;
    p->var_threshold = 16.0f;
    p->var_threshold_gen = 9.0f;
    p->background_ratio = 0.9f;
    p->var_init = 15.0f;
    p->var_min = 4.0f;
    p->var_max = 75.0f; /* 5*var_init */
    p->ct = 0.05f;
    p->detect_shadows = 1;
    p->shadow_value = 127;
    p->shadow_threshold = 0.5f;
}

ORC_API orc_mog *orc_mog_create(int rows, int cols, const orc_mog_params *p)
{
    orc_mog *m = (orc_mog *)calloc(1, sizeof(orc_mog));
    if (p)
        m->p = *p;
    else
        orc_mog_default_params(&m->p);
    m->rows = rows;
    m->cols = cols;
    m->K = m->p.nmixtures;
    size_t n = (size_t)rows * cols;
    m->modes_used = (uint8_t *)calloc(n, 1);
    m->weight = (float *)calloc(n * m->K, sizeof(float));
    m->variance = (float *)calloc(n * m->K, sizeof(float));
    m->mean = (float *)calloc(n * m->K * 3, sizeof(float));
    return m;
}

ORC_API void orc_mog_destroy(orc_mog *m)
{
    if (!m) return;
    free(m->modes_used);
    free(m->weight);
    free(m->variance);
    free(m->mean);
    free(m);
}

ORC_API void orc_mog_reset(orc_mog *m)
{
    size_t n = (size_t)m->rows * m->cols;
    m->nframes = 0;
    memset(m->modes_used, 0, n);
    memset(m->weight, 0, n * m->K * sizeof(float));
    memset(m->variance, 0, n * m->K * sizeof(float));
    memset(m->mean, 0, n * m->K * 3 * sizeof(float));
}

ORC_API void orc_mog_state(orc_mog *m, uint8_t **modes, float **w, float **var, float **mean)
{
    if (modes) *modes = m->modes_used;
    if (w) *w = m->weight;
    if (var) *var = m->variance;
    if (mean) *mean = m->mean;
}

static int mog_shadow(const float *x, int nmodes, const float *w, const float *var,
                      const float *mean, float Tb, float TB, float tau)
{
    float tw = 0.0f;
    for (int mode = 0; mode < nmodes; ++mode, mean += 3) {
        float num = 0.0f, den = 0.0f;
        for (int c = 0; c < 3; ++c) {
            num += x[c] * mean[c];
            den += mean[c] * mean[c];
        }
        if (den == 0.0f) return 0;
        if (num <= den && num >= tau * den) {
            float a = num / den;
            float dist2a = 0.0f;
            for (int c = 0; c < 3; ++c) {
                float dD = a * mean[c] - x[c];
                dist2a += dD * dD;
            }
            if (dist2a < Tb * var[mode] * a * a) return 1;
        }
        tw += w[mode];
        if (tw > TB) return 0;
    }
    return 0;
}

/* The effective learning rate of cv::BackgroundSubtractorMOG2Impl::apply, also used by the
 * host side of the product (kept separately there; this copy is the checker). */
ORC_API double orc_mog_effective_rate(int nframes_after_increment, double learning_rate, int history)
{
    if (learning_rate >= 0 && nframes_after_increment > 1) return learning_rate;
    int d = 2 * nframes_after_increment;
    if (d > history) d = history;
    return 1.0 / d;
}

ORC_API void orc_mog_apply(orc_mog *m, const uint8_t *bgr, size_t pitch, uint8_t *mask,
                           size_t mask_pitch, double learning_rate)
{
    const int K = m->K;
    /* needToInitialize = nframes == 0 || learningRate >= 1 (size/type never change here) */
    if (m->nframes == 0 || learning_rate >= 1) orc_mog_reset(m);
    ++m->nframes;
    learning_rate = orc_mog_effective_rate(m->nframes, learning_rate, m->p.history);

    const float alphaT = (float)learning_rate;
    const float alpha1 = 1.0f - alphaT;
    const float prune = (float)(-learning_rate * (double)m->p.ct);
    const float Tb = m->p.var_threshold, TB = m->p.background_ratio, Tg = m->p.var_threshold_gen;
    const float varInit = m->p.var_init, varMin = m->p.var_min, varMax = m->p.var_max;
    const float tau = m->p.shadow_threshold;

    for (int y = 0; y < m->rows; ++y) {
        const uint8_t *src = bgr + (size_t)y * pitch;
        for (int x = 0; x < m->cols; ++x) {
            size_t px = (size_t)y * m->cols + x;
            float data[3] = {(float)src[3 * x], (float)src[3 * x + 1], (float)src[3 * x + 2]};
            float *W = m->weight + px * K;
            float *V = m->variance + px * K;
            float *mean = m->mean + px * K * 3;
            int background = 0, fits = 0;
            int nmodes = m->modes_used[px];
            float totalWeight = 0.0f;
            float *mean_m = mean;
            for (int mode = 0; mode < nmodes; ++mode, mean_m += 3) {
                float weight = alpha1 * W[mode] + prune;
                int swap_count = 0;
                if (!fits) {
                    float var = V[mode];
                    float d0 = mean_m[0] - data[0];
                    float d1 = mean_m[1] - data[1];
                    float d2 = mean_m[2] - data[2];
                    float dist2 = d0 * d0 + d1 * d1 + d2 * d2;
                    if (totalWeight < TB && dist2 < Tb * var) background = 1;
                    if (dist2 < Tg * var) {
                        fits = 1;
                        weight += alphaT;
                        float k = alphaT / weight;
                        mean_m[0] -= k * d0;
                        mean_m[1] -= k * d1;
                        mean_m[2] -= k * d2;
                        float varnew = var + k * (dist2 - var);
                        varnew = (varnew < varMin) ? varMin : varnew; /* MAX(varnew, varMin) */
                        varnew = (varnew > varMax) ? varMax : varnew; /* MIN(varnew, varMax) */
                        V[mode] = varnew;
                        for (int i = mode; i > 0; --i) {
                            if (weight < W[i - 1]) break;
                            swap_count++;
                            float tf;
                            tf = W[i]; W[i] = W[i - 1]; W[i - 1] = tf;
                            tf = V[i]; V[i] = V[i - 1]; V[i - 1] = tf;
                            for (int c = 0; c < 3; ++c) {
                                tf = mean[i * 3 + c];
                                mean[i * 3 + c] = mean[(i - 1) * 3 + c];
                                mean[(i - 1) * 3 + c] = tf;
                            }
                        }
                    }
                }
                if (weight < -prune) {
                    weight = 0.0f;
                    nmodes--;
                }
                W[mode - swap_count] = weight;
                totalWeight += weight;
            }
            float invWeight = 0.0f;
            if (fabsf(totalWeight) > FLT_EPSILON) invWeight = 1.0f / totalWeight;
            for (int mode = 0; mode < nmodes; ++mode) W[mode] *= invWeight;

            if (!fits && alphaT > 0.0f) {
                int mode = (nmodes == K) ? K - 1 : nmodes++;
                if (nmodes == 1)
                    W[mode] = 1.0f;
                else {
                    W[mode] = alphaT;
                    for (int i = 0; i < nmodes - 1; ++i) W[i] *= alpha1;
                }
                mean[mode * 3 + 0] = data[0];
                mean[mode * 3 + 1] = data[1];
                mean[mode * 3 + 2] = data[2];
                V[mode] = varInit;
                for (int i = nmodes - 1; i > 0; --i) {
                    if (alphaT < W[i - 1]) break;
                    float tf;
                    tf = W[i]; W[i] = W[i - 1]; W[i - 1] = tf;
                    tf = V[i]; V[i] = V[i - 1]; V[i - 1] = tf;
                    for (int c = 0; c < 3; ++c) {
                        tf = mean[i * 3 + c];
                        mean[i * 3 + c] = mean[(i - 1) * 3 + c];
                        mean[(i - 1) * 3 + c] = tf;
                    }
                }
            }
            m->modes_used[px] = (uint8_t)nmodes;
            if (mask) {
                uint8_t v;
                if (background)
                    v = 0;
                else if (m->p.detect_shadows && mog_shadow(data, nmodes, W, V, mean, Tb, TB, tau))
                    v = (uint8_t)m->p.shadow_value;
                else
                    v = 255;
                mask[(size_t)y * mask_pitch + x] = v;
            }
        }
    }
}

/* frame.setTo(0, background_mask_ == 0)   BackgroundSubtractorMOG.cpp:125 */
ORC_API void orc_zero_where_mask0(uint8_t *frame, size_t pitch, const uint8_t *mask,
                                  size_t mask_pitch, int rows, int cols, int channels)
{
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x)
            if (mask[(size_t)y * mask_pitch + x] == 0)
                for (int c = 0; c < channels; ++c) frame[(size_t)y * pitch + channels * x + c] = 0;
}

/* ------------------------------------------------------------------------------------ */
/* framefilt col -C HSV: cv::cvtColor(COLOR_BGR2HSV), 8-bit                              */
/* call site: /root/reference/src/framefilter/ColorConvert.cpp:104, Color.h:46-51.        */
/* OpenCV's integer RGB2HSV_b: fixed-point tables with hsv_shift = 12, hrange = 180.      */
/* ------------------------------------------------------------------------------------ */
static int g_sdiv[256], g_hdiv[256], g_tab_init = 0;
static void hsv_tables(void)
{
    if (g_tab_init) return;
    g_sdiv[0] = g_hdiv[0] = 0;
    for (int i = 1; i < 256; ++i) {
        g_sdiv[i] = (int)nearbyint((255 << 12) / (1.0 * i));  /* saturate_cast<int>(double) */
        g_hdiv[i] = (int)nearbyint((180 << 12) / (6.0 * i));
    }
    g_tab_init = 1;
}

ORC_API void orc_hsv_tables(int *sdiv, int *hdiv)
{
    hsv_tables();
    memcpy(sdiv, g_sdiv, sizeof(g_sdiv));
    memcpy(hdiv, g_hdiv, sizeof(g_hdiv));
}

ORC_API void orc_bgr2hsv(const uint8_t *bgr, size_t in_pitch, uint8_t *hsv, size_t out_pitch,
                         int rows, int cols)
{
    hsv_tables();
    for (int y = 0; y < rows; ++y) {
        const uint8_t *s = bgr + (size_t)y * in_pitch;
        uint8_t *d = hsv + (size_t)y * out_pitch;
        for (int x = 0; x < cols; ++x) {
            int b = s[3 * x], g = s[3 * x + 1], r = s[3 * x + 2];
            int v = b, vmin = b;
            if (g > v) v = g;
            if (r > v) v = r;
            if (g < vmin) vmin = g;
            if (r < vmin) vmin = r;
            int diff = v - vmin;
            int vr = (v == r) ? -1 : 0;
            int vg = (v == g) ? -1 : 0;
            int sat = (diff * g_sdiv[v] + (1 << 11)) >> 12;
            int h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
            h = (h * g_hdiv[diff] + (1 << 11)) >> 12;
            if (h < 0) h += 180;
            d[3 * x] = (uint8_t)h;
            d[3 * x + 1] = (uint8_t)sat;
            d[3 * x + 2] = (uint8_t)v;
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* posidet hsv step 1: cv::inRange(frame, Scalar(lo), Scalar(hi), thr)                   */
/* call site: /root/reference/src/positiondetector/HSVDetector.cpp:146-149.              */
/* Inclusive on both ends; bounds compare as integers against u8 data, so 256 == 255 and  */
/* lo > 255 or lo > hi passes nothing.                                                    */
/* ------------------------------------------------------------------------------------ */
ORC_API void orc_inrange3(const uint8_t *src, size_t pitch, uint8_t *dst, size_t dst_pitch, int rows,
                          int cols, const int lo[3], const int hi[3])
{
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            int ok = 1;
            for (int c = 0; c < 3; ++c) {
                int v = src[(size_t)y * pitch + 3 * x + c];
                ok &= (lo[c] <= v) & (v <= hi[c]);
            }
            dst[(size_t)y * dst_pitch + x] = ok ? 255 : 0;
        }
}

/* ------------------------------------------------------------------------------------ */
/* posidet hsv steps 2/3: cv::erode / cv::dilate with getStructuringElement(MORPH_RECT,  */
/* Size(k,k)), default anchor (k/2,k/2) and default border (ignored: +inf for erode, -inf */
/* for dilate). call sites: HSVDetector.cpp:152-156, :253-273.                            */
/* Both sample the input window [x - k/2, x - k/2 + k - 1] (same in y).                   */
/* ------------------------------------------------------------------------------------ */
static void morph_rect(const uint8_t *src, size_t pitch, uint8_t *dst, size_t dst_pitch, int rows,
                       int cols, int k, int is_dilate)
{
    int a = k / 2;
    uint8_t *tmp = (uint8_t *)malloc((size_t)rows * cols);
    /* horizontal */
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            int acc = is_dilate ? 0 : 255;
            int x0 = x - a, x1 = x - a + k - 1;
            if (x0 < 0) x0 = 0;
            if (x1 > cols - 1) x1 = cols - 1;
            for (int j = x0; j <= x1; ++j) {
                int v = src[(size_t)y * pitch + j];
                acc = is_dilate ? (v > acc ? v : acc) : (v < acc ? v : acc);
            }
            tmp[(size_t)y * cols + x] = (uint8_t)acc;
        }
    /* vertical */
    for (int y = 0; y < rows; ++y) {
        int y0 = y - a, y1 = y - a + k - 1;
        if (y0 < 0) y0 = 0;
        if (y1 > rows - 1) y1 = rows - 1;
        for (int x = 0; x < cols; ++x) {
            int acc = is_dilate ? 0 : 255;
            for (int j = y0; j <= y1; ++j) {
                int v = tmp[(size_t)j * cols + x];
                acc = is_dilate ? (v > acc ? v : acc) : (v < acc ? v : acc);
            }
            dst[(size_t)y * dst_pitch + x] = (uint8_t)acc;
        }
    }
    free(tmp);
}

ORC_API void orc_erode_rect(const uint8_t *src, size_t pitch, uint8_t *dst, size_t dst_pitch,
                            int rows, int cols, int k)
{
    morph_rect(src, pitch, dst, dst_pitch, rows, cols, k, 0);
}

ORC_API void orc_dilate_rect(const uint8_t *src, size_t pitch, uint8_t *dst, size_t dst_pitch,
                             int rows, int cols, int k)
{
    morph_rect(src, pitch, dst, dst_pitch, rows, cols, k, 1);
}

/* ------------------------------------------------------------------------------------ */
/* posidet hsv step 4: siftContours                                                      */
/* /root/reference/src/positiondetector/DetectorFunc.cpp:31-66:                          */
/*   findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE); per contour cv::moments(contour); */
/*   keep the largest m00 with min <= m00 < max and m00 > best (strict).                  */
/* Restated as: (1) 4-connected flood of the background from the (zero-padded) image      */
/* border; (2) raster scan for 8-connected components whose raster-first pixel's left     */
/* neighbour is exterior background = the outer borders RETR_EXTERNAL reports;           */
/* (3) Suzuki-Abe border following (step 3 of their Algorithm 1) of each outer border;    */
/* (4) cv::moments' contourMoments (Green's theorem on the integer polygon, doubles).    */
/* cv2 lists contours in reverse raster order of their first pixel, so with the strict    */
/* '>' the LAST component in raster order wins an area tie.                               */
/* ------------------------------------------------------------------------------------ */
typedef struct orc_detection {
    int32_t position_valid;
    int32_t n_components;
    double x, y, area;
} orc_detection;

typedef struct orc_contour_rec {
    int32_t first_index; /* y*cols+x of the raster-first pixel */
    int32_t npoints;
    double m00, m10, m01; /* cv::moments spatial moments of the outer border */
} orc_contour_rec;

/* 8-neighbourhood in clockwise order starting at West (image coords, y down):
 * W, NW, N, NE, E, SE, S, SW */
static const int NBX[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
static const int NBY[8] = {0, -1, -1, -1, 0, 1, 1, 1};

static inline int pix(const uint8_t *m, int rows, int cols, int x, int y)
{
    return (x >= 0 && y >= 0 && x < cols && y < rows) ? (m[(size_t)y * cols + x] != 0) : 0;
}

/* 4-connected flood of background from outside the image; ext[] = 1 for exterior background. */
static void flood_exterior(const uint8_t *bin, int rows, int cols, uint8_t *ext)
{
    size_t n = (size_t)rows * cols;
    int32_t *stack = (int32_t *)malloc(n * sizeof(int32_t));
    size_t sp = 0;
    memset(ext, 0, n);
#define PUSH(xx, yy)                                                      \
    do {                                                                  \
        size_t _i = (size_t)(yy) * cols + (xx);                           \
        if (!bin[_i] && !ext[_i]) {                                       \
            ext[_i] = 1;                                                  \
            stack[sp++] = (int32_t)_i;                                    \
        }                                                                 \
    } while (0)
    for (int x = 0; x < cols; ++x) {
        PUSH(x, 0);
        PUSH(x, rows - 1);
    }
    for (int y = 0; y < rows; ++y) {
        PUSH(0, y);
        PUSH(cols - 1, y);
    }
    while (sp) {
        int32_t i = stack[--sp];
        int x = i % cols, y = i / cols;
        if (x > 0) PUSH(x - 1, y);
        if (x < cols - 1) PUSH(x + 1, y);
        if (y > 0) PUSH(x, y - 1);
        if (y < rows - 1) PUSH(x, y + 1);
    }
#undef PUSH
    free(stack);
}

/* Follow the outer border that starts at (sx,sy) (its left neighbour is background) and
 * accumulate cv::moments' a00/a10/a01 over the closed polygon of visited pixel centres. */
static void follow_outer_border(const uint8_t *bin, int rows, int cols, int sx, int sy,
                                orc_contour_rec *rec)
{
    rec->first_index = sy * cols + sx;
    rec->m00 = rec->m10 = rec->m01 = 0.0;
    /* 3.1: from West, clockwise, find the first foreground neighbour */
    int d1 = -1;
    for (int k = 0; k < 8; ++k) {
        int d = k; /* start at W (index 0), clockwise */
        if (pix(bin, rows, cols, sx + NBX[d], sy + NBY[d])) {
            d1 = d;
            break;
        }
    }
    if (d1 < 0) { /* isolated pixel: one-point contour, all moments 0 */
        rec->npoints = 1;
        return;
    }
    /* generous bound on the number of border steps */
    size_t cap = 64, np = 0;
    int32_t *px = (int32_t *)malloc(cap * sizeof(int32_t));
    int32_t *py = (int32_t *)malloc(cap * sizeof(int32_t));
    int x1 = sx + NBX[d1], y1 = sy + NBY[d1];
    int x2 = x1, y2 = y1; /* (i2,j2) */
    int x3 = sx, y3 = sy; /* (i3,j3) */
    for (;;) {
        if (np == cap) {
            cap *= 2;
            px = (int32_t *)realloc(px, cap * sizeof(int32_t));
            py = (int32_t *)realloc(py, cap * sizeof(int32_t));
        }
        px[np] = x3;
        py[np] = y3;
        ++np;
        /* 3.3: from the element after (i2,j2), counter-clockwise around (i3,j3) */
        int d2 = 0;
        for (int k = 0; k < 8; ++k)
            if (x3 + NBX[k] == x2 && y3 + NBY[k] == y2) d2 = k;
        int x4 = 0, y4 = 0;
        for (int k = 1; k <= 8; ++k) {
            int d = (d2 - k + 16) % 8; /* counter-clockwise = decreasing index */
            if (pix(bin, rows, cols, x3 + NBX[d], y3 + NBY[d])) {
                x4 = x3 + NBX[d];
                y4 = y3 + NBY[d];
                break;
            }
        }
        /* 3.5 */
        if (x4 == sx && y4 == sy && x3 == x1 && y3 == y1) break;
        x2 = x3; y2 = y3;
        x3 = x4; y3 = y4;
    }
    /* cv::moments(contour) -> contourMoments(): Green's theorem, integer points in doubles */
    double a00 = 0, a10 = 0, a01 = 0;
    double xi_1 = px[np - 1], yi_1 = py[np - 1];
    for (size_t i = 0; i < np; ++i) {
        double xi = px[i], yi = py[i];
        double dxy = xi_1 * yi - xi * yi_1;
        a00 += dxy;
        a10 += dxy * (xi_1 + xi);
        a01 += dxy * (yi_1 + yi);
        xi_1 = xi;
        yi_1 = yi;
    }
    if (fabs(a00) > FLT_EPSILON) {
        double db1_2, db1_6;
        if (a00 > 0) {
            db1_2 = 0.5;
            db1_6 = 0.16666666666666666666666666666667;
        } else {
            db1_2 = -0.5;
            db1_6 = -0.16666666666666666666666666666667;
        }
        rec->m00 = a00 * db1_2;
        rec->m10 = a10 * db1_6;
        rec->m01 = a01 * db1_6;
    }
    rec->npoints = (int32_t)np;
    free(px);
    free(py);
}

/* 8-connected component labels, label = linear index of the raster-first pixel, -1 = bg.
 * (cv2.connectedComponents(connectivity=8) after canonical relabel.) */
ORC_API void orc_label8(const uint8_t *mask, size_t pitch, int rows, int cols, int32_t *labels)
{
    size_t n = (size_t)rows * cols;
    int32_t *stack = (int32_t *)malloc(n * sizeof(int32_t));
    for (size_t i = 0; i < n; ++i) labels[i] = -1;
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            size_t i = (size_t)y * cols + x;
            if (!mask[(size_t)y * pitch + x] || labels[i] >= 0) continue;
            size_t sp = 0;
            labels[i] = (int32_t)i;
            stack[sp++] = (int32_t)i;
            while (sp) {
                int32_t j = stack[--sp];
                int jx = j % cols, jy = j / cols;
                for (int k = 0; k < 8; ++k) {
                    int nx = jx + NBX[k], ny = jy + NBY[k];
                    if (nx < 0 || ny < 0 || nx >= cols || ny >= rows) continue;
                    size_t q = (size_t)ny * cols + nx;
                    if (mask[(size_t)ny * pitch + nx] && labels[q] < 0) {
                        labels[q] = (int32_t)i;
                        stack[sp++] = (int32_t)q;
                    }
                }
            }
        }
    free(stack);
}

/* External contours of a binary mask (non-zero = foreground) in RASTER order of their first
 * pixel (cv2 returns the reverse). Returns the count; fills at most max_recs records. */
ORC_API int orc_external_contours(const uint8_t *mask, size_t pitch, int rows, int cols,
                                  orc_contour_rec *recs, int max_recs)
{
    size_t n = (size_t)rows * cols;
    uint8_t *bin = (uint8_t *)calloc(n, 1);
    uint8_t *ext = (uint8_t *)malloc(n);
    uint8_t *seen = (uint8_t *)calloc(n, 1); /* pixels of components already reported/skipped */
    int32_t *stack = (int32_t *)malloc(n * sizeof(int32_t));
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) bin[(size_t)y * cols + x] = mask[(size_t)y * pitch + x] != 0;
    flood_exterior(bin, rows, cols, ext);
    int count = 0;
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            size_t i = (size_t)y * cols + x;
            if (!bin[i] || seen[i]) continue;
            /* raster-first pixel of a new 8-connected component: mark the whole component */
            size_t sp = 0;
            seen[i] = 1;
            stack[sp++] = (int32_t)i;
            while (sp) {
                int32_t j = stack[--sp];
                int jx = j % cols, jy = j / cols;
                for (int k = 0; k < 8; ++k) {
                    int nx = jx + NBX[k], ny = jy + NBY[k];
                    if (nx < 0 || ny < 0 || nx >= cols || ny >= rows) continue;
                    size_t q = (size_t)ny * cols + nx;
                    if (bin[q] && !seen[q]) {
                        seen[q] = 1;
                        stack[sp++] = (int32_t)q;
                    }
                }
            }
            /* outer border is external iff the pixel to the left is exterior background */
            int external = (x == 0) || ext[i - 1];
            if (!external) continue;
            if (count < max_recs) follow_outer_border(bin, rows, cols, x, y, &recs[count]);
            ++count;
        }
    free(bin);
    free(ext);
    free(seen);
    free(stack);
    return count;
}

ORC_API void orc_sift_contours(const uint8_t *mask, size_t pitch, int rows, int cols, double min_area,
                               double max_area, orc_detection *out)
{
    int cap = 1024;
    orc_contour_rec *recs = (orc_contour_rec *)malloc((size_t)cap * sizeof(orc_contour_rec));
    int n = orc_external_contours(mask, pitch, rows, cols, recs, cap);
    if (n > cap) {
        cap = n;
        recs = (orc_contour_rec *)realloc(recs, (size_t)cap * sizeof(orc_contour_rec));
        n = orc_external_contours(mask, pitch, rows, cols, recs, cap);
    }
    double object_area = 0;
    out->position_valid = 0;
    out->x = out->y = 0;
    out->n_components = n;
    /* cv2 order = reverse raster order; DetectorFunc.cpp:47-62 iterates that list */
    for (int i = n - 1; i >= 0; --i) {
        double a = recs[i].m00;
        if (a >= min_area && a < max_area && a > object_area) {
            out->x = recs[i].m10 / a;
            out->y = recs[i].m01 / a;
            out->position_valid = 1;
            object_area = a;
        }
    }
    out->area = object_area;
    free(recs);
}

/* The 2x2-cell identity the CUDA path uses, restated on the CPU so
 * the tests can check it against the border-following version above on arbitrary masks:
 * F = component + its holes; a 2x2 block of pixel centres with 4 pixels in F adds area 1 at
 * its centre, with exactly 3 adds area 1/2 at the mean of the three centres.
 * Outputs integer sums per external component in raster order:
 * s00 = 2*m00, s10 = 6*m10, s01 = 6*m01. */
ORC_API int orc_cell_moments(const uint8_t *mask, size_t pitch, int rows, int cols, int32_t *first_index,
                             int64_t *s00, int64_t *s10, int64_t *s01, int max_recs)
{
    size_t n = (size_t)rows * cols;
    uint8_t *bin = (uint8_t *)calloc(n, 1);
    uint8_t *ext = (uint8_t *)malloc(n);
    int32_t *lab = (int32_t *)malloc(n * sizeof(int32_t));
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) bin[(size_t)y * cols + x] = mask[(size_t)y * pitch + x] != 0;
    flood_exterior(bin, rows, cols, ext);
    for (size_t i = 0; i < n; ++i) bin[i] = !ext[i]; /* G = foreground + holes */
    orc_label8(bin, cols, rows, cols, lab);
    /* compact ids in raster order of first pixel */
    int32_t *slot = (int32_t *)malloc(n * sizeof(int32_t));
    int count = 0;
    for (size_t i = 0; i < n; ++i) {
        slot[i] = -1;
        if (lab[i] == (int32_t)i) {
            if (count < max_recs) {
                first_index[count] = (int32_t)i;
                s00[count] = s10[count] = s01[count] = 0;
            }
            slot[i] = count++;
        }
    }
    for (int y = 0; y + 1 < rows; ++y)
        for (int x = 0; x + 1 < cols; ++x) {
            int p[4] = {bin[(size_t)y * cols + x], bin[(size_t)y * cols + x + 1],
                        bin[(size_t)(y + 1) * cols + x], bin[(size_t)(y + 1) * cols + x + 1]};
            int xs[4] = {x, x + 1, x, x + 1}, ys[4] = {y, y, y + 1, y + 1};
            int c = p[0] + p[1] + p[2] + p[3];
            if (c < 3) continue;
            int L = -1;
            for (int k = 0; k < 4; ++k)
                if (p[k]) {
                    L = slot[lab[(size_t)ys[k] * cols + xs[k]]];
                    break;
                }
            if (L < 0 || L >= max_recs) continue;
            if (c == 4) {
                s00[L] += 2;
                s10[L] += 6 * (int64_t)x + 3;
                s01[L] += 6 * (int64_t)y + 3;
            } else {
                int64_t sx = 0, sy = 0;
                for (int k = 0; k < 4; ++k)
                    if (p[k]) {
                        sx += xs[k];
                        sy += ys[k];
                    }
                s00[L] += 1;
                s10[L] += sx;
                s01[L] += sy;
            }
        }
    free(bin);
    free(ext);
    free(lab);
    free(slot);
    return count;
}

/* ------------------------------------------------------------------------------------ */
/* framefilt bsub: /root/reference/src/framefilter/BackgroundSubtractor.cpp:87-100        */
/* first frame -> background; alpha>0: accumulateWeighted(frame, bg_f, alpha) then        */
/* convertTo(8U) (round-half-even, saturate); out = saturate(frame - bg).                 */
/* accumulateWeighted (non-IPP path): dst = src*a + dst*(1-a) in float, a,b from double.  */
/* ------------------------------------------------------------------------------------ */
typedef struct orc_bsub {
    int rows, cols, ch, set;
    double alpha;
    uint8_t *bg;
    float *bgf;
} orc_bsub;

ORC_API orc_bsub *orc_bsub_create(int rows, int cols, int ch, double alpha)
{
    orc_bsub *b = (orc_bsub *)calloc(1, sizeof(orc_bsub));
    b->rows = rows;
    b->cols = cols;
    b->ch = ch;
    b->alpha = alpha;
    b->bg = (uint8_t *)calloc((size_t)rows * cols * ch, 1);
    b->bgf = (float *)calloc((size_t)rows * cols * ch, sizeof(float));
    return b;
}

ORC_API void orc_bsub_destroy(orc_bsub *b)
{
    if (!b) return;
    free(b->bg);
    free(b->bgf);
    free(b);
}

ORC_API void orc_bsub_apply(orc_bsub *b, const uint8_t *in, size_t in_pitch, uint8_t *out,
                            size_t out_pitch)
{
    const int rowbytes = b->cols * b->ch;
    if (!b->set) {
        for (int y = 0; y < b->rows; ++y)
            for (int i = 0; i < rowbytes; ++i) {
                uint8_t v = in[(size_t)y * in_pitch + i];
                b->bg[(size_t)y * rowbytes + i] = v;
                b->bgf[(size_t)y * rowbytes + i] = (float)v;
            }
        b->set = 1;
    }
    if (b->alpha > 0.0) {
        const float a = (float)b->alpha, bb = 1.0f - a;
        for (int y = 0; y < b->rows; ++y)
            for (int i = 0; i < rowbytes; ++i) {
                size_t k = (size_t)y * rowbytes + i;
                float v = (float)in[(size_t)y * in_pitch + i] * a + b->bgf[k] * bb;
                b->bgf[k] = v;
                long r = lrintf(v); /* cvRound: round half to even */
                b->bg[k] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
            }
    }
    for (int y = 0; y < b->rows; ++y)
        for (int i = 0; i < rowbytes; ++i) {
            int d = (int)in[(size_t)y * in_pitch + i] - (int)b->bg[(size_t)y * rowbytes + i];
            out[(size_t)y * out_pitch + i] = (uint8_t)(d < 0 ? 0 : d);
        }
}

/* ------------------------------------------------------------------------------------ */
/* posifilt kalman + posicom mean.                              */
/* KalmanFilter2D::filter / initializeFilter / initializeStaticMatracies                */
/* (src/positionfilter/KalmanFilter2D.cpp:95-200) on cv::KalmanFilter(4, 2, 0, CV_64F): */
/*   predict(): statePre = A*statePost; errorCovPre = A*errorCovPost*A' + Q;            */
/*              statePost = statePre; errorCovPost = errorCovPre; returns statePre      */
/*   correct(z): S = H*errorCovPre*H' + R; gain' = pinv_SVD(S)*(H*errorCovPre);         */
/*              statePost = statePre + gain*(z - H*statePre);                           */
/*              errorCovPost = errorCovPre - gain*(H*errorCovPre)                       */
/* The reference's quirks are kept: the filter reports the PREDICTED (prior) state;     */
/* initializeFilter sets errorCovPre (overwritten by the next predict) and leaves       */
/* errorCovPost as it was; with the default timeout 0 the count test `0 >= 0` switches  */
/* the filter off again in the same call, so every output is invalid; an invalid sample */
/* inside the timeout re-uses the last valid measurement in correct().  Before the      */
/* first prediction the reported state is whatever `cv::Mat_<double>{4, 1, CV_64F}`     */
/* holds (KalmanFilter2D.h:63) -- the fill value 6.0 with the OpenCV 3.0-3.2 reading;   */
/* those outputs are flagged invalid and only the flags are part of the contract.       */
/* MeanPosition::combine (src/positioncombiner/MeanPosition.cpp:60-118).                */
/* ------------------------------------------------------------------------------------ */
typedef struct orc_position {
    int32_t position_valid, velocity_valid, heading_valid, reserved;
    double x, y, vx, vy, hx, hy;
} orc_position;

typedef struct orc_kalman {
    double dt, sig_accel, sig_noise;
    int found, not_found, not_found_thr;
    double A[4][4], H[2][4], Q[4][4], R[2][2];
    double state_pre[4], state_post[4], P_pre[4][4], P_post[4][4];
    double predicted[4], meas[2];
} orc_kalman;

ORC_API orc_kalman *orc_kalman_create(double dt, double timeout, double sig_accel, double sig_noise)
{
    orc_kalman *k = (orc_kalman *)calloc(1, sizeof(orc_kalman));
    k->dt = dt;
    k->sig_accel = sig_accel;
    k->sig_noise = sig_noise;
    k->not_found_thr = (int)(timeout / dt); /* KalmanFilter2D.cpp:74-76 */
    /* cv::KalmanFilter::init: A = I, Q = I, R = I, H = 0, everything else 0 */
    for (int i = 0; i < 4; ++i) k->A[i][i] = k->Q[i][i] = 1.0;
    k->R[0][0] = k->R[1][1] = 1.0;
    for (int i = 0; i < 4; ++i) k->predicted[i] = 6.0;
    k->meas[0] = k->meas[1] = 6.0;
    return k;
}
ORC_API void orc_kalman_destroy(orc_kalman *k) { free(k); }

static void kal_static(orc_kalman *k) /* initializeStaticMatracies, :162-200 */
{
    const double dt = k->dt, sa = k->sig_accel;
    memset(k->A, 0, sizeof k->A);
    for (int i = 0; i < 4; ++i) k->A[i][i] = 1.0;
    k->A[0][1] = dt;
    k->A[2][3] = dt;
    memset(k->H, 0, sizeof k->H);
    k->H[0][0] = 1.0;
    k->H[1][2] = 1.0;
    memset(k->Q, 0, sizeof k->Q);
    k->Q[0][0] = sa * sa * (dt * dt * dt * dt) / 4.0;
    k->Q[0][1] = sa * sa * (dt * dt * dt) / 2.0;
    k->Q[1][0] = sa * sa * (dt * dt * dt) / 2.0;
    k->Q[1][1] = sa * sa * (dt * dt);
    k->Q[2][2] = sa * sa * (dt * dt * dt * dt) / 4.0;
    k->Q[2][3] = sa * sa * (dt * dt * dt) / 2.0;
    k->Q[3][2] = sa * sa * (dt * dt * dt) / 2.0;
    k->Q[3][3] = sa * sa * (dt * dt);
    memset(k->R, 0, sizeof k->R);
    k->R[0][0] = k->R[1][1] = k->sig_noise * k->sig_noise;
}

/* pseudo-inverse of a symmetric 2x2 through its eigen-decomposition, singular values below
 * 2*DBL_EPSILON*sum treated as zero -- what cv::solve(..., DECOMP_SVD) (SVD::backSubst) does */
static void pinv_sym2(const double S[2][2], double Si[2][2])
{
    const double a = S[0][0], b = 0.5 * (S[0][1] + S[1][0]), d = S[1][1];
    const double tr = a + d, df = a - d;
    const double rad = sqrt(df * df + 4.0 * b * b);
    double l1 = 0.5 * (tr + rad), l2 = 0.5 * (tr - rad);
    double v1x, v1y; /* unit eigenvector of l1 */
    if (fabs(b) > 0.0) {
        v1x = l1 - d;
        v1y = b;
        const double n = sqrt(v1x * v1x + v1y * v1y);
        v1x /= n;
        v1y /= n;
    } else if (a >= d) {
        v1x = 1.0, v1y = 0.0;
    } else {
        v1x = 0.0, v1y = 1.0;
    }
    const double v2x = -v1y, v2y = v1x;
    const double thr = 2.0 * DBL_EPSILON * (fabs(l1) + fabs(l2));
    const double i1 = fabs(l1) > thr ? 1.0 / l1 : 0.0, i2 = fabs(l2) > thr ? 1.0 / l2 : 0.0;
    Si[0][0] = i1 * v1x * v1x + i2 * v2x * v2x;
    Si[0][1] = Si[1][0] = i1 * v1x * v1y + i2 * v2x * v2y;
    Si[1][1] = i1 * v1y * v1y + i2 * v2y * v2y;
}

ORC_API void orc_kalman_filter(orc_kalman *k, orc_position *p) /* KalmanFilter2D::filter, :95-145 */
{
    if (p->position_valid) {
        k->meas[0] = p->x;
        k->meas[1] = p->y;
        k->not_found = 0;
        if (!k->found) { /* initializeFilter, :147-170 */
            kal_static(k);
            memset(k->P_pre, 0, sizeof k->P_pre);
            for (int i = 0; i < 4; ++i) k->P_pre[i][i] = 1000.0;
            const double s0[4] = {k->meas[0], 0.0, k->meas[1], 0.0};
            memcpy(k->state_pre, s0, sizeof s0);
            memcpy(k->state_post, s0, sizeof s0);
        }
        k->found = 1;
    } else {
        k->not_found++;
    }
    if (k->not_found >= k->not_found_thr) k->found = 0;
    if (k->found) {
        double T[4][4];
        /* predict */
        for (int i = 0; i < 4; ++i) {
            double s = 0.0;
            for (int j = 0; j < 4; ++j) s += k->A[i][j] * k->state_post[j];
            k->state_pre[i] = s;
        }
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                double s = 0.0;
                for (int q = 0; q < 4; ++q) s += k->A[i][q] * k->P_post[q][j];
                T[i][j] = s;
            }
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                double s = 0.0;
                for (int q = 0; q < 4; ++q) s += T[i][q] * k->A[j][q];
                k->P_pre[i][j] = s + k->Q[i][j];
            }
        memcpy(k->state_post, k->state_pre, sizeof k->state_pre);
        memcpy(k->P_post, k->P_pre, sizeof k->P_pre);
        memcpy(k->predicted, k->state_pre, sizeof k->state_pre);
        /* correct */
        double HP[2][4], S[2][2], Si[2][2], G[4][2], innov[2];
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 4; ++j) {
                double s = 0.0;
                for (int q = 0; q < 4; ++q) s += k->H[i][q] * k->P_pre[q][j];
                HP[i][j] = s;
            }
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) {
                double s = 0.0;
                for (int q = 0; q < 4; ++q) s += HP[i][q] * k->H[j][q];
                S[i][j] = s + k->R[i][j];
            }
        pinv_sym2(S, Si);
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 2; ++j) /* gain = (Si * HP)' */
                G[i][j] = Si[j][0] * HP[0][i] + Si[j][1] * HP[1][i];
        for (int i = 0; i < 2; ++i) {
            double s = 0.0;
            for (int q = 0; q < 4; ++q) s += k->H[i][q] * k->state_pre[q];
            innov[i] = k->meas[i] - s;
        }
        for (int i = 0; i < 4; ++i) k->state_post[i] = k->state_pre[i] + (G[i][0] * innov[0] + G[i][1] * innov[1]);
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) k->P_post[i][j] = k->P_pre[i][j] - (G[i][0] * HP[0][j] + G[i][1] * HP[1][j]);
    }
    p->x = k->predicted[0];
    p->vx = k->predicted[1];
    p->y = k->predicted[2];
    p->vy = k->predicted[3];
    p->position_valid = p->velocity_valid = k->found ? 1 : 0;
}

/* MeanPosition::combine (MeanPosition.cpp:60-118); heading_anchor < 0: no heading generation */
ORC_API void orc_mean_combine(const orc_position *src, int n, int heading_anchor, orc_position *out)
{
    const double md = 1.0 / (double)n;
    memset(out, 0, sizeof *out);
    out->position_valid = out->velocity_valid = out->heading_valid = 1;
    for (int i = 0; i < n; ++i) {
        const orc_position *p = &src[i];
        if (p->position_valid) {
            out->x += md * p->x;
            out->y += md * p->y;
        } else
            out->position_valid = 0;
        if (p->velocity_valid) {
            out->vx += md * p->vx;
            out->vy += md * p->vy;
        } else
            out->velocity_valid = 0;
        if (heading_anchor >= 0) {
            if (out->position_valid) {
                out->hx += p->x - src[heading_anchor].x;
                out->hy += p->y - src[heading_anchor].y;
            } else
                out->heading_valid = 0;
        } else {
            if (p->heading_valid) {
                out->hx += p->hx;
                out->hy += p->hy;
            } else
                out->heading_valid = 0;
        }
    }
    if (out->heading_valid) {
        const double mag = sqrt(pow(out->hx, 2.0) + pow(out->hy, 2.0));
        out->hx = out->hx / mag;
        out->hy = out->hy / mag;
    }
}
