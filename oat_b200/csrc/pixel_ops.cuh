// pixel_ops.cuh -- the small elementwise kernels around the fused path: stand-alone
// BGR->HSV (framefilt col), inRange -> bits (posidet hsv on an HSV frame), u8 mask <-> bits,
// framefilt bsub, the synthetic frame generator, and utility reductions.
#pragma once
#include "common.cuh"
#include "mog_fused.cuh"

namespace oat {

// framefilt col -C HSV  (src/framefilter/ColorConvert.cpp:101-107).  One thread per pixel.
__global__ void bgr2hsv_kernel(const uint8_t *__restrict__ bgr, size_t in_pitch, uint8_t *__restrict__ hsv,
                               size_t out_pitch, int rows, int cols, const int *__restrict__ lut_g)
{
    __shared__ int lut[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) lut[i] = lut_g[i];
    __syncthreads();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)rows * cols) return;
    const int y = (int)(t / cols), x = (int)(t % cols);
    const uint8_t *s = bgr + (size_t)y * in_pitch + 3 * x;
    int h, sa, v;
    bgr2hsv_px(s[0], s[1], s[2], lut, h, sa, v);
    uint8_t *d = hsv + (size_t)y * out_pitch + 3 * x;
    d[0] = (uint8_t)h;
    d[1] = (uint8_t)sa;
    d[2] = (uint8_t)v;
}

// cv::inRange on a 3-channel frame -> 1 bit/pixel (HSVDetector.cpp:146-149). One warp per word.
__global__ void inrange_bits_kernel(const uint8_t *__restrict__ img, size_t pitch, BitGeom g, int lo0, int lo1,
                                    int lo2, int hi0, int hi1, int hi2, uint32_t *__restrict__ bits)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int ppx = g.pitch_px();
    const int y = (int)(t / ppx), x = (int)(t % ppx);
    bool in = false;
    if (y < g.rows && x < g.cols) {
        const uint8_t *s = img + (size_t)y * pitch + 3 * x;
        const int a = s[0], b = s[1], c = s[2];
        in = (lo0 <= a) & (a <= hi0) & (lo1 <= b) & (b <= hi1) & (lo2 <= c) & (c <= hi2);
    }
    const uint32_t w = __ballot_sync(0xffffffffu, in);
    if ((threadIdx.x & 31) == 0 && y < g.rows) bits[(size_t)y * g.wpr + (x >> 5)] = w;
}

// cv::inRange on a 1-channel frame -> bits (posidet thresh, src/positiondetector/SimpleThreshold.cpp:169-172)
__global__ void inrange1_bits_kernel(const uint8_t *__restrict__ img, size_t pitch, BitGeom g, int lo, int hi,
                                     uint32_t *__restrict__ bits)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int ppx = g.pitch_px();
    const int y = (int)(t / ppx), x = (int)(t % ppx);
    bool in = false;
    if (y < g.rows && x < g.cols) {
        const int v = img[(size_t)y * pitch + x];
        in = (lo <= v) & (v <= hi);
    }
    const uint32_t w = __ballot_sync(0xffffffffu, in);
    if ((threadIdx.x & 31) == 0 && y < g.rows) bits[(size_t)y * g.wpr + (x >> 5)] = w;
}

// framefilt thresh (src/framefilter/Threshold.cpp:67-81): grey = cvtColor(BGR2GRAY) (8-bit fixed point:
// (3735 b + 19235 g + 9798 r + 16384) >> 15), keep = inRange(grey, lo, hi), frame.setTo(0, keep == 0).
// framefilt mask (src/framefilter/FrameMasker.cpp:71-75): frame.setTo(0, roi == 0).  One thread per pixel.
__global__ void keep_where_kernel(const uint8_t *__restrict__ in, size_t in_pitch, uint8_t *__restrict__ out, size_t out_pitch,
                                  int rows, int cols, int ch, const uint8_t *__restrict__ roi, size_t roi_pitch, int lo, int hi)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)rows * cols) return;
    const int y = (int)(t / cols), x = (int)(t % cols);
    const uint8_t *s = in + (size_t)y * in_pitch + (size_t)ch * x;
    bool keep;
    if (roi) {
        keep = roi[(size_t)y * roi_pitch + x] != 0;
    } else {
        const int grey = ch == 3 ? ((3735 * s[0] + 19235 * s[1] + 9798 * s[2] + 16384) >> 15) : s[0];
        keep = (lo <= grey) & (grey <= hi);
    }
    uint8_t *d = out + (size_t)y * out_pitch + (size_t)ch * x;
    for (int c = 0; c < ch; ++c) d[c] = keep ? s[c] : 0;
}

// posidet diff (src/positiondetector/DifferenceDetector.cpp:154-173): bit = |frame - last| > threshold (cv::absdiff +
// cv::threshold THRESH_BINARY), last <- frame.  first != 0: the reference's first call has no previous image and sifts
// the raw frame itself (threshold_frame_ = frame.clone()): bit = frame != 0.
__global__ void absdiff_bits_kernel(const uint8_t *__restrict__ img, size_t pitch, uint8_t *__restrict__ last, BitGeom g,
                                    int thresh, int first, uint32_t *__restrict__ bits)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int ppx = g.pitch_px();
    const int y = (int)(t / ppx), x = (int)(t % ppx);
    bool in = false;
    if (y < g.rows && x < g.cols) {
        const int v = img[(size_t)y * pitch + x];
        uint8_t *l = last + (size_t)y * g.cols + x;
        const int d = v > *l ? v - *l : *l - v;
        in = first ? (v != 0) : (d > thresh);
        *l = (uint8_t)v;
    }
    const uint32_t w = __ballot_sync(0xffffffffu, in);
    if ((threadIdx.x & 31) == 0 && y < g.rows) bits[(size_t)y * g.wpr + (x >> 5)] = w;
}
// cv::blur's BORDER_REFLECT_101 with an EVEN k x k box (anchor k/2): at x = 0 (y = 0) the mirrored sample -k/2 -> +k/2
// falls outside the plain window, so column 0 (row 0) additionally sees column (row) k/2.  vertical == 0: fix column 0
// of `out` from `in` (the un-dilated rows); vertical != 0: fix row 0 of `out` from row k/2 of `in`.
__global__ void reflect_even_fix_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, BitGeom g, int a, int vertical)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (!vertical) {
        if (t >= g.rows || a >= g.cols) return;
        const uint32_t bit = (in[(size_t)t * g.wpr + (a >> 5)] >> (a & 31)) & 1u;
        if (bit) out[(size_t)t * g.wpr] |= 1u;
    } else {
        if (t >= g.wpr || a >= g.rows) return;
        out[t] |= in[(size_t)a * g.wpr + t];
    }
}

// u8 mask (non-zero = foreground) -> bits
__global__ void mask_to_bits_kernel(const uint8_t *__restrict__ mask, size_t pitch, BitGeom g,
                                    uint32_t *__restrict__ bits)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int ppx = g.pitch_px();
    const int y = (int)(t / ppx), x = (int)(t % ppx);
    bool in = false;
    if (y < g.rows && x < g.cols) in = mask[(size_t)y * pitch + x] != 0;
    const uint32_t w = __ballot_sync(0xffffffffu, in);
    if ((threadIdx.x & 31) == 0 && y < g.rows) bits[(size_t)y * g.wpr + (x >> 5)] = w;
}

// bits -> u8 {0,255}
__global__ void bits_to_mask_kernel(const uint32_t *__restrict__ bits, BitGeom g, uint8_t *__restrict__ mask,
                                    size_t pitch)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int ppx = g.pitch_px();
    const int y = (int)(t / ppx), x = (int)(t % ppx);
    if (y < g.rows && x < g.cols)
        mask[(size_t)y * pitch + x] = ((bits[(size_t)y * g.wpr + (x >> 5)] >> (x & 31)) & 1u) ? 255 : 0;
}

// framefilt bsub (src/framefilter/BackgroundSubtractor.cpp:87-100), one thread per byte.
// first != 0: this frame becomes the background.  alpha > 0: accumulateWeighted then
// convertTo(8U) (round half to even, saturate).  out = saturate(frame - background).
__global__ void bsub_kernel(const uint8_t *__restrict__ in, size_t in_pitch, uint8_t *__restrict__ out,
                            size_t out_pitch, uint8_t *__restrict__ bg, float *__restrict__ bgf, int rows,
                            int rowbytes, int first, float a, float b, int adapt)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)rows * rowbytes) return;
    const int y = (int)(t / rowbytes), i = (int)(t % rowbytes);
    const int v = in[(size_t)y * in_pitch + i];
    int bgv;
    float f;
    if (first) {
        bgv = v;
        f = (float)v;
    } else {
        bgv = bg[t];
        f = bgf[t];
    }
    if (adapt) {
        f = __fadd_rn(__fmul_rn((float)v, a), __fmul_rn(f, b));
        int r = __float2int_rn(f);
        bgv = r < 0 ? 0 : (r > 255 ? 255 : r);
    }
    if (first || adapt) {
        bg[t] = (uint8_t)bgv;
        bgf[t] = f;
    }
    const int d = v - bgv;
    out[(size_t)y * out_pitch + i] = (uint8_t)(d < 0 ? 0 : d);
}

__global__ void bsub_set_bg_kernel(const uint8_t *__restrict__ img, size_t pitch, uint8_t *__restrict__ bg,
                                   float *__restrict__ bgf, int rows, int rowbytes)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)rows * rowbytes) return;
    const int y = (int)(t / rowbytes), i = (int)(t % rowbytes);
    const uint8_t v = img[(size_t)y * pitch + i];
    bg[t] = v;
    bgf[t] = (float)v;
}

// ---- synthetic stream; same integer arithmetic as oracle/synth.py ---------
__device__ __forceinline__ uint32_t fmix32(uint32_t h)
{
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}
__host__ __device__ inline uint32_t fmix32_hd(uint32_t h)
{
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

__global__ void synth_kernel(uint8_t *__restrict__ dst, size_t pitch, int rows, int cols, uint32_t kbg,
                             uint32_t knz, int cx, int cy, int r, int has_disc)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)rows * cols) return;
    const int y = (int)(t / cols), x = (int)(t % cols);
    const int dx = x - cx, dy = y - cy;
    const bool in_disc = has_disc && (dx * dx + dy * dy <= r * r);
    uint8_t *d = dst + (size_t)y * pitch + 3 * x;
    const uint32_t disc[3] = {40u, 220u, 60u};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const uint32_t idx = ((uint32_t)y * (uint32_t)cols + (uint32_t)x) * 3u + (uint32_t)c;
        const int bg = 40 + (int)(fmix32(idx ^ kbg) % 81u);
        const int nz = (int)(fmix32(idx ^ knz) % 7u) - 3;
        d[c] = in_disc ? (uint8_t)disc[c] : (uint8_t)(bg + nz);
    }
}

// sum of live modes over the image (padding excluded) -> *out (must be zeroed first)
__global__ void live_modes_kernel(const uint8_t *__restrict__ nmodes, BitGeom g, unsigned long long *out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int ppx = g.pitch_px();
    const int y = (int)(t / ppx), x = (int)(t % ppx);
    unsigned v = 0;
    if (y < g.rows && x < g.cols) v = nmodes[t];
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, (unsigned long long)v);
}

// state egress in OpenCV's layout (test/diagnostic): one thread per pixel
__global__ void state_export_kernel(const float *__restrict__ state, size_t plane, const uint8_t *__restrict__ nmodes,
                                    BitGeom g, int K, uint8_t *__restrict__ modes_out, float *__restrict__ weight,
                                    float *__restrict__ variance, float *__restrict__ mean)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)g.rows * g.cols) return;
    const int y = (int)(t / g.cols), x = (int)(t % g.cols);
    const size_t p = (size_t)y * g.pitch_px() + x;
    if (modes_out) modes_out[t] = nmodes[p];
    for (int m = 0; m < K; ++m) {
        if (weight) weight[t * K + m] = state[(size_t)(m * 5 + 0) * plane + p];
        if (variance) variance[t * K + m] = state[(size_t)(m * 5 + 1) * plane + p];
        if (mean)
            for (int c = 0; c < 3; ++c) mean[(t * K + m) * 3 + c] = state[(size_t)(m * 5 + 2 + c) * plane + p];
    }
}

// L2 flush: overwrite a buffer larger than L2
// Device-to-device staging of a frame into another pitch (the stream's OAT_STREAM_COPY of a device frame, the
// re-pitch of a linearly uploaded host frame): the copy engine needs ~6-10 us per such copy on this part -- host time
// as well as device time (tools/copy_rate.py) -- which is most of a 13 us frame period.  Rows and pitches are multiples
// of 4 bytes; 16-byte accesses when everything is 16-byte aligned, 4-byte ones otherwise.  Grid-stride, no shared
// memory, 32 registers: co-resides with the resident fused kernel's CTAs.
template <typename V>
__global__ void __launch_bounds__(256) repitch_kernel(uint8_t *__restrict__ dst, size_t dpitch, const uint8_t *__restrict__ src, size_t spitch,
                                                      unsigned row_units, unsigned rows)
{
    const unsigned long long total = (unsigned long long)row_units * rows;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned y = (unsigned)(i / row_units), u = (unsigned)(i - (unsigned long long)y * row_units);
        const V v = __ldcs(reinterpret_cast<const V *>(src + (size_t)y * spitch) + u);
        reinterpret_cast<V *>(dst + (size_t)y * dpitch)[u] = v;
    }
}

__global__ void flush_kernel(uint4 *buf, size_t n, uint32_t v)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        buf[i] = make_uint4(v, v, v, v);
}

}  // namespace oat
