// common.cuh -- shared device/host helpers for liboatgpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/oatgpu.h"

namespace oat {

// ---- bit-packed binary image ---------------------------------------------------------------
// Every binary mask on the detect tail is 1 bit/pixel: word j of row y holds pixels
// x = 32*j .. 32*j+31, LSB = lowest x.  Rows are padded to a whole number of words
// (pitch_px = 32 * wpr) and padding bits are always 0 ("canonical").
struct BitGeom {
    int rows, cols;
    int wpr;  // words per row
    __host__ __device__ int pitch_px() const { return wpr * 32; }
    __host__ __device__ uint32_t valid_mask(int j) const
    {
        int rem = cols - 32 * j;
        return rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
    }
};

static inline int div_up(int a, int b) { return (a + b - 1) / b; }

// ---- streaming loads/stores ----------------------------------------------------------------
// Frame ingress/egress is touched exactly once per frame: read through the non-coherent path
// without allocating in L1, and mark evict-first so it does not displace the GMM state (which
// IS re-read next frame and, at 1080p with few live modes, fits the 126 MB L2).
__device__ __forceinline__ uint32_t ld_stream_u32(const void *p)
{
    return __ldcs(reinterpret_cast<const unsigned int *>(p));  // ld.global.cs: evict-first
}
__device__ __forceinline__ uint8_t ld_stream_u8(const void *p)
{
    return __ldcs(reinterpret_cast<const unsigned char *>(p));
}
__device__ __forceinline__ void st_stream_u32(void *p, uint32_t v)
{
    __stcs(reinterpret_cast<unsigned int *>(p), v);  // st.global.cs
}
// GMM state: read-modify-write by the same thread; loads are served by L2, never by L1 (ld.cg) -- there
// is no reuse inside a frame, the planes can stay L2-resident between frames, and a launch that overlaps
// its predecessor (tile-granular chaining, mog_pipe.cuh) must not see a line an earlier launch left in L1.
__device__ __forceinline__ float4 ld_state_f4(const float *p)
{
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_state_f4(float *p, float4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float ld_state_f1(const float *p)
{
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_state_f1(float *p, float v)
{
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

}  // namespace oat
