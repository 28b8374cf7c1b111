// Micro-benchmark of the pipelined fused kernel alone (development tool, not the judged bench).
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo [-DPIPE_...] pipe_bench.cu -o pipe_bench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../oat_b200/csrc/mog_pipe.cuh"
using namespace oat;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
__global__ void init_state(float *state, size_t plane, uint8_t *nm, const uint8_t *bgr)
{
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plane) return;
    state[p] = 1.f; state[plane + p] = 15.f;
    for (int c = 0; c < 3; ++c) state[(2 + c) * plane + p] = (float)bgr[3 * p + c];
    nm[p] = 1;
}
__global__ void gen(uint8_t *bgr, size_t n, uint32_t seed, int noise)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h = (uint32_t)i * 2654435761u; h ^= h >> 15; h *= 0x85ebca6bu; h ^= h >> 13;
    uint32_t g = ((uint32_t)i + seed * 7919u) * 0x9e3779b9u; g ^= g >> 16; g *= 0xc2b2ae35u; g ^= g >> 13;
    bgr[i] = (uint8_t)(40 + h % 81 + (noise ? (int)(g % 7) - 3 : 0));
}
__global__ void flushk(uint4 *b, size_t n, uint32_t v) { for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = make_uint4(v, v, v, v); }
__global__ void __launch_bounds__(PIPE_THREADS, 2) nullk(const __grid_constant__ PipeArgs pa) { extern __shared__ uint8_t sm[]; if (pa.ntiles < 0) sm[threadIdx.x] = 1; }
int main(int argc, char **argv)
{
    int rows = argc > 1 ? atoi(argv[1]) : 1080, cols = argc > 2 ? atoi(argv[2]) : 1920, iters = argc > 3 ? atoi(argv[3]) : 100;
    int cps = argc > 4 ? atoi(argv[4]) : 2;
    float alpha = argc > 5 ? atof(argv[5]) : 0.01f;
    size_t plane = (size_t)rows * cols;
    float *state; uint8_t *nm, *bgr[9]; uint32_t *bits; uint4 *fl; int *lut;
    CK(cudaMalloc(&state, plane * 25 * 4)); CK(cudaMalloc(&nm, plane)); CK(cudaMalloc(&bits, plane / 8 + 64));
    CK(cudaMalloc(&lut, 2048)); CK(cudaMemset(lut, 0, 2048));
    for (int i = 0; i < 9; ++i) { CK(cudaMalloc(&bgr[i], plane * 3)); gen<<<(plane * 3 + 255) / 256, 256>>>(bgr[i], plane * 3, i, i > 0); }
    init_state<<<(plane + 255) / 256, 256>>>(state, plane, nm, bgr[0]);
    size_t fbytes = (size_t)256 << 20; CK(cudaMalloc(&fl, fbytes));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    PipeArgs pa{}; FusedArgs &a = pa.f;
    a.in_pitch = 3 * cols; a.rows = rows; a.cols = cols; a.wpr = cols / 32; a.state = state; a.plane = plane; a.nmodes = nm;
    a.c.aT = alpha; a.c.a1 = 1.f - alpha; a.c.prune = (float)(-(double)alpha * 0.05); a.c.Tb = 16; a.c.TB = 0.9f; a.c.Tg = 9; a.c.varInit = 15; a.c.varMin = 4; a.c.varMax = 75; a.c.tau = 0.5f;
    a.c.detect_shadows = 1; a.c.shadow_value = 127; a.do_hsv = 1; a.lo[0] = 40; a.hi[0] = 80; a.lo[1] = 100; a.hi[1] = 256; a.lo[2] = 100; a.hi[2] = 256;
    a.thr_bits = bits; a.hsv_lut = lut;
    pa.ntiles = (int)((plane + PIPE_TILE - 1) / PIPE_TILE); pa.div_magic = 0xffffffffffffffffull / (unsigned long long)cols + 1ull; pa.zero_in = 0;
    int grid = pa.ntiles < cps * prop.multiProcessorCount ? pa.ntiles : cps * prop.multiProcessorCount; pa.grid_tiles = grid;
    auto kern = mog_pipe_kernel<5, false, true>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM_BYTES));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<cudaEvent_t> E0(iters + 5), E1(iters + 5);
    for (auto &e : E0) cudaEventCreate(&e);
    for (auto &e : E1) cudaEventCreate(&e);
    CK(cudaFuncSetAttribute(nullk, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM_BYTES));
    if (getenv("CARVE")) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CK(cudaFuncSetAttribute(flushk, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CK(cudaFuncSetAttribute(nullk, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
    for (int which = 0; which < 2; ++which)
    for (int mode = 0; mode < 2; ++mode) {  // 0: L2 flushed before every launch, 1: back to back; all queued, one sync
        for (int it = 0; it < iters + 5; ++it) {
            a.bgr = bgr[1 + it % 8];
            if (mode == 0) flushk<<<1184, 256>>>(fl, fbytes / 16, it);
            cudaEventRecord(E0[it]);
            if (which == 0) kern<<<grid, PIPE_THREADS, PIPE_SMEM_BYTES>>>(pa); else nullk<<<grid, PIPE_THREADS, PIPE_SMEM_BYTES>>>(pa);
            cudaEventRecord(E1[it]);
        }
        CK(cudaDeviceSynchronize());
        double tot = 0, best = 1e9;
        for (int it = 5; it < iters + 5; ++it) { float ms; cudaEventElapsedTime(&ms, E0[it], E1[it]); tot += ms; if (ms < best) best = ms; }
        double us = 1e3 * tot / iters;
        printf("%s %dx%d stages=%d cps=%d grid=%d %s (queued): mean %.2f us best %.2f us -> %.0f GB/s (45 B/px)\n", which ? "NULL " : "PIPE ", cols, rows, PIPE_STAGES, cps, grid,
               mode == 0 ? "flushed" : "hot", us, best * 1e3, 45.0 * plane / us / 1e3);
    }
    {   // S states round-robin, back to back, ONE event pair: true HBM-fed steady state incl. launch gaps
        const int S = 8;
        std::vector<float *> st(S); std::vector<uint8_t *> nms(S);
        for (int i = 0; i < S; ++i) { CK(cudaMalloc(&st[i], plane * 5 * 4)); CK(cudaMalloc(&nms[i], plane)); init_state<<<(plane + 255) / 256, 256>>>(st[i], plane, nms[i], bgr[0]); }
        CK(cudaDeviceSynchronize());
        for (int rep = 0; rep < 2; ++rep) {
            const int n = 10 * S;
            cudaEventRecord(e0);
            for (int it = 0; it < n; ++it) { PipeArgs q = pa; q.f.state = st[it % S]; q.f.nmodes = nms[it % S]; q.f.bgr = bgr[1 + it % 8]; kern<<<grid, PIPE_THREADS, PIPE_SMEM_BYTES>>>(q); }
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("BACK2BACK %d states x %d launches, one event pair: %.2f us per launch -> %.0f GB/s (45 B/px)\n", S, n, 1e3 * ms / n, 45.0 * plane / (1e3 * ms / n) / 1e3);
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
