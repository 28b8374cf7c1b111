#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
struct Big { char b[512]; };
__global__ void nullk(const __grid_constant__ Big p, int flag) { extern __shared__ uint8_t sm[]; if (flag < 0) sm[threadIdx.x] = p.b[3]; }
__global__ void nullsmall(int flag) { extern __shared__ uint8_t sm[]; if (flag < 0) sm[threadIdx.x] = 1; }
int main()
{
    const int iters = 200;
    std::vector<cudaEvent_t> E0(iters), E1(iters);
    for (auto &e : E0) cudaEventCreate(&e);
    for (auto &e : E1) cudaEventCreate(&e);
    CK(cudaFuncSetAttribute(nullk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(nullsmall, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    Big big{};
    struct Cfg { int grid, threads, smem, bigarg; } cfgs[] = {
        {296, 288, 98816, 1}, {296, 288, 98816, 0}, {296, 288, 0, 0}, {296, 256, 0, 0}, {148, 288, 98816, 0}, {148, 1024, 0, 0}, {1, 32, 0, 0},
        {592, 160, 49664, 0}, {4050, 128, 0, 0}, {135, 256, 24576, 0}};
    for (auto c : cfgs) {
        for (int it = 0; it < iters; ++it) {
            cudaEventRecord(E0[it]);
            if (c.bigarg) nullk<<<c.grid, c.threads, c.smem>>>(big, 1); else nullsmall<<<c.grid, c.threads, c.smem>>>(1);
            cudaEventRecord(E1[it]);
        }
        CK(cudaDeviceSynchronize());
        double tot = 0; float best = 1e9;
        for (int it = 5; it < iters; ++it) { float ms; cudaEventElapsedTime(&ms, E0[it], E1[it]); tot += ms; if (ms < best) best = ms; }
        printf("null grid=%d threads=%d smem=%d bigarg=%d: mean %.2f us best %.2f us\n", c.grid, c.threads, c.smem, c.bigarg, 1e3 * tot / (iters - 5), best * 1e3);
    }
    return 0;
}
