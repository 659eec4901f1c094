// Which scheduler (SM sub-partition) does warp w of a block land on?  One block
// of 16 warps per SM; only the warps in `mask` run an FP64-bound loop.  Warps
// that share a scheduler share its FP64 pipe, so the time tells the placement.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512) probe(double *sink, unsigned mask, int iters) {
    const int warp = threadIdx.x >> 5;
    if (!((mask >> warp) & 1u)) return;
    double a[8];
    for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 1e-3 + k;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = fma(a[k], m, c);
    double s = 0; for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 12345.678) sink[0] = s;
}
int main() {
    double *sink; cudaMalloc(&sink, 8);
    unsigned masks[] = {0x0001, 0x000F, 0x1111, 0x0033, 0x0303, 0x0101, 0x0005, 0x0011, 0x00FF, 0x3333, 0x0F0F, 0x3FFF, 0xFFFF, 0x0FFF, 0xCFFF};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (unsigned mask : masks) {
        probe<<<148, 512>>>(sink, mask, 1000);
        cudaEventRecord(e0);
        probe<<<148, 512>>>(sink, mask, 200000);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        int nw = __builtin_popcount(mask);
        printf("mask 0x%04x (%2d warps): %.3f ms  -> %.2f x the one-warp time per... cycles/DFMA/warp %.2f\n", mask, nw, ms, 0.0, ms * 1e-3 * 1.965e9 / (200000.0 * 8));
    }
    return 0;
}
