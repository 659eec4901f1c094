// FP64 pipe micro-probe (development tool): dependent-issue latency and
// per-SMSP throughput of DFMA on this GPU, measured with clock64().
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double *out, long long *cyc, int iters) {
    double x[ILP];
    for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x * 1e-3 + k;
    const double a = 1.0000001, b = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
    }
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < ILP; ++k) s += x[k];
    if (threadIdx.x == 0 && blockIdx.x == 0) { *cyc = t1 - t0; }
    if (s == 1.2345) out[0] = s;
}
template <int ILP> void run(int warps) {
    double *d; long long *c, h; cudaMalloc(&d, 8); cudaMalloc(&c, 8);
    int iters = 4096;
    chain<ILP><<<1, 32 * warps>>>(d, c, iters);
    chain<ILP><<<1, 32 * warps>>>(d, c, iters);
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("warps/SM=%2d (per SMSP %.2f) ILP=%d: %.2f cycles per DFMA per warp, %.2f DFMA warp-inst/cycle/SMSP\n",
           warps, warps / 4.0, ILP, (double)h / iters / ILP, (double)iters * ILP * warps / 4.0 / h);
}
int main() {
    run<1>(1); run<2>(1); run<4>(1); run<8>(1);
    run<1>(4); run<1>(8); run<1>(16); run<1>(32);
    run<2>(4); run<2>(8); run<2>(16);
    run<4>(4); run<4>(8);
    return 0;
}
