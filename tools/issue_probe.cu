// Issue-port probe (development tool): does a non-FP64 instruction issue "for
// free" in the shadow of a DFMA (which holds the 16-lane FP64 pipe for two
// cycles), or does every DFMA cost the scheduler two issue cycles?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/issue_probe.cu -o build/issue_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int NF64, int NOTHER, int KIND>
__global__ void probe(double *sink, int iters) {
    double a[8]; float f[8]; int q[8];
    for (int k = 0; k < 8; ++k) { a[k] = threadIdx.x * 1e-3 + k; f[k] = threadIdx.x + k; q[k] = threadIdx.x + k; }
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < NF64; ++k) a[k] = fma(a[k], m, c);
#pragma unroll
            for (int k = 0; k < NOTHER; ++k) {
                if (KIND == 0) f[k] = fmaf(f[k], 1.0001f, 0.5f);            // FFMA (fp32 pipe)
                if (KIND == 1) q[k] = q[k] * 3 + 1;                          // IMAD
                if (KIND == 2) q[k] = (q[k] ^ (q[k] >> 3)) + 7;              // ALU (LOP3/SHF/IADD)
            }
        }
    }
    double s = 0; for (int k = 0; k < 8; ++k) s += a[k] + f[k] + q[k];
    if (s == 12345.678) sink[0] = s;
}

template <int NF64, int NOTHER, int KIND>
void run(const char *name, int warps_per_sm) {
    double *sink; cudaMalloc(&sink, 8);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dim3 grid(148), block(warps_per_sm * 32);
    probe<NF64, NOTHER, KIND><<<grid, block>>>(sink, 100);
    cudaEventRecord(e0);
    probe<NF64, NOTHER, KIND><<<grid, block>>>(sink, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9;             // assumes max clock
    const double per_iter = cyc / iters / 4.0 / (warps_per_sm / 4.0);   // cycles per inner round per warp-slot
    printf("%-28s warps/sched %2d  f64 %d other %d : %.2f cycles per round per scheduler-warp (f64*2=%d, f64*2+other=%d, max(f64*2,f64+other)=%d)\n",
           name, warps_per_sm / 4, NF64, NOTHER, per_iter, NF64 * 2, NF64 * 2 + NOTHER,
           NF64 * 2 > NF64 + NOTHER ? NF64 * 2 : NF64 + NOTHER);
    cudaFree(sink);
}

int main() {
    for (int w : {16, 32}) {
        run<8, 0, 0>("dfma only", w);
        run<8, 4, 0>("dfma + ffma", w);
        run<8, 8, 0>("dfma + ffma", w);
        run<8, 4, 1>("dfma + imad", w);
        run<8, 8, 1>("dfma + imad", w);
        run<8, 8, 2>("dfma + alu x3", w);
        run<4, 8, 0>("dfma + ffma", w);
        run<0, 8, 0>("ffma only", w);
        run<0, 8, 1>("imad only", w);
    }
    return 0;
}
