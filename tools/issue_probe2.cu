// Issue-port probe 2: alternating DFMA / X patterns, X = FFMA, IMAD, LOP3, FSEL, ISETP+SEL, LDS, MUFU.
#include <cstdio>
#include <cuda_runtime.h>

#define D(k) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[k]) : "d"(m), "d"(c));
#define XF(k) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(fm), "f"(fc));
#define XI(k) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(q[k]) : "r"(im), "r"(ic));
#define XL(k) asm volatile("xor.b32 %0, %0, %1;" : "+r"(q[k]) : "r"(im));
#define XS(k) asm volatile("selp.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(fc), "r"(pp));   // needs pred; use setp inside
#define XM(k) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[k]));
#define XLD(k) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f[k]) : "r"(saddr + 4 * k));

template <int PAT>
__global__ void probe(double *sink, int iters) {
    __shared__ float sm[256];
    sm[threadIdx.x & 255] = threadIdx.x;
    __syncthreads();
    unsigned saddr = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 4;
    double a[8]; float f[8]; int q[8];
    for (int k = 0; k < 8; ++k) { a[k] = threadIdx.x * 1e-3 + k; f[k] = threadIdx.x * 1e-3f + k; q[k] = threadIdx.x + k; }
    double m = 1.0000001, c = 1e-9; float fm = 1.0001f, fc = 0.5f; int im = 3, ic = 1;
    for (int it = 0; it < iters; ++it) {
        if (PAT == 0) { D(0) D(1) D(2) D(3) D(4) D(5) D(6) D(7) }
        if (PAT == 1) { D(0) XF(0) D(1) XF(1) D(2) XF(2) D(3) XF(3) D(4) XF(4) D(5) XF(5) D(6) XF(6) D(7) XF(7) }
        if (PAT == 2) { D(0) D(1) D(2) D(3) D(4) D(5) D(6) D(7) XF(0) XF(1) XF(2) XF(3) XF(4) XF(5) XF(6) XF(7) }
        if (PAT == 3) { D(0) XI(0) D(1) XI(1) D(2) XI(2) D(3) XI(3) D(4) XI(4) D(5) XI(5) D(6) XI(6) D(7) XI(7) }
        if (PAT == 4) { D(0) XL(0) D(1) XL(1) D(2) XL(2) D(3) XL(3) D(4) XL(4) D(5) XL(5) D(6) XL(6) D(7) XL(7) }
        if (PAT == 5) { D(0) XL(0) XI(0) D(1) XL(1) XI(1) D(2) XL(2) XI(2) D(3) XL(3) XI(3) D(4) XL(4) XI(4) D(5) XL(5) XI(5) D(6) XL(6) XI(6) D(7) XL(7) XI(7) }
        if (PAT == 6) { D(0) XM(0) D(1) XM(1) D(2) XM(2) D(3) XM(3) D(4) XM(4) D(5) XM(5) D(6) XM(6) D(7) XM(7) }
        if (PAT == 7) { D(0) XLD(0) D(1) XLD(1) D(2) XLD(2) D(3) XLD(3) D(4) XLD(4) D(5) XLD(5) D(6) XLD(6) D(7) XLD(7) }
        if (PAT == 8) { D(0) XF(0) XF(1) D(1) XF(2) XF(3) D(2) XF(4) XF(5) D(3) XF(6) XF(7) D(4) XF(0) XF(1) D(5) XF(2) XF(3) D(6) XF(4) XF(5) D(7) XF(6) XF(7) }
        if (PAT == 9) { D(0) D(1) XF(0) D(2) D(3) XF(1) D(4) D(5) XF(2) D(6) D(7) XF(3) }
        if (PAT == 10) { XL(0) XL(1) XL(2) XL(3) XL(4) XL(5) XL(6) XL(7) }
        if (PAT == 11) { D(0) XL(0) XL(1) D(1) XL(2) XL(3) D(2) XL(4) XL(5) D(3) XL(6) XL(7) D(4) XL(0) XL(1) D(5) XL(2) XL(3) D(6) XL(4) XL(5) D(7) XL(6) XL(7) }
    }
    double s = 0; for (int k = 0; k < 8; ++k) s += a[k] + f[k] + q[k];
    if (s == 12345.678) sink[0] = s;
}

template <int PAT>
void run(const char *name, int nd, int nx, int warps_per_sm) {
    double *sink; cudaMalloc(&sink, 8);
    const int iters = 40000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<PAT><<<148, warps_per_sm * 32>>>(sink, 100);
    cudaEventRecord(e0);
    probe<PAT><<<148, warps_per_sm * 32>>>(sink, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9 / iters / (warps_per_sm / 4.0);
    printf("%-34s w/sched %d: D=%d X=%d  %.2f cycles/round  (2D=%d, 2D+X=%d, max(2D,D+X)=%d)\n", name, warps_per_sm / 4, nd, nx, cyc, 2 * nd, 2 * nd + nx, 2 * nd > nd + nx ? 2 * nd : nd + nx);
    cudaFree(sink);
}

int main() {
    for (int w : {4, 16, 28}) {
        run<0>("D only", 8, 0, w);
        run<1>("D F alternating", 8, 8, w);
        run<2>("8D then 8F", 8, 8, w);
        run<3>("D IMAD alternating", 8, 8, w);
        run<4>("D LOP alternating", 8, 8, w);
        run<5>("D LOP IMAD", 8, 16, w);
        run<6>("D MUFU alternating", 8, 8, w);
        run<7>("D LDS alternating", 8, 8, w);
        run<8>("D F F", 8, 16, w);
        run<9>("D D F", 8, 4, w);
        run<10>("LOP only", 0, 8, w);
        run<11>("D LOP LOP", 8, 16, w);
    }
    return 0;
}
