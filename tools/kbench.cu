// Stand-alone kernel timing harness (development tool, not part of the
// product): builds the headline FHN problem on the device, then times
// pb200_eval through the C ABI with CUDA events.  Compile one executable per
// kernel variant (-D flags) and run them back to back in one gpurun call.
//
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a \
//        [-DPB200_...] tools/kbench.cu pints_b200/csrc/*.cu -o /tmp/kbench
//   /tmp/kbench [n=65536] [reps=20] [block=0] [spread=0.05] [n_times=1000]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "../include/pints_b200.h"

#if PB200_TIMELINE
extern "C" int pb200_debug_timeline(long long *host, int n_words);
#endif
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
    printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
#define PB(x) do { int rc = (x); if (rc) { printf("pb200 error %d: %s\n", rc, pb200_last_error()); exit(1);} } while (0)

static unsigned long long lcg = 88172645463325252ull;
static double urand() { lcg ^= lcg << 13; lcg ^= lcg >> 7; lcg ^= lcg << 17; return (lcg >> 11) * (1.0 / 9007199254740992.0); }
static double nrand() { double u1 = urand() + 1e-300, u2 = urand(); return sqrt(-2 * log(u1)) * cos(6.283185307179586 * u2); }

int main(int argc, char **argv) {
    long n = argc > 1 ? atol(argv[1]) : 65536;
    int reps = argc > 2 ? atoi(argv[2]) : 20;
    int block = argc > 3 ? atoi(argv[3]) : 0;
    double spread = argc > 4 ? atof(argv[4]) : 0.05;
    int nt = argc > 5 ? atoi(argv[5]) : 1000;
    const int no = 2;
    long slen = pb200_series_len(nt, no);
    std::vector<double> series(slen, 0.0);
    for (int j = 0; j < nt; ++j) series[3 * j] = 20.0 * j / (nt - 1);
    for (int k = 0; k < PB200_SENTINEL_ROWS; ++k) series[3 * (nt + k)] = INFINITY;   // sentinel rows
    double *d_series; CK(cudaMalloc(&d_series, slen * 8));
    CK(cudaMemcpy(d_series, series.data(), slen * 8, cudaMemcpyHostToDevice));
    double mc[2] = {-1.0, 1.0};
    double sigma = 0.5;
    double off = -0.5 * nt * log(2 * M_PI) - nt * log(sigma), mul = -1 / (2 * sigma * sigma);
    double oc[4] = {off, off, mul, mul};
    pb200_problem *p0, *p;
    PB(pb200_problem_create(PB200_MODEL_FHN, PB200_OBJ_GAUSSIAN_KNOWN_SIGMA, nt, no, d_series, mc, 2, oc, 4, 1.0, nullptr, nullptr, 0, 0.0, &p0));
    // clean trajectory at the generating parameters, then add noise
    double truth[3] = {0.1, 0.5, 3.0};
    double *d_one, *d_traj; CK(cudaMalloc(&d_one, 24)); CK(cudaMalloc(&d_traj, (size_t)nt * no * 8));
    CK(cudaMemcpy(d_one, truth, 24, cudaMemcpyHostToDevice));
    PB(pb200_simulate(p0, d_one, 1, 3, d_traj, nullptr, nullptr, nullptr));
    std::vector<double> traj((size_t)nt * no);
    CK(cudaMemcpy(traj.data(), d_traj, traj.size() * 8, cudaMemcpyDeviceToHost));
    for (int j = 0; j < nt; ++j) for (int o = 0; o < no; ++o) series[3 * j + 1 + o] = traj[j * no + o] + 0.5 * nrand();
    CK(cudaMemcpy(d_series, series.data(), slen * 8, cudaMemcpyHostToDevice));
    PB(pb200_problem_create(PB200_MODEL_FHN, PB200_OBJ_GAUSSIAN_KNOWN_SIGMA, nt, no, d_series, mc, 2, oc, 4, 1.0, nullptr, nullptr, 0, 0.0, &p));
    std::vector<double> xs((size_t)n * 3);
    for (long i = 0; i < n; ++i) {
        if (spread > 0) for (int k = 0; k < 3; ++k) xs[i * 3 + k] = truth[k] * (1 + spread * nrand());
        else { xs[i * 3] = urand(); xs[i * 3 + 1] = urand(); xs[i * 3 + 2] = 1.0 + 4.0 * urand(); }   // wide box
    }
    int presort = argc > 7 ? atoi(argv[7]) : 0;     // 1: host-side sort by c (what a device sort would achieve)
    if (presort) {
        std::vector<long> idx(n); for (long i = 0; i < n; ++i) idx[i] = i;
        std::sort(idx.begin(), idx.end(), [&](long a, long b) { return xs[a * 3 + 2] < xs[b * 3 + 2]; });
        std::vector<double> ys(xs.size());
        for (long i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) ys[i * 3 + k] = xs[idx[i] * 3 + k];
        xs.swap(ys);
    }
    int devsort = argc > 8 ? atoi(argv[8]) : 0;     // 1: pb200_eval_sorted
    void *d_ws = nullptr; long ws_bytes = 0;
    if (devsort) { ws_bytes = pb200_sort_workspace_bytes(n); CK(cudaMalloc(&d_ws, ws_bytes)); }
    double *d_x, *d_s; int *d_steps;
    CK(cudaMalloc(&d_x, xs.size() * 8)); CK(cudaMalloc(&d_s, n * 8)); CK(cudaMalloc(&d_steps, n * 4));
    CK(cudaMemcpy(d_x, xs.data(), xs.size() * 8, cudaMemcpyHostToDevice));
    int lanes = argc > 6 ? atoi(argv[6]) : 0;
    int wide = argc > 9 ? atoi(argv[9]) : 0;
    pb200_solver_opts opts = {0, 0, 0, 0, block, lanes, wide};
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) PB(pb200_eval_sorted(p, d_x, n, 3, d_s, d_steps, &opts, d_ws, ws_bytes, nullptr));
    CK(cudaDeviceSynchronize());
    std::vector<float> ms(reps);
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        PB(pb200_eval_sorted(p, d_x, n, 3, d_s, nullptr, &opts, d_ws, ws_bytes, nullptr));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms[r], e0, e1));
    }
    std::sort(ms.begin(), ms.end());
    std::vector<double> s(n); std::vector<int> st(n);
    CK(cudaMemcpy(s.data(), d_s, n * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(st.data(), d_steps, n * 4, cudaMemcpyDeviceToHost));
    double sum = 0, steps = 0; long bad = 0;
    for (long i = 0; i < n; ++i) { if (std::isfinite(s[i])) sum += s[i]; else ++bad; steps += st[i]; }
    double flops = 234.0 * steps + 24.0 * nt * (double)n;
    double med = ms[reps / 2];
    printf("sort=%d%d lanes=%d n=%ld block=%d spread=%g nt=%d  median %.4f ms  min %.4f ms  %.3e evals/s  %.2f TFLOP/s  mean_steps %.1f  checksum %.10e  nonfinite %ld\n",
           presort, devsort, lanes, n, block, spread, nt, med, ms[0], n / (med * 1e-3), flops / (med * 1e-3) / 1e12, steps / n, sum / (n - bad), bad);
#if PB200_TIMELINE
    {
        // per-warp start / end clocks of the LAST launch (migrating layout only)
        std::vector<long long> tl(160 * 16 * 4);
        if (pb200_debug_timeline(tl.data(), (int)tl.size()) == 0) {
            long long t0 = -1;
            for (int b = 0; b < 147; ++b) for (int w = 0; w < 16; ++w) { long long v = tl[(b * 16 + w) * 4]; if (v && (t0 < 0 || v < t0)) t0 = v; }
            std::vector<double> bend;
            double sched_sum[4] = {0, 0, 0, 0};
            for (int b = 0; b < 147; ++b) {
                long long e = 0;
                for (int w = 0; w < 16; ++w) { long long v = tl[(b * 16 + w) * 4 + 1]; if (v > e) e = v; }
                bend.push_back((double)(e - t0));
                for (int sc = 0; sc < 4; ++sc) { long long m = 0; for (int w = sc; w < 16; w += 4) { long long v = tl[(b * 16 + w) * 4 + 1]; if (v > m) m = v; } sched_sum[sc] += (double)(m - t0); }
            }
            std::vector<double> sorted_end = bend; std::sort(sorted_end.begin(), sorted_end.end());
            double mean = 0; for (double v : bend) mean += v; mean /= bend.size();
            printf("timeline: block end (cycles after first start) min %.0f  p10 %.0f  median %.0f  mean %.0f  p90 %.0f  max %.0f\n",
                   sorted_end.front(), sorted_end[14], sorted_end[73], mean, sorted_end[132], sorted_end.back());
            printf("timeline: mean end of the last warp per scheduler: %.0f %.0f %.0f %.0f\n", sched_sum[0] / 147, sched_sum[1] / 147, sched_sum[2] / 147, sched_sum[3] / 147);
            for (int b : {0, 36, 73, 110, 146}) {
                printf("timeline: block %3d sm %3lld:", b, tl[(b * 16) * 4 + 2]);
                for (int w = 0; w < 16; ++w) printf(" %lld-%lld(%lld)", (tl[(b * 16 + w) * 4] - t0) / 1000, (tl[(b * 16 + w) * 4 + 1] - t0) / 1000, tl[(b * 16 + w) * 4 + 3]);
                printf("\n");
            }
        }
    }
#endif
    return 0;
}
