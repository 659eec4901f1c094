// Reading side of the fused score exchange of a multi-GPU job (ScatterDev in
// pb200_common.cuh has the writing side, which lives inside the evaluation
// kernels).  Replaces the result queue of ParallelEvaluator
// (pints/_evaluation.py:316-337): there, the parent polls a multiprocessing
// queue for (index, score) tuples and files each score under its index; here
// every rank's kernel has stored {score, row} pairs into every rank's staging
// array over NVLink, and this kernel files them under their rows once all
// ranks have raised their flags.
#include "pb200_common.cuh"

namespace {

constexpr int FINISH_THREADS = 256;

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// flags[q] >= epoch * EXCHANGE_UNIT: every block of rank q's launch of this
// generation has added its share, i.e. rank q has stored the whole of its block
// (and everything before the blocks' fences is visible after the acquire).
// One thread per rank spins; the block barrier hands the visibility on to the
// other threads, which read the pairs through L2 (ld.global.cg: the lines were
// written by peers, never cached in this SM's L1).
__global__ void __launch_bounds__(FINISH_THREADS)
exchange_finish_kernel(const double2 *__restrict__ stage,
                       const unsigned long long *__restrict__ flags, int n_ranks,
                       int64_t per_rank, int64_t rows_per_rank, unsigned long long epoch,
                       int64_t n_total, double *__restrict__ out) {
    if (threadIdx.x < n_ranks) {
        const unsigned long long t0 = global_timer_ns();
        const unsigned long long target = epoch << PB200_EXCHANGE_UNIT_LOG2;
        while (ld_acquire_sys(flags + threadIdx.x) < target) {
            if (global_timer_ns() - t0 > 30000000000ull) __trap();   // a peer died
            __nanosleep(64);
        }
    }
    __syncthreads();
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
         k < rows_per_rank * n_ranks; k += stride) {
        const int64_t q = k / rows_per_rank, slot = k - q * rows_per_rank;
        const int64_t base = q * rows_per_rank;
        const int64_t rows = n_total - base < rows_per_rank ? n_total - base : rows_per_rank;
        if (slot >= rows) continue;
        const double2 pair = __ldcg(stage + q * per_rank + slot);
        const int64_t row = __double_as_longlong(pair.y);
        if (row >= 0 && row < rows) out[base + row] = pair.x;
    }
}

// a rank without rows still tells everybody that it has reached this generation
__global__ void exchange_signal_kernel(const ScatterDev sc) {
    __threadfence_system();
    for (int r = 0; r < sc.n; ++r)
        red_add_relaxed_sys(sc.flag[r], 1ull << PB200_EXCHANGE_UNIT_LOG2);
}

}  // namespace

int launch_exchange_signal(const ScatterDev &sc, cudaStream_t st) {
    exchange_signal_kernel<<<1, 1, 0, st>>>(sc);
    return pb200_check_cuda(cudaGetLastError(), "exchange_signal_kernel launch");
}

int launch_exchange_finish(const void *stage, const void *flags, int n_ranks, int64_t per_rank,
                           int64_t rows_per_rank, unsigned long long epoch, int64_t n_total,
                           double *d_scores_all, cudaStream_t st) {
    const int64_t pairs = rows_per_rank * n_ranks;
    int64_t grid = (pairs + FINISH_THREADS - 1) / FINISH_THREADS;
    if (grid > 592) grid = 592;                  // 4 blocks per SM, grid-stride beyond
    if (grid < 1) grid = 1;
    exchange_finish_kernel<<<static_cast<unsigned>(grid), FINISH_THREADS, 0, st>>>(
        static_cast<const double2 *>(stage), static_cast<const unsigned long long *>(flags),
        n_ranks, per_rank, rows_per_rank, epoch, n_total, d_scores_all);
    return pb200_check_cuda(cudaGetLastError(), "exchange_finish_kernel launch");
}
