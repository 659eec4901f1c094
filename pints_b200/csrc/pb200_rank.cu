// Ranking of a population's scores on the device: argsort of n doubles
// (n <= PB200_ARGSORT_MAX), ascending, NaN last, ties by index -- what
// XNES.tell needs before it can weight the samples by rank
// (pints/_optimisers/_xnes.py:159-163 `order = np.argsort(fx)`).
//
// A sample sort sized for populations of 10^3 .. 2.6 10^5 (a general radix
// sort of 64-bit keys takes eight passes over the data; this one reads the
// scores twice):
//   1. one block sorts 2048 evenly spaced samples (bitonic network in shared
//      memory) and keeps every 8th as a splitter: 256 buckets whose sizes are
//      balanced whatever the distribution of the scores (quantiles, not value
//      ranges; equal scores are told apart by their index);
//   2. every element finds its bucket (binary search over the splitters in
//      shared memory), buckets are counted and filled through atomic cursors;
//   3. one block per bucket sorts it in shared memory (the (score, index) pair
//      is the key, so the unordered filling of step 2 does not matter and the
//      result is deterministic) and writes its slice of the permutation.
// A bucket that does not fit a block's shared memory (never seen; it needs an
// adversarial sample) sets *status = 1 and the caller sorts another way.
#include "pb200_common.cuh"

namespace {

constexpr int SAMPLES = 2048;
constexpr int BUCKETS = 256;
constexpr int BUCKET_CAP = 4096;           // elements one block sorts
constexpr int SORT_THREADS = 1024;

struct SortHeader {
    uint32_t count[BUCKETS];
    uint32_t cursor[BUCKETS];
};

// monotone map double -> uint64 (NaN above +inf)
__device__ __forceinline__ uint64_t sortable(double v) {
    if (v != v) return ~0ull;
    uint64_t b = static_cast<uint64_t>(__double_as_longlong(v));
    if (b == 0x8000000000000000ull) b = 0;          // -0.0 ranks as +0.0
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ bool pair_less(uint64_t ka, uint32_t ia, uint64_t kb, uint32_t ib) {
    return ka < kb || (ka == kb && ia < ib);
}

// bitonic sort of p (power of two) (key, index) pairs in shared memory.
// Thread t owns the t-th pair of every stage; for partner distances j <= 32
// the 32 pairs of a warp lie in one aligned block of 64 elements that no other
// warp touches, so those stages (51 of the 66 for 2048 pairs) only need a warp
// barrier; a block barrier separates them from the stages with j >= 64.
__device__ void bitonic_sort(uint64_t *key, uint32_t *idx, int p) {
    for (int k = 2; k <= p; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < p / 2; t += blockDim.x) {
                // the t-th pair of this stage: i has bit j clear
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool up = (i & k) == 0;
                const uint64_t ka = key[i], kb = key[l];
                const uint32_t ia = idx[i], ib = idx[l];
                const bool swap = up ? pair_less(kb, ib, ka, ia) : pair_less(ka, ia, kb, ib);
                if (swap) { key[i] = kb; key[l] = ka; idx[i] = ib; idx[l] = ia; }
            }
            const int next_j = j > 1 ? (j >> 1) : k;
            if (j >= 64 || next_j >= 64) __syncthreads(); else __syncwarp();
        }
    }
}

__device__ __forceinline__ int pow2_at_least(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

// n <= BUCKET_CAP: the whole array in one block
__global__ void __launch_bounds__(SORT_THREADS)
argsort_small_kernel(const double *__restrict__ v, int64_t *__restrict__ order, int n) {
    extern __shared__ __align__(16) unsigned char raw[];
    uint64_t *key = reinterpret_cast<uint64_t *>(raw);
    uint32_t *idx = reinterpret_cast<uint32_t *>(key + BUCKET_CAP);
    const int p = pow2_at_least(n);
    for (int i = threadIdx.x; i < p; i += blockDim.x) {
        key[i] = i < n ? sortable(v[i]) : ~0ull;
        idx[i] = i < n ? i : 0xffffffffu;
    }
    __syncthreads();
    bitonic_sort(key, idx, p);
    for (int i = threadIdx.x; i < n; i += blockDim.x) order[i] = idx[i];
}

// step 1: splitters from evenly spaced samples; also clears the header
__global__ void __launch_bounds__(SORT_THREADS)
argsort_sample_kernel(const double *__restrict__ v, int n, uint64_t *__restrict__ split_key,
                      uint32_t *__restrict__ split_idx, SortHeader *__restrict__ hdr) {
    extern __shared__ __align__(16) unsigned char raw[];
    uint64_t *key = reinterpret_cast<uint64_t *>(raw);
    uint32_t *idx = reinterpret_cast<uint32_t *>(key + SAMPLES);
    for (int s = threadIdx.x; s < SAMPLES; s += blockDim.x) {
        const uint32_t i = static_cast<uint32_t>((static_cast<int64_t>(s) * n) / SAMPLES);
        key[s] = sortable(v[i]);
        idx[s] = i;
    }
    for (int b = threadIdx.x; b < BUCKETS; b += blockDim.x) { hdr->count[b] = 0; hdr->cursor[b] = 0; }
    __syncthreads();
    bitonic_sort(key, idx, SAMPLES);
    for (int b = threadIdx.x; b < BUCKETS - 1; b += blockDim.x) {
        split_key[b] = key[(b + 1) * (SAMPLES / BUCKETS)];
        split_idx[b] = idx[(b + 1) * (SAMPLES / BUCKETS)];
    }
}

// bucket of element (k, i): number of splitters <= it
__device__ __forceinline__ int bucket_of(uint64_t k, uint32_t i, const uint64_t *sk,
                                         const uint32_t *si) {
    int lo = 0, hi = BUCKETS - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (pair_less(k, i, sk[mid], si[mid])) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// step 2a: bucket of every element, bucket sizes
__global__ void argsort_count_kernel(const double *__restrict__ v, int n,
                                     const uint64_t *__restrict__ split_key,
                                     const uint32_t *__restrict__ split_idx,
                                     uint8_t *__restrict__ bucket, SortHeader *__restrict__ hdr) {
    __shared__ uint64_t sk[BUCKETS];
    __shared__ uint32_t si[BUCKETS];
    __shared__ uint32_t local[BUCKETS];
    for (int b = threadIdx.x; b < BUCKETS; b += blockDim.x) {
        sk[b] = b < BUCKETS - 1 ? split_key[b] : ~0ull;
        si[b] = b < BUCKETS - 1 ? split_idx[b] : 0xffffffffu;
        local[b] = 0;
    }
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int b = bucket_of(sortable(v[i]), static_cast<uint32_t>(i), sk, si);
        bucket[i] = static_cast<uint8_t>(b);
        atomicAdd(&local[b], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < BUCKETS; b += blockDim.x)
        if (local[b]) atomicAdd(&hdr->count[b], local[b]);
}

// step 2b: fill the buckets (any order inside a bucket)
__global__ void argsort_scatter_kernel(const double *__restrict__ v, int n,
                                       const uint8_t *__restrict__ bucket,
                                       SortHeader *__restrict__ hdr, uint64_t *__restrict__ tmp_key,
                                       uint32_t *__restrict__ tmp_idx) {
    __shared__ uint32_t start[BUCKETS];
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int b = 0; b < BUCKETS; ++b) { start[b] = acc; acc += hdr->count[b]; }
    }
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int b = bucket[i];
        const uint32_t pos = start[b] + atomicAdd(&hdr->cursor[b], 1u);
        tmp_key[pos] = sortable(v[i]);
        tmp_idx[pos] = static_cast<uint32_t>(i);
    }
}

// step 3: one block per bucket
__global__ void __launch_bounds__(SORT_THREADS)
argsort_bucket_kernel(const SortHeader *__restrict__ hdr, const uint64_t *__restrict__ tmp_key,
                      const uint32_t *__restrict__ tmp_idx, int64_t *__restrict__ order,
                      int32_t *__restrict__ status) {
    extern __shared__ __align__(16) unsigned char raw[];
    uint64_t *key = reinterpret_cast<uint64_t *>(raw);
    uint32_t *idx = reinterpret_cast<uint32_t *>(key + BUCKET_CAP);
    uint32_t &first = idx[BUCKET_CAP];              // one word past the pairs
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int q = 0; q < b; ++q) acc += hdr->count[q];
        first = acc;
    }
    __syncthreads();
    const int cnt = static_cast<int>(hdr->count[b]);
    if (cnt > BUCKET_CAP) {
        if (threadIdx.x == 0) *status = 1;
        return;
    }
    if (cnt == 0) return;
    const int p = pow2_at_least(cnt);
    for (int i = threadIdx.x; i < p; i += blockDim.x) {
        key[i] = i < cnt ? tmp_key[first + i] : ~0ull;
        idx[i] = i < cnt ? tmp_idx[first + i] : 0xffffffffu;
    }
    __syncthreads();
    bitonic_sort(key, idx, p);
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) order[first + i] = idx[i];
}

constexpr size_t PAIR_SMEM = static_cast<size_t>(BUCKET_CAP) * (8 + 4) + 16;

size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

}  // namespace

extern "C" {

int64_t pb200_argsort_workspace_bytes(int64_t n) {
    if (n < 0) return 0;
    // header, splitters (key + index), bucket ids, scattered keys and indices
    return static_cast<int64_t>(align256(sizeof(SortHeader)) + align256(BUCKETS * 8) +
                                align256(BUCKETS * 4) + align256(static_cast<size_t>(n)) +
                                align256(static_cast<size_t>(n) * 8) +
                                align256(static_cast<size_t>(n) * 4));
}

int pb200_argsort_f64(const double *d_values, int64_t *d_order, int64_t n, void *d_ws,
                      int64_t ws_bytes, int32_t *d_status, void *stream) {
    if (n < 0 || n > PB200_ARGSORT_MAX) {
        pb200_set_error("pb200_argsort_f64: n = %lld outside [0, %d]", (long long)n,
                        PB200_ARGSORT_MAX);
        return PB200_EINVAL;
    }
    if (n == 0) return PB200_OK;
    if (!d_values || !d_order || !d_status) {
        pb200_set_error("pb200_argsort_f64: NULL pointer");
        return PB200_EINVAL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = pb200_check_cuda(cudaMemsetAsync(d_status, 0, sizeof(int32_t), st), "memset status");
    if (rc) return rc;
    // more than 48 KB of dynamic shared memory: opt in once per device
    static bool configured_on[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    bool &configured = configured_on[dev & 63];
    if (!configured) {
        rc = pb200_check_cuda(cudaFuncSetAttribute(argsort_small_kernel,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(PAIR_SMEM)), "smem attribute");
        if (rc) return rc;
        rc = pb200_check_cuda(cudaFuncSetAttribute(argsort_bucket_kernel,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(PAIR_SMEM)), "smem attribute");
        if (rc) return rc;
        rc = pb200_check_cuda(cudaFuncSetAttribute(argsort_sample_kernel,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(SAMPLES * 12)), "smem attribute");
        if (rc) return rc;
        configured = true;
    }
    const int ni = static_cast<int>(n);
    if (n <= BUCKET_CAP) {
        argsort_small_kernel<<<1, SORT_THREADS, PAIR_SMEM, st>>>(d_values, d_order, ni);
        return pb200_check_cuda(cudaGetLastError(), "argsort_small_kernel launch");
    }
    if (!d_ws || ws_bytes < pb200_argsort_workspace_bytes(n)) {
        pb200_set_error("pb200_argsort_f64: workspace of %lld bytes needed",
                        (long long)pb200_argsort_workspace_bytes(n));
        return PB200_EINVAL;
    }
    char *w = static_cast<char *>(d_ws);
    SortHeader *hdr = reinterpret_cast<SortHeader *>(w);          w += align256(sizeof(SortHeader));
    uint64_t *split_key = reinterpret_cast<uint64_t *>(w);        w += align256(BUCKETS * 8);
    uint32_t *split_idx = reinterpret_cast<uint32_t *>(w);        w += align256(BUCKETS * 4);
    uint8_t *bucket = reinterpret_cast<uint8_t *>(w);             w += align256(static_cast<size_t>(n));
    uint64_t *tmp_key = reinterpret_cast<uint64_t *>(w);          w += align256(static_cast<size_t>(n) * 8);
    uint32_t *tmp_idx = reinterpret_cast<uint32_t *>(w);
    const int blocks = static_cast<int>((n + 255) / 256 < 592 ? (n + 255) / 256 : 592);
    argsort_sample_kernel<<<1, SORT_THREADS, SAMPLES * 12, st>>>(d_values, ni, split_key, split_idx,
                                                                 hdr);
    argsort_count_kernel<<<blocks, 256, 0, st>>>(d_values, ni, split_key, split_idx, bucket, hdr);
    argsort_scatter_kernel<<<blocks, 256, 0, st>>>(d_values, ni, bucket, hdr, tmp_key, tmp_idx);
    argsort_bucket_kernel<<<BUCKETS, SORT_THREADS, PAIR_SMEM, st>>>(hdr, tmp_key, tmp_idx, d_order,
                                                                    d_status);
    return pb200_check_cuda(cudaGetLastError(), "argsort kernels launch");
}

}  // extern "C"
