// Literal per-call kernels behind the three functions_1.py signatures (batch-of-1 use by the drop-in
// module; the batched pipeline uses k_calibrate instead).
#pragma once
#include "mua_common.cuh"

namespace mua {

// online_histogram_w_sat_based_nb_of_samples (functions_1.py:27-68) for one channel:
// i = min(max(H,1), n); x[:i] saturated IN PLACE with `>= m -> m` (:45-46);
// counts[v] = #{t < i : x_t == v}; first[v] = first position of v (dict insertion order, :48-53).
__global__ void __launch_bounds__(256) k_online_hist(uint8_t* __restrict__ x, int64_t n, int64_t H, int m,
                                                     unsigned int* __restrict__ counts, int* __restrict__ first) {
    __shared__ unsigned int s_cnt[256];
    __shared__ int s_first[256];
    const int tid = threadIdx.x;
    s_cnt[tid] = 0;
    s_first[tid] = 0x7FFFFFFF;
    __syncthreads();
    const int64_t i = n < (H < 1 ? 1 : H) ? n : (H < 1 ? 1 : H);
    for (int64_t t = tid; t < i; t += blockDim.x) {
        int v = x[t];
        if (v >= m) { v = m; x[t] = (uint8_t)m; }
        atomicAdd(&s_cnt[v], 1u);
        atomicMin(&s_first[v], (int)t);
    }
    __syncthreads();
    counts[tid] = s_cnt[tid];
    first[tid] = s_first[tid];
}

// approx_sort (functions_1.py:75-90) for `count` histograms of length n: idx = argsort(rank) i.e.
// idx[rank_of(p, s, n)] = s with p = first argmax (:77).
template <typename TH>
__global__ void __launch_bounds__(128) k_approx_sort(const TH* __restrict__ hist, int n, int64_t count,
                                                     int64_t* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const TH* h = hist + i * n;
    int p = 0;
    TH best = h[0];
    for (int s = 1; s < n; ++s)
        if (h[s] > best) { best = h[s]; p = s; }
    for (int s = 0; s < n; ++s) idx[i * n + rank_of(p, s, n)] = s;
}

}  // namespace mua
