// Batched all-pairs evaluator support: the bookkeeping of model_tester.py:595-658
// ("count steps until the state is in the target attractor, give up after max_steps") for E rollouts
// that are stepped together.  Small memory-bound kernels; the step itself is pbn_step.
#pragma once
#include "pbn_common.cuh"

namespace pbn {

// env.in_target(state) for every instance (model_tester.py:616).
template <int W>
__global__ void __launch_bounds__(256) in_target_kernel(const __grid_constant__ NetParams n,
                                                       const uint64_t* __restrict__ state,
                                                       const int32_t* __restrict__ target_id, uint8_t* __restrict__ out,
                                                       int64_t n_envs) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_envs; e += (int64_t)gridDim.x * blockDim.x) {
    uint64_t s[W];
#pragma unroll
    for (int w = 0; w < W; ++w) s[w] = state[e * W + w];
    const int a = target_id[e];
    out[e] = (a >= 0 && a < n.n_attr && (n.ahash_tags != nullptr ? in_attractor_hashed<W>(n, a, s) : in_attractor<W>(n.attr_offset, n.attr_care, n.attr_val, a, s))) ? 1 : 0;
  }
}

// One loop iteration of model_tester.py:616-638 after the step: for rollouts still running,
// count += 1; count > max_steps -> failed (count stays max_steps + 1, the reference books 101);
// else if the new state is in the target -> finished with `count` steps.
__global__ void __launch_bounds__(256) rollout_track_kernel(const uint8_t* __restrict__ terminated,
                                                           uint8_t* __restrict__ active, int32_t* __restrict__ count,
                                                           int32_t max_steps, int64_t n_envs,
                                                           unsigned int* __restrict__ n_active) {
  unsigned int mine = 0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_envs; e += (int64_t)gridDim.x * blockDim.x) {
    if (!active[e]) continue;
    const int32_t c = count[e] + 1;
    count[e] = c;
    if (c > max_steps || terminated[e]) active[e] = 0; else ++mine;
  }
  if (n_active != nullptr) {
    mine = __reduce_add_sync(0xFFFFFFFFu, mine);
    if ((threadIdx.x & 31u) == 0 && mine) atomicAdd(n_active, mine);
  }
}

// result_matrix[src, tgt] += count; data[count] += 1 (model_tester.py:645-652); rollout e belongs to pair
// pair_id[e] = src * A + tgt.
__global__ void __launch_bounds__(256) rollout_reduce_kernel(const int32_t* __restrict__ count,
                                                            const int32_t* __restrict__ pair_id, int64_t n_envs,
                                                            int32_t n_pairs, int32_t max_steps,
                                                            unsigned long long* __restrict__ matrix,
                                                            unsigned long long* __restrict__ hist) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_envs; e += (int64_t)gridDim.x * blockDim.x) {
    int32_t c = count[e];
    c = c < 0 ? 0 : (c > max_steps + 1 ? max_steps + 1 : c);
    const int32_t p = pair_id[e];
    if (p >= 0 && p < n_pairs) atomicAdd(&matrix[p], (unsigned long long)c);
    atomicAdd(&hist[c], 1ull);
  }
}

}  // namespace pbn
