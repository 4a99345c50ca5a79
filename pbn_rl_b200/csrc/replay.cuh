// Device-resident replay ring and fused observation unpack.
//
// The reference keeps transitions in a python list (bdq_model/memory.py:22-70) and rebuilds float
// tensors from tuples for every policy update (bdq_model/__init__.py:100-109) and every action
// (bdq_model/__init__.py:92-93).  Here transitions stay packed in HBM (8W+4 B observe half,
// bins+4+1+8W B commit half per env-step) and the float [2,B,N] tensors the Q-network consumes are
// produced by one gather+unpack kernel.  All kernels are HBM-bound copies with coalesced accesses.
#pragma once
#include "pbn_common.cuh"

namespace pbn {

// slot of env e for a push that starts at ring position head
__device__ __forceinline__ int64_t ring_slot(int64_t head, int64_t e, int64_t cap) {
  int64_t s = head + e;
  return s >= cap ? s % cap : s;
}

// Before the step: slot <- (state, target_id).
__global__ void __launch_bounds__(256) replay_observe_kernel(pbn_replay r, int64_t head, const uint64_t* __restrict__ state,
                                                            const int32_t* __restrict__ target_id, int W, int64_t n) {
  const int64_t nth = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = tid; i < n * W; i += nth) {
    const int64_t e = W == 1 ? i : (i >> 1), w = W == 1 ? 0 : (i & 1);
    r.state[ring_slot(head, e, r.capacity) * W + w] = state[i];
  }
  for (int64_t e = tid; e < n; e += nth) r.target_id[ring_slot(head, e, r.capacity)] = target_id ? target_id[e] : -1;
}

// After the step: slot <- (actions, reward, done, next_state).
__global__ void __launch_bounds__(256) replay_commit_kernel(pbn_replay r, int64_t head, const uint8_t* __restrict__ actions,
                                                           const float* __restrict__ reward,
                                                           const uint8_t* __restrict__ terminated,
                                                           const uint8_t* __restrict__ truncated,
                                                           const uint64_t* __restrict__ next_state, int W, int bins,
                                                           int64_t n) {
  const int64_t nth = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = tid; i < n * W; i += nth) {
    const int64_t e = W == 1 ? i : (i >> 1), w = W == 1 ? 0 : (i & 1);
    r.next_state[ring_slot(head, e, r.capacity) * W + w] = next_state[i];
  }
  for (int64_t i = tid; i < n * bins; i += nth) {
    const int64_t e = i / bins, k = i - e * bins;
    r.actions[ring_slot(head, e, r.capacity) * bins + k] = actions ? actions[i] : (uint8_t)0;
  }
  for (int64_t e = tid; e < n; e += nth) {
    const int64_t s = ring_slot(head, e, r.capacity);
    r.reward[s] = reward[e];
    r.done[s] = (uint8_t)((terminated[e] & 1u) | ((truncated ? truncated[e] & 1u : 0u) << 1));
  }
}

// Packed -> float tensors of the Q-network.  obs[0,b,:] = bits of state[index[b]], obs[1,b,:] = bits of
// the first state of attractor target_id[index[b]] ('*' -> 0; the `target` of env.reset(),
// bdq_model/__init__.py:161); next_obs likewise from next_state.  index == nullptr: identity.
// One thread per (b, gene): coalesced float stores, the packed words are broadcast from L1.
__global__ void __launch_bounds__(256) gather_unpack_kernel(NetParams n, const uint64_t* __restrict__ state,
                                                           const uint64_t* __restrict__ next_state,
                                                           const int32_t* __restrict__ target_id,
                                                           const int64_t* __restrict__ index, int W, int64_t B,
                                                           float* __restrict__ obs, float* __restrict__ next_obs) {
  const int N = n.n_genes;
  const int64_t total = B * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / N;
    const int g = (int)(i - b * N);
    const int64_t src = index ? index[b] : b;
    const int w = g >> 6, sh = g & 63;
    float tbit = 0.0f;
    const int32_t a = target_id ? target_id[src] : -1;
    if (a >= 0 && a < n.n_attr) {
      const int64_t s0 = n.attr_offset[a];
      tbit = (float)(((n.attr_val[s0 * W + w] & n.attr_care[s0 * W + w]) >> sh) & 1ull);
    }
    if (obs) {
      obs[i] = (float)((state[src * W + w] >> sh) & 1ull);
      obs[total + i] = tbit;
    }
    if (next_obs) {
      next_obs[i] = (float)((next_state[src * W + w] >> sh) & 1ull);
      next_obs[total + i] = tbit;
    }
  }
}

// The scalar columns of a sampled batch: actions as int64 [B,bins] (gather index of
// bdq_model/__init__.py:106), reward [B], done [B] as float (the reference's `masks`).
__global__ void __launch_bounds__(256) gather_scalars_kernel(pbn_replay r, const int64_t* __restrict__ index, int bins,
                                                            int64_t B, int64_t* __restrict__ actions,
                                                            float* __restrict__ reward, float* __restrict__ done) {
  const int64_t nth = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (actions)
    for (int64_t i = tid; i < B * bins; i += nth) {
      const int64_t b = i / bins, k = i - b * bins;
      actions[i] = (int64_t)r.actions[index[b] * bins + k];
    }
  for (int64_t b = tid; b < B; b += nth) {
    const int64_t s = index[b];
    if (reward) reward[b] = r.reward[s];
    if (done) done[b] = r.done[s] != 0 ? 1.0f : 0.0f;
  }
}

}  // namespace pbn
