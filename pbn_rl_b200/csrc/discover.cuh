// Attractor discovery / steady-state statistics support: a visit-count hash table in
// HBM.  Massive perturbation-free rollouts (pbn_step) end in the network's attractors; counting the
// distinct end states on the device replaces the reference's per-instance bookkeeping
// (env.all_attractors growing during training, bdq_model/__init__.py:182-184; compute_ssd_hist's
// state histogram, train_pbn_28.py:257).  Open addressing, linear probing, capacity a power of two.
#pragma once
#include "pbn_common.cuh"

namespace pbn {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

// Fingerprint of a state: never 0 (0 marks an empty slot).  Documented so the host can re-derive it.
template <int W>
__device__ __forceinline__ uint64_t state_tag(const uint64_t (&s)[W]) {
  uint64_t h = mix64(s[0] + 0x9E3779B97F4A7C15ull);
  if (W == 2) h = mix64(h ^ (s[1] + 0xD1B54A32D192ED03ull));
  return h == 0 ? 1ull : h;
}

// counts[slot(state[e])] += 1 for every instance with mask[e] != 0 (all if mask == nullptr).
// tags[cap] (0 = empty), slot_state[cap*W], counts[cap]; *overflow counts states that found no slot
// within kMaxProbe probes (table too full).
template <int W>
__global__ void __launch_bounds__(256) visit_count_kernel(const uint64_t* __restrict__ state, const uint8_t* __restrict__ mask,
                                                         int64_t n_envs, unsigned long long* __restrict__ tags,
                                                         uint64_t* __restrict__ slot_state,
                                                         unsigned long long* __restrict__ counts, uint64_t cap_mask,
                                                         unsigned int* __restrict__ overflow) {
  constexpr int kMaxProbe = 4096;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_envs; e += (int64_t)gridDim.x * blockDim.x) {
    if (mask != nullptr && mask[e] == 0) continue;
    uint64_t s[W];
#pragma unroll
    for (int w = 0; w < W; ++w) s[w] = state[e * W + w];
    const uint64_t tag = state_tag<W>(s);
    uint64_t slot = mix64(tag) & cap_mask;
    bool done = false;
    for (int probe = 0; probe < kMaxProbe && !done; ++probe, slot = (slot + 1) & cap_mask) {
      unsigned long long cur = tags[slot];
      if (cur == 0ull) {
        cur = atomicCAS(&tags[slot], 0ull, (unsigned long long)tag);
        if (cur == 0ull) {  // claimed: publish the state (every later writer of this tag would write the same words)
#pragma unroll
          for (int w = 0; w < W; ++w) slot_state[slot * W + w] = s[w];
          cur = tag;
        }
      }
      if (cur == tag) {
        atomicAdd(&counts[slot], 1ull);
        done = true;
      }
    }
    if (!done) atomicAdd(overflow, 1u);
  }
}

// Per state of a list: which genes CAN become 1 / CAN become 0 in one perturbation-free update (over all
// predictor choices) -- the successor descriptor of the state-transition graph (graph.genSTG(),
// print_graph.py:15-21): successors = {fixed bits} x {0,1}^(genes that can do both).
template <int W>
__global__ void __launch_bounds__(256) successor_sets_kernel(const __grid_constant__ NetParams n,
                                                            const uint64_t* __restrict__ state, int64_t n_states,
                                                            uint64_t* __restrict__ can1, uint64_t* __restrict__ can0) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_states; e += (int64_t)gridDim.x * blockDim.x) {
    uint64_t s[W], c1[W], c0[W];
#pragma unroll
    for (int w = 0; w < W; ++w) { s[w] = state[e * W + w]; c1[w] = 0; c0[w] = 0; }
    for (int g = 0; g < n.n_genes; ++g) {
      for (int f = n.func_offset[g]; f < n.func_offset[g + 1]; ++f) {
        const FuncDesc d = n.funcs[f];
        uint32_t idx = 0;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const uint32_t in = ((j < 4 ? d.in03 >> (8 * j) : d.in47 >> (8 * (j - 4))) & 0xFFu);
          idx |= (uint32_t)((s[in >> 6] >> (in & 63u)) & 1ull) << j;
        }
        const uint64_t lut = ((uint64_t)d.lut_hi << 32) | d.lut_lo;   // replicated over unused inputs
        const uint64_t v = (lut >> idx) & 1ull;
        c1[g >> 6] |= v << (g & 63);
        c0[g >> 6] |= (v ^ 1ull) << (g & 63);
      }
    }
#pragma unroll
    for (int w = 0; w < W; ++w) { can1[e * W + w] = c1[w]; can0[e * W + w] = c0[w]; }
  }
}

}  // namespace pbn
