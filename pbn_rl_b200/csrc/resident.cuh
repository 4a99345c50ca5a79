// Boundary kernels of the plane-resident env state (csrc/step_planes.cuh): row-format arrays <-> resident block.
// One warp per (tile, column): lane b holds env  tile*1024 + 128*(b>>2) + 4*column + (b&3);  a plane word is one
// __ballot_sync over the column's 32 envs.  Run once per import/export, not per step.
#pragma once
#include "pbn_common.cuh"

namespace pbn {

constexpr int kResTidPlanes = 8, kResTPlanes = 16;
constexpr uint32_t kResNoTarget = 255u;
__host__ __device__ inline int resident_rows(int n_genes) { return 2 * n_genes + kResTidPlanes + kResTPlanes; }

template <int W>
__global__ void __launch_bounds__(256) resident_import_kernel(const __grid_constant__ NetParams n, uint32_t* __restrict__ res,
                                                             const uint64_t* __restrict__ state,
                                                             const int32_t* __restrict__ target_id,
                                                             const uint16_t* __restrict__ t, int64_t n_envs) {
  const int N = n.n_genes, rows = resident_rows(N);
  const uint32_t lane = threadIdx.x & 31u;
  const int64_t n_cols = ((n_envs + 1023) >> 10) * 32;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int max_tid = n.n_attr < (int)kResNoTarget ? n.n_attr : (int)kResNoTarget;
  for (int64_t col = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; col < n_cols; col += warps) {
    const int64_t tile = col >> 5;
    const uint32_t L = (uint32_t)(col & 31);
    const int64_t env = tile * 1024 + 128 * (lane >> 2) + 4 * L + (lane & 3u);
    const bool ok = env < n_envs;
    uint64_t s[W], tg[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
      s[w] = ok ? state[env * W + w] : 0ull;
      tg[w] = 0ull;
    }
    uint32_t tid = kResNoTarget;
    if (ok && target_id != nullptr) {
      const int v = target_id[env];
      if (v >= 0 && v < max_tid) tid = (uint32_t)v;
    }
    if (tid != kResNoTarget) {
      const int e0 = n.attr_offset[tid];   // the attractor's first state, '*' -> 0 (env.reset()'s `target`)
#pragma unroll
      for (int w = 0; w < W; ++w) tg[w] = n.attr_val[(size_t)e0 * W + w];
    }
    const uint32_t tt = (ok && t != nullptr) ? t[env] : 0u;
    uint32_t* blk = res + tile * (int64_t)rows * 32 + L;
    for (int g = 0; g < N; ++g) {
      const uint32_t ws = __ballot_sync(0xFFFFFFFFu, (s[W == 1 ? 0 : (g >> 6)] >> (g & 63)) & 1ull);
      const uint32_t wt = __ballot_sync(0xFFFFFFFFu, (tg[W == 1 ? 0 : (g >> 6)] >> (g & 63)) & 1ull);
      if (lane == (uint32_t)(g & 31)) {
        blk[g * 32] = ws;
        blk[(N + g) * 32] = wt;
      }
    }
    for (int k = 0; k < kResTidPlanes; ++k) {
      const uint32_t wk = __ballot_sync(0xFFFFFFFFu, (tid >> k) & 1u);
      if (lane == (uint32_t)k) blk[(2 * N + k) * 32] = wk;
    }
    for (int k = 0; k < kResTPlanes; ++k) {
      const uint32_t wk = __ballot_sync(0xFFFFFFFFu, (tt >> k) & 1u);
      if (lane == (uint32_t)k) blk[(2 * N + kResTidPlanes + k) * 32] = wk;
    }
  }
}

template <int W>
__global__ void __launch_bounds__(256) resident_export_kernel(const __grid_constant__ NetParams n, const uint32_t* __restrict__ res,
                                                             uint64_t* __restrict__ state, int32_t* __restrict__ target_id,
                                                             uint16_t* __restrict__ t, int64_t n_envs) {
  const int N = n.n_genes, rows = resident_rows(N);
  const uint32_t lane = threadIdx.x & 31u;
  const int64_t n_cols = ((n_envs + 1023) >> 10) * 32;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t col = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; col < n_cols; col += warps) {
    const int64_t tile = col >> 5;
    const uint32_t L = (uint32_t)(col & 31);
    const int64_t env = tile * 1024 + 128 * (lane >> 2) + 4 * L + (lane & 3u);
    if (env >= n_envs) continue;
    const uint32_t* blk = res + tile * (int64_t)rows * 32 + L;
    if (state != nullptr) {
      uint64_t s[W];
#pragma unroll
      for (int w = 0; w < W; ++w) s[w] = 0ull;
      for (int g = 0; g < N; ++g) s[W == 1 ? 0 : (g >> 6)] |= (uint64_t)((blk[g * 32] >> lane) & 1u) << (g & 63);
#pragma unroll
      for (int w = 0; w < W; ++w) state[env * W + w] = s[w];
    }
    if (target_id != nullptr) {
      uint32_t tid = 0u;
      for (int k = 0; k < kResTidPlanes; ++k) tid |= ((blk[(2 * N + k) * 32] >> lane) & 1u) << k;
      target_id[env] = tid == kResNoTarget ? -1 : (int32_t)tid;
    }
    if (t != nullptr) {
      uint32_t tt = 0u;
      for (int k = 0; k < kResTPlanes; ++k) tt |= ((blk[(2 * N + kResTidPlanes + k) * 32] >> lane) & 1u) << k;
      t[env] = (uint16_t)tt;
    }
  }
}

}  // namespace pbn
