// Microbenchmark: Philox4x32-10 blocks per cycle per SM sub-partition at various ILP / occupancy.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../pbn_rl_b200/csrc/philox.cuh"
using namespace pbn;
template <int ILP>
__global__ void k(uint32_t* out, int iters, uint32_t k0, uint32_t k1) {
  uint32_t acc = 0;
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = 0; i < iters; ++i) {
    Philox4 r[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) r[j] = philox4x32_10(id, i * ILP + j, acc & 1, 7, k0, k1);
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc ^= r[j].x ^ r[j].y ^ r[j].z ^ r[j].w;
  }
  out[id] = acc;
}
template <int ILP>
void run(int warps_per_sm, int iters) {
  const int threads = 128, blocks = 148 * warps_per_sm / 4;
  uint32_t* out; cudaMalloc(&out, sizeof(uint32_t) * threads * blocks);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP><<<blocks, threads>>>(out, 10, 1, 2);
  cudaEventRecord(e0);
  k<ILP><<<blocks, threads>>>(out, iters, 1, 2);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double blocks_per_smsp = (double)iters * ILP * warps_per_sm / 4.0;
  const double cycles = ms * 1e-3 * 1.965e9;
  printf("ILP %d warps/SM %2d: %.1f cycles per Philox block per SMSP (%.2f us)\n", ILP, warps_per_sm, cycles / blocks_per_smsp, ms * 1e3);
  cudaFree(out);
}
int main() {
  for (int w : {4, 8, 16, 28, 32, 64}) { run<1>(w, 2000); run<2>(w, 1000); run<4>(w, 500); }
  return 0;
}
