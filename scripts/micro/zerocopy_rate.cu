// Zero-copy (mapped pinned host memory) write/read rate from SM threads vs the copy engine.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o zerocopy_rate zerocopy_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void wr(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
int main() {
  const size_t bytes = 16u << 20, n = bytes / 16;
  uint4 *h, *d, *hd;
  cudaHostAlloc(&h, bytes, cudaHostAllocDefault);
  cudaMalloc(&d, bytes);
  cudaMemset(d, 1, bytes);
  cudaHostGetDevicePointer(&hd, h, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int grid : {8, 32, 148, 592}) for (int mode = 0; mode < 3; ++mode) {
    float best = 1e9f;
    for (int it = 0; it < 6; ++it) {
      cudaEventRecord(e0);
      if (mode == 0) wr<<<grid, 256>>>(hd, d, n);           // SM writes to host
      else if (mode == 1) wr<<<grid, 256>>>(d, hd, n);      // SM reads from host
      else cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (it && ms < best) best = ms;
    }
    printf("grid %4d %-14s %.3f ms  %.1f GB/s\n", grid, mode == 0 ? "sm->host write" : mode == 1 ? "sm<-host read" : "memcpy D2H", best, bytes / best / 1e6);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
