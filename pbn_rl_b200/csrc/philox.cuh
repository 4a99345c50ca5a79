// Philox4x32-10 (Salmon et al., SC'11), counter-based: no RNG state in HBM.
// Each round is 2 IMAD.WIDE (fma pipe) + 2 three-input XORs (one LOP3 each, alu pipe);
// the key schedule is warp-uniform and folds into immediates/uniform registers.
#pragma once
#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif

namespace pbn {

struct Philox4 {
  uint32_t x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0;
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += W0;
    k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

// Same function with the key schedule precomputed (rk[2r], rk[2r+1] = round-r keys): when rk lives in
// the kernel parameter block the keys become constant-bank operands of the XORs.
__device__ __forceinline__ Philox4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                    const uint32_t (&rk)[20]) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0;
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ rk[2 * r];
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ rk[2 * r + 1];
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
  }
  return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ Philox4 philox_stream_rk(uint64_t id, uint64_t step, uint32_t kind, uint32_t idx,
                                                    const uint32_t (&rk)[20]) {
  const uint32_t c3 = ((uint32_t)(step >> 32) & 0xFFFFu) | (kind << 28) | (idx << 16);
  return philox4x32_10_rk((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)step, c3, rk);
}

// Counter layout shared by every stream of the library (include/pbn_b200.h "Random streams").
__device__ __forceinline__ Philox4 philox_stream(uint64_t id, uint64_t step, uint32_t kind, uint32_t idx,
                                                 uint32_t k0, uint32_t k1) {
  const uint32_t c3 = ((uint32_t)(step >> 32) & 0xFFFFu) | (kind << 28) | (idx << 16);
  return philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)step, c3, k0, k1);
}

}  // namespace pbn
