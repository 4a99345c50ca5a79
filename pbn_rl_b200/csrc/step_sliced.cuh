// Bit-sliced step kernel: the fast path, specialised per network at load time (NVRTC).
//
// Layout of the work
//   A warp owns a tile of 1024 consecutive env instances.  Lane L owns the 32 envs
//       env(L, j, c) = tile_base + 128*j + 4*L + c,   j = 0..7, c = 0..3,   slice bit b = 4*j + c
//   so every global access is a coalesced vector access (4 consecutive envs per lane: 32 B of
//   state, 4*BINS action bytes, 16 B of target ids / rewards, 8 B of step counters, 4 B of flags).
//   The lane transposes its 32 packed states into N bit-planes (plane i, bit b = gene i of env b):
//   from then on one 32-bit logic instruction advances 32 envs at once.
//
// What is generated per network (net_gen.cuh, emitted by pbn_b200.cu from the truth tables)
//   pbn_update(): for every gene, its predictor functions as LOP3 trees over the input planes --
//   the truth tables are the LOP3 immediates -- plus the predictor-selection logic.
//
// Random streams (DESIGN.md "Sliced random stream"; CPU twin: oracle/pbn_oracle.py sliced_stream)
//   group id = (env >> 10) * 32 + ((env >> 2) & 31)  (= tile * 32 + lane), slice bit as above.
//   SELECT  words, in gene order: K=1 none, K=2 one word, K=4 two words (b0,b1), K=3 six words
//           (b0,b1,c0,c1,d0,d1): pair value 3 is rejected and replaced by the next pair; slots still
//           rejected after the third pair draw 2-bit pairs from the FIX stream (gene order, bit
//           order, 16 pairs per word LSB first) until one is != 3.  sel = b0 + 2*b1.
//   PERTURB words: geometric skipping over slots q = gene*32 + bit with survival table S[0..32N].
//   RESET   per env (global env id), as in the scalar kernel.
//
// Algorithmic HBM bytes per env-step: 33 (N <= 64) / 49 (N <= 128), see step_scalar.cuh.
#pragma once
#include "pbn_common.cuh"

#ifndef PBN_N
#error "net_gen.cuh must be included first"
#endif

namespace pbn {

#define PBN_RNG_FIX 3u

constexpr int kSlots = 32 * PBN_N;            // perturbation slots per lane-tile
constexpr int kNW = PBN_NW32;                 // 32-bit words per state
constexpr int kW64 = (PBN_N <= 64) ? 1 : 2;   // 64-bit words per state in HBM
constexpr uint32_t kLastMask = (PBN_N % 32) ? ((1u << (PBN_N % 32)) - 1u) : 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t s) {
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));  // shift counts > 31 give 0
  return r;
}

template <int IMM>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(IMM));
  return r;
}

// a ? b : c, bitwise
__device__ __forceinline__ uint32_t bmux(uint32_t a, uint32_t b, uint32_t c) { return lop3<0xCA>(a, b, c); }

// In-place 32x32 bit transpose: afterwards a[i] bit b == (old a[b]) bit i.
__device__ __forceinline__ void transpose32(uint32_t (&a)[32]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const uint32_t lo = a[k], hi = a[k + 16];
    a[k] = __byte_perm(lo, hi, 0x5410);
    a[k + 16] = __byte_perm(lo, hi, 0x7632);
  }
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    if (k & 8) continue;
    const uint32_t lo = a[k], hi = a[k + 8];
    a[k] = __byte_perm(lo, hi, 0x6240);
    a[k + 8] = __byte_perm(lo, hi, 0x7351);
  }
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    if (k & 4) continue;
    const uint32_t lo = a[k], hi = a[k + 4];
    a[k] = bmux(0x0F0F0F0Fu, lo, hi << 4);
    a[k + 4] = bmux(0x0F0F0F0Fu, lo >> 4, hi);
  }
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    if (k & 2) continue;
    const uint32_t lo = a[k], hi = a[k + 2];
    a[k] = bmux(0x33333333u, lo, hi << 2);
    a[k + 2] = bmux(0x33333333u, lo >> 2, hi);
  }
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    if (k & 1) continue;
    const uint32_t lo = a[k], hi = a[k + 1];
    a[k] = bmux(0x55555555u, lo, hi << 1);
    a[k + 1] = bmux(0x55555555u, lo >> 1, hi);
  }
}

// Per-lane random-stream context.
struct SlicedRng {
  uint64_t gid;    // global slice-group id
  uint64_t step;
  uint32_t k0, k1;
  // FIX-stream reservoir (2-bit pairs)
  uint32_t fix_word, fix_cnt, fix_next;
  Philox4 fix_blk;
  // PERTURB-stream cursor
  uint32_t pert_next;
  Philox4 pert_blk;

  __device__ __forceinline__ Philox4 block(uint32_t kind, uint32_t idx) const {
    return philox_stream(gid, step, kind, idx, k0, k1);
  }
  __device__ __forceinline__ static uint32_t pick(const Philox4& b, uint32_t q) {
    return q == 0 ? b.x : q == 1 ? b.y : q == 2 ? b.z : b.w;
  }
  __device__ __forceinline__ uint32_t next_fix_pair() {
    if (fix_cnt == 0) {
      if ((fix_next & 3u) == 0u) fix_blk = block(PBN_RNG_FIX, fix_next >> 2);
      fix_word = pick(fix_blk, fix_next & 3u);
      ++fix_next;
      fix_cnt = 16;
    }
    const uint32_t pr = fix_word & 3u;
    fix_word >>= 2;
    --fix_cnt;
    return pr;
  }
  // Resolve the slots of a K=3 gene that are still rejected (b0 = b1 = 1 there) after the three
  // unconditional pair draws.
  __device__ __forceinline__ void fix3(uint32_t& b0, uint32_t& b1, uint32_t rej) {
    while (rej) {
      const uint32_t pr = next_fix_pair();
      if (pr != 3u) {
        const uint32_t m = rej & (0u - rej);
        rej ^= m;
        b0 ^= (pr & 1u) ? 0u : m;
        b1 ^= (pr & 2u) ? 0u : m;
      }
    }
  }
  __device__ __forceinline__ uint32_t next_pert_word() {
    if ((pert_next & 3u) == 0u) pert_blk = block(PBN_RNG_PERTURB, pert_next >> 2);
    const uint32_t u = pick(pert_blk, pert_next & 3u);
    ++pert_next;
    return u;
  }
};

// K=3 selection planes from six words: three unconditional pair draws, then the FIX stream.
__device__ __forceinline__ void sel3(SlicedRng& rng, uint32_t b0, uint32_t b1, uint32_t c0, uint32_t c1,
                                     uint32_t d0, uint32_t d1, uint32_t& s0, uint32_t& s1) {
  uint32_t rej = b0 & b1;
  b0 = bmux(rej, c0, b0);
  b1 = bmux(rej, c1, b1);
  rej = rej & c0 & c1;
  b0 = bmux(rej, d0, b0);
  b1 = bmux(rej, d1, b1);
  rej = rej & d0 & d1;
  rng.fix3(b0, b1, rej);
  s0 = b0;
  s1 = b1;
}

// Parity entry point: predictor choices arrive per (env, gene) as bytes (slow, tests only).
struct InjectedSel {
  const uint8_t* sel;
  int64_t e0, E;
  __device__ __forceinline__ void get(int gene, uint32_t K, uint32_t& s0, uint32_t& s1) const {
    s0 = 0u;
    s1 = 0u;
    for (int b = 0; b < 32; ++b) {
      const int64_t env = e0 + 128 * (b >> 2) + (b & 3);
      uint32_t v = (env < E) ? sel[env * PBN_N + gene] : 0u;
      v = v < K ? v : K - 1u;
      s0 |= (v & 1u) << b;
      s1 |= ((v >> 1) & 1u) << b;
    }
  }
};

#if PBN_INJECTED
#define PBN_SEL_BLOCK(name, idx)
#define PBN_SEL2(g, w0, s0, s1) sel.get(g, 2u, s0, s1)
#define PBN_SEL3(g, w0, w1, w2, w3, w4, w5, s0, s1) sel.get(g, 3u, s0, s1)
#define PBN_SEL4(g, w0, w1, s0, s1) sel.get(g, 4u, s0, s1)
#else
#define PBN_SEL_BLOCK(name, idx) const Philox4 name = sel.block(PBN_RNG_SELECT, idx)
#define PBN_SEL2(g, w0, s0, s1) do { s0 = (w0); s1 = 0u; } while (0)
#define PBN_SEL3(g, w0, w1, w2, w3, w4, w5, s0, s1) sel3(sel, w0, w1, w2, w3, w4, w5, s0, s1)
#define PBN_SEL4(g, w0, w1, s0, s1) do { s0 = (w0); s1 = (w1); } while (0)
#endif

}  // namespace pbn

#include "net_update.inc"  // generated: pbn::pbn_update(x, o, rng)

namespace pbn {

struct TileStats {
  uint32_t steps, eps, term, trunc, len, flips, pert;
};

// Next event distance of the perturbation stream (cold path: one call per perturbed gene + 1).
__device__ __noinline__ int pert_skip(SlicedRng& rng, const uint32_t* s_surv) {
  const uint32_t u = rng.next_pert_word();
  if (u < s_surv[kSlots]) return kSlots + 1;
  return count_below_survival(s_surv, kSlots, u) + 1;
}

// FULL: the tile has all 1024 envs (vector loads/stores); otherwise every access is guarded.
__device__ __forceinline__ void tile_step(const StepParams& p, const uint32_t* s_surv, const float* s_rew,
                                          const int32_t* aoffs, const uint64_t* acare, const uint64_t* aval,
                                          int64_t tile, uint64_t step_ctr, const bool FULL, TileStats& st) {
  const pbn_step_args& a = p.a;
  const NetParams& n = p.n;
  const uint32_t lane = threadIdx.x & 31u;
  const int64_t E = a.n_envs;
  const int64_t e0 = tile * 1024 + 4 * (int64_t)lane;  // local env index of (j = 0, c = 0)

  uint32_t x[kNW][32];
  uint32_t nfp[4] = {0u, 0u, 0u, 0u};  // 4-bit flip counts, env b -> nfp[b >> 3] bits 4*(b&7)..

  // ---- A. load states + actions, apply the interventions ---------------------------------
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int64_t e = e0 + 128 * j;
    uint64_t s[4][kW64];
    if (FULL) {
      const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(a.state + e * kW64);
#pragma unroll
      for (int q = 0; q < 2 * kW64; ++q) {
        const ulonglong2 v = sp[q];
        s[(2 * q) / kW64][(2 * q) % kW64] = v.x;
        s[(2 * q + 1) / kW64][(2 * q + 1) % kW64] = v.y;
      }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int w = 0; w < kW64; ++w) s[c][w] = (e + c < E) ? a.state[(e + c) * kW64 + w] : 0ull;
    }
    uint32_t aw[PBN_BINS];
#pragma unroll
    for (int k = 0; k < PBN_BINS; ++k) aw[k] = 0u;
    if (a.actions != nullptr) {
      if (FULL) {
        const uint32_t* ap = reinterpret_cast<const uint32_t*>(a.actions + e * PBN_BINS);
#pragma unroll
        for (int k = 0; k < PBN_BINS; ++k) aw[k] = ap[k];
      } else {
#pragma unroll
        for (int q = 0; q < 4 * PBN_BINS; ++q) {
          const int64_t env = e + q / PBN_BINS;
          const uint32_t v = (env < E) ? a.actions[e * PBN_BINS + q] : 0u;
          aw[q >> 2] |= v << (8 * (q & 3));
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t fl[kNW];
#pragma unroll
      for (int w = 0; w < kNW; ++w) fl[w] = 0u;
#pragma unroll
      for (int k = 0; k < PBN_BINS; ++k) {
        const int q = c * PBN_BINS + k;
        const uint32_t act = (aw[q >> 2] >> (8 * (q & 3))) & 0xFFu;
#pragma unroll
        for (int w = 0; w < kNW; ++w) fl[w] |= shl_clamp(1u, act - 1u - 32u * w);
      }
      fl[kNW - 1] &= kLastMask;
      uint32_t nf = 0;
#pragma unroll
      for (int w = 0; w < kNW; ++w) nf += __popc(fl[w]);
      const int b = 4 * j + c;
      nfp[b >> 3] |= nf << (4 * (b & 7));
#pragma unroll
      for (int w = 0; w < kNW; ++w) {
        const uint32_t sw = (uint32_t)(s[c][w >> 1] >> (32 * (w & 1)));
        x[w][b] = sw ^ fl[w];
      }
    }
  }

  // ---- B. rows -> bit-planes --------------------------------------------------------------
#pragma unroll
  for (int w = 0; w < kNW; ++w) transpose32(x[w]);

  // ---- C. predictor selection + synchronous update (generated) ---------------------------
  SlicedRng rng;
  rng.gid = (uint64_t)(((a.env_offset >> 10) + tile) * 32 + lane);
  rng.step = step_ctr;
  rng.k0 = n.k0;
  rng.k1 = n.k1;
  rng.fix_word = 0;
  rng.fix_cnt = 0;
  rng.fix_next = 0;
  rng.pert_next = 0;
  uint32_t o[kNW][32];
#if PBN_INJECTED
  InjectedSel isel{a.sel, e0, E};
  pbn_update(x, o, isel);
#else
  pbn_update(x, o, rng);
#endif

  // ---- D. perturbation ---------------------------------------------------------------------
  uint32_t npert = 0;
  const int pert_mode = n.pert_mode;  // warp-uniform
  if (pert_mode != PBN_PERT_NONE) {
    uint32_t M = 0u;
#if PBN_INJECTED
    {
      if (a.pert_mask != nullptr) {
        // rows -> planes of the injected masks (slow path)
        uint32_t q[kNW][32];
#pragma unroll
        for (int b = 0; b < 32; ++b) {
          const int64_t env = e0 + 128 * (b >> 2) + (b & 3);
#pragma unroll
          for (int w = 0; w < kNW; ++w) {
            const uint64_t v = (env < E) ? a.pert_mask[env * kW64 + (w >> 1)] : 0ull;
            q[w][b] = (uint32_t)(v >> (32 * (w & 1)));
          }
        }
#pragma unroll
        for (int w = 0; w < kNW; ++w) transpose32(q[w]);
#pragma unroll
        for (int w = 0; w < kNW; ++w)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (32 * w + i >= PBN_N) continue;
            M |= q[w][i];
            npert += __popc(q[w][i]);
            if (pert_mode == PBN_PERT_A) x[w][i] ^= q[w][i];
            else if (pert_mode == PBN_PERT_B) o[w][i] ^= q[w][i];
            else o[w][i] = bmux(q[w][i], ~x[w][i], o[w][i]);
          }
      }
    }
#else
    if (n.pert_rng) {
      int pos = -1 + pert_skip(rng, s_surv);
#pragma unroll
      for (int w = 0; w < kNW; ++w)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (32 * w + i >= PBN_N) continue;
          uint32_t q = 0u;
          while ((pos >> 5) == 32 * w + i) {
            q |= 1u << (pos & 31);
            pos += pert_skip(rng, s_surv);
          }
          M |= q;
          npert += __popc(q);
          if (pert_mode == PBN_PERT_A) x[w][i] ^= q;
          else if (pert_mode == PBN_PERT_B) o[w][i] ^= q;
          else o[w][i] = bmux(q, ~x[w][i], o[w][i]);
        }
    }
#endif
    if (pert_mode == PBN_PERT_A) {
      // envs with any perturbed gene keep s1 XOR pert (x already holds it), the others take f
#pragma unroll
      for (int w = 0; w < kNW; ++w)
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (32 * w + i < PBN_N) o[w][i] = bmux(M, x[w][i], o[w][i]);
    }
  }

  // ---- E. bit-planes -> rows ---------------------------------------------------------------
#pragma unroll
  for (int w = 0; w < kNW; ++w) transpose32(o[w]);

  // ---- F. target test, counters, reward, stores ---------------------------------------------
  uint32_t H = 0u, TR = 0u, VALID = 0u;
  uint32_t len_sum = 0u, flips = 0u;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int64_t e = e0 + 128 * j;
    int32_t tg[4] = {-1, -1, -1, -1};
    uint32_t tt[4] = {0u, 0u, 0u, 0u};
    if (a.target_id != nullptr) {
      if (FULL) {
        const int4 v = *reinterpret_cast<const int4*>(a.target_id + e);
        tg[0] = v.x; tg[1] = v.y; tg[2] = v.z; tg[3] = v.w;
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) tg[c] = (e + c < E) ? a.target_id[e + c] : -1;
      }
    }
    if (a.t != nullptr) {
      if (FULL) {
        const uint2 v = *reinterpret_cast<const uint2*>(a.t + e);
        tt[0] = v.x & 0xFFFFu; tt[1] = v.x >> 16; tt[2] = v.y & 0xFFFFu; tt[3] = v.y >> 16;
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) tt[c] = (e + c < E) ? a.t[e + c] : 0u;
      }
    }
    uint64_t y[4][kW64];
    float rw[4];
    uint32_t hbits = 0u, tbits = 0u;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int b = 4 * j + c;
#pragma unroll
      for (int w = 0; w < kW64; ++w) {
        const uint32_t lo = o[2 * w][b];
        const uint32_t hi = (2 * w + 1 < kNW) ? o[(2 * w + 1 < kNW) ? 2 * w + 1 : 0][b] : 0u;
        y[c][w] = ((uint64_t)hi << 32) | lo;
      }
      bool hit = false;
      if (tg[c] >= 0 && tg[c] < n.n_attr) hit = in_attractor<kW64>(aoffs, acare, aval, tg[c], y[c]);
      const uint32_t t1 = tt[c] < 65535u ? tt[c] + 1u : 65535u;
      const bool trunc = !hit && n.horizon > 0 && t1 >= (uint32_t)n.horizon;
      const uint32_t nf = (nfp[b >> 3] >> (4 * (b & 7))) & 0xFu;
      rw[c] = s_rew[nf + (hit ? 9u : 0u)];
      const bool valid = FULL || (e + c < E);
      hbits |= (hit && valid ? 1u : 0u) << c;
      tbits |= (trunc && valid ? 1u : 0u) << c;
      VALID |= (valid ? 1u : 0u) << b;
      if ((hit || trunc) && valid) len_sum += t1;
      if (valid) flips += nf;
      tt[c] = t1;
    }
    H |= hbits << (4 * j);
    TR |= tbits << (4 * j);
    const uint32_t hbytes = (hbits * 0x00204081u) & 0x01010101u;
    const uint32_t tbytes = (tbits * 0x00204081u) & 0x01010101u;
    if (FULL) {
      ulonglong2* sp = reinterpret_cast<ulonglong2*>(a.state + e * kW64);
#pragma unroll
      for (int q = 0; q < 2 * kW64; ++q)
        sp[q] = make_ulonglong2(y[(2 * q) / kW64][(2 * q) % kW64], y[(2 * q + 1) / kW64][(2 * q + 1) % kW64]);
      if (a.final_state != nullptr) {
        ulonglong2* fp = reinterpret_cast<ulonglong2*>(a.final_state + e * kW64);
#pragma unroll
        for (int q = 0; q < 2 * kW64; ++q)
          fp[q] = make_ulonglong2(y[(2 * q) / kW64][(2 * q) % kW64], y[(2 * q + 1) / kW64][(2 * q + 1) % kW64]);
      }
      if (a.reward != nullptr) *reinterpret_cast<float4*>(a.reward + e) = make_float4(rw[0], rw[1], rw[2], rw[3]);
      if (a.terminated != nullptr) *reinterpret_cast<uint32_t*>(a.terminated + e) = hbytes;
      if (a.truncated != nullptr) *reinterpret_cast<uint32_t*>(a.truncated + e) = tbytes;
      if (a.t != nullptr) *reinterpret_cast<uint2*>(a.t + e) = make_uint2(tt[0] | (tt[1] << 16), tt[2] | (tt[3] << 16));
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (e + c >= E) continue;
#pragma unroll
        for (int w = 0; w < kW64; ++w) {
          a.state[(e + c) * kW64 + w] = y[c][w];
          if (a.final_state != nullptr) a.final_state[(e + c) * kW64 + w] = y[c][w];
        }
        if (a.reward != nullptr) a.reward[e + c] = rw[c];
        if (a.terminated != nullptr) a.terminated[e + c] = (uint8_t)((hbits >> c) & 1u);
        if (a.truncated != nullptr) a.truncated[e + c] = (uint8_t)((tbits >> c) & 1u);
        if (a.t != nullptr) a.t[e + c] = (uint16_t)tt[c];
      }
    }
  }

  // ---- G. auto-reset of finished envs (sparse: scattered writes after the vector stores) ----
  uint32_t D = (H | TR) & VALID;
  st.steps += __popc(VALID);
  st.eps += __popc(D);
  st.term += __popc(H & VALID);
  st.trunc += __popc(TR & VALID);
  st.len += len_sum;
  st.flips += flips;
  st.pert += npert;
  if (a.flags & PBN_STEP_AUTORESET) {
    while (D) {
      const int b = __ffs(D) - 1;
      D &= D - 1u;
      const int64_t env = e0 + 128 * (b >> 2) + (b & 3);
      const Philox4 r = philox_stream((uint64_t)(a.env_offset + env), step_ctr, PBN_RNG_RESET, 0, n.k0, n.k1);
      uint64_t s[kW64];
      int src, tgt;
      reset_draw<kW64>(n, r, s, src, tgt);
#pragma unroll
      for (int w = 0; w < kW64; ++w) a.state[env * kW64 + w] = s[w];
      a.target_id[env] = tgt;
      if (a.source_id != nullptr) a.source_id[env] = src;
      a.t[env] = 0;
    }
  }
}

extern "C" __global__ void __launch_bounds__(PBN_THREADS, PBN_MIN_BLOCKS)
pbn_step_sliced(const __grid_constant__ StepParams p, const SlicedSmemLayout L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const NetParams& n = p.n;
  const pbn_step_args& a = p.a;
  uint32_t* s_surv = reinterpret_cast<uint32_t*>(smem_raw + L.surv_off);
  float* s_rew = reinterpret_cast<float*>(smem_raw + L.rew_off);
  uint64_t* s_acare = reinterpret_cast<uint64_t*>(smem_raw + L.acare_off);
  uint64_t* s_aval = reinterpret_cast<uint64_t*>(smem_raw + L.aval_off);
  int32_t* s_aoffs = reinterpret_cast<int32_t*>(smem_raw + L.aoffs_off);

  if (n.pert_rng)
    for (int i = threadIdx.x; i <= kSlots; i += blockDim.x) s_surv[i] = n.surv_sliced[i];
  if (threadIdx.x < 18) {
    const uint32_t nf = threadIdx.x % 9u;
    const bool hit = threadIdx.x >= 9;
    const float base = __fadd_rn(n.r_step, __fmul_rn(n.r_action, (float)nf));
    s_rew[threadIdx.x] = __fadd_rn(base, hit ? n.r_success : 0.0f);
  }
  if (L.attractors_in_smem) {
    for (int i = threadIdx.x; i < n.n_attr_states * kW64; i += blockDim.x) {
      s_acare[i] = n.attr_care[i];
      s_aval[i] = n.attr_val[i];
    }
    for (int i = threadIdx.x; i <= n.n_attr; i += blockDim.x) s_aoffs[i] = n.attr_offset[i];
  }
  __syncthreads();
  const int32_t* aoffs = L.attractors_in_smem ? s_aoffs : n.attr_offset;
  const uint64_t* acare = L.attractors_in_smem ? s_acare : n.attr_care;
  const uint64_t* aval = L.attractors_in_smem ? s_aval : n.attr_val;

  const uint64_t step_ctr = effective_step(a);
  const int64_t n_tiles = (a.n_envs + 1023) >> 10;
  const int warps_per_block = blockDim.x >> 5;
  TileStats st = {0u, 0u, 0u, 0u, 0u, 0u, 0u};
  for (int64_t tile = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); tile < n_tiles;
       tile += (int64_t)gridDim.x * warps_per_block) {
    tile_step(p, s_surv, s_rew, aoffs, acare, aval, tile, step_ctr, (tile + 1) * 1024 <= a.n_envs, st);
  }
  if (a.stats != nullptr) {
    const uint32_t v[7] = {st.steps, st.eps, st.term, st.trunc, st.len, st.flips, st.pert};
#pragma unroll
    for (int q = 0; q < 7; ++q) {
      const uint32_t x = __reduce_add_sync(0xFFFFFFFFu, v[q]);
      if ((threadIdx.x & 31u) == 0u && x != 0u) atomicAdd(&a.stats[q], (unsigned long long)x);
    }
  }
  bump_device_step(a, p.ticket);
}

}  // namespace pbn
