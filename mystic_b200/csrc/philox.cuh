// Philox4x32-10 counter RNG (Salmon, Moraes, Dror, Shaw, SC'11) and the draw protocol of the
// DE engine.  The same protocol is restated in numpy in oracle/de_oracle.py so Philox-mode
// trajectories can be checked bit for bit on the CPU.
//
// counter = (block, global candidate index, generation, stream), key = 64-bit seed.
//   stream 0 (donors):   64-bit words r_{2b} = x0<<32|x1, r_{2b+1} = x2<<32|x3 of block b.
//                        donor i: rank = mulhi64(r_i, NP-1-i) among the indices not yet excluded
//                        ({candidate, earlier donors}); start index = mulhi64(r_k, D).
//                        Replaces random.sample / random.randrange (mystic/strategy.py:27,42).
//   stream 1 (crossover): word i = lane i%4 of block i/4; Bernoulli(CR) is `word < thr`,
//                        thr = floor(CR * 2^32) in [0, 2^32].  Replaces random.random() < CR
//                        (strategy.py:49,80).
//   stream 2 (init):     u = (r64 >> 11) * 2^-53, two doubles per block
//                        (abstract_solver.py:532-534 random.uniform).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define MB2_HD __host__ __device__ __forceinline__
#else
#define MB2_HD inline
#endif

namespace mb2 {

enum : uint32_t { STREAM_DONORS = 0, STREAM_XOVER = 1, STREAM_INIT = 2 };

struct u32x4 {
  uint32_t x, y, z, w;
};

MB2_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

MB2_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}

MB2_HD u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                           uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += W0;
    k1 += W1;
  }
  return u32x4{c0, c1, c2, c3};
}

// k distinct donors != cand, uniform over ordered k-tuples, and the crossover start index.
// K <= 5.  `cand_local` indexes the shard the donors are drawn from (size np_local);
// `cand_global` keys the counter so the draws do not depend on how rows are laid out.
template <int K>
MB2_HD void draw_donors(uint32_t k0, uint32_t k1, uint32_t generation, uint32_t cand_global,
                        int64_t cand_local, int64_t np_local, int dim, int64_t* donors,
                        int* nstart) {
  uint64_t r[6];
#pragma unroll
  for (int b = 0; b < (K + 2) / 2; ++b) {
    const u32x4 v = philox4x32_10((uint32_t)b, cand_global, generation, STREAM_DONORS, k0, k1);
    r[2 * b] = ((uint64_t)v.x << 32) | v.y;
    r[2 * b + 1] = ((uint64_t)v.z << 32) | v.w;
  }
  // sorted exclusion list, padded with +inf
  int64_t ex[K + 1];
  ex[0] = cand_local;
#pragma unroll
  for (int i = 1; i <= K; ++i) ex[i] = INT64_MAX;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    int64_t idx = (int64_t)mulhi64(r[i], (uint64_t)(np_local - 1 - i));
#pragma unroll
    for (int e = 0; e <= i; ++e) idx += (idx >= ex[e]) ? 1 : 0;
    donors[i] = idx;
    // insert idx into the sorted list ex[0..i] -> ex[0..i+1]
    int64_t carry = idx;
#pragma unroll
    for (int e = 0; e <= i + 1; ++e) {
      if (e <= i + 1) {
        const int64_t cur = ex[e];
        if (carry < cur) {
          ex[e] = carry;
          carry = cur;
        }
      }
    }
  }
  *nstart = (int)mulhi64(r[K], (uint64_t)dim);
}

// Ranks -> indices: donor i is the rank-th index among those not yet excluded ({candidate,
// earlier donors}); r[0..K-1] are the donor words, r[K] the word of the start index.
template <int K>
MB2_HD void resolve_donors(const uint64_t* r, int64_t cand_local, int64_t np_local, int dim, int64_t* donors,
                           int* nstart) {
  int64_t ex[K + 1];  // sorted exclusion list, padded with +inf
  ex[0] = cand_local;
#pragma unroll
  for (int i = 1; i <= K; ++i) ex[i] = INT64_MAX;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    int64_t idx = (int64_t)mulhi64(r[i], (uint64_t)(np_local - 1 - i));
#pragma unroll
    for (int e = 0; e <= i; ++e) idx += (idx >= ex[e]) ? 1 : 0;
    donors[i] = idx;
    int64_t carry = idx;
#pragma unroll
    for (int e = 0; e <= i + 1; ++e) {
      const int64_t cur = ex[e];
      if (carry < cur) {
        ex[e] = carry;
        carry = cur;
      }
    }
  }
  *nstart = (int)mulhi64(r[K], (uint64_t)dim);
}

#if defined(__CUDACC__)
// Warp-collective variant of draw_donors (same results): lane b computes Philox block b, the
// 64-bit words are gathered by shuffle, then every lane resolves ranks to indices redundantly.
// All 32 lanes of the warp must call it.
template <int K>
__device__ __forceinline__ void draw_donors_warp(uint32_t k0, uint32_t k1, uint32_t generation, uint32_t cand_global,
                                                 int64_t cand_local, int64_t np_local, int dim, int lane,
                                                 int64_t* donors, int* nstart) {
  const u32x4 v = philox4x32_10((uint32_t)lane, cand_global, generation, STREAM_DONORS, k0, k1);
  const uint64_t even = ((uint64_t)v.x << 32) | v.y, odd = ((uint64_t)v.z << 32) | v.w;
  uint64_t r[6];
#pragma unroll
  for (int b = 0; b < (K + 2) / 2; ++b) {
    r[2 * b] = __shfl_sync(0xffffffffu, even, b);
    r[2 * b + 1] = __shfl_sync(0xffffffffu, odd, b);
  }
  int64_t ex[K + 1];
  ex[0] = cand_local;
#pragma unroll
  for (int i = 1; i <= K; ++i) ex[i] = INT64_MAX;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    int64_t idx = (int64_t)mulhi64(r[i], (uint64_t)(np_local - 1 - i));
#pragma unroll
    for (int e = 0; e <= i; ++e) idx += (idx >= ex[e]) ? 1 : 0;
    donors[i] = idx;
    int64_t carry = idx;
#pragma unroll
    for (int e = 0; e <= i + 1; ++e) {
      const int64_t cur = ex[e];
      if (carry < cur) {
        ex[e] = carry;
        carry = cur;
      }
    }
  }
  *nstart = (int)mulhi64(r[K], (uint64_t)dim);
}
#endif

MB2_HD uint64_t crossover_threshold(double cr) {
  if (!(cr > 0.0)) return 0ull;
  if (cr >= 1.0) return 4294967296ull;
  return (uint64_t)(cr * 4294967296.0);  // floor for positive values
}

MB2_HD double u64_to_unit(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

}  // namespace mb2
