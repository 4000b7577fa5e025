// Philox4x32-10 counter RNG (Salmon, Moraes, Dror, Shaw, SC'11) and the draw protocol of the
// DE engine.  The same protocol is restated in numpy in oracle/de_oracle.py so Philox-mode
// trajectories can be checked bit for bit on the CPU.
//
// counter = (block, global candidate index, generation, stream), key = 64-bit seed.
// Draw protocol, version 2 (MB2_RNG_PROTOCOL; version 1 spent one 32-bit word per crossover decision):
//   stream 0 (donors):   64-bit words r_{2b} = x0<<32|x1, r_{2b+1} = x2<<32|x3 of block b.
//                        donor i: rank = mulhi64(r_i, NP-1-i) among the indices not yet excluded
//                        ({candidate, earlier donors}); start index = mulhi64(r_k, D).
//                        Replaces random.sample / random.randrange (mystic/strategy.py:27,42).
//                        Exponential crossover (strategy.py:48-56: mutate while random() < CR, at most D
//                        times): the number of leading successes is ONE geometric draw from the low 32
//                        bits w of r_k:  L = min(D, #{l >= 1 : w < T[l]}),  T[0] = 2^32,
//                        T[l] = (T[l-1] * thr) >> 32,  thr = floor(CR * 2^32)  -- integer arithmetic only,
//                        so the CPU oracle reproduces it bit for bit; P(L >= l) = T[l] / 2^32 ~ CR^l,
//                        the distribution of the reference's loop (w and the start index share r_k: their
//                        joint law differs from independence by < D / 2^32).
//   stream 1 (crossover, binomial strategies only: strategy.py:79-85, one decision per dimension):
//                        16-bit fields, eight per block: dimension 8q+e reads field e of block q = bits
//                        16*(e&1) .. 16*(e&1)+15 of word e>>1; Bernoulli(CR) is `field < thr16`,
//                        thr16 = floor(CR * 2^16) in [0, 2^16].  Replaces random.random() < CR.
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
constexpr int kRngProtocol = 2;

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
                        int* nstart, uint32_t* xword = nullptr) {
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
  if (xword != nullptr) *xword = (uint32_t)r[K];  // geometric draw of the exponential crossover length
}

// Ranks -> indices: donor i is the rank-th index among those not yet excluded ({candidate,
// earlier donors}); r[0..K-1] are the donor words, r[K] the word of the start index.
template <int K>
MB2_HD void resolve_donors(const uint64_t* r, int64_t cand_local, int64_t np_local, int dim, int64_t* donors,
                           int* nstart, uint32_t* xword = nullptr) {
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
  if (xword != nullptr) *xword = (uint32_t)r[K];
}

#if defined(__CUDACC__)
// Warp-collective variant of draw_donors (same results): lane b computes Philox block b, the
// 64-bit words are gathered by shuffle, then every lane resolves ranks to indices redundantly.
// All 32 lanes of the warp must call it.
template <int K>
__device__ __forceinline__ void draw_donors_warp(uint32_t k0, uint32_t k1, uint32_t generation, uint32_t cand_global,
                                                 int64_t cand_local, int64_t np_local, int dim, int lane,
                                                 int64_t* donors, int* nstart, uint32_t* xword = nullptr) {
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
  if (xword != nullptr) *xword = (uint32_t)r[K];
}
#endif

MB2_HD uint64_t crossover_threshold(double cr) {
  if (!(cr > 0.0)) return 0ull;
  if (cr >= 1.0) return 4294967296ull;
  return (uint64_t)(cr * 4294967296.0);  // floor for positive values
}

// binomial crossover: Bernoulli(CR) on a 16-bit field
MB2_HD uint32_t crossover_threshold16(double cr) {
  if (!(cr > 0.0)) return 0u;
  if (cr >= 1.0) return 65536u;
  return (uint32_t)(cr * 65536.0);  // floor for positive values
}

// field e (0..7) of a crossover block: the decision of dimension 8q+e
MB2_HD uint32_t xover_field(const u32x4& w, int e) {
  const uint32_t word = (e & 4) ? ((e & 2) ? w.w : w.z) : ((e & 2) ? w.y : w.x);
  return (e & 1) ? (word >> 16) : (word & 0xffffu);
}
// the two fields of word `word` as (even dimension, odd dimension)
MB2_HD bool field_lo_hit(uint32_t word, uint32_t thr16) { return (word & 0xffffu) < thr16; }
MB2_HD bool field_hi_hit(uint32_t word, uint32_t thr16) { return (word >> 16) < thr16; }

// Exponential crossover length table: T[l] for l = 1..nmax (tab[l-1] = T[l]), nmax = min(D, last l with
// T[l] > 0).  thr >= 2^32 (CR >= 1): every draw succeeds, the length is D (no table; returns -1).
// `tab` must hold `dim` entries.
inline int build_exp_length_table(uint64_t thr, int dim, uint32_t* tab) {
  if (thr >= 4294967296ull) return -1;
  uint64_t t = 4294967296ull;
  int n = 0;
  while (n < dim) {
    t = (t * thr) >> 32;
    if (t == 0) break;
    tab[n++] = (uint32_t)t;
  }
  return n;
}

struct ExpLength {
  const uint32_t* tab;  // T[1..nmax]
  int nmax;             // < 0: CR >= 1, the length is always D
  float c;              // 1 / -log2(CR): first guess of the length from log2 of the word
};

#if defined(__CUDACC__)
// L = #{l in 1..nmax : w < T[l]} (T decreasing): a guess from the word's logarithm, corrected against the table
__device__ __forceinline__ int exp_length_v2(uint32_t w, const ExpLength& X, int dim) {
  if (X.nmax < 0) return dim;
  int n = __float2int_rz((32.0f - __log2f((float)w + 0.5f)) * X.c);  // saturating conversion
  n = n < 0 ? 0 : (n > X.nmax ? X.nmax : n);
  while (n < X.nmax && w < __ldg(X.tab + n)) ++n;       // T[n+1] = tab[n]
  while (n > 0 && !(w < __ldg(X.tab + n - 1))) --n;     // T[n]   = tab[n-1]
  return n;
}
#endif

MB2_HD double u64_to_unit(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

}  // namespace mb2
