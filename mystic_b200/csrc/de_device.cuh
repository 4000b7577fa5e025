// Device-side building blocks of the fused DE generation kernel: unfused FP64 arithmetic,
// streaming loads, warp reductions, the mystic.models cost functions and the penalty terms.
// Reference citations are relative to the mystic repository root.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <math_constants.h>

#include "philox.cuh"

namespace mb2 {

// ---- unfused arithmetic: python evaluates a + F*(b-c) with separately rounded operations;
// nvcc would contract to FMA, so the reference-order formulas use the _rn intrinsics.
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }

// ---- streaming global loads (read once per generation: do not allocate in L1)
__device__ __forceinline__ double2 ld_stream2(const double* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ double ld_stream1(const double* p) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream2(double* p, double a, double b) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

// 4 consecutive elements starting at i0 (multiple of 4) of a row of length D.
// VEC: D is even, the row base is 16-byte aligned.  Missing elements read as 0.
template <bool VEC>
__device__ __forceinline__ void load4(const double* __restrict__ row, int i0, int D, double (&v)[4]) {
  if (VEC) {
    double2 a = ld_stream2(row + i0);
    v[0] = a.x;
    v[1] = a.y;
    if (i0 + 2 < D) {
      double2 b = ld_stream2(row + i0 + 2);
      v[2] = b.x;
      v[3] = b.y;
    } else {
      v[2] = 0.0;
      v[3] = 0.0;
    }
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = (i0 + e < D) ? ld_stream1(row + i0 + e) : 0.0;
  }
}

template <bool VEC>
__device__ __forceinline__ void store4(double* __restrict__ row, int i0, int D, const double (&v)[4]) {
  if (VEC) {
    st_stream2(row + i0, v[0], v[1]);
    if (i0 + 2 < D) st_stream2(row + i0 + 2, v[2], v[3]);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (i0 + e < D) row[i0 + e] = v[e];
  }
}

// ---- cooperative groups of G warps (G in {1,2,4,8}) that own one candidate row ------------
// Reductions are CANONICAL: element i belongs to virtual thread j = i mod 256; virtual thread
// sums are combined by a fixed xor tree over the 32 lanes and then over the 8 virtual warps as
// ((w0+w1)+(w2+w3))+((w4+w5)+(w6+w7)).  A group of G warps emulates the 256 virtual threads with
// M = 8/G accumulators per thread, so every kernel (any G) produces bit-identical energies.
template <int GW>
struct Group {
  static constexpr int kWarps = GW;          // warps per group (1, 2, 4 or 8)
  static constexpr int kThreads = 32 * GW;
  static constexpr int kAcc = 8 / GW;        // accumulators per thread (virtual threads emulated)
  int tig;        // thread index in the group
  int wig;        // warp index in the group
  int lane;
  int bar_id;     // named barrier of the group (unused when GW == 1)
  double* scratch;  // 16 doubles of shared memory per group
};

template <int GW>
__device__ __forceinline__ void group_sync(const Group<GW>& g) {
  if (GW == 1)
    __syncwarp();
  else
    asm volatile("bar.sync %0, %1;" ::"r"(g.bar_id), "n"(32 * GW) : "memory");
}

// barrier + OR reduction of a predicate over the group
template <int GW>
__device__ __forceinline__ bool group_any(const Group<GW>& g, bool v) {
  if (GW == 1) return __any_sync(0xffffffffu, v);
  unsigned r;
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.u32 q, %1, 0;\n"
      "bar.red.or.pred p, %2, %3, q;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(r)
      : "r"((unsigned)v), "r"(g.bar_id), "n"(32 * GW)
      : "memory");
  return r != 0;
}

// Reduction of the 8/GW accumulators of every thread over the group.  Within a warp the NA
// accumulators are reduced together by a butterfly: at offsets 16, 8, 4 each lane keeps half of
// its values and hands the other half to its partner, then the single remaining value is
// finished with the usual xor steps.  Every partial sum is the same value the plain per-
// accumulator xor tree (offsets 16,8,4,2,1) would hold, so results are bit-identical to it, at
// 9 instead of 40 shuffles for NA = 8.
template <bool PROD, int GW>
__device__ __forceinline__ double group_reduce(const Group<GW>& g, const double (&acc)[8 / GW], int slot) {
  constexpr int NA = 8 / GW;
  double v[NA];
#pragma unroll
  for (int m = 0; m < NA; ++m) v[m] = acc[m];
  int o = 16;
#pragma unroll
  for (int n = NA; n > 1; n >>= 1) {
    const bool hi = (g.lane & o) != 0;
#pragma unroll
    for (int j = 0; j < n / 2; ++j) {
      const double send = hi ? v[j] : v[j + n / 2];
      const double keep = hi ? v[j + n / 2] : v[j];
      const double recv = __shfl_xor_sync(0xffffffffu, send, o);
      v[j] = PROD ? dmul(keep, recv) : dadd(keep, recv);
    }
    o >>= 1;
  }
  double r0 = v[0];
#pragma unroll
  for (int oo = 16 / NA; oo > 0; oo >>= 1) {
    const double w = __shfl_xor_sync(0xffffffffu, r0, oo);
    r0 = PROD ? dmul(r0, w) : dadd(r0, w);
  }
  // the lane whose low bits are zero holds accumulator m = lane / (32/NA), i.e. virtual warp wig + GW*m
  if ((g.lane & (32 / NA - 1)) == 0) g.scratch[slot * 8 + g.wig + GW * (g.lane / (32 / NA))] = r0;
  group_sync(g);
  const double* s = g.scratch + slot * 8;
  double r;
  if (PROD)
    r = dmul(dmul(dmul(s[0], s[1]), dmul(s[2], s[3])), dmul(dmul(s[4], s[5]), dmul(s[6], s[7])));
  else
    r = dadd(dadd(dadd(s[0], s[1]), dadd(s[2], s[3])), dadd(dadd(s[4], s[5]), dadd(s[6], s[7])));
  group_sync(g);  // scratch may be reused by the next reduction
  return r;
}

// ---- cost parameters
struct CostParams {
  // chebyshev (models/_model_helper.py:16-33)
  int cheb_M;             // number of points in [-1, 1]
  const double* cheb_x;   // [M] the reference's accumulated abscissae (x=-1; x+=dx)
  double cheb_tp, cheb_tm;  // polyeval(target, +1.2), polyeval(target, -1.2)
  // DMMA path: Vandermonde fragments vfrag[(tile*KS + s)*32 + lane] = x_i^(n-k), k = 4s + lane%4,
  // i = 8*tile + lane/4 (0 for k >= n or i >= M), ntiles = ceil(M/8)
  const double* cheb_vfrag;
  int cheb_ntiles;
  // griewangk (models/storn.py:135-139): per-dimension {sqrt(i+1), RN(1/sqrt(i+1))}, both
  // correctly rounded on the host, so the kernel needs neither a square root nor a divide
  const double2* grw_tab;
  int grw_bounded;  // 1: the strict range keeps every |x_i| <= kCosFastMax (guard-free factors)
  // data-fit residual costs (forward_model.py:204-260): metric sum(((F(params)(x_i) - obs_i)/sigma)^2)
  // over the cheb_M evaluation points cheb_x[] (the DMMA variant reads cheb_vfrag as well)
  const double* fit_obs;  // [M] observations
  double fit_sigma;
  int sep_kind;  // COST_SEPARABLE: which of the SEP_* functions
  // COST_ZIMMERMANN stands for the family of closed-form functions one thread evaluates (fixed_cost)
  int fit_kind;               // COST_LORENTZIAN instantiations: 0 = lorentzian, 1 = dual exponential decay (FIT_*)
  const double* fit_sigma_v;  // [M] per-point sigma (nullptr: the scalar fit_sigma)
  int fixed_kind;      // FIX_*
  double fixed_c[3];   // branins: b = 5.1/(4*pi**2), c = 5./pi, f = 1./(8*pi), computed on the host like Python does
};

// c / s with r = RN(1/s): two Markstein refinement steps on top of c*r give the correctly
// rounded IEEE quotient (checked against div.rn.f64 by mb2_selftest_division) at 5 instructions
// instead of the ~28 of the generic divide.  Falls back to the IEEE divide for non-finite or
// tiny/huge operands where the residuals could over/underflow.
__device__ __forceinline__ double div_by_table(double c, double s, double r) {
  const double q0 = __dmul_rn(c, r);
  const double q1 = __fma_rn(__fma_rn(-q0, s, c), r, q0);
  const double q2 = __fma_rn(__fma_rn(-q1, s, c), r, q1);
  // |c| near the overflow threshold (residuals overflow -> NaN), infinities and NaN take the IEEE
  // divide; s >= 1, so the quotient itself never overflows for finite c
  if (!(fabs(c) < 1e300)) return c / s;
  return q2;
}

// cos(a) for |a| <= kCosFastMax, built for the griewangk product (models/storn.py:139): reduce by
// multiples of pi with a two-part Cody-Waite step (exact first FMA), then ONE even polynomial on
// [-pi/2, pi/2] (Remez fit of cos(sqrt(u)), degree 9 in u = t^2, truncation 6e-21) and the sign
// from the parity of the multiple.  ABSOLUTE error <= 1.6e-16 (checked against mpmath), which is
// what a factor of the product needs: a factor near zero carries its error into the product
// scaled by the other factors (<= 1).  16 instructions instead of ~50 for the general cos with
// its sin/cos kernel selection; no branches.
constexpr double kCosFastMax = 1.0e5;
__device__ __forceinline__ double cos_bounded(double a) {
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52: the sum's low mantissa bits hold rint(a/pi)
  const double big = __fma_rn(a, 0.31830988618379069, kMagic);
  const int q = __double2loint(big);
  const double j = __dsub_rn(big, kMagic);
  double t = __fma_rn(-j, 3.1415926535897931, a);
  t = __fma_rn(-j, 1.2246467991473532e-16, t);
  const double u = __dmul_rn(t, t);
  double p = -1.51196479263702045e-16;
  p = __fma_rn(p, u, 4.77687040278621769e-14);
  p = __fma_rn(p, u, -1.14706700869031000e-11);
  p = __fma_rn(p, u, 2.08767556647971772e-09);
  p = __fma_rn(p, u, -2.75573192096311045e-07);
  p = __fma_rn(p, u, 2.48015873014924644e-05);
  p = __fma_rn(p, u, -1.38888888888885295e-03);
  p = __fma_rn(p, u, 4.16666666666666574e-02);
  p = __fma_rn(p, u, -5.00000000000000000e-01);
  p = __fma_rn(p, u, 1.0);
  return __hiloint2double(__double2hiint(p) ^ (q << 31), __double2loint(p));
}

// One factor of the griewangk product: cos(c / sqrt(i+1)) with {s, r} = {sqrt(i+1), RN(1/sqrt(i+1))}.
// `bounded`: |c| <= kCosFastMax is guaranteed by the strict range, so the guards are dropped (a
// tiny |c| whose residuals underflow still gives a quotient that is tiny, and cos of it is 1).
__device__ __forceinline__ double grw_factor(double c, double s, double r, bool bounded) {
  if (bounded) {
    const double q0 = __dmul_rn(c, r);
    const double q1 = __fma_rn(__fma_rn(-q0, s, c), r, q0);
    return cos_bounded(__fma_rn(__fma_rn(-q1, s, c), r, q1));
  }
  const double q = div_by_table(c, s, r);
  return (fabs(q) <= kCosFastMax) ? cos_bounded(q) : cos(q);
}

enum : int {
  COST_ROSEN = 0, COST_CORANA = 1, COST_GRIEWANGK = 2, COST_ZIMMERMANN = 3, COST_CHEBYSHEV = 4,
  COST_POLYFIT = 6, COST_LORENTZIAN = 8,  // numbered like mb2_cost (5 and 7 are the DMMA variants)
  COST_SEPARABLE = 9                      // mb2_cost 9..12: one kernel body, cp.sep_kind selects the function
};

// Separable test functions: two running sums over the coordinates and a closing formula.
//   sphere     dejong.py:63-66     f = sum(c*c)
//   rastrigin  pohlheim.py:132     10.*N + sum(c*c - 10.*cos(2*pi*c))
//   schwefel   pohlheim.py:71      sum(-c * sin(sqrt(abs(c))))
//   ackley     pohlheim.py:199-204 -20*exp(-0.2*sqrt(sum(c*c)/N)) + (-exp(sum(cos(2*pi*c))/N) + 20 + exp(1))
// Terms follow the reference's operation order; sin/cos/exp are CUDA's (<= 2 ulp from glibc's).
//   step       dejong.py:184-192   30. + sum(floor(c) | 30*(c - 5.12) | 30*(5.12 - c))
//   michal     pohlheim.py:237-240 -sum(sin(c_i) * sin((i+1)*c_i**2/pi)**20)
//   powers     pohlheim.py:162-163 sum(abs(c_i)**(i+2))
enum : int { SEP_SPHERE = 0, SEP_RASTRIGIN = 1, SEP_SCHWEFEL = 2, SEP_ACKLEY = 3, SEP_STEP = 4, SEP_MICHAL = 5, SEP_POWERS = 6 };
__device__ __forceinline__ void sep_accumulate(int kind, double c, int i, double& s1, double& s2) {
  const double two_pi = 6.283185307179586;  // 2*pi in binary64
  if (kind == SEP_STEP) {
    double t;
    if (fabs(c) <= 5.12) t = floor(c);
    else if (c > 5.12) t = dmul(30.0, dsub(c, 5.12));
    else t = dmul(30.0, dsub(5.12, c));  // also where a NaN lands, like the reference's else branch
    s1 = dadd(s1, t);
  } else if (kind == SEP_MICHAL) {
    const double w = sin(__ddiv_rn(dmul((double)(i + 1), dmul(c, c)), 3.141592653589793));
    const double w2 = dmul(w, w), w4 = dmul(w2, w2), w8 = dmul(w4, w4), w16 = dmul(w8, w8);
    s1 = dadd(s1, dmul(sin(c), dmul(w16, w4)));  // w**20 by squaring: <= 3 ulp from pow()
  } else if (kind == SEP_POWERS) {
    s1 = dadd(s1, pow(fabs(c), (double)(i + 2)));
  } else if (kind == SEP_SPHERE) {
    s1 = dadd(s1, dmul(c, c));
  } else if (kind == SEP_RASTRIGIN) {
    s1 = dadd(s1, dsub(dmul(c, c), dmul(10.0, cos(dmul(two_pi, c)))));
  } else if (kind == SEP_SCHWEFEL) {
    s1 = dadd(s1, dmul(-c, sin(sqrt(fabs(c)))));
  } else {
    s1 = dadd(s1, dmul(c, c));
    s2 = dadd(s2, cos(dmul(two_pi, c)));
  }
}
__device__ __forceinline__ double sep_finish(int kind, double s1, double s2, int n) {
  if (kind == SEP_RASTRIGIN) return dadd(dmul(10.0, (double)n), s1);
  if (kind == SEP_STEP) return dadd(30.0, s1);
  if (kind == SEP_MICHAL) return -s1;
  if (kind == SEP_ACKLEY) {
    const double f0 = dmul(-20.0, exp(dmul(-0.2, sqrt(s1 / (double)n))));
    const double f1 = dadd(dadd(-exp(s2 / (double)n), 20.0), 2.718281828459045);  // math.exp(1)
    return dadd(f0, f1);
  }
  return s1;
}

// Horner exactly as mystic/math/poly.py:37-39: val = 0*x; for c in coeffs: val = c + val*x
__device__ __forceinline__ double horner(const double* __restrict__ c, int n, double x) {
  double val = dmul(0.0, x);
  for (int j = 0; j < n; ++j) val = dadd(c[j], dmul(val, x));
  return val;
}

// One squared residual of a data-fit cost (forward_model.py:241-258: x = F(params)(evalpts);
// x = x - observations, or (x - observations)/sigma with a scalar or per-point sigma; metric sum(x*x)),
// elementwise in numpy's order.
// polynomial model (models/poly.py:36-54): numpy.poly1d(coeffs)(x) is numpy.polyval's
// `y = zeros_like(x); for pv in p: y = y*x + pv` (leading zero coefficients change nothing)
__device__ __forceinline__ double polyfit_value(const double* __restrict__ c, int n, double x) {
  double y = 0.0;
  for (int j = 0; j < n; ++j) y = dadd(dmul(y, x), c[j]);
  return y;
}
// lorentzian peak (models/lorentzian.py:32-37), coeffs = (a1,a2,a3,A0,E0,G0,n):
//   (a1 + a2*x + a3*x*x + A0*(G0/(2*pi))/((x-E0)*(x-E0) + (G0/2)*(G0/2)))/n
__device__ __forceinline__ double lorentz_value(const double* __restrict__ p, double x) {
  const double a1 = p[0], a2 = p[1], a3 = p[2], A0 = p[3], E0 = p[4], G0 = p[5], nn = p[6];
  const double k = dmul(A0, __ddiv_rn(G0, 6.283185307179586));  // 2*pi in binary64
  const double hg = __ddiv_rn(G0, 2.0);
  const double d = dsub(x, E0);
  const double den = dadd(dmul(d, d), dmul(hg, hg));
  const double s = dadd(dadd(a1, dmul(a2, x)), dmul(dmul(a3, x), x));
  return __ddiv_rn(dadd(s, __ddiv_rn(k, den)), nn);
}
// dual exponential decay (models/br8.py:27-32), coeffs = (a1,a2,a3,a4,a5): a1 + a2*exp(-t/a4) + a3*exp(-t/a5)
// (CUDA's exp is within 1 ulp of numpy's)
enum : int { FIT_LORENTZIAN = 0, FIT_DECAY = 1 };
__device__ __forceinline__ double decay_value(const double* __restrict__ p, double t) {
  const double e4 = exp(__ddiv_rn(-t, p[3])), e5 = exp(__ddiv_rn(-t, p[4]));
  return dadd(dadd(p[0], dmul(p[1], e4)), dmul(p[2], e5));
}
template <int COST>
__device__ __forceinline__ double fit_term(const double* __restrict__ xs, int D, const CostParams& cp, int i) {
  const double x = __ldg(cp.cheb_x + i), obs = __ldg(cp.fit_obs + i);
  const double v = (COST == COST_POLYFIT) ? polyfit_value(xs, D, x) : (cp.fit_kind == FIT_DECAY ? decay_value(xs, x) : lorentz_value(xs, x));
  double r = dsub(v, obs);
  const double sigma = (cp.fit_sigma_v != nullptr) ? __ldg(cp.fit_sigma_v + i) : cp.fit_sigma;
  if (sigma != 1.0) r = __ddiv_rn(r, sigma);  // AbstractModel's default sigma = 1.0 divides exactly
  return dmul(r, r);
}

// Closed-form test functions evaluated by ONE thread in the reference's operation order (Python floats:
// `**` is libm pow, here x*x for squares, CUDA's pow / sin / cos / exp / log otherwise: <= 2 ulp from glibc's).
//   zimmermann storn.py:182-191      branins   pohlheim.py:278-282   easom    pohlheim.py:312-313
//   goldstein  pohlheim.py:352-355   peaks     nag.py:55-59          venkat91 venkataraman.py:50-52
//   fosc3d     wolfram.py:60-64      nmin51    wolfram.py:101-103    shekel   dejong.py:266-278
//   ellipsoid  pohlheim.py:101-102 (any D)     paviani   schittkowski.py:64-71 (any D)
enum : int { FIX_ZIMMERMANN = 0, FIX_BRANINS = 1, FIX_EASOM = 2, FIX_GOLDSTEIN = 3, FIX_PEAKS = 4, FIX_VENKAT91 = 5,
             FIX_FOSC3D = 6, FIX_NMIN51 = 7, FIX_SHEKEL = 8, FIX_ELLIPSOID = 9, FIX_PAVIANI = 10 };

static __device__ __noinline__ double fixed_cost(int kind, const double* x, int D, const CostParams& cp) {
  const double kPi = 3.141592653589793;
  if (kind == FIX_ELLIPSOID) {
    double pre = 0.0, acc = 0.0;  // sum(sum(coeffs[0:i+1])**2 for i in range(N))
    for (int i = 0; i < D; ++i) {
      pre = dadd(pre, x[i]);
      acc = dadd(acc, dmul(pre, pre));
    }
    return acc;
  }
  if (kind == FIX_PAVIANI) {
    double s = 0.0, p = x[0];
    for (int i = 0; i < D; ++i) {
      const double v = x[i];
      if (v >= 10.0 || v <= 2.0) {
        s = CUDART_INF;
        break;
      }
      const double a = log(dsub(v, 2.0)), b = log(dsub(10.0, v));
      s = dadd(s, dadd(dmul(a, a), dmul(b, b)));
    }
    for (int i = 1; i < D; ++i) p = dmul(p, x[i]);
    return dsub(s, pow(p, 0.2));
  }
  const double x0 = x[0], x1 = (D > 1) ? x[1] : 0.0;
  switch (kind) {
    case FIX_BRANINS: {
      const double b = cp.fixed_c[0], c = cp.fixed_c[1], f = cp.fixed_c[2];
      const double t = dsub(dadd(dsub(x1, dmul(b, dmul(x0, x0))), dmul(c, x0)), 6.0);
      const double f0 = dmul(1.0, dmul(t, t));
      const double f1 = dadd(dmul(dmul(10.0, dsub(1.0, f)), cos(x0)), 10.0);
      return dadd(f0, f1);
    }
    case FIX_EASOM: {
      const double a = dsub(x0, kPi), b = dsub(x1, kPi);
      return dmul(dmul(-cos(x0), cos(x1)), exp(-dadd(dmul(a, a), dmul(b, b))));
    }
    case FIX_GOLDSTEIN: {
      const double s0 = dmul(x0, x0), s1 = dmul(x1, x1);
      double f0 = dsub(19.0, dmul(14.0, x0));
      f0 = dadd(f0, dmul(3.0, s0));
      f0 = dsub(f0, dmul(14.0, x1));
      f0 = dadd(f0, dmul(dmul(6.0, x0), x1));
      f0 = dadd(f0, dmul(3.0, s1));
      double f1 = dsub(18.0, dmul(32.0, x0));
      f1 = dadd(f1, dmul(12.0, s0));
      f1 = dadd(f1, dmul(48.0, x1));
      f1 = dsub(f1, dmul(dmul(36.0, x0), x1));
      f1 = dadd(f1, dmul(27.0, s1));
      const double u = dadd(dadd(x0, x1), 1.0), v = dsub(dmul(2.0, x0), dmul(3.0, x1));
      return dmul(dadd(1.0, dmul(dmul(u, u), f0)), dadd(30.0, dmul(dmul(v, v), f1)));
    }
    case FIX_PEAKS: {
      const double xx = dmul(x0, x0), yy = dmul(x1, x1);
      const double y1 = dadd(x1, 1.0), xm = dsub(1.0, x0), xp = dadd(x0, 1.0);
      const double t1 = dmul(dmul(3.0, dmul(xm, xm)), exp(dsub(-xx, dmul(y1, y1))));
      const double in2 = dsub(dsub(dmul(x0, 1. / 5.), pow(x0, 3.0)), pow(x1, 5.0));
      const double t2 = dmul(dmul(10.0, in2), exp(dsub(-xx, yy)));
      const double t3 = dmul(1. / 3., exp(dsub(-dmul(xp, xp), yy)));
      return dsub(dsub(t1, t2), t3);
    }
    case FIX_VENKAT91: {
      const double a = dsub(x0, 4.0), b = dsub(x1, 4.0);
      const double R = sqrt(dadd(dadd(dmul(a, a), dmul(b, b)), 0.1));
      return dmul(-20.0, sin(R)) / R;
    }
    case FIX_FOSC3D: {
      const double func = dadd(dmul(-4.0, exp(dsub(dmul(-x0, x0), dmul(x1, x1)))), dmul(sin(dmul(6.0, x0)), sin(dmul(5.0, x1))));
      const double pen = (x1 < 0.0) ? dmul(dmul(100.0, x1), x1) : 0.0;
      return dadd(func, pen);
    }
    case FIX_NMIN51: {
      double r = dadd(exp(sin(dmul(50.0, x0))), sin(dmul(60.0, exp(x1))));
      r = dadd(r, sin(dmul(70.0, sin(x0))));
      r = dadd(r, sin(sin(dmul(80.0, x1))));
      r = dsub(r, sin(dmul(10.0, dadd(x0, x1))));
      return dadd(r, dmul(1. / 4., dadd(dmul(x0, x0), dmul(x1, x1))));
    }
    case FIX_SHEKEL: {
      double r = 0.0;
      for (int i = 0; i < 25; ++i) {
        const double a1 = -32.0 + 16.0 * (i % 5), a2 = -32.0 + 16.0 * (i / 5);
        const double z = dadd(dadd(dmul(1.0, (double)i), pow(dsub(x0, a1), 6.0)), pow(dsub(x1, a2), 6.0));
        r = dadd(r, (z != 0.0) ? 1.0 / z : CUDART_INF);
      }
      return 1.0 / dadd(0.002, r);
    }
    default: {  // FIX_ZIMMERMANN (D == 2)
      const double f8 = dsub(dsub(9.0, x0), x1);
      double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
      if (x0 < 0.0) c0 = dmul(-100.0, x0);
      if (x1 < 0.0) c1 = dmul(-100.0, x1);
      const double a = dsub(x0, 3.0), b = dsub(x1, 2.0);
      const double xx = dadd(dmul(a, a), dmul(b, b));
      if (xx > 16.0) c2 = dmul(100.0, dsub(xx, 16.0));
      const double pr = dmul(x0, x1);
      if (pr > 14.0) c3 = dmul(100.0, dsub(pr, 14.0));
      double m = f8;  // python max(f8,c0,c1,c2,c3): replace only when strictly greater
      if (c0 > m) m = c0;
      if (c1 > m) m = c1;
      if (c2 > m) m = c2;
      if (c3 > m) m = c3;
      return m;
    }
  }
}

// Group-cooperative cost of one row `xs` (shared memory, D entries).  Every thread of the group
// returns the same value.  Per-element terms follow the reference's operation order without
// FMA, so each term is bit-identical to the reference; only the summation tree differs
// (<= 1e-12 relative).
template <int COST, int GW>
__device__ __forceinline__ double group_cost(const Group<GW>& g, const double* xs, int D, const CostParams& cp) {
  constexpr int NA = 8 / GW;       // accumulators per thread
  constexpr int NT = 32 * GW;      // threads per group
  if (COST == COST_ROSEN) {
    // dejong.py:98-101: sum(100*(x[1:]-x[:-1]**2)**2 + (1-x[:-1])**2)
    double acc[NA];
#pragma unroll
    for (int m = 0; m < NA; ++m) acc[m] = 0.0;
    const int n = D - 1;
    for (int i0 = g.tig; i0 < n; i0 += 256) {
#pragma unroll
      for (int m = 0; m < NA; ++m) {
        const int i = i0 + m * NT;
        if (i < n) {
          const double a = xs[i], b = xs[i + 1];
          const double t = dsub(b, dmul(a, a));
          const double u = dsub(1.0, a);
          acc[m] = dadd(acc[m], dadd(dmul(100.0, dmul(t, t)), dmul(u, u)));
        }
      }
    }
    return group_reduce<false, GW>(g, acc, 0);
  } else if (COST == COST_CORANA) {
    // storn.py:77-96, strictly sequential over 4 terms; threads compute redundantly
    const double d[4] = {1., 1000., 10., 100.};
    double x[4] = {0., 0., 0., 0.};
    if (D >= 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = xs[j];
    } else {
      // x[_d.index(i)] = _x[i] with _d = [0,3,1,2]: 1 -> x0; 2 -> x0,x2; 3 -> x0,x2,x3
      if (D >= 1) x[0] = xs[0];
      if (D >= 2) x[2] = xs[1];
      if (D >= 3) x[3] = xs[2];
    }
    double r = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double xj = x[j];
      const double sg = (xj > 0.0) ? 1.0 : ((xj < 0.0) ? -1.0 : 0.0);  // numpy.sign
      const double zj = dmul(dmul(floor(dadd(fabs(xj / 0.2), 0.49999)), sg), 0.2);
      double term;
      if (fabs(dsub(xj, zj)) < 0.05) {
        const double sz = (zj > 0.0) ? 1.0 : ((zj < 0.0) ? -1.0 : 0.0);
        const double w = dsub(zj, dmul(0.05, sz));
        term = dmul(dmul(0.15, dmul(w, w)), d[j]);  // 0.15 * pow(w,2) * d[j]
      } else {
        term = dmul(dmul(d[j], xj), xj);  // d[j]*x[j]*x[j]
      }
      r = dadd(r, term);
    }
    return r;
  } else if (COST == COST_GRIEWANGK) {
    // storn.py:135-139: sum(c*c)/4000 - prod(cos(c_i/sqrt(i+1))) + 1
    double sa[NA], pa[NA];
#pragma unroll
    for (int m = 0; m < NA; ++m) {
      sa[m] = 0.0;
      pa[m] = 1.0;
    }
    if ((D & 255) == 0 && cp.grw_bounded) {  // whole blocks of 256 virtual threads, bounded arguments: no guards
      for (int i0 = g.tig; i0 < D; i0 += 256) {
#pragma unroll
        for (int m = 0; m < NA; ++m) {
          const int i = i0 + m * NT;
          const double c = xs[i];
          sa[m] = dadd(sa[m], dmul(c, c));
          const double2 t = __ldg(cp.grw_tab + i);  // {sqrt(i+1), 1/sqrt(i+1)}
          pa[m] = dmul(pa[m], grw_factor(c, t.x, t.y, true));
        }
      }
    } else {
      for (int i0 = g.tig; i0 < D; i0 += 256) {
#pragma unroll
        for (int m = 0; m < NA; ++m) {
          const int i = i0 + m * NT;
          if (i < D) {
            const double c = xs[i];
            sa[m] = dadd(sa[m], dmul(c, c));
            const double2 t = __ldg(cp.grw_tab + i);
            pa[m] = dmul(pa[m], grw_factor(c, t.x, t.y, cp.grw_bounded != 0));
          }
        }
      }
    }
    const double s = group_reduce<false, GW>(g, sa, 0);
    const double p = group_reduce<true, GW>(g, pa, 1);
    return dadd(dsub(s / 4000.0, p), 1.0);
  } else if (COST == COST_ZIMMERMANN) {
    return fixed_cost(cp.fixed_kind, xs, D, cp);  // every thread evaluates the row
  } else if (COST == COST_SEPARABLE) {
    double sa[NA], sb[NA];
#pragma unroll
    for (int m = 0; m < NA; ++m) {
      sa[m] = 0.0;
      sb[m] = 0.0;
    }
    const int kind = cp.sep_kind;
    for (int i0 = g.tig; i0 < D; i0 += 256) {
#pragma unroll
      for (int m = 0; m < NA; ++m) {
        const int i = i0 + m * NT;
        if (i < D) sep_accumulate(kind, xs[i], i, sa[m], sb[m]);
      }
    }
    const double s1 = group_reduce<false, GW>(g, sa, 0);
    const double s2 = (kind == SEP_ACKLEY) ? group_reduce<false, GW>(g, sb, 1) : 0.0;
    return sep_finish(kind, s1, s2, D);
  } else if (COST == COST_POLYFIT || COST == COST_LORENTZIAN) {
    double acc[NA];
#pragma unroll
    for (int m = 0; m < NA; ++m) acc[m] = 0.0;
    const int M = cp.cheb_M;
    for (int i0 = g.tig; i0 < M; i0 += 256) {
#pragma unroll
      for (int m = 0; m < NA; ++m) {
        const int i = i0 + m * NT;
        if (i < M) acc[m] = dadd(acc[m], fit_term<COST>(xs, D, cp, i));
      }
    }
    return group_reduce<false, GW>(g, acc, 0);
  } else {  // COST_CHEBYSHEV
    // _model_helper.py:16-33
    double acc[NA];
#pragma unroll
    for (int m = 0; m < NA; ++m) acc[m] = 0.0;
    const int M = cp.cheb_M;
    for (int i0 = g.tig; i0 < M; i0 += 256) {
#pragma unroll
      for (int m = 0; m < NA; ++m) {
        const int i = i0 + m * NT;
        if (i < M) {
          const double px = horner(xs, D, __ldg(cp.cheb_x + i));
          if (px < -1.0 || px > 1.0) {
            const double u = dsub(1.0, px);
            acc[m] = dadd(acc[m], dmul(u, u));
          }
        }
      }
    }
    double r = group_reduce<false, GW>(g, acc, 0);
    double px = dsub(horner(xs, D, 1.2), cp.cheb_tp);
    if (px < 0.0) r = dadd(r, dmul(px, px));
    px = dsub(horner(xs, D, -1.2), cp.cheb_tm);
    if (px < 0.0) r = dadd(r, dmul(px, px));
    return r;
  }
}

// ---- penalty: stacked mystic.penalty decorators with linear conditions G.x - h
struct PenaltyParams {
  int n;                // number of terms (0 = no penalty: `lambda x: 0.0`, abstract_solver.py:140)
  const int* ptype;     // [n]
  const double* G;      // [n, D]
  const double* h;      // [n]
  const double* kscale; // [n]  k*h**n
  const double* mult;   // [n]  lagrange kinds: the multiplier accumulated from the stored iterations (beta / lambda)
};

// one penalty term of the stack for the condition value pf = G.x - h (mystic/penalty.py; k = k*h**n):
//   0 quadratic_equality :88    k*pf**2            1 linear_equality :147    k*|pf|
//   2 quadratic_inequality :388 2k*max(0,pf)**2    3 linear_inequality :447  2k*|max(0,pf)|
//   4 uniform_equality :208     k if pf else 0     5 uniform_inequality :266 k if pf > 0 else 0
//   6 barrier_inequality :326   inf if pf > 0 else -.5/k*log(-pf)   (numpy.log: log(0) = -inf, so pf == 0 gives +inf)
//   7 lagrange_inequality :517  k*mpf**2 + beta*mpf, mpf = max(-beta/(2k), pf)     8 lagrange_equality :585  k*pf**2 + lam*pf
// Python's max(a, b) keeps a unless b > a.
__device__ __forceinline__ double penalty_term(int ty, double k, double mult, double pf) {
  const double mpos = (pf > 0.0) ? pf : 0.0;
  switch (ty) {
    case 0: return dmul(k, dmul(pf, pf));
    case 1: return dmul(k, fabs(pf));
    case 2: return dmul(2.0 * k, dmul(mpos, mpos));
    case 3: return dmul(2.0 * k, fabs(mpos));
    case 4: return (pf != 0.0) ? k : 0.0;
    case 5: return (pf > 0.0) ? k : 0.0;
    case 6: return (pf > 0.0) ? CUDART_INF : dmul(__ddiv_rn(-0.5, k), log(-pf));
    case 7: {
      const double floor_ = __ddiv_rn(-mult, dmul(2.0, k));
      const double mpf = (pf > floor_) ? pf : floor_;
      return dadd(dmul(k, dmul(mpf, mpf)), dmul(mult, mpf));
    }
    default: return dadd(dmul(k, dmul(pf, pf)), dmul(mult, pf));
  }
}

// term 0 innermost; value = t_{n-1} + (... + (t_0 + 0.0))  (penalty.py:88,147,388,447)
template <int GW>
__device__ __forceinline__ double group_penalty(const Group<GW>& g, const double* xs, int D, const PenaltyParams& pp) {
  constexpr int NA = 8 / GW;
  constexpr int NT = 32 * GW;
  double total = 0.0;
  for (int t = 0; t < pp.n; ++t) {
    const double* gr = pp.G + (int64_t)t * D;
    double acc[NA];
#pragma unroll
    for (int m = 0; m < NA; ++m) acc[m] = 0.0;
    for (int i0 = g.tig; i0 < D; i0 += 256) {
#pragma unroll
      for (int m = 0; m < NA; ++m) {
        const int i = i0 + m * NT;
        if (i < D) acc[m] = dadd(acc[m], dmul(__ldg(gr + i), xs[i]));
      }
    }
    const double pf = dsub(group_reduce<false, GW>(g, acc, 0), __ldg(pp.h + t));
    const double v = penalty_term(__ldg(pp.ptype + t), __ldg(pp.kscale + t), __ldg(pp.mult + t), pf);
    total = dadd(v, total);
  }
  return total;
}

}  // namespace mb2
