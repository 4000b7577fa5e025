/*
 * de_oracle.c -- plain C restatement of mystic's DifferentialEvolutionSolver2 generation step.
 *
 * TEST INFRASTRUCTURE ONLY (checker and CPU baseline; see oracle/README.md).  Nothing in the
 * product package mystic_b200/ links or loads this file.  It follows oracle/de_oracle.py
 * function by function and is checked bit-for-bit against it and against the recordings of the
 * unmodified reference (tests/test_oracle_c.py), so: PARITY PINNED.
 *
 * Citations are file:line in the mystic repository.  OpenMP parallelises over candidates; the
 * arithmetic of every candidate is sequential and compiled with -ffp-contract=off, so results
 * do not depend on the thread count.
 *
 * build: make -C oracle   (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { BASE_BEST1 = 0, BASE_RAND1 = 1, BASE_RANDTOBEST1 = 2, BASE_BEST2 = 3, BASE_RAND2 = 4 };
enum { XOVER_EXP = 0, XOVER_BIN = 1 };
enum { BOUNDS_NONE = 0, BOUNDS_REJECT = 1, BOUNDS_CLAMP = 2 };
enum { COST_ROSEN = 0, COST_CORANA = 1, COST_GRIEWANGK = 2, COST_ZIMMERMANN = 3, COST_CHEBYSHEV = 4,
       COST_POLYFIT = 6, COST_LORENTZIAN = 8, COST_SPHERE = 9, COST_RASTRIGIN = 10, COST_SCHWEFEL = 11,
       COST_ACKLEY = 12, COST_STEP = 13, COST_MICHAL = 14, COST_POWERS = 15,
       COST_BRANINS = 16, COST_EASOM = 17, COST_GOLDSTEIN = 18, COST_PEAKS = 19, COST_VENKAT91 = 20, COST_FOSC3D = 21,
       COST_NMIN51 = 22, COST_SHEKEL = 23, COST_ELLIPSOID = 24, COST_PAVIANI = 25, COST_DECAY = 26 };

typedef struct {
  int32_t dim;
  int32_t strategy; /* ids of mystic/strategy.py in file order, 0..9 */
  int64_t np;
  double F, CR;
  int32_t bounds_mode;
  int32_t cost;
  const double* lo;
  const double* hi;
  /* chebyshev (_model_helper.py:16-33) */
  int32_t cheb_M;
  int32_t cheb_order;
  const double* cheb_target;
  /* stacked linear penalties, term 0 innermost (penalty.py) */
  int32_t n_pen;
  int32_t pad_;
  const int32_t* pen_type;
  const double* pen_G;
  const double* pen_h;
  const double* pen_k; /* k*h**n */
  /* data-fit residual costs (forward_model.py:204-260): M points, observations, sigma */
  int32_t fit_M;
  int32_t pad2_;
  const double* fit_pts;
  const double* fit_obs;
  double fit_sigma;
  const double* fit_sigma_v; /* per-point sigma (models/br8.py:47), NULL: the scalar */
  const double* pen_mult;    /* lagrange penalties: accumulated multiplier per term (beta / lam), NULL: zeros */
} orc_config;

/* strategy id -> base, crossover, donors (strategy.py:34-311; only Best1Bin is binomial) */
static void strategy_info(int s, int* base, int* xover, int* k) {
  static const int B[10] = {BASE_BEST1, BASE_BEST1, BASE_RAND1, BASE_RANDTOBEST1, BASE_BEST2,
                            BASE_RAND2, BASE_RAND1, BASE_RANDTOBEST1, BASE_BEST2, BASE_RAND2};
  static const int K[5] = {2, 3, 2, 4, 5};
  *base = B[s];
  *xover = (s == 1) ? XOVER_BIN : XOVER_EXP;
  *k = K[*base];
}

/* ---------------------------------------------------------------- Philox4x32-10 */
static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static uint64_t mulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }

static void draw_donors(uint64_t seed, uint32_t gen, uint32_t cand_global, int64_t cand, int64_t np, int dim, int k,
                        int64_t* donors, int* nstart, uint32_t* xword) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint64_t r[6];
  for (int b = 0; b < (k + 2) / 2; ++b) {
    uint32_t v[4];
    philox((uint32_t)b, cand_global, gen, 0u, k0, k1, v);
    r[2 * b] = ((uint64_t)v[0] << 32) | v[1];
    r[2 * b + 1] = ((uint64_t)v[2] << 32) | v[3];
  }
  int64_t ex[6];
  int nex = 1;
  ex[0] = cand;
  for (int i = 0; i < k; ++i) {
    int64_t idx = (int64_t)mulhi64(r[i], (uint64_t)(np - 1 - i));
    for (int e = 0; e < nex; ++e)
      if (idx >= ex[e]) ++idx;
    donors[i] = idx;
    int pos = nex;
    while (pos > 0 && ex[pos - 1] > idx) {
      ex[pos] = ex[pos - 1];
      --pos;
    }
    ex[pos] = idx;
    ++nex;
  }
  *nstart = (int)mulhi64(r[k], (uint64_t)dim);
  *xword = (uint32_t)r[k]; /* low half: the geometric draw of the exponential crossover length (protocol 2) */
}

static uint64_t crossover_threshold(double cr) {
  if (!(cr > 0.0)) return 0;
  if (cr >= 1.0) return 4294967296ull;
  return (uint64_t)floor(cr * 4294967296.0);
}

/* binomial crossover, protocol 2 (mystic_b200/csrc/philox.cuh): Bernoulli(CR) on 16-bit fields */
static uint32_t crossover_threshold16(double cr) {
  if (!(cr > 0.0)) return 0;
  if (cr >= 1.0) return 65536u;
  return (uint32_t)floor(cr * 65536.0);
}

/* exponential crossover, protocol 2: T[1..nmax] with T[0] = 2^32, T[l] = (T[l-1] * thr) >> 32;
 * tab[l-1] = T[l]; returns nmax, or -1 when every draw succeeds (CR >= 1) */
static int build_exp_length_table(uint64_t thr, int dim, uint32_t* tab) {
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

/* L = #{l in 1..nmax : w < T[l]} */
static int exp_length(uint32_t w, const uint32_t* tab, int nmax, int dim) {
  if (nmax < 0) return dim;
  int lo = 0, hi = nmax; /* invariant: w < T[l] for l <= lo, w >= T[l] for l > hi */
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (w < tab[mid - 1]) lo = mid; else hi = mid - 1;
  }
  return lo;
}

/* ---------------------------------------------------------------- costs */
/* numpy.sum of a contiguous double vector: pairwise, 8-lane blocks of 128
 * (numpy/core/src/umath/loops_utils.h.src DOUBLE_pairwise_sum) */
static double np_pairwise_sum(const double* a, int64_t n) {
  if (n < 8) {
    double res = 0.;
    for (int64_t i = 0; i < n; ++i) res += a[i];
    return res;
  } else if (n <= 128) {
    double r[8];
    int64_t i;
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  } else {
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
  }
}

static double cost_rosen(const double* x, int D, double* scratch) {
  /* dejong.py:98-101 */
  if (D < 2) return 0.0;
  for (int i = 0; i < D - 1; ++i) {
    double t = x[i + 1] - x[i] * x[i];
    double u = 1 - x[i];
    scratch[i] = 100.0 * (t * t) + u * u;
  }
  return np_pairwise_sum(scratch, D - 1);
}

/* forward_model.py:241-258 with models/poly.py:36-54 (numpy.polyval order) or
 * models/lorentzian.py:32-37; x = (F - obs)/sigma; numpy.sum(x*x).  scratch: fit_M doubles */
static double cost_datafit(const double* c, int D, const orc_config* cfg, double* scratch) {
  const int M = cfg->fit_M;
  for (int i = 0; i < M; ++i) {
    const double x = cfg->fit_pts[i];
    double y;
    if (cfg->cost == COST_POLYFIT) {
      y = 0.0;
      for (int j = 0; j < D; ++j) y = y * x + c[j];
    } else if (cfg->cost == COST_DECAY) { /* models/br8.py:27-32 */
      y = c[0] + c[1] * exp(-x / c[3]) + c[2] * exp(-x / c[4]);
    } else {
      const double a1 = c[0], a2 = c[1], a3 = c[2], A0 = c[3], E0 = c[4], G0 = c[5], n = c[6];
      const double pi = 3.141592653589793;
      y = (a1 + a2 * x + a3 * x * x + A0 * (G0 / (2 * pi)) / ((x - E0) * (x - E0) + (G0 / 2) * (G0 / 2))) / n;
    }
    const double d = (y - cfg->fit_obs[i]) / (cfg->fit_sigma_v ? cfg->fit_sigma_v[i] : cfg->fit_sigma);
    scratch[i] = d * d;
  }
  return np_pairwise_sum(scratch, M);
}

static double sign_(double v) { return (v > 0.0) ? 1.0 : ((v < 0.0) ? -1.0 : 0.0); }

static double cost_corana(const double* c, int D) {
  /* storn.py:77-96 */
  static const double d[4] = {1., 1000., 10., 100.};
  double x[4] = {0., 0., 0., 0.};
  if (D >= 4) {
    for (int j = 0; j < 4; ++j) x[j] = c[j];
  } else {
    if (D >= 1) x[0] = c[0];
    if (D >= 2) x[2] = c[1];
    if (D >= 3) x[3] = c[2];
  }
  double r = 0.0;
  for (int j = 0; j < 4; ++j) {
    double zj = floor(fabs(x[j] / 0.2) + 0.49999) * sign_(x[j]) * 0.2;
    if (fabs(x[j] - zj) < 0.05)
      r += 0.15 * pow(zj - 0.05 * sign_(zj), 2) * d[j];
    else
      r += d[j] * x[j] * x[j];
  }
  return r;
}

static double cost_griewangk(const double* c, int D) {
  /* storn.py:135-139 (plain left-to-right sum; see oracle/de_oracle.py:cost_griewangk) */
  double s = 0.0, p = 1.0;
  for (int i = 0; i < D; ++i) s += c[i] * c[i];
  for (int i = 0; i < D; ++i) p = p * cos(c[i] / sqrt(i + 1.0));
  return s / 4000. - p + 1;
}

/* dejong.py:63-66, pohlheim.py:71,132,199-204 (plain left-to-right sums) */
static double cost_separable(const double* c, int D, int cost) {
  const double pi = 3.141592653589793;
  double s1 = (cost == COST_STEP) ? 30. : 0.0, s2 = 0.0;
  for (int i = 0; i < D; ++i) {
    switch (cost) {
      case COST_SPHERE: s1 += c[i] * c[i]; break;
      case COST_RASTRIGIN: s1 += c[i] * c[i] - 10. * cos(2 * pi * c[i]); break;
      case COST_SCHWEFEL: s1 += -c[i] * sin(sqrt(fabs(c[i]))); break;
      case COST_STEP:
        if (fabs(c[i]) <= 5.12) s1 += floor(c[i]);
        else if (c[i] > 5.12) s1 += 30 * (c[i] - 5.12);
        else s1 += 30 * (5.12 - c[i]);
        break;
      case COST_MICHAL: s1 += sin(c[i]) * pow(sin((i + 1) * pow(c[i], 2) / pi), 20); break;
      case COST_POWERS: s1 += pow(fabs(c[i]), i + 2); break;
      default: s1 += c[i] * c[i]; s2 += cos(2 * pi * c[i]); break;
    }
  }
  if (cost == COST_RASTRIGIN) return 10. * D + s1;
  if (cost == COST_MICHAL) return -s1;
  if (cost == COST_ACKLEY) {
    const double f0 = -20. * exp(-0.2 * sqrt(s1 / D));
    const double f1 = -exp(s2 / D) + 20. + exp(1.0);
    return f0 + f1;
  }
  return s1;
}

/* closed-form functions of a few variables: the reference's expressions on doubles (`**` is libm pow):
 * pohlheim.py:278-282, 312-313, 352-355, 101-102; nag.py:55-59; venkataraman.py:50-52; wolfram.py:60-64,
 * 101-103; dejong.py:266-278; schittkowski.py:64-71.  ellipsoid: plain sums (the reference's builtin sum()
 * is compensated since CPython 3.12, see oracle/de_oracle.py) */
static double cost_fixed(const double* c, int D, int cost) {
  const double pi = 3.141592653589793;
  if (cost == COST_ELLIPSOID) {
    double pre = 0.0, acc = 0.0;
    for (int i = 0; i < D; ++i) {
      pre += c[i];
      acc += pow(pre, 2);
    }
    return acc;
  }
  if (cost == COST_PAVIANI) {
    double s = 0.0, p = c[0];
    for (int i = 0; i < D; ++i) {
      if (c[i] >= 10. || c[i] <= 2.) {
        s = INFINITY;
        break;
      }
      s += pow(log(c[i] - 2.), 2) + pow(log(10. - c[i]), 2);
    }
    for (int i = 1; i < D; ++i) p = p * c[i];
    return s - pow(p, .2);
  }
  const double x = c[0], y = c[1];
  switch (cost) {
    case COST_BRANINS: {
      const double a = 1., b = 5.1 / (4 * pow(pi, 2)), cc = 5. / pi, d = 6., e = 10., f = 1. / (8 * pi);
      const double f0 = a * pow(y - b * pow(x, 2) + cc * x - d, 2);
      const double f1 = e * (1 - f) * cos(x) + e;
      return f0 + f1;
    }
    case COST_EASOM: return -cos(x) * cos(y) * exp(-(pow(x - pi, 2) + pow(y - pi, 2)));
    case COST_GOLDSTEIN: {
      const double f0 = 19 - 14 * x + 3 * pow(x, 2) - 14 * y + 6 * x * y + 3 * pow(y, 2);
      const double f1 = 18 - 32 * x + 12 * pow(x, 2) + 48 * y - 36 * x * y + 27 * pow(y, 2);
      return (1 + pow(x + y + 1, 2) * f0) * (30 + pow(2 * x - 3 * y, 2) * f1);
    }
    case COST_PEAKS:
      return 3. * pow(1. - x, 2) * exp(-pow(x, 2) - pow(y + 1., 2)) -
             10. * (x * (1. / 5.) - pow(x, 3) - pow(y, 5)) * exp(-pow(x, 2) - pow(y, 2)) -
             1. / 3. * exp(-pow(x + 1., 2) - pow(y, 2));
    case COST_VENKAT91: {
      const double R = pow(pow(x - 4., 2) + pow(y - 4., 2) + 0.1, .5);
      return -20. * sin(R) / R;
    }
    case COST_FOSC3D: {
      const double func = -4. * exp(-x * x - y * y) + sin(6. * x) * sin(5. * y);
      return (y < 0) ? func + 100. * y * y : func + 0;
    }
    case COST_NMIN51:
      return exp(sin(50. * x)) + sin(60. * exp(y)) + sin(70. * sin(x)) + sin(sin(80. * y)) - sin(10. * (x + y)) +
             1. / 4. * (pow(x, 2) + pow(y, 2));
    default: {  /* COST_SHEKEL */
      const double A[5] = {-32., -16., 0., 16., 32.};
      double r = 0.0;
      for (int i = 0; i < 25; ++i) {
        const double z = 1.0 * i + pow(x - A[i % 5], 6) + pow(y - A[i / 5], 6);
        r += (z != 0) ? 1.0 / z : INFINITY;
      }
      return 1.0 / (0.002 + r);
    }
  }
}

static double cost_zimmermann(const double* c) {
  /* storn.py:182-191 */
  double x0 = c[0], x1 = c[1];
  double f8 = 9 - x0 - x1;
  double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  if (x0 < 0) c0 = -100 * x0;
  if (x1 < 0) c1 = -100 * x1;
  double xx = (x0 - 3.) * (x0 - 3) + (x1 - 2.) * (x1 - 2);
  if (xx > 16) c2 = 100 * (xx - 16);
  if (x0 * x1 > 14) c3 = 100 * (x0 * x1 - 14.);
  double m = f8;
  if (c0 > m) m = c0;
  if (c1 > m) m = c1;
  if (c2 > m) m = c2;
  if (c3 > m) m = c3;
  return m;
}

static double polyeval(const double* c, int n, double x) {
  /* math/poly.py:37-39 */
  double val = 0 * x;
  for (int j = 0; j < n; ++j) val = c[j] + val * x;
  return val;
}

static double cost_chebyshev(const double* c, int D, const orc_config* cfg) {
  /* _model_helper.py:16-33 */
  double result = 0.0, x = -1.0;
  const int M = cfg->cheb_M;
  const double dx = 2.0 / (M - 1);
  for (int i = 0; i < M; ++i) {
    double px = polyeval(c, D, x);
    if (px < -1 || px > 1) result += (1 - px) * (1 - px);
    x += dx;
  }
  double px = polyeval(c, D, 1.2) - polyeval(cfg->cheb_target, cfg->cheb_order + 1, 1.2);
  if (px < 0) result += px * px;
  px = polyeval(c, D, -1.2) - polyeval(cfg->cheb_target, cfg->cheb_order + 1, -1.2);
  if (px < 0) result += px * px;
  return result;
}

static double penalty_value(const double* x, int D, const orc_config* cfg) {
  double total = 0.0;
  for (int t = 0; t < cfg->n_pen; ++t) {
    const double* g = cfg->pen_G + (int64_t)t * D;
    double acc = 0.0;
    for (int j = 0; j < D; ++j) acc = acc + g[j] * x[j];
    double pf = acc - cfg->pen_h[t];
    double k = cfg->pen_k[t], v;
    const double mult = cfg->pen_mult ? cfg->pen_mult[t] : 0.0;
    switch (cfg->pen_type[t]) {
      case 0: v = k * pow(pf, 2); break;                        /* penalty.py:88  */
      case 1: v = k * fabs(pf); break;                          /* penalty.py:147 */
      case 2: { double m = pf > 0. ? pf : 0.; v = (2 * k) * pow(m, 2); break; }  /* :388 */
      case 3: { double m = pf > 0. ? pf : 0.; v = (2 * k) * fabs(m); break; }    /* :447 */
      case 4: v = (pf != 0.) ? k : 0.; break;                   /* :208 uniform_equality: _k if pf else 0.0 */
      case 5: v = (pf > 0.) ? k : 0.; break;                    /* :266 uniform_inequality */
      case 6: v = (pf > 0.) ? INFINITY : (-.5 / k) * log(-pf); break; /* :322-326 barrier (numpy.log: log(0) = -inf) */
      case 7: {                                                 /* :516-517 lagrange_inequality */
        const double lo = -mult / (2. * k);
        const double mpf = (pf > lo) ? pf : lo;                 /* Python max(lo, pf) */
        v = k * pow(mpf, 2) + mult * mpf;
        break;
      }
      default: v = k * pow(pf, 2) + mult * pf; break;           /* :583 lagrange_equality */
    }
    total = v + total;
  }
  return total;
}

/* (inf if out of range else cost) + penalty: differential_evolution.py:515-525, tools.py:386-389,420-426 */
static double energy_of(const double* x, const orc_config* cfg, double* scratch) {
  const int D = cfg->dim;
  int oob = 0;
  if (cfg->bounds_mode != BOUNDS_NONE)
    for (int i = 0; i < D; ++i) oob |= (x[i] < cfg->lo[i]) || (x[i] > cfg->hi[i]);
  double raw = INFINITY;
  if (!oob) {
    switch (cfg->cost) {
      case COST_ROSEN: raw = cost_rosen(x, D, scratch); break;
      case COST_CORANA: raw = cost_corana(x, D); break;
      case COST_GRIEWANGK: raw = cost_griewangk(x, D); break;
      case COST_ZIMMERMANN: raw = cost_zimmermann(x); break;
      case COST_POLYFIT:
      case COST_DECAY:
      case COST_LORENTZIAN: raw = cost_datafit(x, D, cfg, scratch); break;
      case COST_SPHERE:
      case COST_RASTRIGIN:
      case COST_SCHWEFEL:
      case COST_ACKLEY:
      case COST_STEP:
      case COST_MICHAL:
      case COST_POWERS: raw = cost_separable(x, D, cfg->cost); break;
      case COST_BRANINS: case COST_EASOM: case COST_GOLDSTEIN: case COST_PEAKS: case COST_VENKAT91: case COST_FOSC3D:
      case COST_NMIN51: case COST_SHEKEL: case COST_ELLIPSOID: case COST_PAVIANI: raw = cost_fixed(x, D, cfg->cost); break;
      default: raw = cost_chebyshev(x, D, cfg); break;
    }
  }
  return raw + penalty_value(x, D, cfg);
}

/* selection + best: differential_evolution.py:603-613.  Sequential (ascending candidate). */
static void select_best(int64_t np, int D, const double* trial_e, const double* e_cur, const double* pop_next,
                        double* best_e, int64_t* best_idx, double* best_x, int64_t* n_inf, int64_t* n_acc) {
  int64_t ninf = 0, nacc = 0, w = -1;
  double be = *best_e;
  for (int64_t c = 0; c < np; ++c) {
    if (isinf(trial_e[c])) ++ninf;
    if (trial_e[c] < e_cur[c]) {
      ++nacc;
      if (trial_e[c] < be) {
        be = trial_e[c];
        w = c;
      }
    }
  }
  if (w >= 0) {
    *best_e = be;
    *best_idx = w;
    memcpy(best_x, pop_next + w * D, sizeof(double) * D);
  } else {
    *best_idx = -1;
  }
  *n_inf = ninf;
  *n_acc = nacc;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* bench.py's CPU legs set the thread count themselves: torchrun exports OMP_NUM_THREADS=1 */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n >= 1) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* generation 0: clip into the strict range (differential_evolution.py:516-520), evaluate, select.
 * pop_next receives the (clipped) population, e_next its energies; trial_e_out [np] scratch/out. */
int orc_step0(const orc_config* cfg, const double* pop_cur, double* pop_next, double* e_next, double* best_x,
              double* best_e, int64_t* best_idx, double* trial_e_out, int64_t* n_inf, int64_t* n_acc) {
  const int D = cfg->dim;
  const int64_t np = cfg->np;
#pragma omp parallel
  {
    double* scratch = (double*)malloc(sizeof(double) * ((D > cfg->fit_M ? D : cfg->fit_M) + 1));
#pragma omp for schedule(static)
    for (int64_t c = 0; c < np; ++c) {
      double* row = pop_next + c * D;
      for (int i = 0; i < D; ++i) {
        double v = pop_cur[c * D + i];
        if (cfg->bounds_mode != BOUNDS_NONE) v = fmin(fmax(v, cfg->lo[i]), cfg->hi[i]);
        row[i] = v;
      }
      double E = energy_of(row, cfg, scratch);
      trial_e_out[c] = E;
      e_next[c] = (E < INFINITY) ? E : INFINITY;
    }
    free(scratch);
  }
  double* inf_cur = (double*)malloc(sizeof(double) * np);
  for (int64_t c = 0; c < np; ++c) inf_cur[c] = INFINITY;
  *best_e = INFINITY;
  memcpy(best_x, pop_next, sizeof(double) * D);
  select_best(np, D, trial_e_out, inf_cur, pop_next, best_e, best_idx, best_x, n_inf, n_acc);
  free(inf_cur);
  return 0;
}

/* one generation g >= 1.  rng_mode 0: Philox(seed, generation, offset); 1: replayed draws
 * (donors [np,5] int32, nstart [np] int32, uni [np, dim+1] double, NaN = not drawn).
 * trial_out [np,dim] and trial_e_out [np] receive every trial (trial_out may be NULL). */
int orc_step(const orc_config* cfg, const double* pop_cur, double* pop_next, const double* e_cur, double* e_next,
             double* best_x, double* best_e, int64_t* best_idx, int rng_mode, uint64_t seed, uint32_t generation,
             int64_t offset, const int32_t* r_donors, const int32_t* r_nstart, const double* r_uni,
             double* trial_out, double* trial_e_out, int64_t* n_inf, int64_t* n_acc) {
  const int D = cfg->dim;
  const int64_t np = cfg->np;
  int base, xover, k;
  strategy_info(cfg->strategy, &base, &xover, &k);
  const double F = cfg->F, CR = cfg->CR;
  const uint32_t thr16 = crossover_threshold16(CR);
  uint32_t* xtab = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(D > 0 ? D : 1));
  const int xnmax = build_exp_length_table(crossover_threshold(CR), D, xtab);
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma omp parallel
  {
    double* t = (double*)malloc(sizeof(double) * (D + 1));
    double* scratch = (double*)malloc(sizeof(double) * ((D > cfg->fit_M ? D : cfg->fit_M) + 1));
    unsigned char* mask = (unsigned char*)malloc(D);
#pragma omp for schedule(static)
    for (int64_t c = 0; c < np; ++c) {
      int64_t r[5] = {0, 0, 0, 0, 0};
      int n;
      uint32_t xword = 0;
      const uint32_t cg = (uint32_t)(offset + c);
      if (rng_mode == 1) {
        for (int j = 0; j < k; ++j) r[j] = r_donors[c * 5 + j];
        n = r_nstart[c];
      } else {
        draw_donors(seed, generation, cg, c, np, D, k, r, &n, &xword);
      }
      /* crossover mask */
      if (xover == XOVER_BIN) {
        /* strategy.py:79-85 */
        uint32_t v[4] = {0, 0, 0, 0};
        for (int i = 0; i < D; ++i) {
          int hit;
          if (rng_mode == 1) {
            hit = r_uni[c * (int64_t)(D + 1) + i] < CR;
          } else {
            /* dimension 8q+e: bits 16*(e&1)..+15 of word e>>1 of block q */
            if ((i & 7) == 0) philox((uint32_t)(i >> 3), cg, generation, 1u, k0, k1, v);
            const uint32_t word = v[(i & 7) >> 1];
            hit = ((i & 1) ? (word >> 16) : (word & 0xffffu)) < thr16;
          }
          mask[i] = (unsigned char)(hit || i == n);
        }
      } else {
        /* strategy.py:48-56: L leading successes, capped at D */
        int L = 0;
        if (rng_mode == 1) {
          while (L < D && r_uni[c * (int64_t)(D + 1) + L] < CR) ++L;
        } else {
          L = exp_length(xword, xtab, xnmax, D); /* one geometric draw (protocol 2) */
        }
        memset(mask, 0, D);
        for (int j = 0; j < L; ++j) mask[(n + j) % D] = 1;
      }
      /* mutation, reference operation order (strategy.py formulas) */
      const double* p = pop_cur + c * D;
      const double* p0 = pop_cur + r[0] * D;
      const double* p1 = pop_cur + r[1] * D;
      const double* p2 = pop_cur + r[2] * D;
      const double* p3 = pop_cur + r[3] * D;
      const double* p4 = pop_cur + r[4] * D;
      for (int i = 0; i < D; ++i) {
        double v = p[i];
        if (mask[i]) {
          switch (base) {
            case BASE_BEST1: v = best_x[i] + F * (p0[i] - p1[i]); break;
            case BASE_RAND1: v = p0[i] + F * (p1[i] - p2[i]); break;
            case BASE_RANDTOBEST1: v = p[i] + (F * (best_x[i] - p[i]) + F * (p0[i] - p1[i])); break;
            case BASE_BEST2: v = best_x[i] + F * (p0[i] + p1[i] - p2[i] - p3[i]); break;
            default: v = p0[i] + F * (p1[i] + p2[i] - p3[i] - p4[i]); break;
          }
        }
        if (cfg->bounds_mode == BOUNDS_CLAMP) { /* symbolic.py:1139-1144 */
          v = (v > cfg->lo[i]) ? v : cfg->lo[i];
          v = (v < cfg->hi[i]) ? v : cfg->hi[i];
        }
        t[i] = v;
      }
      const double E = energy_of(t, cfg, scratch);
      trial_e_out[c] = E;
      if (trial_out) memcpy(trial_out + c * D, t, sizeof(double) * D);
      const int acc = E < e_cur[c];
      e_next[c] = acc ? E : e_cur[c];
      memcpy(pop_next + c * D, acc ? t : p, sizeof(double) * D);
    }
    free(t);
    free(scratch);
    free(mask);
  }
  free(xtab);
  select_best(np, D, trial_e_out, e_cur, pop_next, best_e, best_idx, best_x, n_inf, n_acc);
  return 0;
}

/* Philox stream 2: device-side SetRandomInitialPoints twin (abstract_solver.py:532-534) */
int orc_init_population(double* pop, int64_t np, int D, int64_t offset, const double* lo, const double* hi,
                        uint64_t seed) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < np; ++c) {
    for (int b = 0; b < (D + 1) / 2; ++b) {
      uint32_t v[4];
      philox((uint32_t)b, (uint32_t)(offset + c), 0u, 2u, k0, k1, v);
      uint64_t r0 = ((uint64_t)v[0] << 32) | v[1], r1 = ((uint64_t)v[2] << 32) | v[3];
      int j = 2 * b;
      pop[c * D + j] = lo[j] + (hi[j] - lo[j]) * ((double)(r0 >> 11) * (1.0 / 9007199254740992.0));
      if (j + 1 < D) pop[c * D + j + 1] = lo[j + 1] + (hi[j + 1] - lo[j + 1]) * ((double)(r1 >> 11) * (1.0 / 9007199254740992.0));
    }
  }
  return 0;
}
