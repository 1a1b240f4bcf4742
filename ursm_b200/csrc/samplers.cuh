// Device samplers for the URSM Gibbs sweeps.  Everything is __host__ __device__
// so tests/host_samplers.cu can check the arithmetic on the CPU with the very
// same code the kernels inline.
//
//   pg_tilt/pg_attempt/pg1_draw   w ~ PG(1, psi)   e_step_gibbs.py:318-323 (pypolyagamma)
//   s_threshold      S ~ Bern(expit(...))      e_step_gibbs.py:326-342, threshold form
//   normal_pair      2x N(0,1)                 e_step_gibbs.py:376 (multivariate_normal)
//   gamma_mt         Gamma(shape,1)            e_step_gibbs.py:144 (dirichlet)
//   binomial         Bin(n,p)                  e_step_gibbs.py:137 (multinomial = chain of binomials)
#pragma once
#include <math.h>
#include "rng.cuh"

namespace ursm {

// fast single-precision special functions on the device: the bare MUFU approximations
// with denormals flushed (ex2/lg2/rcp/rsqrt.approx.ftz, 2-ulp class) -- one MUFU and one
// multiply each.  The non-ftz forms the __expf/__logf/__fdividef intrinsics compile to
// without -use_fast_math wrap every MUFU in a denormal-range rescaling (3-4 extra
// instructions each, ~60 per PG attempt); nothing here needs denormals: every
// quantity that can underflow is a probability mass that is then exactly 0.  libm on the host.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float mufu_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float mufu_lg2(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float mufu_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float mufu_rsqrt(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
#define URSM_EXPF(x) ::ursm::mufu_ex2(1.4426950408889634f * (x))
#define URSM_LOGF(x) (0.6931471805599453f * ::ursm::mufu_lg2(x))
#define URSM_DIV(a, b) ((a) * ::ursm::mufu_rcp(b))
#define URSM_RCP(b) ::ursm::mufu_rcp(b)
#define URSM_RSQRT(x) ::ursm::mufu_rsqrt(x)
#else
#define URSM_EXPF(x) expf(x)
#define URSM_LOGF(x) logf(x)
#define URSM_DIV(a, b) ((a) / (b))
#define URSM_RCP(b) (1.0f / (b))
#define URSM_RSQRT(x) (1.0f / sqrtf(x))
#endif

// ---------------------------------------------------------------------------
// PG(1, psi) = J*(1, z)/4, z = |psi|/2  (Polson, Scott & Windle 2013, sec. 4;
// Devroye 2009) -- the distribution pypolyagamma 1.1.1 samples for n == 1.
//
// Same target, same alternating-series representation f(x) ~ exp(-z^2 x/2) *
// sum_n (-1)^n a_n(x), same split point t = 0.64, but restated as ONE rejection
// sampler whose every attempt costs a fixed amount of work and a fixed number of
// random words (one Philox block), so that a warp can run attempts in lock step
// while each lane walks through its own genes (csrc/sc_gibbs.cu):
//
//   piece R (x > t)   envelope (pi/2) exp(-fz x), fz = pi^2/8 + z^2/2; mass
//                     p = (pi/2) exp(-fz t)/fz; x = t + Exp(1)/fz;
//                     accept w.p. S_R(x) = 1 - 3q + 5q^3 - 7q^6, q = exp(-pi^2 x)
//   piece L, z <  1/t envelope a_0(x) WITHOUT the tilt (truncated Levy), mass
//                     q0 = 4 Phi(-1/sqrt t); x = 1/Z^2, Z ~ N(0,1) | |Z| > 1/sqrt t
//                     (inversion); accept w.p. exp(-z^2 x/2) S_L(x), q = exp(-4/x)
//   piece L, z >= 1/t envelope a_0(x) exp(-z^2 x/2) on (0, inf), mass 2 exp(-z)
//                     (inverse Gaussian(1/z, 1), Michael-Schucany-Haas root);
//                     accept w.p. 1[x <= t] S_L(x)
//
// The pieces are chosen with probability proportional to their masses; any
// rejection restarts the whole attempt.  Devroye's original nests exact samplers
// for the tilted/truncated left piece inside the mixture; folding those inner
// rejections into the outer one changes only the envelope (a looser left piece
// with a closed-form mass), not the target -- and removes the erfc/erfcx from
// the mixture weight.  The alternating series is summed to n = 3 instead of being
// squeezed term by term: with the regime chosen by x the ratios are a_1/a_0 <=
// 5.8e-3, a_4/a_0 < 1e-53, far below fp32 resolution.
// Acceptance per attempt: 0.999 at psi = 0, 0.90 at |psi| = 2, >= 0.73 (worst,
// |psi| ~ 3.1), -> 1 for large |psi|.
// ---------------------------------------------------------------------------
#define URSM_PG_T 0.64f
#define URSM_PG_INV_T 1.5625f
#define URSM_PG_C0 0.10564977366685535f  // Phi(-1/sqrt t)

struct PgTilt {
  float hz2, inv_fz, mass_right, inv_left, mu;
  bool ig;  // left piece by inverse Gaussian (z >= 1/t) or by truncated Levy
};

URSM_HD PgTilt pg_tilt(float psi) {
  PgTilt g;
  const float t = URSM_PG_T;
  const float z = 0.5f * fabsf(psi);
  g.hz2 = 0.5f * z * z;
  const float fz = 1.2337005501361697f + g.hz2;
  g.inv_fz = URSM_RCP(fz);
  g.ig = z >= URSM_PG_INV_T;
  // q/p with the cosh(z) factors cancelled; the exponent is >= 0.008 for every z
  const float ex = g.ig ? (t * fz - z) : (t * fz);
  const float c = g.ig ? 1.2732395447351628f /* 4/pi */ : 0.26903493f /* 2 q0/pi */;
  const float qdivp = c * fz * URSM_EXPF(ex);  // +inf for large z -> mass 0
  // mass_right = 1/(1 + q/p); 1/(1 - mass_right) = 1 + p/q (one reciprocal each, no
  // cancellation)
  g.mass_right = URSM_RCP(1.0f + qdivp);
  g.inv_left = 1.0f + URSM_RCP(qdivp);
  g.mu = URSM_RCP(fmaxf(z, 1e-20f));
  return g;
}

// erfinv(x)/x as a function of w = -log(1 - x^2): Giles' single-precision
// approximation (2012); the tail branch (w >= 5, |x| > 0.9966) is a real branch, taken
// by about 1% of the lanes
URSM_HD float erfinv_over_x(float w) {
  if (w < 5.0f) {
    const float wc = w - 2.5f;
    float p = 2.81022636e-08f;
    p = fmaf(p, wc, 3.43273939e-07f);
    p = fmaf(p, wc, -3.5233877e-06f);
    p = fmaf(p, wc, -4.39150654e-06f);
    p = fmaf(p, wc, 0.00021858087f);
    p = fmaf(p, wc, -0.00125372503f);
    p = fmaf(p, wc, -0.00417768164f);
    p = fmaf(p, wc, 0.246640727f);
    return fmaf(p, wc, 1.50140941f);
  }
  const float wt = w * URSM_RSQRT(w) - 3.0f;
  float r = -0.000200214257f;
  r = fmaf(r, wt, 0.000100950558f);
  r = fmaf(r, wt, 0.00134934322f);
  r = fmaf(r, wt, -0.00367342844f);
  r = fmaf(r, wt, 0.00573950773f);
  r = fmaf(r, wt, -0.0076224613f);
  r = fmaf(r, wt, 0.00943887047f);
  r = fmaf(r, wt, 1.00167406f);
  return fmaf(r, wt, 2.83297682f);
}

// Z^2 for Z = Phi^{-1}(v), v in (0,1)
URSM_HD float ndtri_sq(float v) {
  const float e = erfinv_over_x(-URSM_LOGF(4.0f * v * (1.0f - v))) * (2.0f * v - 1.0f);
  return 2.0f * e * e;
}

// One attempt from three uniforms in (0,1).  Returns true and x ~ J*(1, z) on acceptance.
URSM_HD bool pg_attempt(const PgTilt& g, float u0, float u1, float u2, float& x) {
  const float t = URSM_PG_T;
  const bool right = u0 < g.mass_right;
  // one logarithm serves both pieces: Exp(1) = -log u1 on the right, the erfinv
  // argument -log(1 - (2v-1)^2) on the left
  const float v = g.ig ? u1 : u1 * URSM_PG_C0;
  const float lg = -URSM_LOGF(right ? u1 : 4.0f * v * (1.0f - v));
  // piece R
  const float xr = fmaf(lg, g.inv_fz, t);
  // piece L: zn2 = Z^2, Z = Phi^{-1}(v)
  const float e = erfinv_over_x(lg) * (2.0f * v - 1.0f);
  const float zn2 = 2.0f * e * e;
  //   inverse Gaussian(mu, 1) by Michael-Schucany-Haas (smaller root written
  //   without cancellation); the root choice reuses u0 | left
  const float my = g.mu * zn2;
  const float disc = fmaf(my, my, 4.0f * my);
  float xi = URSM_DIV(g.mu, 1.0f + 0.5f * my + 0.5f * disc * URSM_RSQRT(fmaxf(disc, 1e-30f)));
  const float up = (u0 - g.mass_right) * g.inv_left;  // uniform given "left"
  if (up * (g.mu + xi) > g.mu) xi = URSM_DIV(g.mu * g.mu, xi);
  //   truncated Levy: x = 1/Z^2
  const float inv_xl = g.ig ? URSM_DIV(1.0f, xi) : zn2;
  const float xl = g.ig ? xi : URSM_DIV(1.0f, zn2);
  x = right ? xr : xl;
  const float q = URSM_EXPF(right ? -9.869604401089358f * xr : -4.0f * inv_xl);
  const float q3 = q * q * q;
  const float series = 1.0f - 3.0f * q + 5.0f * q3 - 7.0f * q3 * q3;
  const float tilt = (right || g.ig) ? ((right || xl <= t) ? 1.0f : 0.0f) : URSM_EXPF(-g.hz2 * xl);
  return u2 <= tilt * series;
}

// Convenience draw (host tests only; the kernels run pg_attempt in their own warp loop, which has
// no attempt limit): attempts until accepted.
template <class RNG>
URSM_HD float pg1_draw(float psi, RNG& rng) {
  const PgTilt g = pg_tilt(psi);
  for (int guard = 0; guard < 10000; ++guard) {
    const float u0 = rng.uniform(), u1 = rng.uniform(), u2 = rng.uniform();
    float x;
    if (pg_attempt(g, u0, u1, u2, x)) return 0.25f * x;
  }
  return 0.25f * URSM_PG_T;
}

// ---------------------------------------------------------------------------
// Dropout indicator in threshold form (hard part 1).
//   S_new = 1  <=>  u < expit(psi - R log1p(a / s_other))  <=>  s_other > T
// with T = a / expm1((psi - logit u)/R)  (T = +inf when the argument is <= 0).
// R == 0 (cell without reads; the reference's `sum_other == 0` branch,
// e_step_gibbs.py:334-335) decides on logit(u) < psi alone: T = -inf / +inf.
// ---------------------------------------------------------------------------
// expm1(x) for x > 0: Taylor to x^6 below 1/4 (relative error < 5e-8), exp - 1 above
URSM_HD float expm1_pos(float x) {
  const float p = x * fmaf(x, fmaf(x, fmaf(x, fmaf(x, fmaf(x, 1.0f / 720, 1.0f / 120), 1.0f / 24),
                                          1.0f / 6), 0.5f), 1.0f);
  return (x < 0.25f) ? p : URSM_EXPF(x) - 1.0f;
}

// `invR` = 1/R for R > 0, anything for R <= 0
URSM_HD float s_threshold(float psi, float u, float a, float R, float invR) {
  const float lg = URSM_LOGF(URSM_DIV(u, 1.0f - u));  // logit(u); 1-u is exact for 23-bit u
  const float D = psi - lg;
  if (R <= 0.0f) return (D > 0.0f) ? -INFINITY : INFINITY;
  if (D <= 0.0f) return INFINITY;
  return URSM_DIV(a, expm1_pos(D * invR));  // expm1 == +inf -> 0: always on
}
URSM_HD float s_threshold(float psi, float u, float a, float R) {
  return s_threshold(psi, u, a, R, R > 0.0f ? 1.0f / R : 0.0f);
}

// ---------------------------------------------------------------------------
// Normal pair (Box-Muller), double precision: feeds the 2x2 (kappa, tau) draw.
// ---------------------------------------------------------------------------
template <class RNG>
URSM_HD void normal_pair(RNG& rng, double& n0, double& n1) {
  double u1 = rng.uniform_d(), u2 = rng.uniform_d();
  double r = sqrt(-2.0 * log(u1));
  double s, c;
#if defined(__CUDA_ARCH__)
  sincospi(2.0 * u2, &s, &c);
#else
  s = sin(6.283185307179586 * u2);
  c = cos(6.283185307179586 * u2);
#endif
  n0 = r * c;
  n1 = r * s;
}

// Gamma(shape, 1), Marsaglia & Tsang (2000); shape < 1 via the U^(1/shape) boost.
// `failed` (optional) is set if the rejection loop runs out of tries (probability ~ 0.05^1000:
// a broken generator, never a legitimate draw); the value returned then is the mode, not a draw.
template <class RNG>
URSM_HD double gamma_mt(double shape, RNG& rng, bool* failed = nullptr) {
  double boost = 1.0;
  if (shape < 1.0) {
    boost = pow(rng.uniform_d(), 1.0 / shape);
    shape += 1.0;
  }
  double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (int guard = 0; guard < 1000; ++guard) {
    double x, dummy;
    normal_pair(rng, x, dummy);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    double u = rng.uniform_d();
    double x2 = x * x;
    if (u < 1.0 - 0.0331 * x2 * x2) return boost * d * v;
    if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return boost * d * v;
  }
  if (failed) *failed = true;
  return boost * d;
}

// ---------------------------------------------------------------------------
// Binomial(n, p): sequential inversion when n*min(p,1-p) <= 30, BTPE
// (Kachitvichyanukul & Schmeiser 1988) otherwise -- exact for every (n, p),
// which cuRAND does not offer (hard part 4).
// ---------------------------------------------------------------------------
URSM_HD double btpe_stirling(double v2) {
  return (13680.0 - (462.0 - (132.0 - (99.0 - 140.0 / v2) / v2) / v2) / v2) / 166320.0;
}

template <class RNG>
URSM_HD int binomial_inversion(int n, double r, RNG& rng) {
  // r <= 0.5, n r <= 30
  double q = 1.0 - r;
  double qn = exp(n * log1p(-r));
  double s = r / q;
  double bound = fmin((double)n, n * r + 10.0 * sqrt(n * r * q + 1.0));
  for (int guard = 0; guard < 64; ++guard) {
    double px = qn, u = rng.uniform_d();
    int x = 0;
    bool ok = true;
    while (u > px) {
      ++x;
      if (x > bound) { ok = false; break; }
      u -= px;
      px = ((n - x + 1) * s * px) / x;
    }
    if (ok) return x;
  }
  return (int)(n * r);
}

template <class RNG>
URSM_HD int binomial_btpe(int n, double r, RNG& rng) {
  double q = 1.0 - r;
  double nrq = n * r * q;
  double fm = n * r + r;
  int m = (int)floor(fm);
  double p1 = floor(2.195 * sqrt(nrq) - 4.6 * q) + 0.5;
  double xm = m + 0.5, xl = xm - p1, xr = xm + p1;
  double c = 0.134 + 20.5 / (15.3 + m);
  double a = (fm - xl) / (fm - xl * r);
  double laml = a * (1.0 + a / 2.0);
  a = (xr - fm) / (xr * q);
  double lamr = a * (1.0 + a / 2.0);
  double p2 = p1 * (1.0 + 2.0 * c);
  double p3 = p2 + c / laml;
  double p4 = p3 + c / lamr;
  for (int guard = 0; guard < 10000; ++guard) {
    double u = rng.uniform_d() * p4, v = rng.uniform_d();
    int y;
    if (u <= p1) {  // triangle
      return (int)floor(xm - p1 * v + u);
    } else if (u <= p2) {  // parallelogram
      double x = xl + (u - p1) / c;
      v = v * c + 1.0 - fabs(m - x + 0.5) / p1;
      if (v > 1.0) continue;
      y = (int)floor(x);
    } else if (u <= p3) {  // left exponential tail
      y = (int)floor(xl + log(v) / laml);
      if (y < 0) continue;
      v = v * (u - p2) * laml;
    } else {  // right exponential tail
      y = (int)floor(xr - log(v) / lamr);
      if (y > n) continue;
      v = v * (u - p3) * lamr;
    }
    int k = abs(y - m);
    if (k <= 20 || k >= nrq / 2.0 - 1.0) {
      // explicit evaluation of f(y)/f(m)
      double s = r / q, aa = s * (n + 1.0), F = 1.0;
      if (m < y) {
        for (int i = m + 1; i <= y; ++i) F *= (aa / i - s);
      } else if (m > y) {
        for (int i = y + 1; i <= m; ++i) F /= (aa / i - s);
      }
      if (v > F) continue;
      return y;
    }
    // squeezes on log f(y)/f(m)
    double rho = (k / nrq) * ((k * (k / 3.0 + 0.625) + 0.16666666666666666) / nrq + 0.5);
    double tt = -(double)k * k / (2.0 * nrq);
    double A = log(v);
    if (A < tt - rho) return y;
    if (A > tt + rho) continue;
    double x1 = y + 1.0, f1 = m + 1.0, zz = n + 1.0 - m, w = n - y + 1.0;
    double bound = xm * log(f1 / x1) + (n - m + 0.5) * log(zz / w) +
                   (y - m) * log(w * r / (x1 * q)) + btpe_stirling(f1 * f1) / f1 +
                   btpe_stirling(zz * zz) / zz + btpe_stirling(x1 * x1) / x1 +
                   btpe_stirling(w * w) / w;
    if (A > bound) continue;
    return y;
  }
  return m;
}

template <class RNG>
URSM_HD int binomial(int n, double p, RNG& rng) {
  if (n <= 0 || p <= 0.0) return 0;
  if (p >= 1.0) return n;
  bool flip = p > 0.5;
  double r = flip ? 1.0 - p : p;
  int x = (n * r <= 30.0) ? binomial_inversion(n, r, rng) : binomial_btpe(n, r, rng);
  return flip ? n - x : x;
}

// ---------------------------------------------------------------------------
// Binomial(n, p) as a resumable state machine in fp32 for the bulk split kernel
// (csrc/bk_gibbs.cu::bk_split_kernel): lanes of a warp interleave "a few inversion
// steps" / "one BTRS attempt" instead of each running its own data-dependent loop.
//   n r <= 20  sequential inversion from 0 (as binomial_inversion above), fp32;
//   n r  > 20  BTRS, Hormann's transformed rejection with squeeze (1993): the
//              86% quick-accept path in fp32, the exact log-ratio test in fp64.
// `p` and its complement are passed separately (both are ratios of positive sums in
// the caller), so 1 - p is never formed by subtraction.
// ---------------------------------------------------------------------------
#define URSM_BINOM_TABLE 1024
#define URSM_BINOM_MODE_MAX 32.0f

#if defined(__CUDA_ARCH__)
#define URSM_LG2(x) ::ursm::mufu_lg2(x)
#define URSM_EX2(x) ::ursm::mufu_ex2(x)
#else
#define URSM_LG2(x) log2f(x)
#define URSM_EX2(x) exp2f(x)
#endif

enum : int { kBinDone = 0, kBinInv = 1, kBinBtrs = 2, kBinRetry = 3 };

struct BinomState {
  int n, x, mode, result, bound;
  float r, rc, px, u, s, c;  // s = r/(1-r), c = (n+1) s
  bool flip;
};

// n r above which BTRS replaces inversion (BTRS needs n r >= 10; pmf(0) >= e^-28 stays a
// normal fp32 number up to here)
#define URSM_BINOM_INV_MAX 20.0f

URSM_HD void binom_inv_start(BinomState& b, float u) {
  // pmf(0) = (1-r)^n
  // log(1 - r): series below r = 1/16 (relative error < 1e-8), log of the complement above
  const float r = b.r;
  const float ser = -r * fmaf(r, fmaf(r, fmaf(r, fmaf(r, fmaf(r, 1.0f / 6, 0.2f), 0.25f), 1.0f / 3),
                                      0.5f), 1.0f);
  b.px = URSM_EXPF((float)b.n * ((r < 0.0625f) ? ser : URSM_LOGF(b.rc)));
  b.x = 0;
  b.u = u;
  b.mode = kBinInv;
}

URSM_HD void binom_begin(BinomState& b, int n, float p, float pc, float u) {
  b.n = n;
  b.flip = p > 0.5f;
  b.r = b.flip ? pc : p;
  b.rc = b.flip ? p : pc;
  b.result = 0;
  if (n <= 0 || !(b.r > 0.0f)) {  // degenerate: everything to one side
    b.result = b.flip ? n : 0;
    b.mode = kBinDone;
    return;
  }
  const float nr = (float)n * b.r;
  if (nr <= URSM_BINOM_INV_MAX) {
    b.s = URSM_DIV(b.r, b.rc);
    b.c = (float)(n + 1) * b.s;
    const float var1 = fmaf(nr, b.rc, 1.0f);
    const int guard = (int)(nr + 10.0f * var1 * URSM_RSQRT(var1));
    b.bound = guard < n ? guard : n;
    binom_inv_start(b, u);
  } else {
    b.mode = kBinBtrs;
  }
}

// pmf(x)/pmf(x-1) = (n-x+1) s / x = c/x - s.  kBinDone when u falls inside the current
// atom, kBinRetry when the search ran past the guard bound (rounding deficit of the fp32
// pmf sum).
// Four CDF steps as straight-line predicated code (the kernel's inner loop): no exit test,
// mode test or branch per step -- a lane that has landed just stops advancing.  The guard
// bound is checked once per chunk (a lane can overrun it by at most three atoms, which
// are legitimate atoms of the same inversion; beyond n the pmf is <= 0 and nothing lands).
URSM_HD void binom_inv_chunk4(BinomState& b) {
  bool landed = false;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    landed = landed || (b.u <= b.px);
    if (!landed) {
      b.u -= b.px;
      ++b.x;
      b.px *= fmaf(b.c, URSM_RCP((float)b.x), -b.s);
    }
  }
  if (landed) {
    b.result = b.flip ? b.n - b.x : b.x;
    b.mode = kBinDone;
  } else if (b.x > b.bound) {
    b.mode = kBinRetry;
  }
}

URSM_HD float btrs_fc(int k) {  // Stirling series tail: log(k!) - Stirling's formula
  switch (k) {
    case 0: return 0.08106146679532726f;
    case 1: return 0.04134069595540929f;
    case 2: return 0.02767792568499834f;
    case 3: return 0.02079067210376509f;
    case 4: return 0.01664469118982119f;
    case 5: return 0.01387612882307075f;
    case 6: return 0.01189670994589177f;
    case 7: return 0.01041126526197209f;
    case 8: return 0.009255462182712733f;
    case 9: return 0.008330563433362871f;
  }
  // values <= 0.0076: fp32 keeps them to 1e-9 absolute, far below the test's resolution
  const float ik = URSM_DIV(1.0f, (float)k + 1.0f), ik2 = ik * ik;
  return (1.0f / 12 - (1.0f / 360 - 1.0f / 1260 * ik2) * ik2) * ik;
}

// one BTRS attempt from two uniforms; kBinDone on acceptance.  `lf2` (log2 k! table, may be
// null) makes the exact test O(1) for n inside the table.
URSM_HD void binom_btrs_attempt(BinomState& b, float U, float V, const double* lf2 = nullptr) {
  const float n = (float)b.n, r = b.r;
  const float spq = sqrtf(n * r * b.rc);
  const float bb = 1.15f + 2.53f * spq;
  const float aa = -0.0873f + 0.0248f * bb + 0.01f * r;
  const float c = n * r + 0.5f;
  const float vr = 0.92f - 4.2f / bb;
  const float u = U - 0.5f;
  const float us = 0.5f - fabsf(u);
  const float kf = floorf((2.0f * aa / us + bb) * u + c);
  if (!(kf >= 0.0f && kf <= n)) return;
  const int k = (int)kf;
  bool ok = (us >= 0.07f && V <= vr);
  if (!ok) {
    // exact test: V alpha / (a/us^2 + b) <= f(k)/f(m), m = mode
    const float alpha = (2.83f + URSM_DIV(5.1f, bb)) * spq;
    const float lhs = URSM_DIV(V * alpha, URSM_DIV(aa, us * us) + bb);
    const int m = (int)floorf((n + 1.0f) * r);
    const int d = k - m;
    if (lf2 != nullptr && b.n < URSM_BINOM_TABLE) {
      // log2 f(k)/f(m) = log2 m! + log2 (n-m)! - log2 k! - log2 (n-k)! + (k - m) log2(r/(1-r))
      const double lg = (lf2[m] + lf2[b.n - m]) - (lf2[k] + lf2[b.n - k]);
      const float rhs = fmaf((float)d, URSM_LG2(URSM_DIV(r, b.rc)), (float)lg);
      ok = URSM_LG2(lhs) <= rhs;
    } else if (d >= -32 && d <= 32) {
      // f(k)/f(m) as an explicit product of at most 32 ratios (fp32, as BTPE does near
      // the mode): f(i)/f(i-1) = (n-i+1)/i * r/(1-r)
      const float odds = URSM_DIV(r, b.rc);
      float F = 1.0f;
      if (d > 0) {
        for (int i = m + 1; i <= k; ++i) F *= URSM_DIV((float)(b.n - i + 1) * odds, (float)i);
      } else {
        for (int i = k + 1; i <= m; ++i) F *= URSM_DIV((float)i, (float)(b.n - i + 1) * odds);
      }
      ok = lhs <= F;
    } else {
      // far from the mode: Stirling form of log f(k)/f(m), in fp64
      const double rd = (double)r / (double)b.rc;  // odds
      const double nd = b.n;
      const double rhs = (m + 0.5) * log((m + 1.0) / (rd * (nd - m + 1.0))) +
                         (nd + 1.0) * log((nd - m + 1.0) / (nd - k + 1.0)) +
                         (k + 0.5) * log(rd * (nd - k + 1.0) / (k + 1.0)) +
                         (double)(btrs_fc(m) + btrs_fc(b.n - m) - btrs_fc(k) - btrs_fc(b.n - k));
      ok = log((double)lhs) <= rhs;
    }
  }
  if (ok) {
    b.result = b.flip ? b.n - k : k;
    b.mode = kBinDone;
  }
}

// ---------------------------------------------------------------------------
// Binomial(n, r), 0 < r <= 1/2, by inversion FROM THE MODE (Kemp 1986's modal search):
// the atoms are visited in the fixed order m, m+1, m-1, m+2, m-2, ... (m = floor((n+1) r))
// and the uniform is compared with their masses one after the other -- a valid inversion
// for any visiting order that does not depend on the uniform.  Expected atoms visited
// ~ 1 + 1.6 sqrt(n r (1-r)) instead of 1 + n r for the search from 0, which is what the
// bulk split kernel's cost per item is made of (counts of ~50 reads split over 5-10 types:
// n r ~ 5-10).  pmf(m) = C(n,m) r^m (1-r)^(n-m) comes from a table of log2 k! (double, in
// shared memory in the kernel) and two MUFU logarithms; neighbours by the ratio recurrences
//     pmf(x)/pmf(x-1) = c/x - s,   pmf(x-1)/pmf(x) = x / ((n-x+1) s),  s = r/(1-r), c = (n+1) s.
// Used for n < URSM_BINOM_TABLE and n r <= URSM_BINOM_MODE_MAX; relative error of the fp32
// masses ~ 1e-7 (n r + 10).  Returns the draw, or -1 when the uniform lies beyond the sum of
// the fp32 masses (probability ~1e-6): the caller retries with a fresh uniform, which keeps
// the draw unbiased.
// ---------------------------------------------------------------------------
URSM_HD int binom_mode_try(int n, float r, float rc, float u, const double* lf2) {
  const float nf = (float)n;
  int m = (int)((nf + 1.0f) * r);
  m = m < n ? m : n;
  const float mf = (float)m;
  // log2(1 - r): series below r = 1/16 (relative error < 1e-8), MUFU on the complement above
  const float ser = -r * fmaf(r, fmaf(r, fmaf(r, fmaf(r, fmaf(r, 1.0f / 6, 0.2f), 0.25f), 1.0f / 3),
                                      0.5f), 1.0f);
  const float l2rc = (r < 0.0625f) ? 1.4426950408889634f * ser : URSM_LG2(rc);
  const float logc = (float)(lf2[n] - lf2[m] - lf2[n - m]);  // log2 C(n, m)
  const float pm = URSM_EX2(fmaf(nf - mf, l2rc, fmaf(mf, URSM_LG2(r), logc)));
  if (u <= pm) return m;
  u -= pm;
  const float t = URSM_RCP(r * rc);
  const float s = r * r * t, is = rc * rc * t;  // r/(1-r) and its reciprocal
  const float c = (nf + 1.0f) * s;
  float pu = pm, pd = pm, fu = mf, fd = mf;    // masses and positions of the two cursors
  while (true) {
    const bool up = fu < nf && pu > 1e-30f, dn = fd > 0.0f && pd > 1e-30f;
    if (up) {
      fu += 1.0f;
      pu *= fmaf(c, URSM_RCP(fu), -s);
      if (u <= pu) return (int)fu;
      u -= pu;
    }
    if (dn) {
      pd *= fd * is * URSM_RCP(nf - fd + 1.0f);
      fd -= 1.0f;
      if (u <= pd) return (int)fd;
      u -= pd;
    }
    if (!up && !dn) return -1;
  }
}

// log2 k!, k = 0 .. URSM_BINOM_TABLE-1 (host side; the kernels stage a copy in shared memory)
inline void binom_fill_table(double* lf2) {
  for (int k = 0; k < URSM_BINOM_TABLE; ++k) lf2[k] = lgamma((double)k + 1.0) / 0.6931471805599453;
}

// which sampler the bulk split kernel uses for Bin(n, r), r <= 1/2
URSM_HD bool binom_use_mode(int n, float nr) {
  return n < URSM_BINOM_TABLE && nr <= URSM_BINOM_MODE_MAX;
}

// run the samplers to completion (host tests; same arithmetic and dispatch as the kernel)
template <class RNG>
URSM_HD int binomial_f32(int n, float p, float pc, RNG& rng, const double* lf2) {
  BinomState b;
  const bool flip = p > 0.5f;
  const float r = flip ? pc : p, rc = flip ? p : pc;
  if (n > 0 && r > 0.0f && binom_use_mode(n, (float)n * r)) {
    int x = binom_mode_try(n, r, rc, rng.uniform(), lf2);
    for (int guard = 0; guard < 1000 && x < 0; ++guard) x = binom_mode_try(n, r, rc, rng.uniform(), lf2);
    return flip ? n - x : x;
  }
  binom_begin(b, n, p, pc, rng.uniform());
  for (int guard = 0; guard < 100000 && b.mode != kBinDone; ++guard) {
    if (b.mode == kBinInv) {
      binom_inv_chunk4(b);
    } else if (b.mode == kBinBtrs) {
      const float U = rng.uniform(), V = rng.uniform();
      binom_btrs_attempt(b, U, V, lf2);
    } else {  // kBinRetry
      binom_inv_start(b, rng.uniform());
    }
  }
  return b.result;
}

}  // namespace ursm
