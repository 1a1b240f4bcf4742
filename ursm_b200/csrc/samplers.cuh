// Device samplers for the URSM Gibbs sweeps.  Everything is __host__ __device__
// so tests/host_samplers.cu can check the arithmetic on the CPU with the very
// same code the kernels inline.
//
//   pg1_devroye      w ~ PG(1, psi)            e_step_gibbs.py:318-323 (pypolyagamma)
//   s_threshold      S ~ Bern(expit(...))      e_step_gibbs.py:326-342, threshold form
//   normal_pair      2x N(0,1)                 e_step_gibbs.py:376 (multivariate_normal)
//   gamma_mt         Gamma(shape,1)            e_step_gibbs.py:144 (dirichlet)
//   binomial         Bin(n,p)                  e_step_gibbs.py:137 (multinomial = chain of binomials)
#pragma once
#include <math.h>
#include "rng.cuh"

namespace ursm {

#if defined(__CUDA_ARCH__)
#define URSM_EXPF(x) __expf(x)
#define URSM_LOGF(x) __logf(x)
#else
#define URSM_EXPF(x) expf(x)
#define URSM_LOGF(x) logf(x)
#endif

// ---------------------------------------------------------------------------
// erfcx for the host build (device has erfcxf)
// ---------------------------------------------------------------------------
URSM_HD float erfcx_f(float t) {
#if defined(__CUDA_ARCH__)
  return erfcxf(t);
#else
  double x = t;
  if (x < 5.0) return (float)(exp(x * x) * erfc(x));
  double i2 = 1.0 / (x * x);
  return (float)(0.5641895835477563 / x * (1.0 - 0.5 * i2 + 0.75 * i2 * i2 - 1.875 * i2 * i2 * i2));
#endif
}

// ---------------------------------------------------------------------------
// PG(1, psi): Devroye's alternating-series sampler for the Jacobi J*(1, z)
// distribution, z = |psi|/2 (Polson, Scott & Windle 2013, sec. 4) -- the
// algorithm pypolyagamma 1.1.1 dispatches to for n == 1.  Restated for fp32:
//  * the mixture weight q/p is written so the z^2 terms cancel analytically
//    (erfcx instead of log Phi + exp), which removes the catastrophic
//    cancellation a literal transcription has for |psi| >~ 50;
//  * the series test runs on the RATIOS a_n/a_0 = (2n+1) e^{n(n+1)} with
//    e = exp(-2/x) (x <= t) or exp(-pi^2 x/2) (x > t): one exponential per
//    proposal, no underflow of a_0 for tiny x (large |psi|);
//  * the Michael-Schucany-Haas root uses the cancellation-free form.
// ---------------------------------------------------------------------------
#define URSM_PG_T 0.64f
#define URSM_PG_INV_T 1.5625f

// P(proposal is drawn from the exponential right tail)
URSM_HD float pg_mass_texpon(float z) {
  const float t = URSM_PG_T;
  const float kPi2_8 = 1.2337005501361697f;            // pi^2/8
  const float kC = 1.0083530457f;                         // exp(t pi^2/8 - 1/(2t))
  const float kInvSqrt2t = 0.8838834764831844f;         // 1/sqrt(2t)
  const float k4OverPi = 1.2732395447351628f;
  float fz = kPi2_8 + 0.5f * z * z;
  float ua = (t * z + 1.0f) * kInvSqrt2t;               // -a/sqrt2 >= 0
  float ea = kC * 0.5f * erfcx_f(ua);                   // exp(xa)/fz
  float eb;
  if (z < URSM_PG_INV_T) {
    float ub = (1.0f - t * z) * kInvSqrt2t;             // -b/sqrt2 > 0
    eb = kC * 0.5f * erfcx_f(ub);
  } else {
    float bs = (t * z - 1.0f) * kInvSqrt2t;             // b/sqrt2 >= 0
    float Phi = 1.0f - 0.5f * erfcf(bs);
    eb = URSM_EXPF(t * fz - z) * Phi;                   // may overflow to +inf -> mass 0
  }
  float qdivp = k4OverPi * fz * (ea + eb);
  return 1.0f / (1.0f + qdivp);
}

template <class RNG>
URSM_HD float pg_rtigauss(float z, RNG& rng) {
  const float t = URSM_PG_T;
  if (z < URSM_PG_INV_T) {
    // mu = 1/z > t: exponential-pair proposal for the truncated IG
    for (int guard = 0; guard < 1000; ++guard) {
      float e1 = -URSM_LOGF(rng.uniform());
      float e2 = -URSM_LOGF(rng.uniform());
      if (e1 * e1 > 2.0f * e2 / t) continue;
      float x = 1.0f + e1 * t;
      x = t / (x * x);
      float alpha = URSM_EXPF(-0.5f * z * z * x);
      if (rng.uniform() <= alpha) return x;
    }
    return t;  // unreachable in practice
  }
  float mu = 1.0f / z;
  for (int guard = 0; guard < 1000; ++guard) {
    // chi^2_1 via Box-Muller radius and angle
    float u1 = rng.uniform(), u2 = rng.uniform();
    float c;
#if defined(__CUDA_ARCH__)
    c = __cosf(6.283185307179586f * u2);
#else
    c = cosf(6.283185307179586f * u2);
#endif
    float y = -2.0f * URSM_LOGF(u1) * c * c;
    float my = mu * y;
    // smaller root of the MSH quadratic, written without cancellation
    float x = mu / (1.0f + 0.5f * my + 0.5f * sqrtf(4.0f * my + my * my));
    if (rng.uniform() > mu / (mu + x)) x = mu * mu / x;
    if (x <= t) return x;
  }
  return t;
}

template <class RNG>
URSM_HD float pg1_devroye(float psi, RNG& rng) {
  const float t = URSM_PG_T;
  float z = 0.5f * fabsf(psi);
  float fz = 1.2337005501361697f + 0.5f * z * z;
  float mass = pg_mass_texpon(z);
  for (int guard = 0; guard < 1000; ++guard) {
    float x;
    if (rng.uniform() < mass)
      x = t + (-URSM_LOGF(rng.uniform())) / fz;
    else
      x = pg_rtigauss(z, rng);
    // e = exp(-2/x) on the left piece, exp(-pi^2 x / 2) on the right piece
    float e = (x > t) ? URSM_EXPF(-4.934802200544679f * x) : URSM_EXPF(-2.0f / x);
    float q = e * e;
    float p = q, m = q;          // p = e^{n(n+1)}, m = e^{2n}
    float s = 1.0f;
    float u = rng.uniform();
    float coef = 3.0f;
    bool accept = false, done = false;
    for (int n = 1; n < 24 && !done; ++n) {
      float r = coef * p;
      if (n & 1) {
        s -= r;
        if (u <= s) { accept = true; done = true; }
      } else {
        s += r;
        if (u > s) done = true;
      }
      m *= q;
      p *= m;
      coef += 2.0f;
    }
    if (accept) return 0.25f * x;
  }
  return 0.25f * t;
}

// ---------------------------------------------------------------------------
// Dropout indicator in threshold form (hard part 1).
//   S_new = 1  <=>  u < expit(psi - R log1p(a / s_other))  <=>  s_other > T
// with T = a / expm1((psi - logit u)/R)  (T = +inf when the argument is <= 0).
// R == 0 (cell without reads; the reference's `sum_other == 0` branch,
// e_step_gibbs.py:334-335) decides on logit(u) < psi alone: T = -inf / +inf.
// ---------------------------------------------------------------------------
URSM_HD float s_threshold(float psi, float u, float a, float R) {
  float lg = URSM_LOGF(u) - log1pf(-u);  // logit(u)
  float D = psi - lg;
  if (R <= 0.0f) return (D > 0.0f) ? -INFINITY : INFINITY;
  if (D <= 0.0f) return INFINITY;
  float e = expm1f(D / R);
  return a / e;  // e == +inf -> 0: always on
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
template <class RNG>
URSM_HD double gamma_mt(double shape, RNG& rng) {
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

}  // namespace ursm
