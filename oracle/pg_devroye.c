/* oracle/pg_devroye.c -- CPU restatement of the Polya-Gamma PG(1,z) sampler that the
 * reference calls through pypolyagamma 1.1.1 (un-vendored C++/Cython/OpenMP
 * dependency, /root/reference/README.md:22; call sites e_step_gibbs.py:12,195,204,
 * 209,323).  TEST INFRASTRUCTURE ONLY: linked by oracle/ursm_oracle.py for the
 * `cpu_baseline` / `--impl reference` timing leg and the tests; nothing under
 * ursm_b200/ uses it.
 *
 * Algorithm: Devroye's alternating-series method for J*(1,z), z=|psi|/2, as
 * published in Polson, Scott & Windle (2013) sec. 4 -- the method pypolyagamma's
 * hybrid sampler dispatches to when n == 1.  Double precision,
 * written literally (log Phi + exp), like the oracle's Python version of the same
 * function (ursm_oracle.py::pg1_devroye).  Parity for this piece is distributional
 * ("unpinned": the reference ships no golden vector at this boundary); it is
 * validated against the closed-form PG(1,z) moments in tests/test_oracle_c.py.
 *
 * pgdrawvpar semantics (e_step_gibbs.py:200-209,323): `nthreads` generators, entries
 * of one call split across OpenMP threads, thread t draws from generator t.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PG_T 0.64
#define HALF_PI2 1.2337005501361697 /* pi^2/8 */

typedef struct {
  uint64_t s[4];
} pg_rng; /* xoshiro256** */

static uint64_t splitmix(uint64_t* x) {
  uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static void rng_seed(pg_rng* r, uint64_t seed) {
  for (int i = 0; i < 4; ++i) r->s[i] = splitmix(&seed);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rng_next(pg_rng* r) {
  uint64_t* s = r->s;
  uint64_t result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
  s[2] ^= s[0];
  s[3] ^= s[1];
  s[1] ^= s[2];
  s[0] ^= s[3];
  s[2] ^= t;
  s[3] = rotl(s[3], 45);
  return result;
}
static inline double rng_unif(pg_rng* r) { /* (0,1) */
  return ((rng_next(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
static inline double rng_expo(pg_rng* r) { return -log(rng_unif(r)); }
static inline double rng_norm(pg_rng* r) {
  double u1 = rng_unif(r), u2 = rng_unif(r);
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

static double log_ndtr(double x) {
  if (x > -5.0) return log(0.5 * erfc(-x / 1.4142135623730951));
  double x2 = x * x;
  return -0.5 * x2 - log(-x) - 0.9189385332046727 +
         log1p(-1 / x2 + 3 / (x2 * x2) - 15 / (x2 * x2 * x2) + 105 / (x2 * x2 * x2 * x2));
}

double pg_mass_texpon(double z) {
  const double t = PG_T;
  double fz = HALF_PI2 + 0.5 * z * z;
  double b = sqrt(1.0 / t) * (t * z - 1), a = -sqrt(1.0 / t) * (t * z + 1);
  double x0 = log(fz) + fz * t;
  double xb = x0 - z + log_ndtr(b), xa = x0 + z + log_ndtr(a);
  double qdivp = 4 / M_PI * (exp(xb) + exp(xa));
  return 1.0 / (1.0 + qdivp); /* +inf -> 0 */
}

static double pg_coef(int n, double x) {
  double k = (n + 0.5) * M_PI;
  if (x > PG_T) return k * exp(-0.5 * k * k * x);
  if (x > 0) return exp(-1.5 * (log(0.5 * M_PI) + log(x)) + log(k) - 2.0 * (n + 0.5) * (n + 0.5) / x);
  return 0.0;
}

static double rtigauss(double z, pg_rng* r) {
  const double t = PG_T;
  if (1.0 / t > z) {
    for (;;) {
      double e1 = rng_expo(r), e2 = rng_expo(r);
      while (e1 * e1 > 2 * e2 / t) {
        e1 = rng_expo(r);
        e2 = rng_expo(r);
      }
      double x = 1 + e1 * t;
      x = t / (x * x);
      if (rng_unif(r) <= exp(-0.5 * z * z * x)) return x;
    }
  }
  double mu = 1.0 / z;
  for (;;) {
    double y = rng_norm(r);
    y *= y;
    double mu_y = mu * y;
    double x = mu + 0.5 * mu * mu_y - 0.5 * mu * sqrt(4 * mu_y + mu_y * mu_y);
    if (rng_unif(r) > mu / (mu + x)) x = mu * mu / x;
    if (x <= t) return x;
  }
}

static double pg1_draw(double psi, pg_rng* r) {
  double z = fabs(psi) * 0.5, fz = HALF_PI2 + 0.5 * z * z;
  double mass = pg_mass_texpon(z);
  for (;;) {
    double x = (rng_unif(r) < mass) ? PG_T + rng_expo(r) / fz : rtigauss(z, r);
    double s = pg_coef(0, x), y = rng_unif(r) * s;
    for (int n = 1;; ++n) {
      if (n & 1) {
        s -= pg_coef(n, x);
        if (y <= s) return 0.25 * x;
      } else {
        s += pg_coef(n, x);
        if (y > s) break;
      }
    }
  }
}

/* ---- C ABI used by ursm_oracle.py through ctypes ------------------------- */
typedef struct {
  int nthreads;
  pg_rng* gens;
} pg_pool;

pg_pool* pg_pool_create(int nthreads, const uint64_t* seeds) {
  pg_pool* p = (pg_pool*)malloc(sizeof(pg_pool));
  p->nthreads = nthreads;
  p->gens = (pg_rng*)malloc(sizeof(pg_rng) * nthreads);
  for (int t = 0; t < nthreads; ++t) rng_seed(&p->gens[t], seeds[t]);
  return p;
}
void pg_pool_destroy(pg_pool* p) {
  if (p) {
    free(p->gens);
    free(p);
  }
}
int pg_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
/* out[i] ~ PG(1, z[i]); entries split over the pool's generators */
void pg_drawv(pg_pool* p, int64_t n, const double* z, double* out) {
  int nt = p->nthreads;
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
  {
    int t = omp_get_thread_num(), T = omp_get_num_threads();
    pg_rng* g = &p->gens[t];
    int64_t lo = n * t / T, hi = n * (t + 1) / T;
    for (int64_t i = lo; i < hi; ++i) out[i] = pg1_draw(z[i], g);
  }
#else
  (void)nt;
  for (int64_t i = 0; i < n; ++i) out[i] = pg1_draw(z[i], &p->gens[0]);
#endif
}
