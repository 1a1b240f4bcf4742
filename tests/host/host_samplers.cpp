// Host build of ursm_b200/csrc/samplers.cuh (the same source the kernels inline),
// exported with a C ABI so tests/test_samplers_host.py can check the sampler
// arithmetic on the CPU.  Test infrastructure only.
#include <stdint.h>
#include "../../ursm_b200/csrc/samplers.cuh"

using namespace ursm;

extern "C" {
void h_philox(uint64_t key, uint32_t x, uint32_t y, uint32_t z, uint32_t* out, int nblocks) {
  Philox g(key, x, y, z);
  for (int i = 0; i < 4 * nblocks; ++i) out[i] = g.next();
}
double h_pg_mass(double z) { return pg_tilt((float)(2.0 * z)).mass_right; }
double h_ndtri_sq(double v) { return ndtri_sq((float)v); }
void h_pg_draws(double psi, int n, uint64_t seed, double* out) {
  for (int i = 0; i < n; ++i) {
    Philox g(seed, (uint32_t)i, 7u, 0u);
    out[i] = pg1_draw((float)psi, g);
  }
}
double h_s_threshold(double psi, double u, double a, double R) {
  return s_threshold((float)psi, (float)u, (float)a, (float)R);
}
void h_binomial(int n, double p, int count, uint64_t seed, int* out) {
  for (int i = 0; i < count; ++i) {
    Philox g(seed, (uint32_t)i, 1u, 0u);
    out[i] = binomial(n, p, g);
  }
}
void h_binomial_f32(int n, double p, int count, uint64_t seed, int* out) {
  static double lf2[URSM_BINOM_TABLE];
  static bool ready = false;
  if (!ready) {
    binom_fill_table(lf2);
    ready = true;
  }
  for (int i = 0; i < count; ++i) {
    Philox g(seed, (uint32_t)i, 5u, 0u);
    out[i] = binomial_f32(n, (float)p, (float)(1.0 - p), g, lf2);
  }
}
// how many uniforms out of `count` fall beyond the fp32 mass sum (-1 returns) for Bin(n, r)
int h_binom_mode_misses(int n, double r, int count, uint64_t seed) {
  static double lf2[URSM_BINOM_TABLE];
  binom_fill_table(lf2);
  Philox g(seed, 0u, 6u, 0u);
  int miss = 0;
  for (int i = 0; i < count; ++i)
    if (binom_mode_try(n, (float)r, (float)(1.0 - r), g.uniform(), lf2) < 0) ++miss;
  return miss;
}
void h_gamma(double shape, int count, uint64_t seed, double* out) {
  for (int i = 0; i < count; ++i) {
    Philox g(seed, (uint32_t)i, 2u, 0u);
    out[i] = gamma_mt(shape, g);
  }
}
void h_normal(int count, uint64_t seed, double* out) {
  for (int i = 0; i < count; ++i) {
    Philox g(seed, (uint32_t)i, 3u, 0u);
    normal_pair(g, out[2 * i], out[2 * i + 1]);
  }
}
void h_uniform(int count, uint64_t seed, double* out) {
  Philox g(seed, 0u, 0u, 0u);
  for (int i = 0; i < count; ++i) out[i] = g.uniform();
}
// replay helpers for the GPU parity tests: the uniforms / normals the production
// single-cell kernel consumes for (cell, sweep), and the float32 threshold
void h_gene_uniforms(uint64_t key, uint32_t cell, uint32_t sweep, int N, double* out) {
  for (int i = 0; i < N; ++i) {
    uint32_t r0, r1, r2, r3;  // the S uniform is word 3 of the gene's first attempt block
    philox_block(key, (uint32_t)i, cell, kPurposeGene | sweep, 0u, r0, r1, r2, r3);
    out[i] = u01(r3);
  }
}
void h_cell_normals(uint64_t key, uint32_t cell, uint32_t sweep, double* out2) {
  Philox g(key, 0xffffffffu, cell, kPurposeCell | sweep);
  normal_pair(g, out2[0], out2[1]);
}
void h_init_bits(uint64_t key, uint32_t cell, int N, int* out) {
  for (int i = 0; i < N; ++i) {
    Philox g(key, (uint32_t)i, cell, kPurposeInit);
    out[i] = (int)(g.next() >> 31);
  }
}
uint64_t h_sc_key(uint64_t seed, uint64_t em_iter) {
  return seed + 0x9E3779B97F4A7C15ull * (em_iter + 1);
}
}
