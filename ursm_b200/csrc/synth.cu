// Synthetic inputs generated on the device with the distributions of the reference's demo
// generator (demo/demo_simulate_data.py:45-114), so that BASELINE.json's
// configs never have to exist as CSV text or cross PCIe:
//   bulk   W_j ~ Dir(alpha), depth_j ~ Poisson(d N), X_j ~ Mult(depth_j, A W_j)          (:45-69)
//   cells  kappa_l ~ N(kappa, sd), tau_l ~ N(t N, f t N), S_li ~ Bern(expit(kappa_l + tau_l A[i,G_l])),
//          depth_l ~ Poisson(d N), Y_l ~ Mult(depth_l, normalise(A[:,G_l] S_l))          (:72-114)
// A multinomial with a Poisson total is a vector of independent Poissons, which is how the
// counts are drawn here: exactly the same joint law in one pass.  Same distributions as the
// reference generator, not the same stream: every draw comes from a counter-based Philox
// stream keyed on (seed, gene, GLOBAL row), so a shard of rows is the same whatever rank draws it.
#include "common.cuh"
#include "samplers.cuh"
#include "../../include/ursm_b200.h"

namespace ursm {

enum : uint32_t {
  kPurposeSynW = 8u << 28,     // per sample: Dirichlet mixing weights
  kPurposeSynX = 9u << 28,     // per (sample, gene): bulk count
  kPurposeSynCell = 10u << 28, // per cell: kappa, tau
  kPurposeSynS = 11u << 28,    // per (cell, gene): dropout indicator
  kPurposeSynY = 12u << 28,    // per (cell, gene): single-cell count
};

// Poisson(lam) in double precision: inversion from 0 below 12, PTRS (Hoermann 1993, the
// transformed rejection numpy also uses) above.
template <class RNG>
__device__ double poisson_draw(double lam, RNG& rng) {
  if (!(lam > 0.0)) return 0.0;
  if (lam < 12.0) {
    double p = exp(-lam), cdf = p;
    const double u = rng.uniform_d();
    double x = 0.0;
    while (u > cdf && p > 1e-300) {
      x += 1.0;
      p *= lam / x;
      cdf += p;
    }
    return x;
  }
  const double slam = sqrt(lam), loglam = log(lam);
  const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
  const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
  for (int it = 0; it < 1000; ++it) {
    const double U = rng.uniform_d() - 0.5, V = rng.uniform_d();
    const double us = 0.5 - fabs(U);
    const double k = floor((2.0 * a / us + b) * U + lam + 0.43);
    if (us >= 0.07 && V <= vr) return k;
    if (k < 0.0 || (us < 0.013 && V > us)) continue;
    if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + k * loglam - lgamma(k + 1.0)) return k;
  }
  return floor(lam);  // (1000 rejections in a row: never)
}

// W_j ~ Dir(alpha): K gammas per sample, one thread per sample
__global__ void syn_w_kernel(const double* alpha, long long M, int K, long long row_offset,
                             unsigned long long key, double* W) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  Philox rng(key, 0u, (uint32_t)(row_offset + j), kPurposeSynW);
  double s = 0.0;
  for (int k = 0; k < K; ++k) {
    const double g = gamma_mt(alpha[k], rng);
    W[j * K + k] = g;
    s += g;
  }
  for (int k = 0; k < K; ++k) W[j * K + k] /= s;
}

// X_ji ~ Poisson(scale * sum_k W_jk A_ik); grid = (gene blocks, samples)
__global__ void syn_bulk_kernel(const double* A /*[K][N]*/, const double* W /*[M][K]*/, long long M,
                                int N, int K, double scale, long long row_offset,
                                unsigned long long key, float* X) {
  extern __shared__ double wj[];
  const long long j = blockIdx.y;
  for (int k = threadIdx.x; k < K; k += blockDim.x) wj[k] = W[j * K + k];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double aw = 0.0;
  for (int k = 0; k < K; ++k) aw = fma(wj[k], A[(size_t)k * N + i], aw);
  Philox rng(key, (uint32_t)i, (uint32_t)(row_offset + j), kPurposeSynX);
  X[(size_t)j * N + i] = (float)poisson_draw(scale * aw, rng);
}

// one CTA per cell
constexpr int kSynThreads = 256;
__global__ void __launch_bounds__(kSynThreads)
syn_cells_kernel(const double* A /*[K][N]*/, const int* G, long long L, int N, double depth,
                 double kappa, double kappa_sd, double tau_mean, double tau_sd, long long row_offset,
                 unsigned long long key, float* Y, unsigned char* S, double* kap_out, double* tau_out) {
  __shared__ double red[32];
  __shared__ double cell[2];
  __shared__ int any_s;
  const long long l = blockIdx.x;
  const int tid = threadIdx.x;
  const uint32_t gl = (uint32_t)(row_offset + l);
  if (tid == 0) {
    Philox rng(key, 0u, gl, kPurposeSynCell);
    double n0, n1;
    normal_pair(rng, n0, n1);
    cell[0] = kappa + kappa_sd * n0;
    cell[1] = tau_mean + tau_sd * n1;
    kap_out[l] = cell[0];
    tau_out[l] = cell[1];
    any_s = 0;
  }
  __syncthreads();
  const double kap = cell[0], tau = cell[1];
  const double* a = A + (size_t)G[l] * N;
  // dropout indicators and the mass of the genes that are "on"
  double tot = 0.0;
  for (int i = tid; i < N; i += kSynThreads) {
    Philox rng(key, (uint32_t)i, gl, kPurposeSynS);
    const double prob = 1.0 / (1.0 + exp(-(kap + tau * a[i])));
    const unsigned char s = rng.uniform_d() < prob ? 1 : 0;
    S[(size_t)l * N + i] = s;
    if (s) tot += a[i];
  }
  tot = block_sum(tot, red);
  const bool dead = !(tot > 0.0);  // no expressed gene at all: fall back to the profile itself
  if (dead) {
    double t2 = 0.0;
    for (int i = tid; i < N; i += kSynThreads) t2 += a[i];
    tot = block_sum(t2, red);
  }
  const double scale = depth / tot;
  int mine = 0;
  for (int i = tid; i < N; i += kSynThreads) {
    const double p = (dead || S[(size_t)l * N + i]) ? a[i] : 0.0;
    Philox rng(key, (uint32_t)i, gl, kPurposeSynY);
    const double y = poisson_draw(scale * p, rng);
    Y[(size_t)l * N + i] = (float)y;
    mine |= y > 0.0;
  }
  if (mine) any_s = 1;
  __syncthreads();
  if (!any_s) {
    // every cell gets at least one read (the reference divides by the depth, m_step.py:125):
    // one read on the most probable gene
    double best = -1.0;
    int arg = 0;
    for (int i = tid; i < N; i += kSynThreads) {
      const double p = (dead || S[(size_t)l * N + i]) ? a[i] : 0.0;
      if (p > best) {
        best = p;
        arg = i;
      }
    }
    __shared__ double bv[kSynThreads];
    __shared__ int bi[kSynThreads];
    bv[tid] = best;
    bi[tid] = arg;
    __syncthreads();
    if (tid == 0) {
      for (int t = 1; t < kSynThreads; ++t)
        if (bv[t] > bv[0] || (bv[t] == bv[0] && bi[t] < bi[0])) {
          bv[0] = bv[t];
          bi[0] = bi[t];
        }
      Y[(size_t)l * N + bi[0]] = 1.0f;
    }
  }
}

static void upload_A_type_major(const double* h_A, int N, int K, DevBuf<double>& A) {
  std::vector<double> At((size_t)K * N);
  for (int i = 0; i < N; ++i)
    for (int k = 0; k < K; ++k) At[(size_t)k * N + i] = h_A[(size_t)i * K + k];
  A.alloc(At.size());
  URSM_CUDA(cudaMemcpy(A.p, At.data(), At.size() * sizeof(double), cudaMemcpyHostToDevice));
}

}  // namespace ursm

using namespace ursm;

extern "C" {

int ursm_synth_bulk(const double* h_A, int32_t N, int32_t K, int64_t M, const double* h_alpha,
                    double depth_per_gene, uint64_t seed, int64_t row_offset, float* d_X, double* h_W) {
  URSM_API_BEGIN
  URSM_REQUIRE(h_A && h_alpha && d_X && N > 0 && K > 0 && M > 0 && M <= 65535 * 64LL,
               "ursm_synth_bulk: bad argument");
  DevBuf<double> A, alpha, W;
  upload_A_type_major(h_A, N, K, A);
  alpha.alloc(K);
  URSM_CUDA(cudaMemcpy(alpha.p, h_alpha, K * sizeof(double), cudaMemcpyHostToDevice));
  W.alloc((size_t)M * K);
  const unsigned long long key = seed * 0x9E3779B97F4A7C15ull + 0x5851F42D4C957F2Dull;
  syn_w_kernel<<<(unsigned)((M + 127) / 128), 128>>>(alpha.p, M, K, row_offset, key, W.p);
  URSM_LAUNCH_CHECK();
  for (long long j0 = 0; j0 < M; j0 += 65535) {  // grid.y limit
    const long long rows = std::min<long long>(65535, M - j0);
    syn_bulk_kernel<<<dim3((N + 255) / 256, (unsigned)rows), 256, K * sizeof(double)>>>(
        A.p, W.p + j0 * K, rows, N, K, depth_per_gene * N, row_offset + j0, key, d_X + (size_t)j0 * N);
    URSM_LAUNCH_CHECK();
  }
  if (h_W) URSM_CUDA(cudaMemcpy(h_W, W.p, (size_t)M * K * sizeof(double), cudaMemcpyDeviceToHost));
  URSM_CUDA(cudaDeviceSynchronize());
  URSM_API_END
}

int ursm_synth_cells(const double* h_A, int32_t N, int32_t K, const int64_t* h_G, int64_t L,
                     uint64_t seed, int64_t row_offset, double depth_per_gene, double kappa,
                     double kappa_sd, double tau_scale, double tau_sd_frac, float* d_Y, uint8_t* d_S,
                     double* h_kappa, double* h_tau) {
  URSM_API_BEGIN
  URSM_REQUIRE(h_A && h_G && d_Y && d_S && N > 0 && K > 0 && L > 0 && L < (1ll << 31),
               "ursm_synth_cells: bad argument");
  DevBuf<double> A, kap, tau;
  DevBuf<int> G;
  upload_A_type_major(h_A, N, K, A);
  std::vector<int> g(L);
  for (long long l = 0; l < L; ++l) {
    URSM_REQUIRE(h_G[l] >= 0 && h_G[l] < K, "ursm_synth_cells: cell type outside {0..K-1}");
    g[l] = (int)h_G[l];
  }
  G.alloc(L);
  URSM_CUDA(cudaMemcpy(G.p, g.data(), L * sizeof(int), cudaMemcpyHostToDevice));
  kap.alloc(L);
  tau.alloc(L);
  const unsigned long long key = seed * 0x9E3779B97F4A7C15ull + 0x14057B7EF767814Full;
  const double tau_mean = tau_scale * N;
  syn_cells_kernel<<<(unsigned)L, kSynThreads>>>(A.p, G.p, L, N, depth_per_gene * N, kappa, kappa_sd,
                                                tau_mean, tau_sd_frac * tau_mean, row_offset, key, d_Y,
                                                d_S, kap.p, tau.p);
  URSM_LAUNCH_CHECK();
  if (h_kappa) URSM_CUDA(cudaMemcpy(h_kappa, kap.p, L * sizeof(double), cudaMemcpyDeviceToHost));
  if (h_tau) URSM_CUDA(cudaMemcpy(h_tau, tau.p, L * sizeof(double), cudaMemcpyDeviceToHost));
  URSM_CUDA(cudaDeviceSynchronize());
  URSM_API_END
}

}  // extern "C"
