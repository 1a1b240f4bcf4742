// Bulk-sample E-step for sm_100a: e_step_gibbs.py:19-164.
//
// Data flow: the reference materialises Z (M x N x K fp64) and AW (N x M)
// every sweep; here neither exists.
//
// The deterministic mode (`mean_approx`, the command line's default) is three streaming passes
// over the counts, each free of cross-lane reductions because the thread owns the index its
// sums run over:
//   bk_mean_rows_kernel   lanes = samples (gene-major copy of X): sum_i A_ik X_ji / AW_ij
//                         -- the numerators of get_nmf_W (:157-164) and, after the W update,
//                         E[Z_jk] = W_jk x that sum (draw_Z_mean :149-155, update_suffStats :69-78)
//   bk_mean_cols_kernel   lanes = genes (X as given): E[Z_ik] = A_ik sum_j W_jk X_ji / AW_ij
// AW_ij is a K-term dot product in registers against a broadcast shared-memory row; X_ji / AW_ij
// is one MUFU reciprocal refined by two Newton steps in fp64 (1 ulp).  fp64 throughout (the
// reference is fp64 and the statistics are pinned to 1e-8): the passes are bound by the fp64
// pipe (2 K + 6 DFMA-class instructions per count), not by HBM.
// (draw_Z (:132-138), the Gibbs mode, is bk_split.cuh)
#include "common.cuh"
#include "samplers.cuh"
#include "sc_shard.h"
#include "bk_split.cuh"
#include "../../include/ursm_b200.h"

#include <algorithm>

namespace ursm {

constexpr int kMeanThreads = 128;
constexpr int kMeanBatch = 8;  // counts in flight per thread

// 1 / a for a normal positive double: MUFU reciprocal of the float, two Newton steps (1 ulp)
__device__ __forceinline__ double rcp_newton(double a) {
  double r = (double)URSM_RCP((float)a);
  r = fma(r, fma(-a, r, 1.0), r);
  r = fma(r, fma(-a, r, 1.0), r);
  return r;
}

// Both kernels are instantiated for KP = K rounded up to one of a few sizes; the shared-memory
// rows and the register copies are zero-padded to KP, so the K-loops carry no bounds tests.
//
// out[j][k] += (premul ? W_jk : 1) x sum_{i in the CTA's gene chunk} A_ik X_ji / AW_ij
// grid = (gene chunks, groups of kMeanThreads samples); thread = one sample
template <int KP>
__global__ void __launch_bounds__(kMeanThreads)
bk_mean_rows_kernel(const float* Xt, const double* A, const double* W, long long M, int Mpad, int N,
                    int K, int gchunk, int premul, double* out) {
  extern __shared__ double mean_smem[];  // [gchunk][KP] A rows of the chunk
  const int tid = threadIdx.x;
  const int g0 = blockIdx.x * gchunk, ng = min(gchunk, N - g0);
  for (int k = 0; k < KP; ++k)
    for (int gl = tid; gl < ng; gl += kMeanThreads)
      mean_smem[gl * KP + k] = k < K ? A[(size_t)k * N + g0 + gl] : 0.0;
  __syncthreads();
  const long long j = (long long)blockIdx.y * kMeanThreads + tid;
  if (j >= Mpad) return;
  const bool on = j < M;
  double w[KP], acc[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    w[k] = (on && k < K) ? W[j * K + k] : 0.0;
    acc[k] = 0.0;
  }
  if (!on) w[0] = 1.0;  // (padding lanes: x = 0, any positive AW)
  const float* xp = Xt + (size_t)g0 * Mpad + j;
  // counts are fetched kMeanBatch genes at a time, one batch ahead of the arithmetic
  float xn[kMeanBatch];
#pragma unroll
  for (int q = 0; q < kMeanBatch; ++q) xn[q] = (q < ng) ? __ldg(xp + (size_t)q * Mpad) : 0.f;
  for (int gb = 0; gb < ng; gb += kMeanBatch) {
    float xb[kMeanBatch];
#pragma unroll
    for (int q = 0; q < kMeanBatch; ++q) {  // this batch; the next one is requested before the arithmetic
      xb[q] = xn[q];
      const int gq = gb + kMeanBatch + q;
      xn[q] = (gq < ng) ? __ldg(xp + (size_t)gq * Mpad) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < kMeanBatch; ++q) {
      if (gb + q >= ng) break;
      const double* Ag = mean_smem + (gb + q) * KP;
      double a[KP];
#pragma unroll
      for (int k = 0; k < KP; ++k) a[k] = Ag[k];
      double aw0 = 0.0, aw1 = 0.0;
#pragma unroll
      for (int k = 0; k + 1 < KP; k += 2) {
        aw0 = fma(w[k], a[k], aw0);
        aw1 = fma(w[k + 1], a[k + 1], aw1);
      }
      if (KP & 1) aw0 = fma(w[KP - 1], a[KP - 1], aw0);
      const double r = (double)xb[q] * rcp_newton(aw0 + aw1);
#pragma unroll
      for (int k = 0; k < KP; ++k) acc[k] = fma(a[k], r, acc[k]);
    }
  }
  if (on) {
#pragma unroll
    for (int k = 0; k < KP; ++k)
      if (k < K) {
        const double v = premul ? w[k] * acc[k] : acc[k];
        if (v != 0.0) atomicAdd(&out[j * K + k], v);
      }
  }
}

// sumI[k][i] += scale x A_ik sum_{j in the CTA's sample chunk} W_jk X_ji / AW_ij
// grid = (groups of kMeanThreads genes, sample chunks); thread = one gene
template <int KP>
__global__ void __launch_bounds__(kMeanThreads)
bk_mean_cols_kernel(const float* X, const double* A, const double* W, long long M, int N, int K,
                    int jchunk, double scale, double* sumI) {
  extern __shared__ double mean_smem[];  // [jchunk][KP] W rows of the chunk
  const int tid = threadIdx.x;
  const long long j0 = (long long)blockIdx.y * jchunk;
  const int nj = (int)min((long long)jchunk, M - j0);
  for (int t = tid; t < nj * KP; t += kMeanThreads) {
    const int jl = t / KP, k = t - jl * KP;
    mean_smem[t] = k < K ? W[(j0 + jl) * K + k] : 0.0;
  }
  __syncthreads();
  const int i = blockIdx.x * kMeanThreads + tid;
  if (i >= N) return;
  double a[KP], acc[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    a[k] = (k < K) ? A[(size_t)k * N + i] : 0.0;
    acc[k] = 0.0;
  }
  const float* xp = X + (size_t)j0 * N + i;
  float xn[kMeanBatch];
#pragma unroll
  for (int q = 0; q < kMeanBatch; ++q) xn[q] = (q < nj) ? __ldg(xp + (size_t)q * N) : 0.f;
  for (int jb = 0; jb < nj; jb += kMeanBatch) {
    float xb[kMeanBatch];
#pragma unroll
    for (int q = 0; q < kMeanBatch; ++q) {
      xb[q] = xn[q];
      const int jq = jb + kMeanBatch + q;
      xn[q] = (jq < nj) ? __ldg(xp + (size_t)jq * N) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < kMeanBatch; ++q) {
      if (jb + q >= nj) break;
      const double* Wj = mean_smem + (jb + q) * KP;
      double w[KP];
#pragma unroll
      for (int k = 0; k < KP; ++k) w[k] = Wj[k];
      double aw0 = 0.0, aw1 = 0.0;
#pragma unroll
      for (int k = 0; k + 1 < KP; k += 2) {
        aw0 = fma(w[k], a[k], aw0);
        aw1 = fma(w[k + 1], a[k + 1], aw1);
      }
      if (KP & 1) aw0 = fma(w[KP - 1], a[KP - 1], aw0);
      const double r = (double)xb[q] * rcp_newton(aw0 + aw1);
#pragma unroll
      for (int k = 0; k < KP; ++k) acc[k] = fma(w[k], r, acc[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < KP; ++k)
    if (k < K) {
      const double v = a[k] * acc[k] * scale;
      if (v != 0.0) atomicAdd(&sumI[(size_t)k * N + i], v);
    }
}

// One Gibbs sweep's W step (:140-146) plus the per-sample part of update_suffStats (:69-78).
// One thread per (sample, type): W_j ~ Dir(alpha + sum_i Z_j.) is K independent gammas
// (Marsaglia-Tsang, fp64, each from its own Philox stream keyed on the global sample id and
// the type) that the sample's K threads then normalise through shared memory in a fixed
// order.  Also writes the fp32 copy of W in the split kernel's layout ([group][type][lane])
// and zeroes the buffer the split kernel accumulates into (and its hand-over counter).
constexpr int kWDrawThreads = 128;

__global__ void __launch_bounds__(kWDrawThreads)
bk_w_draw_kernel(double* W, float* W32g, const double* n_prev, double* n_next,
                 const double* alpha, long long M, int K, long long sample_offset,
                 unsigned long long key, unsigned int sweep, int prev_retained, int retained,
                 int draw, double inv_sample, double* eZjk, double* eLogW, double* eW, int* fail,
                 unsigned int* cold_count) {
  __shared__ double g_s[kWDrawThreads];
  if (cold_count && blockIdx.x == 0 && threadIdx.x == 0) *cold_count = 0u;  // this sweep's hand-over list
  const int per_block = kWDrawThreads / K;  // samples per CTA
  const int jl = threadIdx.x / K, k = threadIdx.x - jl * K;
  const long long j = (long long)blockIdx.x * per_block + jl;
  const bool on = jl < per_block && j < M;
  const size_t e = on ? (size_t)j * K + k : 0;
  double gmm = 0.0;
  if (on) {
    if (prev_retained) eZjk[e] += n_prev[e] * inv_sample;
    if (draw) {
      Philox rng(key, 0xfffffffeu - (uint32_t)k, (uint32_t)(sample_offset + j), kPurposeBulkW | sweep);
      bool failed = false;
      gmm = gamma_mt(alpha[k] + n_prev[e], rng, &failed);
      if (failed) atomicAdd(fail, 1);  // never seen; the host raises if it ever is
    }
  }
  if (draw) {
    g_s[threadIdx.x] = gmm;
    __syncthreads();
    if (on) {
      double s = 0.0;
      for (int q = 0; q < K; ++q) s += g_s[jl * K + q];
      W[e] = gmm / s;
    }
  }
  if (on) {
    const double w = W[e];
    if (W32g) W32g[((size_t)(j >> 5) * K + k) * 32 + (j & 31)] = (float)w;
    if (retained) {
      eW[e] += w * inv_sample;
      eLogW[e] += log(w) * inv_sample;
    }
    if (n_next) n_next[e] = 0.0;
  }
}

// after the last sweep: its sum_i Z joins E[Z_jk] if it was retained
__global__ void bk_gibbs_final_kernel(const double* n_last, double* eZjk, size_t n, double inv) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) eZjk[t] += n_last[t] * inv;
}

// A^T (fp64, [K][N]) -> fp32 by position (`order[pos]` = gene id, null: natural order), and the
// type with the largest A_ik of every position (the split chain draws it last)
__global__ void bk_cast_A_kernel(const double* A, int N, int K, const int* order, float* out,
                                 unsigned char* kmax) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= N) return;
  const int gene = order ? order[pos] : pos;
  int best = 0;
  double vbest = -1.0;
  for (int k = 0; k < K; ++k) {
    const double v = A[(size_t)k * N + gene];
    out[(size_t)k * N + pos] = (float)v;
    if (v > vbest) {
      vbest = v;
      best = k;
    }
  }
  if (kmax) kmax[pos] = (unsigned char)best;
}

// X [M][N] -> Xt [N][Mpad] (gene-major, zero padding; row of gene i = rank[i], null: i),
// 32 x 32 tiles through shared memory
__global__ void bk_transpose_kernel(const float* X, long long M, int N, int Mpad, const int* rank,
                                    float* Xt) {
  __shared__ float tile[32][33];
  const long long j0 = (long long)blockIdx.y * 32;
  const int i0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long j = j0 + r;
    const int i = i0 + threadIdx.x;
    tile[r][threadIdx.x] = (j < M && i < N) ? X[(size_t)j * N + i] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = i0 + r;
    const long long j = j0 + threadIdx.x;
    if (i < N && j < Mpad) Xt[(size_t)(rank ? rank[i] : i) * Mpad + j] = tile[threadIdx.x][r];
  }
}

// how many counts are 1024 or more (the lockstep split kernel hands those chains over)
__global__ void bk_count_deep_kernel(const float* X, size_t n, unsigned long long* out) {
  unsigned long long c = 0;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x)
    c += X[t] >= (float)URSM_BINOM_TABLE;
  c = __reduce_add_sync(0xffffffffu, (unsigned int)c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// W_kj <- W_kj * numer_jk + alpha_k - 1, renormalised over k (:159-164)
__global__ void bk_nmf_update_kernel(double* W, const double* numer, const double* alpha,
                                     long long M, int K) {
  long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  double s = 0.0;
  for (int k = 0; k < K; ++k) {
    double v = W[j * K + k] * numer[j * K + k] + alpha[k] - 1.0;
    W[j * K + k] = v;
    s += v;
  }
  for (int k = 0; k < K; ++k) W[j * K + k] /= s;
}

__global__ void bk_mean_finalize_kernel(const double* W, double* eLogW, double* eW, long long M,
                                        int K) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * K) return;
  eW[t] = W[t];
  eLogW[t] = log(W[t]);
}

// stats tail: sum_j E[log W_kj] (K values), sum_jk E[Z_jk] E[log W_kj], sampler failure count
__global__ void bk_tail_kernel(const double* eZjk, const double* eLogW, long long M, int K,
                               double* tail, const int* fail) {
  __shared__ double red[32];
  if (threadIdx.x == 0) tail[K + 1] = fail ? (double)*fail : 0.0;
  double cross = 0.0;
  for (long long t = threadIdx.x; t < M * K; t += blockDim.x) cross += eZjk[t] * eLogW[t];
  cross = block_sum(cross, red);
  if (threadIdx.x == 0) tail[K] = cross;
  for (int k = 0; k < K; ++k) {
    double s = 0.0;
    for (long long j = threadIdx.x; j < M; j += blockDim.x) s += eLogW[j * K + k];
    s = block_sum(s, red);
    if (threadIdx.x == 0) tail[k] = s;
  }
}

static int bk_mean_kp(int K) {
  static const int sizes[] = {2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32};
  for (int v : sizes)
    if (K <= v) return v;
  return 32;
}
#define URSM_MEAN_DISPATCH(CALL) \
  switch (bk_mean_kp(K)) {       \
    case 2: CALL(2); break;      \
    case 3: CALL(3); break;      \
    case 4: CALL(4); break;      \
    case 5: CALL(5); break;      \
    case 6: CALL(6); break;      \
    case 8: CALL(8); break;      \
    case 10: CALL(10); break;    \
    case 12: CALL(12); break;    \
    case 16: CALL(16); break;    \
    case 20: CALL(20); break;    \
    case 24: CALL(24); break;    \
    default: CALL(32); break;    \
  }

// sum_i A_ik X_ji / AW_ij into out[j][k] (times W_jk when `premul`)
static void bk_launch_mean_rows(BkShard* b, int premul, double* out, cudaStream_t st) {
  const int K = b->K, N = b->N;
  const unsigned gy = (unsigned)((b->MpadN + kMeanThreads - 1) / kMeanThreads);
  // gene chunks: a few waves of CTAs, A rows of a chunk within 32 KB of shared memory
  const int KP = bk_mean_kp(K);
  const int max_chunk = std::max(1, (32 * 1024) / (8 * KP));
  long long gx = std::max<long long>(1, (8LL * b->num_sms + gy - 1) / gy);
  int gchunk = (int)std::min<long long>(max_chunk, std::max<long long>(16, (N + gx - 1) / gx));
  gx = (N + gchunk - 1) / gchunk;
  const size_t smem = (size_t)gchunk * KP * sizeof(double);
#define URSM_ROWS(KM)                                                                              \
  bk_mean_rows_kernel<KM><<<dim3((unsigned)gx, gy), kMeanThreads, smem, st>>>(                     \
      b->XtN.p, b->A.p, b->W.p, b->M, b->MpadN, N, K, gchunk, premul, out)
  URSM_MEAN_DISPATCH(URSM_ROWS);
#undef URSM_ROWS
  URSM_LAUNCH_CHECK();
}

// scale x A_ik sum_j W_jk X_ji / AW_ij into sumI[k][i]
static void bk_launch_mean_cols(BkShard* b, double scale, double* sumI, cudaStream_t st) {
  const int K = b->K, N = b->N;
  const unsigned gx = (unsigned)((N + kMeanThreads - 1) / kMeanThreads);
  const int KP = bk_mean_kp(K);
  const int max_chunk = std::max(1, (32 * 1024) / (8 * KP));
  long long gy = std::max<long long>(1, (8LL * b->num_sms + gx - 1) / gx);
  int jchunk = (int)std::min<long long>(max_chunk, std::max<long long>(16, (b->M + gy - 1) / gy));
  gy = (b->M + jchunk - 1) / jchunk;
  URSM_REQUIRE(gy <= 65535, "bulk mean kernel: too many sample chunks");
  const size_t smem = (size_t)jchunk * KP * sizeof(double);
#define URSM_COLS(KM)                                                                              \
  bk_mean_cols_kernel<KM><<<dim3(gx, (unsigned)gy), kMeanThreads, smem, st>>>(                     \
      b->X, b->A.p, b->W.p, b->M, N, K, jchunk, scale, sumI)
  URSM_MEAN_DISPATCH(URSM_COLS);
#undef URSM_COLS
  URSM_LAUNCH_CHECK();
}

static void bk_time_begin(BkShard* b, cudaStream_t st) {
  if (!b->ev0) {
    URSM_CUDA(cudaEventCreate(&b->ev0));
    URSM_CUDA(cudaEventCreate(&b->ev1));
  }
  URSM_CUDA(cudaEventRecord(b->ev0, st));
}
static void bk_time_end(BkShard* b, cudaStream_t st) {
  URSM_CUDA(cudaEventRecord(b->ev1, st));
  b->timed = true;
}

__global__ void bk_colsum_kernel(const float* X, long long M, int N, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double s = 0.0;
  for (long long j = 0; j < M; ++j) s += (double)X[(size_t)j * N + i];
  out[i] = s;
}

// Inputs of the Gibbs split kernels that never change during a fit: the choice of kernel, the
// gene-major copy of X (by position for the lockstep kernel), the log2 k! table, the hand-over
// list, and the guard for the kernels' 32-bit accumulators.
static void bk_build_gibbs_inputs(BkShard* b) {
  const int N = b->N;
  std::vector<double> colsum(N);
  {
    // sum_j Z_jik <= the gene's total count and sum_i Z_jik <= the sample's total count
    DevBuf<double> cs, rs;
    cs.alloc(N);
    rs.alloc((size_t)b->M);
    bk_colsum_kernel<<<(N + 255) / 256, 256>>>(b->X, b->M, N, cs.p);
    URSM_LAUNCH_CHECK();
    launch_rowsum(b->X, b->M, N, rs.p, 0);
    std::vector<double> hr((size_t)b->M);
    URSM_CUDA(cudaMemcpy(colsum.data(), cs.p, N * sizeof(double), cudaMemcpyDeviceToHost));
    URSM_CUDA(cudaMemcpy(hr.data(), rs.p, hr.size() * sizeof(double), cudaMemcpyDeviceToHost));
    double worst = 0.0;
    for (double v : colsum) worst = std::max(worst, v);
    for (double v : hr) worst = std::max(worst, v);
    URSM_REQUIRE(worst < 4294967296.0, "bulk Gibbs kernel: a gene's or a sample's total count "
                                       "reaches 2^32 (32-bit accumulators)");
  }
  {
    std::vector<double> lf2(URSM_BINOM_TABLE);
    binom_fill_table(lf2.data());
    b->lf2.alloc(URSM_BINOM_TABLE);
    URSM_CUDA(cudaMemcpy(b->lf2.p, lf2.data(), lf2.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  // which split kernel: lockstep (+ hand-over list for the chains that reach BTRS territory) unless
  // a sizeable part of the counts is that deep, where the per-lane kernel does all of them
  unsigned long long n_deep = 0;
  {
    DevBuf<unsigned long long> d;
    d.alloc(1);
    d.zero();
    bk_count_deep_kernel<<<4 * b->num_sms, 256>>>(b->X, (size_t)b->M * N, d.p);
    URSM_LAUNCH_CHECK();
    URSM_CUDA(cudaMemcpy(&n_deep, d.p, sizeof n_deep, cudaMemcpyDeviceToHost));
  }
  const double cells = (double)b->M * N;
  b->lane_kernel = (double)n_deep > 0.05 * cells;
  if (const char* e = getenv("URSM_BK_SPLIT")) {  // "lane" / "lockstep": A/B runs and tests
    if (!strcmp(e, "lane")) b->lane_kernel = true;
    if (!strcmp(e, "lockstep")) b->lane_kernel = false;
  }
  const int ngroups = (int)((b->M + 31) / 32);
  b->Mpad = ngroups * 32;
  b->A32.alloc((size_t)b->K * N);
  b->W32g.alloc((size_t)ngroups * b->K * 32);
  b->W32g.zero();
  b->Xt.alloc((size_t)N * b->Mpad);
  DevBuf<int> rank;
  if (!b->lane_kernel) {
    // positions: genes by decreasing total count, dealt round-robin to the CTAs
    std::vector<int> order(N), rk(N);
    for (int i = 0; i < N; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return colsum[x] > colsum[y]; });
    for (int pos = 0; pos < N; ++pos) rk[order[pos]] = pos;
    b->order.alloc(N);
    rank.alloc(N);
    b->kmax.alloc(N);
    URSM_CUDA(cudaMemcpy(b->order.p, order.data(), N * sizeof(int), cudaMemcpyHostToDevice));
    URSM_CUDA(cudaMemcpy(rank.p, rk.data(), N * sizeof(int), cudaMemcpyHostToDevice));
    // every chain handed over starts at a count of 1024 or more; redraws are ~1e-7 of the draws
    const double cap = (double)n_deep + 4096.0 + 1e-5 * cells * b->K;
    URSM_REQUIRE(cap < 4.0e9, "bulk Gibbs kernel: hand-over list too large");
    b->cold_cap = (unsigned int)cap;
    b->cold_list.alloc(b->cold_cap);
    b->cold_count.alloc(1);
    b->cold_count.zero();
  }
  for (long long j0 = 0; j0 < b->Mpad; j0 += 32LL * 32768) {
    const long long rows = std::min<long long>(32LL * 32768, b->Mpad - j0);
    bk_transpose_kernel<<<dim3((N + 31) / 32, (unsigned)(rows / 32)), dim3(32, 8)>>>(
        b->X + (size_t)j0 * N, b->M - j0, N, b->Mpad, rank.p, b->Xt.p + j0);
    URSM_LAUNCH_CHECK();
  }
  URSM_CUDA(cudaDeviceSynchronize());
}

// Input of the mean-mode row pass that never changes during a fit: X gene-major (natural order)
static void bk_build_mean_inputs(BkShard* b) {
  b->MpadN = (int)((b->M + 31) / 32) * 32;
  b->XtN.alloc((size_t)b->N * b->MpadN);
  for (long long j0 = 0; j0 < b->MpadN; j0 += 32LL * 32768) {
    const long long rows = std::min<long long>(32LL * 32768, b->MpadN - j0);
    bk_transpose_kernel<<<dim3((b->N + 31) / 32, (unsigned)(rows / 32)), dim3(32, 8)>>>(
        b->X + (size_t)j0 * b->N, b->M - j0, b->N, b->MpadN, nullptr, b->XtN.p + j0);
    URSM_LAUNCH_CHECK();
  }
  URSM_CUDA(cudaDeviceSynchronize());
}

static void bk_finish_create(BkShard* b) {
  b->num_sms = device_sm_count();
  URSM_REQUIRE(b->K <= 32, "bulk kernels support K <= 32 cell types");
  const size_t MK = (size_t)b->M * b->K;
  b->A.alloc((size_t)b->K * b->N);
  b->alpha.alloc(b->K);
  b->W.alloc(MK);
  b->nA.alloc(MK);
  b->nB.alloc(MK);
  b->acc.alloc(MK);
  b->eZjk.alloc(MK);
  b->eLogW.alloc(MK);
  b->eW.alloc(MK);
  b->fail.alloc(1);
  b->nA.zero();
  b->nB.zero();
  URSM_CUDA(cudaDeviceSynchronize());
}

}  // namespace ursm

using namespace ursm;

extern "C" {

int ursm_bk_create(ursm_bk** out, const double* h_X, int64_t M, int32_t N, int32_t K,
                   int64_t sample_offset) {
  URSM_API_BEGIN
  URSM_REQUIRE(out && h_X && M > 0 && N > 0 && K > 0, "ursm_bk_create: bad argument");
  BkShard* b = new BkShard();
  b->M = M;
  b->N = N;
  b->K = K;
  b->sample_offset = sample_offset;
  try {
    const size_t n = (size_t)M * N;
    b->Xown.alloc(n);
    b->X = b->Xown.p;
    upload_f64_as_f32(h_X, b->Xown.p, n);
    bk_finish_create(b);
  } catch (...) {
    delete b;
    throw;
  }
  *out = reinterpret_cast<ursm_bk*>(b);
  URSM_API_END
}

int ursm_bk_create_dev(ursm_bk** out, const float* d_X, int64_t M, int32_t N, int32_t K,
                       int64_t sample_offset) {
  URSM_API_BEGIN
  URSM_REQUIRE(out && d_X && M > 0 && N > 0 && K > 0, "ursm_bk_create_dev: bad argument");
  BkShard* b = new BkShard();
  b->M = M;
  b->N = N;
  b->K = K;
  b->sample_offset = sample_offset;
  b->X = d_X;
  try {
    bk_finish_create(b);
  } catch (...) {
    delete b;
    throw;
  }
  *out = reinterpret_cast<ursm_bk*>(b);
  URSM_API_END
}

int ursm_bk_destroy(ursm_bk* bk) {
  URSM_API_BEGIN
  cudaDeviceSynchronize();  // the pool frees below are ordered on the default stream only
  delete reinterpret_cast<BkShard*>(bk);
  URSM_API_END
}

int ursm_bk_profile_sum(ursm_bk* bk, double* d_out, void* stream) {
  URSM_API_BEGIN
  BkShard* b = reinterpret_cast<BkShard*>(bk);
  URSM_REQUIRE(b && d_out, "ursm_bk_profile_sum: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  DevBuf<double> R;
  R.alloc((size_t)b->M);
  launch_rowsum(b->X, b->M, b->N, R.p, st);
  URSM_CUDA(cudaMemsetAsync(d_out, 0, (size_t)b->N * sizeof(double), st));
  launch_normalised_colsum(b->X, nullptr, nullptr, R.p, b->M, b->N, b->num_sms, d_out, st);
  URSM_CUDA(cudaStreamSynchronize(st));  // R is released on the default stream
  URSM_API_END
}

int64_t ursm_bk_stats_len(const ursm_bk* bk) {
  const BkShard* b = reinterpret_cast<const BkShard*>(bk);
  return (int64_t)b->N * b->K + b->K + 2;
}

int ursm_bk_kernel_ms(ursm_bk* bk, float* ms) {
  URSM_API_BEGIN
  BkShard* b = reinterpret_cast<BkShard*>(bk);
  URSM_REQUIRE(b && ms && b->timed, "ursm_bk_kernel_ms: no bulk E-step has been launched");
  URSM_CUDA(cudaEventSynchronize(b->ev1));
  URSM_CUDA(cudaEventElapsedTime(ms, b->ev0, b->ev1));
  URSM_API_END
}

int ursm_bk_set_params(ursm_bk* bk, const double* h_A, const double* h_alpha) {
  URSM_API_BEGIN
  BkShard* b = reinterpret_cast<BkShard*>(bk);
  URSM_REQUIRE(b && h_A && h_alpha, "ursm_bk_set_params: bad argument");
  std::vector<double> At((size_t)b->K * b->N);
  for (int i = 0; i < b->N; ++i)
    for (int k = 0; k < b->K; ++k) At[(size_t)k * b->N + i] = h_A[(size_t)i * b->K + k];
  URSM_CUDA(cudaMemcpy(b->A.p, At.data(), At.size() * sizeof(double), cudaMemcpyHostToDevice));
  URSM_CUDA(cudaMemcpy(b->alpha.p, h_alpha, b->K * sizeof(double), cudaMemcpyHostToDevice));
  b->has_params = true;
  URSM_API_END
}

int ursm_bk_set_state(ursm_bk* bk, const double* h_W, const double* h_Zjk) {
  URSM_API_BEGIN
  BkShard* b = reinterpret_cast<BkShard*>(bk);
  URSM_REQUIRE(b && h_W, "ursm_bk_set_state: bad argument");
  const size_t MK = (size_t)b->M * b->K;
  std::vector<double> Wt(MK);  // reference layout is K x M
  for (long long j = 0; j < b->M; ++j)
    for (int k = 0; k < b->K; ++k) Wt[j * b->K + k] = h_W[(size_t)k * b->M + j];
  URSM_CUDA(cudaMemcpy(b->W.p, Wt.data(), MK * sizeof(double), cudaMemcpyHostToDevice));
  if (h_Zjk)
    URSM_CUDA(cudaMemcpy(b->nA.p, h_Zjk, MK * sizeof(double), cudaMemcpyHostToDevice));
  else
    b->nA.zero();
  URSM_CUDA(cudaDeviceSynchronize());
  b->has_state = true;
  URSM_API_END
}

int ursm_bk_mean_step(ursm_bk* bk, double* d_stats, void* stream) {
  URSM_API_BEGIN
  BkShard* b = reinterpret_cast<BkShard*>(bk);
  URSM_REQUIRE(b && d_stats, "ursm_bk_mean_step: bad argument");
  URSM_REQUIRE(b->has_params && b->has_state, "ursm_bk_mean_step: params/state not set");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t MK = (size_t)b->M * b->K;
  const unsigned mb = (unsigned)((b->M + 127) / 128), mkb = (unsigned)((MK + 255) / 256);
  if (!b->XtN.p) bk_build_mean_inputs(b);  // once per fit: the gene-major copy of X
  bk_time_begin(b, st);
  URSM_CUDA(cudaMemsetAsync(d_stats, 0, (size_t)ursm_bk_stats_len(bk) * sizeof(double), st));
  b->acc.zero(st);
  b->eZjk.zero(st);
  bk_launch_mean_rows(b, 0, b->acc.p, st);  // numerators of get_nmf_W
  bk_nmf_update_kernel<<<mb, 128, 0, st>>>(b->W.p, b->acc.p, b->alpha.p, b->M, b->K);
  URSM_LAUNCH_CHECK();
  bk_launch_mean_rows(b, 1, b->eZjk.p, st);  // E[Z_jk] with the updated W
  bk_launch_mean_cols(b, 1.0, d_stats, st);  // E[Z_ik]
  bk_mean_finalize_kernel<<<mkb, 256, 0, st>>>(b->W.p, b->eLogW.p, b->eW.p, b->M, b->K);
  URSM_LAUNCH_CHECK();
  bk_tail_kernel<<<1, 1024, 0, st>>>(b->eZjk.p, b->eLogW.p, b->M, b->K,
                                     d_stats + (size_t)b->N * b->K, nullptr);
  URSM_LAUNCH_CHECK();
  bk_time_end(b, st);
  URSM_API_END
}

int ursm_bk_gibbs(ursm_bk* bk, int32_t burnin, int32_t sample, int32_t thin, uint64_t seed,
                  uint64_t em_iter, int32_t freeze, double* d_stats, void* stream) {
  URSM_API_BEGIN
  BkShard* b = reinterpret_cast<BkShard*>(bk);
  URSM_REQUIRE(b && d_stats, "ursm_bk_gibbs: bad argument");
  URSM_REQUIRE(b->has_params && b->has_state, "ursm_bk_gibbs: params/state not set");
  URSM_REQUIRE(burnin >= 0 && sample >= 1 && thin >= 1, "ursm_bk_gibbs: bad sweep counts");
  if (!b->Xt.p) bk_build_gibbs_inputs(b);  // once per fit, first Gibbs E-step only
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned long long key = seed + 0x9E3779B97F4A7C15ull * (em_iter + 1) + 0x5bd1e995ull;
  const double inv = 1.0 / sample;
  const size_t MK = (size_t)b->M * b->K;
  const BkSplitGeom g = bk_split_geometry(b->M, b->N, b->K, b->num_sms);
  URSM_REQUIRE(g.rows <= 65535, "bulk split kernel: too many sample groups");
  const size_t smem = b->lane_kernel ? g.smem_lane : g.smem;
  URSM_REQUIRE(smem <= device_smem_optin(), "bulk split kernel: too many cell types for shared memory");
  static size_t configured[2] = {0, 0};
  if (smem > 48 * 1024 && smem > configured[b->lane_kernel]) {
    if (b->lane_kernel)
      URSM_CUDA(cudaFuncSetAttribute(bk_split_lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    else
      URSM_CUDA(cudaFuncSetAttribute(bk_split_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    configured[b->lane_kernel] = smem;
  }
  const unsigned mb = (unsigned)((b->M + (kWDrawThreads / b->K) - 1) / (kWDrawThreads / b->K));
  bk_time_begin(b, st);
  URSM_CUDA(cudaMemsetAsync(d_stats, 0, (size_t)ursm_bk_stats_len(bk) * sizeof(double), st));
  b->eZjk.zero(st);
  b->eLogW.zero(st);
  b->eW.zero(st);
  b->fail.zero(st);
  bk_cast_A_kernel<<<(unsigned)((b->N + 255) / 256), 256, 0, st>>>(b->A.p, b->N, b->K, b->order.p, b->A32.p,
                                                               b->kmax.p);
  URSM_LAUNCH_CHECK();
  BkSplitArgs a;
  memset(&a, 0, sizeof a);
  a.Xt = b->Xt.p;
  a.A32 = b->A32.p;
  a.W32g = b->W32g.p;
  a.lf2 = b->lf2.p;
  a.order = b->order.p;
  a.kmax = b->kmax.p;
  a.cold_list = b->cold_list.p;
  a.cold_count = b->cold_count.p;
  a.cold_cap = b->cold_cap;
  a.fail = b->fail.p;
  a.sumI = d_stats;
  a.M = b->M;
  a.sample_offset = b->sample_offset;
  a.Mpad = b->Mpad;
  a.N = b->N;
  a.K = b->K;
  a.ngroups = g.ngroups;
  a.gpc = g.gpc;
  a.genes_per_cta = g.genes_per_cta;
  a.inv_sample = inv;
  a.key = key;
  a.rk = philox_keys(key);
  const bool freezeW = freeze & 1, freezeZ = freeze & 2;
  const int nsweeps = burnin + sample * thin;
  bool prev_ret = false;
  // nA holds sum_i Z of the previous sweep (the chain persists across E-steps, :99-100,136)
  for (int sweep = 0; sweep < nsweeps; ++sweep) {
    const bool retained = sweep >= burnin && ((sweep - burnin) % thin == 0);
    bk_w_draw_kernel<<<mb, kWDrawThreads, 0, st>>>(
        b->W.p, b->W32g.p, b->nA.p, b->nB.p, b->alpha.p, b->M, b->K, b->sample_offset, key,
        (unsigned)sweep, prev_ret ? 1 : 0, retained ? 1 : 0, freezeW ? 0 : 1, inv, b->eZjk.p,
        b->eLogW.p, b->eW.p, b->fail.p, b->cold_count.p);
    URSM_LAUNCH_CHECK();
    a.n_cur = b->nB.p;
    a.retained = retained ? 1 : 0;
    a.sweep = (unsigned)sweep;
    if (b->lane_kernel) {
      bk_split_lane_kernel<<<dim3(g.gx, g.rows), kSplitThreads, smem, st>>>(a);
      URSM_LAUNCH_CHECK();
    } else {
      bk_split_kernel<false><<<dim3(g.gx, g.rows), kSplitThreads, smem, st>>>(a);
      URSM_LAUNCH_CHECK();
      // grid for the expected list (its capacity when counts reach 1024), early exit beyond it
      bk_split_cold_kernel<<<(b->cold_cap + 127) / 128, 128, 0, st>>>(a);
      URSM_LAUNCH_CHECK();
    }
    if (!freezeZ) std::swap(b->nA.p, b->nB.p);  // (frozen: the sums W is drawn from stay fixed)
    prev_ret = retained;
  }
  if (prev_ret) {  // the last sweep's sum_i Z
    bk_gibbs_final_kernel<<<(unsigned)((MK + 255) / 256), 256, 0, st>>>(b->nA.p, b->eZjk.p, MK, inv);
    URSM_LAUNCH_CHECK();
  }
  bk_tail_kernel<<<1, 1024, 0, st>>>(b->eZjk.p, b->eLogW.p, b->M, b->K,
                                     d_stats + (size_t)b->N * b->K, b->fail.p);
  URSM_LAUNCH_CHECK();
  bk_time_end(b, st);
  URSM_API_END
}

static void copy_MK_to_KM(const BkShard* b, const double* d_src, double* h_dst, bool transpose) {
  const size_t MK = (size_t)b->M * b->K;
  std::vector<double> tmp(MK);
  URSM_CUDA(cudaMemcpy(tmp.data(), d_src, MK * sizeof(double), cudaMemcpyDeviceToHost));
  if (!transpose) {
    memcpy(h_dst, tmp.data(), MK * sizeof(double));
    return;
  }
  for (long long j = 0; j < b->M; ++j)
    for (int k = 0; k < b->K; ++k) h_dst[(size_t)k * b->M + j] = tmp[j * b->K + k];
}

int ursm_bk_get_sample_stats(ursm_bk* bk, double* h_exp_Zjk, double* h_exp_logW,
                             double* h_exp_W) {
  URSM_API_BEGIN
  BkShard* b = reinterpret_cast<BkShard*>(bk);
  URSM_CUDA(cudaDeviceSynchronize());
  if (h_exp_Zjk) copy_MK_to_KM(b, b->eZjk.p, h_exp_Zjk, false);  // M x K
  if (h_exp_logW) copy_MK_to_KM(b, b->eLogW.p, h_exp_logW, true);  // K x M
  if (h_exp_W) copy_MK_to_KM(b, b->eW.p, h_exp_W, true);          // K x M
  URSM_API_END
}

int ursm_bk_get_W(ursm_bk* bk, double* h_W) {
  URSM_API_BEGIN
  BkShard* b = reinterpret_cast<BkShard*>(bk);
  URSM_CUDA(cudaDeviceSynchronize());
  copy_MK_to_KM(b, b->W.p, h_W, true);
  URSM_API_END
}

}  // extern "C"
