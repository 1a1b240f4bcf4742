// Bulk-sample E-step for sm_100a: e_step_gibbs.py:19-164.
//
// Data flow: the reference materialises Z (M x N x K fp64) and
// AW (N x M) every sweep; here neither exists.  One "tile" kernel owns a tile of
// genes (A rows live in registers for the whole kernel) and walks a chunk of bulk
// samples; AW_ij is a K-term dot product in registers, the K-way split of X_ji is
// consumed on the spot into
//    sum_i Z_jik  (warp reduction -> one atomic per warp, sample, type)   and
//    sum_j Z_jik  (register accumulators, flushed once per CTA),
// so HBM sees X exactly once per pass (4 B per (sample, gene)).
//   mode 0: numerators of get_nmf_W (:157-164)
//   mode 1: draw_Z_mean (:149-155) + update_suffStats(1) (:69-78)
// (draw_Z (:132-138), the Gibbs mode, is bk_split_kernel below)
#include "common.cuh"
#include "samplers.cuh"
#include "sc_shard.h"
#include "../../include/ursm_b200.h"

#include <algorithm>

namespace ursm {

constexpr int kBkThreads = 256;

struct BkTileArgs {
  const float* X;      // [M][N]
  const double* A;     // [K][N]
  const double* W;     // [M][K]
  const int* perm;     // [N] split kernel: sorted position -> gene (X is column-sorted)
  long long M;
  int N, K;
  long long sample_offset;
  int chunk;           // samples per grid.y slice
  double* sumJ;        // [M][K]  mode 0: nmf numerators; 1: exp_Zjk; 2: n_next
  double* sumI;        // [K][N]  modes 1, 2 (retained): exp_Zik accumulator
  double scaleI;       // 1/sample
  int accumulate_I;    // mode 2: this sweep is retained
  unsigned long long key;
  PhiloxKeys rk;       // round keys of `key` (constant-bank operands of the Philox rounds)
  unsigned int sweep;
};

// Sums P (a power of two <= 32) values per lane over the 32 lanes of a warp.  Exchange
// `off` = 16, 8, ...: the lane with that bit set keeps the upper half of its live values
// and receives the partner's upper half, the other lane the lower halves.  Returns the
// index of the value whose total ends up in v[0] of this lane (-1 for lanes that hold a
// duplicate).
template <int P>
__device__ __forceinline__ int warp_transpose_reduce(double (&v)[P], int lane) {
  int idx = 0;
  int off = 16;
#pragma unroll
  for (int c = P / 2; c >= 1; c /= 2) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < c; ++i) {
      const double send = up ? v[i] : v[i + c];
      const double recv = __shfl_xor_sync(0xffffffffu, send, off);
      v[i] = (up ? v[i + c] : v[i]) + recv;
    }
    if (up) idx += c;
    off >>= 1;
  }
  bool owner = true;
  for (; off >= 1; off >>= 1) {  // P < 32: the remaining lanes hold partial sums of the same index
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
    if (lane & off) owner = false;
  }
  return owner ? idx : -1;
}

template <int KMAX, int MODE>
__global__ void __launch_bounds__(kBkThreads) bk_tile_kernel(BkTileArgs a) {
  constexpr int P = KMAX <= 4 ? 4 : (KMAX <= 8 ? 8 : (KMAX <= 16 ? 16 : 32));
  extern __shared__ double bk_smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int K = a.K, N = a.N;
  const long long j0 = (long long)blockIdx.y * a.chunk;
  const long long j1 = min(a.M, j0 + a.chunk);
  const int nloc = (int)(j1 - j0);
  double* Wsm = bk_smem;             // [nloc][K] W rows of this CTA's samples
  double* Jsm = bk_smem + nloc * K;  // [nloc][K] CTA-level sums over genes
  // the thread's A row and its sum_j accumulators live in private shared-memory columns
  // (registers are needed for the K-wide butterfly below)
  double* Ar = bk_smem + 2 * a.chunk * K + tid;   // [K][T]
  double* accI = Ar + K * kBkThreads;              // [K][T]
  for (int t = tid; t < nloc * K; t += kBkThreads) {
    Wsm[t] = a.W[j0 * K + t];
    Jsm[t] = 0.0;
  }
  const int gi = blockIdx.x * kBkThreads + tid;
  const bool in = gi < N;
  for (int k = 0; k < K; ++k) {
    Ar[k * kBkThreads] = in ? a.A[(size_t)k * N + gi] : 0.0;
    accI[k * kBkThreads] = 0.0;
  }
  __syncthreads();

  for (int jl = 0; jl < nloc; ++jl) {
    const long long j = j0 + jl;
    const double* Wj = Wsm + jl * K;
    const float xf = in ? a.X[(size_t)j * N + gi] : 0.f;
    const bool active = xf != 0.f;  // zero counts contribute nothing in any mode
    double aw = 0.0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) aw += Wj[k] * Ar[k * kBkThreads];
    const double r = active ? (double)xf / aw : 0.0;
    double v[P];
#pragma unroll
    for (int k = 0; k < P; ++k) {
      v[k] = 0.0;
      if (k < KMAX && k < K) {
        if (MODE == 0) {
          v[k] = Ar[k * kBkThreads] * r;
        } else {
          v[k] = Wj[k] * Ar[k * kBkThreads] * r;
          accI[k * kBkThreads] += v[k];
        }
      }
    }
    // sum over the 32 genes of the warp, all K types at once: a butterfly that halves
    // the number of live values per lane at every exchange (P - 1 + 5 - log2 P shuffles
    // instead of 5 K); lane `l` ends up with the total of type tr_index(l)
    const int kk = warp_transpose_reduce<P>(v, lane);
    if (kk >= 0 && kk < K && v[0] != 0.0) atomicAdd(&Jsm[jl * K + kk], v[0]);
  }
  __syncthreads();
  for (int t = tid; t < nloc * K; t += kBkThreads)
    if (Jsm[t] != 0.0) atomicAdd(&a.sumJ[j0 * K + t], Jsm[t]);
  if (in && MODE == 1) {
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k >= K) break;
      const double t = accI[k * kBkThreads];
      if (t != 0.0) atomicAdd(&a.sumI[(size_t)k * N + gi], t * a.scaleI);
    }
  }
}

// ---------------------------------------------------------------------------
// draw_Z (:132-138) for one Gibbs sweep: Z_ji. ~ Mult(X_ji, W_.j A_i. / AW_ij) as a
// chain of K-1 conditional binomials (what numpy.random.multinomial does).
// A thread owns one gene (its A row in registers) and walks the CTA's chunk of bulk
// samples; the 32 lanes of a warp work on the same (sample, type) item in lock step, so
// set-up, RNG and the consumption of Z are full-width and sum_i Z_jik is a warp
// reduction + one shared-memory atomic per warp.  The data-dependent part -- fp32
// inversion (n r <= 40) or BTRS (samplers.cuh::BinomState) -- diverges only by the
// spread of n r inside the warp, which is small because X is stored with its columns
// SORTED by total count: the 32 genes of a warp have similar depth and sit in the same
// regime.  Z is never stored: sum_j Z_jik accumulates in registers and is flushed once
// per CTA.  Probabilities are ratios of positive masses (the mass of the types still to
// come is tracked in fp64); counts are exact integers; the RNG is keyed on the ORIGINAL
// gene id and the global sample id, so results depend neither on the sort nor on the
// sharding.
// ---------------------------------------------------------------------------
constexpr int kSplitThreads = 256;

__global__ void __launch_bounds__(kSplitThreads, 4) bk_split_kernel(BkTileArgs a) {
  extern __shared__ __align__(16) unsigned char split_smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int K = a.K, N = a.N;
  const long long j0 = (long long)blockIdx.y * a.chunk;
  const long long j1 = min(a.M, j0 + a.chunk);
  const int nloc = (int)(j1 - j0);
  float* Wsm = reinterpret_cast<float*>(split_smem);                      // [chunk][K]
  unsigned int* Jsm = reinterpret_cast<unsigned int*>(Wsm + a.chunk * K);  // [chunk][K]
  float* Ar = reinterpret_cast<float*>(Jsm + a.chunk * K) + tid;            // [K][T], own column
  unsigned int* accI = reinterpret_cast<unsigned int*>(Ar - tid + K * kSplitThreads) + tid;
  for (int t = tid; t < nloc * K; t += kSplitThreads) {
    Wsm[t] = (float)a.W[j0 * K + t];
    Jsm[t] = 0u;
  }
  const int gi = blockIdx.x * kSplitThreads + tid;  // position in the sorted column order
  const bool in = gi < N;
  const int gene = in ? a.perm[gi] : 0;             // original gene id
  for (int k = 0; k < K; ++k) {
    Ar[k * kSplitThreads] = in ? (float)a.A[(size_t)k * N + gene] : 0.f;
    accI[k * kSplitThreads] = 0u;
  }
  __syncthreads();
  const uint32_t ctr2 = kPurposeBulkZ | a.sweep;

  for (int jl = 0; jl < nloc; ++jl) {
    const float* Wj = Wsm + jl * K;
    const float xf = in ? a.X[(size_t)(j0 + jl) * N + gi] : 0.f;
    int rem = (int)xf;
    if (__all_sync(0xffffffffu, rem == 0)) continue;  // nothing to split in this warp
    const uint32_t jg = (uint32_t)(a.sample_offset + j0 + jl);
    // AW_ij and the mass of the types still to come (fp64: exact enough down to the last type)
    double rest = 0.0;
    for (int k = 0; k < K; ++k) rest += (double)(Wj[k] * Ar[k * kSplitThreads]);
    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
      int z;
      if (k == K - 1) {
        z = rem;  // the last type takes what is left
      } else {
        if ((k & 3) == 0) philox_block(a.rk, (uint32_t)gene, jg, ctr2, (uint32_t)(k >> 2), w0, w1, w2, w3);
        const uint32_t wk = (k & 3) == 0 ? w0 : ((k & 3) == 1 ? w1 : ((k & 3) == 2 ? w2 : w3));
        const float qk = Wj[k] * Ar[k * kSplitThreads];
        const double after = fmax(rest - (double)qk, 0.0);
        const float inv = URSM_DIV(1.0f, (float)rest);
        BinomState b;
        // P(type k | not yet assigned) and its complement, each formed from its own mass
        binom_begin(b, rem, qk * inv, (float)after * inv, u01(wk));
        rest = after;
        int extra = 0;  // further Philox blocks of this item: BTRS attempts, inversion restarts
        while (b.mode != kBinDone) {
          if (b.mode == kBinInv) {
            binom_inv_chunk4(b);
          } else {
            uint32_t e0, e1, e2, e3;
            philox_block(a.rk, (uint32_t)gene, jg, ctr2, 0x100u + (uint32_t)(k << 8) + (uint32_t)extra,
                         e0, e1, e2, e3);
            ++extra;
            if (b.mode == kBinBtrs) {
              binom_btrs_attempt(b, u01(e0), u01(e1));
              if (b.mode != kBinDone) binom_btrs_attempt(b, u01(e2), u01(e3));
            } else {  // kBinRetry
              binom_inv_start(b, u01(e0));
            }
          }
        }
        z = b.result;
        rem -= z;
      }
      accI[k * kSplitThreads] += (unsigned int)z;
      const unsigned int tot = __reduce_add_sync(0xffffffffu, (unsigned int)z);
      if (lane == 0 && tot != 0u) atomicAdd(&Jsm[jl * K + k], tot);
    }
  }
  __syncthreads();
  for (int t = tid; t < nloc * K; t += kSplitThreads)
    if (Jsm[t] != 0u) atomicAdd(&a.sumJ[j0 * K + t], (double)Jsm[t]);
  if (in && a.accumulate_I) {
    for (int k = 0; k < K; ++k) {
      const unsigned int v = accI[k * kSplitThreads];
      if (v != 0u) atomicAdd(&a.sumI[(size_t)k * N + gene], (double)v * a.scaleI);
    }
  }
}

static void bk_launch_split(BkShard* b, BkTileArgs& a, cudaStream_t st) {
  a.X = b->Xs.p;
  a.perm = b->perm.p;
  a.A = b->A.p;
  a.W = b->W.p;
  a.M = b->M;
  a.N = b->N;
  a.K = b->K;
  a.sample_offset = b->sample_offset;
  const int K = b->K;
  const int gx = (b->N + kSplitThreads - 1) / kSplitThreads;
  // a chunk's W rows and per-sample sums (8 B per sample and type): at most 24 KB of smem
  const long long max_chunk = std::max<long long>(1, (long long)(24 * 1024) / (8 * (long long)K));
  // many small CTAs: the tiles differ a lot in cost (columns are sorted by depth), so the
  // hardware scheduler needs fine grains to balance them (16 per SM measured best at c2)
  int per_sm = 16;
  if (const char* e = getenv("URSM_BK_CTAS_PER_SM")) per_sm = std::max(1, atoi(e));
  long long want_y = std::max<long long>(1, (long long)(per_sm * b->num_sms + gx - 1) / gx);
  want_y = std::max<long long>(want_y, (b->M + max_chunk - 1) / max_chunk);
  long long ny = std::min<long long>(b->M, want_y);
  a.chunk = (int)((b->M + ny - 1) / ny);
  ny = (b->M + a.chunk - 1) / a.chunk;
  URSM_REQUIRE(ny <= 65535, "bulk split kernel: too many sample chunks");
  const size_t smem = (size_t)a.chunk * K * 8 + (size_t)2 * K * kSplitThreads * 4;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    URSM_CUDA(cudaFuncSetAttribute(bk_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    configured = smem;
  }
  bk_split_kernel<<<dim3(gx, (unsigned)ny), kSplitThreads, smem, st>>>(a);
  URSM_LAUNCH_CHECK();
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

// One Gibbs sweep's W step (:140-146) plus the per-sample part of update_suffStats.
// One thread per (sample, type): W_j ~ Dir(alpha + sum_i Z_j.) is K independent gammas
// (each from its own Philox stream, keyed on the global sample id and the type) that the
// sample's K threads then normalise through shared memory in a fixed order.
constexpr int kWDrawThreads = 128;

__global__ void __launch_bounds__(kWDrawThreads)
bk_w_draw_kernel(double* W, const double* n_prev, double* n_next, const double* alpha,
                 long long M, int K, long long sample_offset, unsigned long long key,
                 unsigned int sweep, int prev_retained, int retained, int draw,
                 double inv_sample, double* eZjk, double* eLogW, double* eW) {
  __shared__ double g_s[kWDrawThreads];
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
      gmm = gamma_mt(alpha[k] + n_prev[e], rng);
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
    if (retained) {
      const double w = W[e];
      eW[e] += w * inv_sample;
      eLogW[e] += log(w) * inv_sample;
    }
    if (n_next) n_next[e] = 0.0;
  }
}

__global__ void bk_mean_finalize_kernel(const double* W, double* eLogW, double* eW, long long M,
                                        int K) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * K) return;
  eW[t] = W[t];
  eLogW[t] = log(W[t]);
}

// stats tail: sum_j E[log W_kj] (K values) and sum_jk E[Z_jk] E[log W_kj]
__global__ void bk_tail_kernel(const double* eZjk, const double* eLogW, long long M, int K,
                               double* tail) {
  __shared__ double red[32];
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

template <int MODE>
static void bk_launch_tile(BkShard* b, BkTileArgs& a, cudaStream_t st) {
  a.X = b->X;
  a.A = b->A.p;
  a.W = b->W.p;
  a.M = b->M;
  a.N = b->N;
  a.K = b->K;
  a.sample_offset = b->sample_offset;
  const int gx = (b->N + kBkThreads - 1) / kBkThreads;
  // enough CTAs to fill the machine a few times over; a chunk's W rows and per-sample
  // sums (2 x chunk x K doubles) get 16 KB next to the per-thread columns (2 x K x T doubles)
  const long long max_chunk = std::max<long long>(1, (16 * 1024) / (16 * (long long)b->K));
  long long want_y = std::max<long long>(1, (long long)(8 * b->num_sms + gx - 1) / gx);
  want_y = std::max<long long>(want_y, (b->M + max_chunk - 1) / max_chunk);
  long long ny = std::min<long long>(b->M, want_y);
  a.chunk = (int)((b->M + ny - 1) / ny);
  ny = (b->M + a.chunk - 1) / a.chunk;
  URSM_REQUIRE(ny <= 65535, "bulk tile kernel: too many sample chunks");
  dim3 grid(gx, (unsigned)ny);
  const int K = b->K;
  size_t smem = (2 * (size_t)a.chunk * K + 2 * (size_t)K * kBkThreads) * sizeof(double);
  URSM_REQUIRE(smem <= 200 * 1024, "bulk tile kernel: too many cell types for shared memory");
#define URSM_TILE_LAUNCH(KM)                                                                    \
  do {                                                                                          \
    if (smem > 48 * 1024)                                                                       \
      URSM_CUDA(cudaFuncSetAttribute(bk_tile_kernel<KM, MODE>,                                  \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    bk_tile_kernel<KM, MODE><<<grid, kBkThreads, smem, st>>>(a);                                \
  } while (0)
  if (K <= 4)
    URSM_TILE_LAUNCH(4);
  else if (K <= 8)
    URSM_TILE_LAUNCH(8);
  else if (K <= 12)
    URSM_TILE_LAUNCH(12);
  else if (K <= 20)
    URSM_TILE_LAUNCH(20);
  else
    URSM_TILE_LAUNCH(32);
#undef URSM_TILE_LAUNCH
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

__global__ void bk_gather_cols_kernel(const float* X, const int* perm, long long M, int N,
                                      float* Xs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const long long j = blockIdx.y;
  if (i < N) Xs[(size_t)j * N + i] = X[(size_t)j * N + perm[i]];
}

// column-sorted copy of X for the Gibbs split kernel (X never changes during a fit)
static void bk_build_sorted(BkShard* b) {
  const int N = b->N;
  DevBuf<double> cs;
  cs.alloc(N);
  bk_colsum_kernel<<<(N + 255) / 256, 256>>>(b->X, b->M, N, cs.p);
  URSM_LAUNCH_CHECK();
  std::vector<double> h(N);
  URSM_CUDA(cudaMemcpy(h.data(), cs.p, N * sizeof(double), cudaMemcpyDeviceToHost));
  std::vector<int> perm(N);
  for (int i = 0; i < N; ++i) perm[i] = i;
  std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) { return h[x] > h[y]; });
  b->perm.alloc(N);
  URSM_CUDA(cudaMemcpy(b->perm.p, perm.data(), N * sizeof(int), cudaMemcpyHostToDevice));
  b->Xs.alloc((size_t)b->M * N);
  for (long long j0 = 0; j0 < b->M; j0 += 65535) {
    const unsigned rows = (unsigned)std::min<long long>(65535, b->M - j0);
    bk_gather_cols_kernel<<<dim3((N + 255) / 256, rows), 256>>>(b->X + (size_t)j0 * N, b->perm.p,
                                                               rows, N, b->Xs.p + (size_t)j0 * N);
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
  return (int64_t)b->N * b->K + b->K + 1;
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
  bk_time_begin(b, st);
  URSM_CUDA(cudaMemsetAsync(d_stats, 0, (size_t)ursm_bk_stats_len(bk) * sizeof(double), st));
  b->acc.zero(st);
  b->eZjk.zero(st);
  BkTileArgs a;
  memset(&a, 0, sizeof a);
  a.sumJ = b->acc.p;
  bk_launch_tile<0>(b, a, st);
  bk_nmf_update_kernel<<<mb, 128, 0, st>>>(b->W.p, b->acc.p, b->alpha.p, b->M, b->K);
  URSM_LAUNCH_CHECK();
  a.sumJ = b->eZjk.p;
  a.sumI = d_stats;
  a.scaleI = 1.0;
  bk_launch_tile<1>(b, a, st);
  bk_mean_finalize_kernel<<<mkb, 256, 0, st>>>(b->W.p, b->eLogW.p, b->eW.p, b->M, b->K);
  URSM_LAUNCH_CHECK();
  bk_tail_kernel<<<1, 1024, 0, st>>>(b->eZjk.p, b->eLogW.p, b->M, b->K,
                                     d_stats + (size_t)b->N * b->K);
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
  if (!b->Xs.p) bk_build_sorted(b);  // once per fit, first Gibbs E-step only
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned mb = (unsigned)((b->M + (kWDrawThreads / b->K) - 1) / (kWDrawThreads / b->K));
  const unsigned long long key = seed + 0x9E3779B97F4A7C15ull * (em_iter + 1) + 0x5bd1e995ull;
  const PhiloxKeys rk = philox_keys(key);
  const double inv = 1.0 / sample;
  bk_time_begin(b, st);
  URSM_CUDA(cudaMemsetAsync(d_stats, 0, (size_t)ursm_bk_stats_len(bk) * sizeof(double), st));
  b->eZjk.zero(st);
  b->eLogW.zero(st);
  b->eW.zero(st);
  const int nsweeps = burnin + sample * thin;
  bool prev_ret = false;
  // nA holds sum_i Z of the previous sweep (chain persists across E-steps, :99-100,136)
  for (int sweep = 0; sweep < nsweeps; ++sweep) {
    const bool retained = sweep >= burnin && ((sweep - burnin) % thin == 0);
    const bool freezeW = freeze & 1, freezeZ = freeze & 2;
    bk_w_draw_kernel<<<mb, kWDrawThreads, 0, st>>>(b->W.p, b->nA.p, b->nB.p, b->alpha.p, b->M, b->K,
                                         b->sample_offset, key, (unsigned)sweep, prev_ret ? 1 : 0,
                                         retained ? 1 : 0, freezeW ? 0 : 1, inv, b->eZjk.p,
                                         b->eLogW.p, b->eW.p);
    URSM_LAUNCH_CHECK();
    BkTileArgs a;
    memset(&a, 0, sizeof a);
    a.sumJ = b->nB.p;
    a.sumI = d_stats;
    a.scaleI = inv;
    a.accumulate_I = retained ? 1 : 0;
    a.key = key;
    a.rk = rk;
    a.sweep = (unsigned)sweep;
    bk_launch_split(b, a, st);
    if (!freezeZ) std::swap(b->nA.p, b->nB.p);
    if (freezeZ && retained) {
      // frozen-Z moment test: E[Z_jk] accumulates the fixed sums
    }
    prev_ret = retained;
  }
  // the last retained sweep's sum_i Z
  bk_w_draw_kernel<<<mb, kWDrawThreads, 0, st>>>(b->W.p, b->nA.p, nullptr, b->alpha.p, b->M, b->K,
                                       b->sample_offset, key, 0u, prev_ret ? 1 : 0, 0, 0, inv,
                                       b->eZjk.p, b->eLogW.p, b->eW.p);
  URSM_LAUNCH_CHECK();
  bk_tail_kernel<<<1, 1024, 0, st>>>(b->eZjk.p, b->eLogW.p, b->M, b->K,
                                     d_stats + (size_t)b->N * b->K);
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
