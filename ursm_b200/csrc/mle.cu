// M-step for sm_100a: m_step.py:13-320 and utils.py:8-31.
//
//  * opt_u (:108-115) + update_gd_coeffConst (:118-125) are the only O(L*N) part
//    and the reference redoes them after EVERY backtracking step of EVERY column.
//    Columns are independent (column k only sees cells of type k), so all K
//    columns advance together: one streaming pass over the uint16 E[S] counters
//    per outer iteration (row dot with A[:,G_l], then type-wise column sums
//    weighted by R_l/u_l), fp64 accumulation.
//  * backtracking (:209-233) with simplex_proj as the projection runs as one CTA
//    per column, entirely on the device (no host round trip per trial step).
//  * simplex_proj (utils.py:8-31) sorts; the projection itself is the unique
//    y -> max(y + lambda, min_y) with sum = 1, so the kernel finds the same
//    lambda by Michelot's finite active-set iteration (block reductions only,
//    no sort).  Same result up to summation-order rounding.
#include "common.cuh"
#include "sc_shard.h"
#include "../../include/ursm_b200.h"

#include <algorithm>

namespace ursm {

struct Mle {
  ScShard* sc = nullptr;
  BkShard* bk = nullptr;
  int N = 0, K = 0;
  double min_A = 1e-6;
  DevBuf<double> A, Anew, y, grad;       // [K][N]
  const double* cAinv = nullptr;         // [K][N] caller-owned (all-reduced)
  const double* cA = nullptr;
  const double* coeffA = nullptr;
  DevBuf<double> u, ratio;               // [L] u_l and R_l / u_l
  DevBuf<int> act_cur, act_next;         // [K] columns in opt_Ak at the start / end of an iteration
  DevBuf<int> niter, base, halv, win;    // [K] outer iterations, halvings already tried, total halvings
  DevBuf<double> cand;                   // [K][kCand][N] candidate columns of one backtracking batch
  DevBuf<double> cres;                   // [K][kCand][4]: accepted?, objective, ||dA||_1, step
  DevBuf<double> delta;                  // [K] last accepted ||dA_k||_1
  DevBuf<double> terms;                  // [4]
  double init_step = 0, conv = 0;
  int maxiter = 0;
  int num_sms = 148;
};

constexpr int kCand = 16;  // trial step sizes of backtracking evaluated concurrently

constexpr int kColThreads = 256;

// MODE 0: out[type][i] += Y_li * E[S_li]            (m_step.py:88-89)
// MODE 1: out[type][i] += E[S_li] * R_l / u_l       (m_step.py:123-125), active types only
// A thread owns four adjacent genes (one 64-bit load of counters per cell) and walks a
// chunk of the type-sorted cells, four cells per trip so that four independent loads
// are in flight; `ratio` = R_l/u_l comes from mle_u_kernel.  HBM-bound: 2 B per
// cell x gene (MODE 0 adds the 4 B count).
constexpr int kColGenes = 4;

template <int MODE>
__global__ void __launch_bounds__(kColThreads)
mle_colsum_kernel(const float* Y, const unsigned short* cnt, const int* G, const int* order,
                  const double* ratio, const int* active, long long L, int N, int chunk,
                  double inv_sample, double* out) {
  const int i = kColGenes * (blockIdx.x * kColThreads + threadIdx.x);
  const long long c0 = (long long)blockIdx.y * chunk, c1 = min(L, c0 + chunk);
  if (i >= N) return;
  const bool vec = (i + 3 < N) && ((N & 3) == 0);  // 64-bit loads need 8-byte aligned rows
  double acc[kColGenes] = {0.0, 0.0, 0.0, 0.0};
  int cur = -1;
  auto flush = [&]() {
    if (cur < 0) return;
#pragma unroll
    for (int q = 0; q < kColGenes; ++q)
      if (acc[q] != 0.0 && i + q < N) atomicAdd(&out[(size_t)cur * N + i + q], acc[q]);
#pragma unroll
    for (int q = 0; q < kColGenes; ++q) acc[q] = 0.0;
  };
  auto load4 = [&](int l, unsigned int& lo, unsigned int& hi) {
    const unsigned short* row = cnt + (size_t)l * N + i;
    if (vec) {
      const uint2 v = *reinterpret_cast<const uint2*>(row);
      lo = v.x;
      hi = v.y;
    } else {
      lo = row[0];
      hi = 0u;
      if (i + 1 < N) lo |= (unsigned int)row[1] << 16;
      if (i + 2 < N) hi = row[2];
      if (i + 3 < N) hi |= (unsigned int)row[3] << 16;
    }
  };
  for (long long c = c0; c < c1; c += 4) {
    int l[4], ty[4];
    unsigned int lo[4], hi[4];
    double w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const bool ok = c + q < c1;
      l[q] = ok ? order[c + q] : -1;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      ty[q] = l[q] >= 0 ? G[l[q]] : -1;
      const bool use = ty[q] >= 0 && (MODE == 0 || active[ty[q]]);
      if (!use) ty[q] = ty[q] >= 0 ? -2 - ty[q] : -1;  // keep the type for flushing, skip the row
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      lo[q] = hi[q] = 0u;
      w[q] = 0.0;
      if (ty[q] >= 0) {
        load4(l[q], lo[q], hi[q]);
        w[q] = (MODE == 1) ? ratio[l[q]] * inv_sample : inv_sample;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (l[q] < 0) continue;
      const int type = ty[q] >= 0 ? ty[q] : -2 - ty[q];
      if (type != cur) {
        flush();
        cur = type;
      }
      if (ty[q] < 0) continue;
      const double e0 = (double)(lo[q] & 0xffffu), e1 = (double)(lo[q] >> 16);
      const double e2 = (double)(hi[q] & 0xffffu), e3 = (double)(hi[q] >> 16);
      if (MODE == 0) {
        const float* y = Y + (size_t)l[q] * N + i;
        acc[0] += (double)y[0] * e0 * w[q];
        if (i + 1 < N) acc[1] += (double)y[1] * e1 * w[q];
        if (i + 2 < N) acc[2] += (double)y[2] * e2 * w[q];
        if (i + 3 < N) acc[3] += (double)y[3] * e3 * w[q];
      } else {
        acc[0] += e0 * w[q];
        acc[1] += e1 * w[q];
        acc[2] += e2 * w[q];
        acc[3] += e3 * w[q];
      }
    }
  }
  flush();
}

// u_l = sum_i A[i, G_l] E[S_li] and ratio_l = R_l / u_l  (one warp per cell; active types)
__global__ void mle_u_kernel(const unsigned short* cnt, const int* G, const double* A,
                             const double* R, const int* active, long long L, int N,
                             double inv_sample, double* u, double* ratio) {
  const long long l = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (l >= L) return;
  const int type = G[l];
  if (!active[type]) return;
  const double* Ak = A + (size_t)type * N;
  const unsigned short* row = cnt + (size_t)l * N;
  double s = 0.0;
  if ((N & 3) == 0) {  // 64-bit loads: four counters per lane and trip
    const uint2* row4 = reinterpret_cast<const uint2*>(row);
    const int n4 = N >> 2;
    double s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int q = lane; q < n4; q += 128) {  // four independent 8-byte loads in flight
      const uint2 z = make_uint2(0u, 0u);
      const int q1 = q + 32, q2 = q + 64, q3 = q + 96;
      const uint2 c0 = row4[q];
      const uint2 c1 = q1 < n4 ? row4[q1] : z;
      const uint2 c2 = q2 < n4 ? row4[q2] : z;
      const uint2 c3 = q3 < n4 ? row4[q3] : z;
      const double* a0 = Ak + 4 * q;
      s += a0[0] * (double)(c0.x & 0xffffu) + a0[1] * (double)(c0.x >> 16) +
           a0[2] * (double)(c0.y & 0xffffu) + a0[3] * (double)(c0.y >> 16);
      if (q1 < n4) {
        const double* a1 = Ak + 4 * q1;
        s1 += a1[0] * (double)(c1.x & 0xffffu) + a1[1] * (double)(c1.x >> 16) +
              a1[2] * (double)(c1.y & 0xffffu) + a1[3] * (double)(c1.y >> 16);
      }
      if (q2 < n4) {
        const double* a2 = Ak + 4 * q2;
        s2 += a2[0] * (double)(c2.x & 0xffffu) + a2[1] * (double)(c2.x >> 16) +
              a2[2] * (double)(c2.y & 0xffffu) + a2[3] * (double)(c2.y >> 16);
      }
      if (q3 < n4) {
        const double* a3 = Ak + 4 * q3;
        s3 += a3[0] * (double)(c3.x & 0xffffu) + a3[1] * (double)(c3.x >> 16) +
              a3[2] * (double)(c3.y & 0xffffu) + a3[3] * (double)(c3.y >> 16);
      }
    }
    s += s1 + s2 + s3;
  } else {
    for (int i = lane; i < N; i += 32) s += Ak[i] * (double)row[i];
  }
  s = warp_sum(s);
  if (lane == 0) {
    const double ul = s * inv_sample;
    u[l] = ul;
    ratio[l] = R[l] / ul;
  }
}

// out[type] = sum_{l in type} R_l log u_l   (compute_elbo :308), every type
__global__ void mle_rlogu_kernel(const int* G, const double* R, const double* u, long long L,
                                 double* out) {
  long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  const double r = R[l];
  if (r != 0.0) atomicAdd(&out[G[l]], r * log(u[l]));
}

__device__ __forceinline__ void block_sum2(double& a, double& b, double* red) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  __syncthreads();
  if (lane == 0) {
    red[warp] = a;
    red[32 + warp] = b;
  }
  __syncthreads();
  double x = (lane < nw) ? red[lane] : 0.0, z = (lane < nw) ? red[32 + lane] : 0.0;
  a = warp_sum(x);
  b = warp_sum(z);
}

// lambda such that sum_i max(y_i + lambda, m) = 1  (utils.py:8-31 without the sort)
__device__ double proj_lambda(const double* y, int N, double m, double* red) {
  double s = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s += y[i];
  s = block_sum(s, red);
  double lam = (1.0 - s) / N;
  double prev = N;
  for (int it = 0; it < 4096; ++it) {
    s = 0.0;
    c = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      const double v = y[i];
      if (v + lam > m) {
        s += v;
        c += 1.0;
      }
    }
    block_sum2(s, c, red);
    if (c == 0.0) break;  // degenerate (N*m >= 1): everything sits on the floor
    lam = (1.0 - (N - c) * m - s) / c;
    if (c == prev) break;
    prev = c;
  }
  return lam;
}

struct StepArgs {
  const double* A;     // [K][N] current columns
  double* cand;        // [K][kCand][N] out: projected trial columns
  double* cres;        // [K][kCand][4]
  const double* cAinv;
  const double* cA;
  const double* coeffA;
  const double* part;  // all-reduced sum_{l in k} E[S] R/u
  const int* act;      // [K]
  const int* base;     // [K] halvings already rejected in earlier batches
  int N;
  double min_A, init_step;
};

__device__ __forceinline__ void block_sum6(double* v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int q = 0; q < 6; ++q) v[q] = warp_sum(v[q]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < 6; ++q) red[q * 32 + warp] = v[q];
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 6; ++q) v[q] = warp_sum((lane < nw) ? red[q * 32 + lane] : 0.0);
}

// One trial of backtracking (:209-233) per CTA: column blockIdx.y, step size
// init_step / 2^(base + blockIdx.x).  The reference halves sequentially until the
// sufficient-decrease test passes; the kCand trials of a batch are independent, so
// they run concurrently and mle_commit_kernel takes the FIRST one that passes -- the
// same step the sequential loop would have accepted.
//   grad_func = -get_grad_A (:262-272), obj_func = -get_obj_A (:275-287),
//   projection = simplex_proj (utils.py:8-31) via proj_lambda.
__global__ void __launch_bounds__(1024) mle_cand_kernel(StepArgs a) {
  __shared__ double red[6 * 32];
  const int k = blockIdx.y, c = blockIdx.x, N = a.N, tid = threadIdx.x, T = blockDim.x;
  if (!a.act[k]) return;
  const double* Ak = a.A + (size_t)k * N;
  double* y = a.cand + ((size_t)k * kCand + c) * N;
  const double* ci = a.cAinv + (size_t)k * N;
  const double* ca = a.cA + (size_t)k * N;
  const double* c0 = a.coeffA + (size_t)k * N;
  const double* pt = a.part + (size_t)k * N;
  const double invN = 1.0 / N;
  const double step = ldexp(a.init_step, -(a.base[k] + c));
  double v[6] = {0, 0, 0, 0, 0, 0};
  for (int i = tid; i < N; i += T) {
    const double x = Ak[i], cc = c0[i] - pt[i];
    const double g = -((ci[i] / x + ca[i] * x + cc) * invN);
    y[i] = x - step * g;
    v[0] += ci[i] * log(x);
    v[1] += ca[i] * x * x;
    v[2] += cc * x;
  }
  block_sum6(v, red);
  const double obj_old = -((v[0] + v[1] / 2.0 + v[2]) * invN);
  __syncthreads();
  const double lam = proj_lambda(y, N, a.min_A, red);
  for (int q = 0; q < 6; ++q) v[q] = 0.0;
  for (int i = tid; i < N; i += T) {
    const double x = Ak[i], cc = c0[i] - pt[i];
    const double g = -((ci[i] / x + ca[i] * x + cc) * invN);
    const double nv = fmax(y[i] + lam, a.min_A);
    const double Gt = (x - nv) / step;
    y[i] = nv;
    v[0] += ci[i] * log(nv);
    v[1] += ca[i] * nv * nv;
    v[2] += cc * nv;
    v[3] += Gt * Gt;
    v[4] += g * Gt;
    v[5] += fabs(nv - x);
  }
  block_sum6(v, red);
  if (tid == 0) {
    const double obj_new = -((v[0] + v[1] / 2.0 + v[2]) * invN);
    const bool halve = obj_new > obj_old + (step * 0.5) * v[3] - step * v[4];
    double* r = a.cres + ((size_t)k * kCand + c) * 4;
    r[0] = halve ? 0.0 : 1.0;
    r[1] = obj_new;
    r[2] = v[5];
    r[3] = step;
  }
}

// Per column (one thread each): accept the first passing trial of the batch (or
// schedule the next batch of halvings) and advance the opt_Ak loop state (:164-181).
// win[k] = index of the accepted trial, -1 if none.
__global__ void mle_decide_kernel(const double* cres, const int* act, int* act_next, int* niter,
                                  int* base, int* halv, double* delta, int* win, int K,
                                  double conv, int maxiter) {
  const int k = threadIdx.x;
  if (k >= K) return;
  win[k] = -1;
  if (!act[k]) {
    act_next[k] = 0;
    return;
  }
  int w = -1;
  for (int c = 0; c < kCand; ++c)
    if (cres[((size_t)k * kCand + c) * 4] != 0.0) {
      w = c;
      break;
    }
  if (w < 0 && base[k] + kCand >= 1024) w = kCand - 1;  // step underflow guard
  if (w < 0) {  // every trial of this batch was rejected: keep halving
    base[k] += kCand;
    act_next[k] = 1;
    return;
  }
  const double d = cres[((size_t)k * kCand + w) * 4 + 2];
  win[k] = w;
  delta[k] = d;
  halv[k] += base[k] + w;
  base[k] = 0;
  const int n = ++niter[k];
  act_next[k] = (d > conv && n < maxiter) ? 1 : 0;
}

// A[k] <- accepted trial column
__global__ void mle_copy_kernel(double* A, const double* cand, const int* win, int N) {
  const int k = blockIdx.x, w = win[k];
  if (w < 0) return;
  const double* src = cand + ((size_t)k * kCand + w) * N;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < N; i += gridDim.y * blockDim.x)
    A[(size_t)k * N + i] = src[i];
}

// part[k][:] <- part_local[k][:] for the columns that were stepped; the per-type
// sum_l R_l log u_l tail is refreshed for every type; then the loop state advances.
__global__ void mle_merge_kernel(const double* part_local, double* part, const int* act, int N,
                                 int K) {
  const int k = blockIdx.y;
  if (k == K) {
    if (blockIdx.x == 0 && threadIdx.x < K)
      part[(size_t)K * N + threadIdx.x] = part_local[(size_t)K * N + threadIdx.x];
    return;
  }
  if (!act[k]) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) part[(size_t)k * N + i] = part_local[(size_t)k * N + i];
}

__global__ void mle_advance_kernel(int* act_cur, const int* act_next, int K) {
  if (threadIdx.x < K) act_cur[threadIdx.x] = act_next[threadIdx.x];
}

// A_k = simplex_proj(v / sum v)   (m_step.py:149-153, utils.py:8-31)
__global__ void __launch_bounds__(1024)
mle_project_kernel(const double* v, double* y, double* out, int N, double m, int normalise) {
  __shared__ double red[96];
  double s = 0.0;
  if (normalise) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) s += v[i];
    s = block_sum(s, red);
  } else {
    s = 1.0;
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) y[i] = v[i] / s;
  __syncthreads();
  const double lam = proj_lambda(y, N, m, red);
  for (int i = threadIdx.x; i < N; i += blockDim.x) out[i] = fmax(y[i] + lam, m);
}

// sums over (k, i): cAinv log A | cA A^2 / 2 | (coeffA - part) A    (compute_elbo :297-306)
__global__ void mle_elbo_kernel(const double* A, const double* cAinv, const double* cA,
                                const double* coeffA, const double* part, size_t n, int use_sc,
                                double* terms) {
  __shared__ double red[96];
  double t0 = 0.0, t1 = 0.0, t2 = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const double v = A[i];
    t0 += cAinv[i] * log(v);
    if (use_sc) {
      t1 += cA[i] * v * v;
      t2 += (coeffA[i] - part[i]) * v;
    }
  }
  t0 = block_sum(t0, red);
  t1 = block_sum(t1, red);
  t2 = block_sum(t2, red);
  if (threadIdx.x == 0) {
    atomicAdd(&terms[0], t0);
    atomicAdd(&terms[1], t1 / 2.0);
    atomicAdd(&terms[2], t2);
  }
}

static void transpose_NK_to_KN(const double* h, std::vector<double>& t, int N, int K) {
  t.resize((size_t)N * K);
  for (int i = 0; i < N; ++i)
    for (int k = 0; k < K; ++k) t[(size_t)k * N + i] = h[(size_t)i * K + k];
}

static int colsum_chunk(const Mle* m, long long L, int gx, int* ny_out) {
  long long want = std::max<long long>(1, (8LL * m->num_sms + gx - 1) / gx);
  long long ny = std::min<long long>(std::min<long long>(L, want), 65535);
  int chunk = (int)((L + ny - 1) / ny);
  *ny_out = (int)((L + chunk - 1) / chunk);
  return chunk;
}

}  // namespace ursm

using namespace ursm;

extern "C" {

int ursm_mle_create(ursm_mle** out, ursm_sc* sc, ursm_bk* bk, int32_t N, int32_t K,
                    double min_A) {
  URSM_API_BEGIN
  URSM_REQUIRE(out && N > 0 && K > 0 && min_A > 0, "ursm_mle_create: bad argument");
  URSM_REQUIRE(K <= 32, "ursm_mle_create: K <= 32 cell types supported");
  Mle* m = new Mle();
  m->sc = reinterpret_cast<ScShard*>(sc);
  m->bk = reinterpret_cast<BkShard*>(bk);
  m->N = N;
  m->K = K;
  m->min_A = min_A;
  try {
    if (m->sc) URSM_REQUIRE(m->sc->N == N && m->sc->K == K, "ursm_mle_create: shard shape mismatch");
    m->num_sms = device_sm_count();
    const size_t KN = (size_t)K * N;
    m->A.alloc(KN);
    m->Anew.alloc(KN);
    m->y.alloc(KN);
    m->grad.alloc(KN);
    m->act_cur.alloc(K);
    m->act_next.alloc(K);
    m->niter.alloc(K);
    m->base.alloc(K);
    m->halv.alloc(K);
    m->win.alloc(K);
    m->delta.alloc(K);
    m->act_cur.zero();
    m->act_next.zero();
    m->cand.alloc(KN * kCand);
    m->cres.alloc((size_t)K * kCand * 4);
    m->terms.alloc(4);
    if (m->sc) {
      m->u.alloc(m->sc->L);
      m->u.zero();
      m->ratio.alloc(m->sc->L);
      m->ratio.zero();
    }
    URSM_CUDA(cudaDeviceSynchronize());
  } catch (...) {
    delete m;
    throw;
  }
  *out = reinterpret_cast<ursm_mle*>(m);
  URSM_API_END
}

int ursm_mle_destroy(ursm_mle* mm) {
  URSM_API_BEGIN
  cudaDeviceSynchronize();  // the pool frees below are ordered on the default stream only
  delete reinterpret_cast<Mle*>(mm);
  URSM_API_END
}

int ursm_mle_set_A(ursm_mle* mm, const double* h_A) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  std::vector<double> t;
  transpose_NK_to_KN(h_A, t, m->N, m->K);
  URSM_CUDA(cudaMemcpy(m->A.p, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice));
  URSM_API_END
}

int ursm_mle_get_A(ursm_mle* mm, double* h_A) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  std::vector<double> t((size_t)m->N * m->K);
  URSM_CUDA(cudaDeviceSynchronize());
  URSM_CUDA(cudaMemcpy(t.data(), m->A.p, t.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (int i = 0; i < m->N; ++i)
    for (int k = 0; k < m->K; ++k) h_A[(size_t)i * m->K + k] = t[(size_t)k * m->N + i];
  URSM_API_END
}

int ursm_mle_begin(ursm_mle* mm, double* d_part, void* stream) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && d_part, "ursm_mle_begin: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  URSM_CUDA(cudaMemsetAsync(d_part, 0, (size_t)m->K * m->N * sizeof(double), st));
  if (m->sc) {
    ScShard* s = m->sc;
    const int gx = (m->N + kColGenes * kColThreads - 1) / (kColGenes * kColThreads);
    int ny;
    const int chunk = colsum_chunk(m, s->L, gx, &ny);
    mle_colsum_kernel<0><<<dim3(gx, ny), kColThreads, 0, st>>>(
        s->Y, s->cnt.p, s->G.p, s->order.p, nullptr, nullptr, s->L, m->N, chunk,
        1.0 / s->sample, d_part);
    URSM_LAUNCH_CHECK();
  }
  URSM_API_END
}

// opt_u + update_gd_coeffConst for the columns flagged in m->act_cur (device flags)
static void mle_pass(Mle* m, double* d_part, cudaStream_t st) {
  ScShard* s = m->sc;
  const size_t KN = (size_t)m->K * m->N;
  URSM_CUDA(cudaMemsetAsync(d_part, 0, (KN + m->K) * sizeof(double), st));
  const double inv = 1.0 / s->sample;
  mle_u_kernel<<<(unsigned)((s->L + 7) / 8), 256, 0, st>>>(s->cnt.p, s->G.p, m->A.p, s->R.p,
                                                           m->act_cur.p, s->L, m->N, inv, m->u.p,
                                                           m->ratio.p);
  URSM_LAUNCH_CHECK();
  const int gx = (m->N + kColGenes * kColThreads - 1) / (kColGenes * kColThreads);
  int ny;
  const int chunk = colsum_chunk(m, s->L, gx, &ny);
  mle_colsum_kernel<1><<<dim3(gx, ny), kColThreads, 0, st>>>(s->Y, s->cnt.p, s->G.p, s->order.p,
                                                            m->ratio.p, m->act_cur.p, s->L, m->N,
                                                            chunk, inv, d_part);
  URSM_LAUNCH_CHECK();
  mle_rlogu_kernel<<<(unsigned)((s->L + 255) / 256), 256, 0, st>>>(s->G.p, s->R.p, m->u.p, s->L,
                                                                  d_part + KN);
  URSM_LAUNCH_CHECK();
}

int ursm_mle_pass_u(ursm_mle* mm, const int32_t* h_active, double* d_part, void* stream) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && m->sc && h_active && d_part, "ursm_mle_pass_u: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  URSM_CUDA(cudaMemcpyAsync(m->act_cur.p, h_active, m->K * sizeof(int), cudaMemcpyHostToDevice, st));
  mle_pass(m, d_part, st);
  URSM_API_END
}

int ursm_mle_set_coeffs(ursm_mle* mm, const double* d_cAinv, const double* d_cA,
                        const double* d_coeffA) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && d_cAinv && d_cA && d_coeffA, "ursm_mle_set_coeffs: bad argument");
  m->cAinv = d_cAinv;
  m->cA = d_cA;
  m->coeffA = d_coeffA;
  URSM_API_END
}

int ursm_mle_opt_begin(ursm_mle* mm, const int32_t* h_active, double init_step, double conv,
                       int32_t maxiter, void* stream) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && h_active && m->cAinv && init_step > 0 && maxiter >= 1,
               "ursm_mle_opt_begin: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  m->init_step = init_step;
  m->conv = conv;
  m->maxiter = maxiter;
  URSM_CUDA(cudaMemcpyAsync(m->act_cur.p, h_active, m->K * sizeof(int), cudaMemcpyHostToDevice, st));
  URSM_CUDA(cudaMemcpyAsync(m->act_next.p, h_active, m->K * sizeof(int), cudaMemcpyHostToDevice, st));
  m->niter.zero(st);
  m->base.zero(st);
  m->halv.zero(st);
  m->delta.zero(st);
  URSM_API_END
}

int ursm_mle_opt_step(ursm_mle* mm, const double* d_cconst_part, void* stream) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && d_cconst_part && m->cAinv && m->maxiter > 0, "ursm_mle_opt_step: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  StepArgs a;
  a.A = m->A.p;
  a.cand = m->cand.p;
  a.cres = m->cres.p;
  a.cAinv = m->cAinv;
  a.cA = m->cA;
  a.coeffA = m->coeffA;
  a.part = d_cconst_part;
  a.act = m->act_cur.p;
  a.base = m->base.p;
  a.N = m->N;
  a.min_A = m->min_A;
  a.init_step = m->init_step;
  const int T = std::max(64, std::min(1024, (m->N / 4 + 31) / 32 * 32));
  mle_cand_kernel<<<dim3(kCand, m->K), T, 0, st>>>(a);
  URSM_LAUNCH_CHECK();
  mle_decide_kernel<<<1, 32, 0, st>>>(m->cres.p, m->act_cur.p, m->act_next.p, m->niter.p,
                                      m->base.p, m->halv.p, m->delta.p, m->win.p, m->K, m->conv,
                                      m->maxiter);
  URSM_LAUNCH_CHECK();
  mle_copy_kernel<<<dim3(m->K, 8), 256, 0, st>>>(m->A.p, m->cand.p, m->win.p, m->N);
  URSM_LAUNCH_CHECK();
  URSM_API_END
}

int ursm_mle_opt_pass(ursm_mle* mm, double* d_part_local, void* stream) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && m->sc && d_part_local, "ursm_mle_opt_pass: bad argument");
  mle_pass(m, d_part_local, (cudaStream_t)stream);
  URSM_API_END
}

int ursm_mle_opt_merge(ursm_mle* mm, const double* d_part_local, double* d_part, void* stream) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && d_part_local && d_part, "ursm_mle_opt_merge: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  mle_merge_kernel<<<dim3((m->N + 255) / 256, m->K + 1), 256, 0, st>>>(d_part_local, d_part,
                                                                       m->act_cur.p, m->N, m->K);
  URSM_LAUNCH_CHECK();
  mle_advance_kernel<<<1, 32, 0, st>>>(m->act_cur.p, m->act_next.p, m->K);
  URSM_LAUNCH_CHECK();
  URSM_API_END
}

int ursm_mle_opt_poll(ursm_mle* mm, int32_t* h_active, int32_t* h_niter, int32_t* h_halvings,
                      double* h_delta, void* stream) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && h_active, "ursm_mle_opt_poll: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t b = m->K * sizeof(int);
  URSM_CUDA(cudaMemcpyAsync(h_active, m->act_cur.p, b, cudaMemcpyDeviceToHost, st));
  if (h_niter) URSM_CUDA(cudaMemcpyAsync(h_niter, m->niter.p, b, cudaMemcpyDeviceToHost, st));
  if (h_halvings) URSM_CUDA(cudaMemcpyAsync(h_halvings, m->halv.p, b, cudaMemcpyDeviceToHost, st));
  if (h_delta)
    URSM_CUDA(cudaMemcpyAsync(h_delta, m->delta.p, m->K * sizeof(double), cudaMemcpyDeviceToHost, st));
  URSM_CUDA(cudaStreamSynchronize(st));
  URSM_API_END
}

int ursm_mle_closed_form(ursm_mle* mm, int32_t k, const double* d_exp_Zik, void* stream) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && d_exp_Zik && k >= 0 && k < m->K, "ursm_mle_closed_form: bad argument");
  const int T = std::max(64, std::min(1024, (m->N + 31) / 32 * 32));
  mle_project_kernel<<<1, T, 0, (cudaStream_t)stream>>>(
      d_exp_Zik + (size_t)k * m->N, m->y.p + (size_t)k * m->N, m->A.p + (size_t)k * m->N, m->N,
      m->min_A, 1);
  URSM_LAUNCH_CHECK();
  URSM_API_END
}

int ursm_mle_elbo_terms(ursm_mle* mm, const double* d_cconst_part, double* h_terms3,
                        void* stream) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && h_terms3 && m->cAinv, "ursm_mle_elbo_terms: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  m->terms.zero(st);
  const size_t KN = (size_t)m->K * m->N;
  const int use_sc = (m->sc && d_cconst_part) ? 1 : 0;
  const unsigned blocks = (unsigned)std::min<size_t>(4 * (size_t)m->num_sms, (KN + 255) / 256);
  mle_elbo_kernel<<<blocks, 256, 0, st>>>(m->A.p, m->cAinv, m->cA, m->coeffA, d_cconst_part, KN,
                                          use_sc, m->terms.p);
  URSM_LAUNCH_CHECK();
  URSM_CUDA(cudaMemcpyAsync(h_terms3, m->terms.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  URSM_CUDA(cudaStreamSynchronize(st));
  URSM_API_END
}

int ursm_mle_get_u(ursm_mle* mm, double* h_u) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && m->sc && h_u, "ursm_mle_get_u: bad argument");
  URSM_CUDA(cudaDeviceSynchronize());
  URSM_CUDA(cudaMemcpy(h_u, m->u.p, m->sc->L * sizeof(double), cudaMemcpyDeviceToHost));
  URSM_API_END
}

int ursm_simplex_proj(const double* h_y, int32_t n, double min_y, double* h_out) {
  URSM_API_BEGIN
  URSM_REQUIRE(h_y && h_out && n > 0, "ursm_simplex_proj: bad argument");
  DevBuf<double> v, y, o;
  v.alloc(n);
  y.alloc(n);
  o.alloc(n);
  URSM_CUDA(cudaMemcpy(v.p, h_y, n * sizeof(double), cudaMemcpyHostToDevice));
  const int T = std::max(64, std::min(1024, (n + 31) / 32 * 32));
  mle_project_kernel<<<1, T>>>(v.p, y.p, o.p, n, min_y, 0);
  URSM_LAUNCH_CHECK();
  URSM_CUDA(cudaMemcpy(h_out, o.p, n * sizeof(double), cudaMemcpyDeviceToHost));
  URSM_API_END
}

}  // extern "C"
