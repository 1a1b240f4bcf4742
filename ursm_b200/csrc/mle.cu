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
  DevBuf<double> u;                      // [L]
  DevBuf<int> active, list;              // [K]
  DevBuf<double> out;                    // [K][4]
  DevBuf<double> terms;                  // [4]
  int num_sms = 148;
};

constexpr int kColThreads = 256;

// MODE 0: out[type][i] += Y_li * E[S_li]            (m_step.py:88-89)
// MODE 1: out[type][i] += E[S_li] * R_l / u_l       (m_step.py:123-125), active types only
template <int MODE>
__global__ void __launch_bounds__(kColThreads)
mle_colsum_kernel(const float* Y, const unsigned short* cnt, const int* G, const int* order,
                  const double* R, const double* u, const int* active, long long L, int N,
                  int chunk, double inv_sample, double* out) {
  const int i = blockIdx.x * kColThreads + threadIdx.x;
  const long long c0 = (long long)blockIdx.y * chunk, c1 = min(L, c0 + chunk);
  if (i >= N) return;
  double acc = 0.0;
  int cur = -1;
  for (long long c = c0; c < c1; ++c) {
    const int l = order[c];
    const int type = G[l];
    if (type != cur) {
      if (cur >= 0 && acc != 0.0) atomicAdd(&out[(size_t)cur * N + i], acc);
      acc = 0.0;
      cur = type;
    }
    if (MODE == 1 && !active[type]) continue;
    const double es = (double)cnt[(size_t)l * N + i] * inv_sample;
    if (MODE == 0)
      acc += (double)Y[(size_t)l * N + i] * es;
    else
      acc += es * (R[l] / u[l]);
  }
  if (cur >= 0 && acc != 0.0) atomicAdd(&out[(size_t)cur * N + i], acc);
}

// u_l = sum_i A[i, G_l] E[S_li]  (one warp per cell; active types only)
__global__ void mle_u_kernel(const unsigned short* cnt, const int* G, const double* A,
                             const int* active, long long L, int N, double inv_sample,
                             double* u) {
  const long long l = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (l >= L) return;
  const int type = G[l];
  if (!active[type]) return;
  const double* Ak = A + (size_t)type * N;
  const unsigned short* row = cnt + (size_t)l * N;
  double s = 0.0;
  for (int i = lane; i < N; i += 32) s += Ak[i] * (double)row[i];
  s = warp_sum(s);
  if (lane == 0) u[l] = s * inv_sample;
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
  double* A;          // [K][N] in/out
  double* Anew;       // [K][N] scratch
  double* y;          // [K][N] scratch
  double* grad;       // [K][N] scratch
  const double* cAinv;
  const double* cA;
  const double* coeffA;
  const double* part;  // all-reduced sum_{l in k} E[S] R/u
  const int* list;     // active columns
  int N;
  double min_A, init_step;
  double* out;         // [K][4]: ||dA||_1, halvings, obj_new, step
};

// One outer iteration of opt_Ak (:157-182) for one column per CTA.
__global__ void __launch_bounds__(1024) mle_step_kernel(StepArgs a) {
  __shared__ double red[96];
  const int k = a.list[blockIdx.x], N = a.N, tid = threadIdx.x, T = blockDim.x;
  double* Ak = a.A + (size_t)k * N;
  double* An = a.Anew + (size_t)k * N;
  double* y = a.y + (size_t)k * N;
  double* gr = a.grad + (size_t)k * N;
  const double* ci = a.cAinv + (size_t)k * N;
  const double* ca = a.cA + (size_t)k * N;
  const double* c0 = a.coeffA + (size_t)k * N;
  const double* pt = a.part + (size_t)k * N;
  const double invN = 1.0 / N;
  // grad_func = -get_grad_A (:262-272), obj_func = -get_obj_A (:275-287)
  double f1 = 0.0, f2 = 0.0, f3 = 0.0;
  for (int i = tid; i < N; i += T) {
    const double v = Ak[i], cc = c0[i] - pt[i];
    gr[i] = -((ci[i] / v + ca[i] * v + cc) * invN);
    f1 += ci[i] * log(v);
    f2 += ca[i] * v * v;
    f3 += cc * v;
  }
  f1 = block_sum(f1, red);
  f2 = block_sum(f2, red);
  f3 = block_sum(f3, red);
  const double obj_old = -((f1 + f2 / 2.0 + f3) * invN);
  double step = a.init_step, obj_new = 0.0;
  int halvings = 0;
  for (;; ++halvings) {
    for (int i = tid; i < N; i += T) y[i] = Ak[i] - step * gr[i];
    __syncthreads();
    const double lam = proj_lambda(y, N, a.min_A, red);
    double g2 = 0.0, gg = 0.0;
    f1 = f2 = f3 = 0.0;
    for (int i = tid; i < N; i += T) {
      const double nv = fmax(y[i] + lam, a.min_A);
      const double Gt = (Ak[i] - nv) / step;
      const double cc = c0[i] - pt[i];
      An[i] = nv;
      g2 += Gt * Gt;
      gg += gr[i] * Gt;
      f1 += ci[i] * log(nv);
      f2 += ca[i] * nv * nv;
      f3 += cc * nv;
    }
    block_sum2(g2, gg, red);
    f1 = block_sum(f1, red);
    f2 = block_sum(f2, red);
    f3 = block_sum(f3, red);
    obj_new = -((f1 + f2 / 2.0 + f3) * invN);
    if (!(obj_new > obj_old + (step * 0.5) * g2 - step * gg) || halvings >= 1000) break;
    step *= 0.5;
  }
  double d = 0.0;
  for (int i = tid; i < N; i += T) {
    d += fabs(An[i] - Ak[i]);
    Ak[i] = An[i];
  }
  d = block_sum(d, red);
  if (tid == 0) {
    a.out[k * 4 + 0] = d;
    a.out[k * 4 + 1] = (double)halvings;
    a.out[k * 4 + 2] = obj_new;
    a.out[k * 4 + 3] = step;
  }
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
  Mle* m = new Mle();
  m->sc = reinterpret_cast<ScShard*>(sc);
  m->bk = reinterpret_cast<BkShard*>(bk);
  m->N = N;
  m->K = K;
  m->min_A = min_A;
  try {
    if (m->sc) URSM_REQUIRE(m->sc->N == N && m->sc->K == K, "ursm_mle_create: shard shape mismatch");
    int dev;
    URSM_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    URSM_CUDA(cudaGetDeviceProperties(&prop, dev));
    m->num_sms = prop.multiProcessorCount;
    const size_t KN = (size_t)K * N;
    m->A.alloc(KN);
    m->Anew.alloc(KN);
    m->y.alloc(KN);
    m->grad.alloc(KN);
    m->active.alloc(K);
    m->list.alloc(K);
    m->out.alloc(4 * (size_t)K);
    m->terms.alloc(4);
    if (m->sc) {
      m->u.alloc(m->sc->L);
      m->u.zero();
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
    const int gx = (m->N + kColThreads - 1) / kColThreads;
    int ny;
    const int chunk = colsum_chunk(m, s->L, gx, &ny);
    mle_colsum_kernel<0><<<dim3(gx, ny), kColThreads, 0, st>>>(
        s->Y, s->cnt.p, s->G.p, s->order.p, s->R.p, nullptr, nullptr, s->L, m->N, chunk,
        1.0 / s->sample, d_part);
    URSM_LAUNCH_CHECK();
  }
  URSM_API_END
}

int ursm_mle_pass_u(ursm_mle* mm, const int32_t* h_active, double* d_part, void* stream) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && m->sc && h_active && d_part, "ursm_mle_pass_u: bad argument");
  ScShard* s = m->sc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t KN = (size_t)m->K * m->N;
  URSM_CUDA(cudaMemcpyAsync(m->active.p, h_active, m->K * sizeof(int), cudaMemcpyHostToDevice, st));
  URSM_CUDA(cudaMemsetAsync(d_part, 0, (KN + m->K) * sizeof(double), st));
  const double inv = 1.0 / s->sample;
  mle_u_kernel<<<(unsigned)((s->L + 7) / 8), 256, 0, st>>>(s->cnt.p, s->G.p, m->A.p, m->active.p,
                                                           s->L, m->N, inv, m->u.p);
  URSM_LAUNCH_CHECK();
  const int gx = (m->N + kColThreads - 1) / kColThreads;
  int ny;
  const int chunk = colsum_chunk(m, s->L, gx, &ny);
  mle_colsum_kernel<1><<<dim3(gx, ny), kColThreads, 0, st>>>(s->Y, s->cnt.p, s->G.p, s->order.p,
                                                            s->R.p, m->u.p, m->active.p, s->L,
                                                            m->N, chunk, inv, d_part);
  URSM_LAUNCH_CHECK();
  mle_rlogu_kernel<<<(unsigned)((s->L + 255) / 256), 256, 0, st>>>(s->G.p, s->R.p, m->u.p, s->L,
                                                                  d_part + KN);
  URSM_LAUNCH_CHECK();
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

int ursm_mle_step_A(ursm_mle* mm, const int32_t* h_active, const double* d_cconst_part,
                    double init_step, double* h_delta, int32_t* h_halvings, void* stream) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && h_active && d_cconst_part && m->cAinv, "ursm_mle_step_A: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<int> list;
  for (int k = 0; k < m->K; ++k)
    if (h_active[k]) list.push_back(k);
  if (list.empty()) return 0;
  URSM_CUDA(cudaMemcpyAsync(m->list.p, list.data(), list.size() * sizeof(int),
                            cudaMemcpyHostToDevice, st));
  StepArgs a;
  a.A = m->A.p;
  a.Anew = m->Anew.p;
  a.y = m->y.p;
  a.grad = m->grad.p;
  a.cAinv = m->cAinv;
  a.cA = m->cA;
  a.coeffA = m->coeffA;
  a.part = d_cconst_part;
  a.list = m->list.p;
  a.N = m->N;
  a.min_A = m->min_A;
  a.init_step = init_step;
  a.out = m->out.p;
  const int T = std::max(64, std::min(1024, (m->N + 31) / 32 * 32));
  mle_step_kernel<<<(unsigned)list.size(), T, 0, st>>>(a);
  URSM_LAUNCH_CHECK();
  std::vector<double> out(4 * (size_t)m->K);
  URSM_CUDA(cudaMemcpyAsync(out.data(), m->out.p, out.size() * sizeof(double),
                            cudaMemcpyDeviceToHost, st));
  URSM_CUDA(cudaStreamSynchronize(st));
  for (int k : list) {
    if (h_delta) h_delta[k] = out[k * 4];
    if (h_halvings) h_halvings[k] = (int)out[k * 4 + 1];
  }
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
