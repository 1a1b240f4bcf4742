// Stand-alone A/B harness for the bulk split kernel (csrc/bk_split.cuh): synthetic bulk counts
// with the distributions of demo_simulate_data.py (A = normalised exp N(0,1), W ~ Dir(2),
// depth 50 N), one launch per "sweep", event-timed; conservation and moment checks of the
// per-sample and per-gene sums against their exact expectations; optional event counters.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o /tmp/bk_probe tools/bk_probe.cu
//   /tmp/bk_probe [M N K depth_per_gene reps]
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include <cuda_runtime.h>

#include "../ursm_b200/csrc/bk_split.cuh"

using namespace ursm;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e = (x);                                                              \
    if (e != cudaSuccess) {                                                           \
      printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e));                \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

// V = 0: per-lane kernel; V = 1: lockstep kernel + cold-list kernel
template <int V, bool P>
static void launch(const BkSplitArgs& a, const BkSplitGeom& g) {
  static bool once = false;
  if (!once) {
    CK(cudaFuncSetAttribute(bk_split_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    CK(cudaFuncSetAttribute(bk_split_lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_lane));
    once = true;
  }
  if (V == 1) {
    CK(cudaMemsetAsync(a.cold_count, 0, sizeof(unsigned int)));
    bk_split_kernel<P><<<dim3(g.gx, g.rows), kSplitThreads, g.smem>>>(a);
    bk_split_cold_kernel<<<(a.cold_cap + 127) / 128, 128>>>(a);
  } else {
    bk_split_lane_kernel<<<dim3(g.gx, g.rows), kSplitThreads, g.smem_lane>>>(a);
  }
}

template <int V, bool P>
static float run(BkSplitArgs a, const BkSplitGeom& g, int reps, int retained, size_t MK, size_t KN) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  a.retained = retained;
  // two warm-up sweeps, then `reps` sweeps enqueued back to back between two events (the queue
  // stays ahead of the device, so launch latency of the host is not in the figure); the
  // accumulators are not reset in between (the last sweep alone is checked by the caller)
  for (int r = 0; r < 2; ++r) {
    a.sweep = r;
    launch<V, P>(a, g);
  }
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r) {
    a.sweep = 2 + r;
    launch<V, P>(a, g);
  }
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaMemset(a.n_cur, 0, MK * sizeof(double)));
  CK(cudaMemset(a.sumI, 0, KN * sizeof(double)));
  a.sweep = 2 + reps;
  launch<V, P>(a, g);
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  return ms / reps;
}

int main(int argc, char** argv) {
  const long long M = argc > 1 ? atoll(argv[1]) : 200;
  const int N = argc > 2 ? atoi(argv[2]) : 5000;
  const int K = argc > 3 ? atoi(argv[3]) : 5;
  const double depth = argc > 4 ? atof(argv[4]) : 50.0;
  const int reps = argc > 5 ? atoi(argv[5]) : 10;
  int dev_sms = 0;
  CK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0));
  std::mt19937_64 gen(12345);
  std::normal_distribution<double> nd(0, 1);
  std::gamma_distribution<double> gd(2.0, 1.0);
  std::vector<double> A((size_t)K * N), W((size_t)M * K);
  for (int k = 0; k < K; ++k) {
    double s = 0;
    for (int i = 0; i < N; ++i) s += (A[(size_t)k * N + i] = std::exp(nd(gen)));
    for (int i = 0; i < N; ++i) A[(size_t)k * N + i] /= s;
  }
  for (long long j = 0; j < M; ++j) {
    double s = 0;
    for (int k = 0; k < K; ++k) s += (W[j * K + k] = gd(gen));
    for (int k = 0; k < K; ++k) W[j * K + k] /= s;
  }
  if (const char* e = getenv("PROBE_SORT_SAMPLES")) {
    // samples with similar mixing weights into the same warp (1: by dominant type, then its weight)
    std::vector<int> idx(M);
    for (long long j = 0; j < M; ++j) idx[j] = (int)j;
    auto key = [&](int j) {
      int best = 0;
      for (int k = 1; k < K; ++k)
        if (W[(size_t)j * K + k] > W[(size_t)j * K + best]) best = k;
      return std::make_pair(best, W[(size_t)j * K + best]);
    };
    if (atoi(e) == 1) std::sort(idx.begin(), idx.end(), [&](int x, int y) { return key(x) < key(y); });
    std::vector<double> W2(W.size());
    for (long long j = 0; j < M; ++j)
      for (int k = 0; k < K; ++k) W2[j * K + k] = W[(size_t)idx[j] * K + k];
    W = W2;
  }
  const int ngroups = (int)((M + 31) / 32), Mpad = ngroups * 32;
  std::vector<float> Xt((size_t)N * Mpad, 0.f), A32((size_t)K * N), W32g((size_t)ngroups * K * 32, 0.f);
  for (size_t t = 0; t < A32.size(); ++t) A32[t] = (float)A[t];
  for (long long j = 0; j < M; ++j)
    for (int k = 0; k < K; ++k) W32g[((size_t)(j >> 5) * K + k) * 32 + (j & 31)] = (float)W[j * K + k];
  // a different (perturbed) W for the data than for the split, as in a running chain
  double total = 0;
  for (long long j = 0; j < M; ++j)
    for (int i = 0; i < N; ++i) {
      double aw = 0;
      for (int k = 0; k < K; ++k) aw += W[j * K + k] * A[(size_t)k * N + i];
      std::poisson_distribution<long long> pd(depth * N * aw);
      const double x = (double)pd(gen);
      Xt[(size_t)i * Mpad + j] = (float)x;
      total += x;
    }
  printf("M=%lld N=%d K=%d depth/gene=%.0f  reads=%.3e  SMs=%d\n", M, N, K, depth, total, dev_sms);
  std::vector<double> lf2(URSM_BINOM_TABLE);
  binom_fill_table(lf2.data());

  const size_t MK = (size_t)M * K, KN = (size_t)K * N;
  BkSplitArgs a;
  memset(&a, 0, sizeof a);
  float *dXt, *dA, *dW;
  double *dl, *dn, *dI;
  unsigned long long* dprobe;
  CK(cudaMalloc(&dXt, Xt.size() * 4));
  CK(cudaMalloc(&dA, A32.size() * 4));
  CK(cudaMalloc(&dW, W32g.size() * 4));
  CK(cudaMalloc(&dl, lf2.size() * 8));
  CK(cudaMalloc(&dn, MK * 8));
  CK(cudaMalloc(&dI, KN * 8));
  CK(cudaMalloc(&dprobe, 16 * 8));
  CK(cudaMemcpy(dXt, Xt.data(), Xt.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dA, A32.data(), A32.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, W32g.data(), W32g.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dl, lf2.data(), lf2.size() * 8, cudaMemcpyHostToDevice));
  const BkSplitGeom g = bk_split_geometry(M, N, K, dev_sms);
  // the lockstep kernel's inputs: genes by decreasing total count, stored by position
  BkSplitArgs a1;
  {
    std::vector<double> cs(N, 0.0);
    for (int i = 0; i < N; ++i)
      for (long long j = 0; j < M; ++j) cs[i] += Xt[(size_t)i * Mpad + j];
    std::vector<int> order(N);
    for (int i = 0; i < N; ++i) order[i] = i;
    if (!getenv("PROBE_NO_ORDER"))
      std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cs[x] > cs[y]; });
    std::vector<unsigned char> km(N);
    std::vector<float> Xs((size_t)N * Mpad), As((size_t)K * N);
    for (int pos = 0; pos < N; ++pos) {
      const int i = order[pos];
      int best = 0;
      for (int k = 1; k < K; ++k)
        if (A32[(size_t)k * N + i] > A32[(size_t)best * N + i]) best = k;
      km[pos] = (unsigned char)best;
      for (int k = 0; k < K; ++k) As[(size_t)k * N + pos] = A32[(size_t)k * N + i];
      memcpy(&Xs[(size_t)pos * Mpad], &Xt[(size_t)i * Mpad], Mpad * sizeof(float));
    }
    int* dorder;
    unsigned char* dkm;
    float *dXs, *dAs;
    CK(cudaMalloc(&dorder, N * 4));
    CK(cudaMalloc(&dkm, N));
    CK(cudaMalloc(&dXs, Xs.size() * 4));
    CK(cudaMalloc(&dAs, As.size() * 4));
    CK(cudaMemcpy(dorder, order.data(), N * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dkm, km.data(), N, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dXs, Xs.data(), Xs.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dAs, As.data(), As.size() * 4, cudaMemcpyHostToDevice));
    memset(&a1, 0, sizeof a1);
    a1.order = dorder;
    if (!getenv("PROBE_NO_KMAX")) a1.kmax = dkm;
    a1.Xt = dXs;
    a1.A32 = dAs;
  }
  {
    const char* cap = getenv("PROBE_COLD_CAP");
    a.cold_cap = cap ? (unsigned int)atoll(cap) : 4096u;
    CK(cudaMalloc(&a.cold_list, (size_t)a.cold_cap * sizeof(uint4)));
    CK(cudaMalloc(&a.cold_count, sizeof(unsigned int)));
  }
  a.Xt = dXt;
  a.A32 = dA;
  a.W32g = dW;
  a.lf2 = dl;
  a.n_cur = dn;
  a.sumI = dI;
  a.M = M;
  a.Mpad = Mpad;
  a.N = N;
  a.K = K;
  a.ngroups = g.ngroups;
  a.gpc = g.gpc;
  a.genes_per_cta = g.genes_per_cta;
  a.inv_sample = 1.0;
  a.key = 0x1234567ull;
  a.rk = philox_keys(a.key);
  {
    BkSplitArgs t = a;
    t.order = a1.order;
    t.kmax = a1.kmax;
    t.Xt = a1.Xt;
    t.A32 = a1.A32;
    a1 = t;
  }
  printf("geometry: gpc %d rows %d gx %d genes/cta %d smem %zu\n", g.gpc, g.rows, g.gx, g.genes_per_cta, g.smem);
  const double msplit = (double)M * N / 1e6;

  for (int variant = 0; variant < 2; ++variant) {
    float t0, t1;
    if (variant == 0) {
      t0 = run<0, false>(a, g, reps, 0, MK, KN);
      t1 = run<0, false>(a, g, reps, 1, MK, KN);
    } else {
      t0 = run<1, false>(a1, g, reps, 0, MK, KN);
      t1 = run<1, false>(a1, g, reps, 1, MK, KN);
    }
    printf("variant %d: burn-in sweep %.1f us (%.1f us / 1e6 splits), retained sweep %.1f us (%.1f)\n",
           variant, t0 * 1e3, t0 * 1e3 / msplit, t1 * 1e3, t1 * 1e3 / msplit);
    // moments of the last launch (retained): per (sample, type) and per (gene, type) sums
    std::vector<double> hn(MK), hI(KN);
    CK(cudaMemcpy(hn.data(), dn, MK * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hI.data(), dI, KN * 8, cudaMemcpyDeviceToHost));
    std::vector<double> en(MK, 0), vn(MK, 0), eI(KN, 0), vI(KN, 0);
    double cons = 0;
    for (long long j = 0; j < M; ++j) {
      double rowx = 0, rown = 0;
      for (int i = 0; i < N; ++i) {
        const double x = Xt[(size_t)i * Mpad + j];
        rowx += x;
        if (x == 0) continue;
        double aw = 0;
        for (int k = 0; k < K; ++k) aw += (double)W32g[((size_t)(j >> 5) * K + k) * 32 + (j & 31)] * A32[(size_t)k * N + i];
        for (int k = 0; k < K; ++k) {
          const double p = (double)W32g[((size_t)(j >> 5) * K + k) * 32 + (j & 31)] * A32[(size_t)k * N + i] / aw;
          en[j * K + k] += x * p;
          vn[j * K + k] += x * p * (1 - p);
          eI[(size_t)k * N + i] += x * p;
          vI[(size_t)k * N + i] += x * p * (1 - p);
        }
      }
      for (int k = 0; k < K; ++k) rown += hn[j * K + k];
      cons = std::max(cons, std::fabs(rowx - rown));
    }
    double zmax = 0, z2 = 0;
    size_t cnt = 0;
    for (size_t t = 0; t < MK; ++t)
      if (vn[t] > 0) {
        const double z = (hn[t] - en[t]) / std::sqrt(vn[t]);
        zmax = std::max(zmax, std::fabs(z));
        z2 += z * z;
        ++cnt;
      }
    double zmaxI = 0, z2I = 0;
    size_t cntI = 0;
    for (size_t t = 0; t < KN; ++t)
      if (vI[t] > 1) {
        const double z = (hI[t] - eI[t]) / std::sqrt(vI[t]);
        zmaxI = std::max(zmaxI, std::fabs(z));
        z2I += z * z;
        ++cntI;
      }
    printf("  conservation max |sum_k n_jk - sum_i X_ji| = %.1f;  per (sample,type): max|z| %.2f mean z^2 %.3f (%zu);"
           "  per (gene,type): max|z| %.2f mean z^2 %.3f (%zu)\n",
           cons, zmax, z2 / cnt, cnt, zmaxI, z2I / cntI, cntI);
  }
  // event counters of the lockstep variant
  CK(cudaMemset(dprobe, 0, 16 * 8));
  a1.probe = dprobe;
  run<1, true>(a1, g, 1, 1, MK, KN);
  unsigned long long pr[16];
  CK(cudaMemcpy(pr, dprobe, sizeof pr, cudaMemcpyDeviceToHost));
  const double nb = 3.0 * (double)M * N * (K - 1);  // 3 launches
  unsigned int cold_n = 0;
  CK(cudaMemcpy(&cold_n, a.cold_count, 4, cudaMemcpyDeviceToHost));
  printf("lockstep probe (fractions of all binomials): zero-search %.4f  modal %.4f  handed over: deep %.5f  redraws %.2e;"
         "  cold list of the last sweep %u (capacity %u)\n",
         pr[0] / nb, pr[1] / nb, pr[2] / nb, pr[3] / nb, cold_n, a.cold_cap);
  return 0;
}
