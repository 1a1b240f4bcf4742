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
#include <vector>

namespace ursm {

constexpr int kMaxPeers = 16;  // ranks of one NVLink domain that exchange partials directly

struct Mle {
  ScShard* sc = nullptr;
  BkShard* bk = nullptr;
  int N = 0, K = 0;
  double min_A = 1e-6;
  DevBuf<double> A, y;                   // [K][N] columns; scratch of the closed-form rule
  const double* cAinv = nullptr;         // [K][N] caller-owned (all-reduced)
  const double* cA = nullptr;
  const double* coeffA = nullptr;
  DevBuf<double> u, ratio;               // [L] u_l; partial sums of the u pass (u_acc)
  DevBuf<int> act_cur, act_next;         // [K] columns in opt_Ak at the start / end of an iteration
  DevBuf<int> type_start;                // [K+1] offsets of the types in the sorted cell order
  DevBuf<int> niter, base, halv;         // [K] outer iterations, halvings already tried, total halvings
  DevBuf<double> cand;                   // [K][kCand][N] candidate columns of one backtracking batch
  DevBuf<double> cres;                   // [K][kCand][4]: accepted?, objective, ||dA||_1, step
  DevBuf<double> delta;                  // [K] last accepted ||dA_k||_1
  DevBuf<double> terms;                  // [4]
  double init_step = 0, conv = 0;
  int maxiter = 0;
  int num_sms = 148;
  // peer-memory exchange of the per-iteration partials (ursm_mle_p2p_*)
  struct P2PBlock* p2p = nullptr;
  // device-resident loop of opt_Ak (ursm_mle_opt_run): its own stream, the captured graph of
  // kGraphIters outer iterations, double-buffered partials, per-column arrival counters
  cudaStream_t gstream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  cudaGraphExec_t gexec = nullptr;
  struct GraphKey {
    double init_step = 0, conv = 0;
    int maxiter = 0, ncand = 0, p2p = 0, cnt_w = 0, sample = 0, y_smem = 0;
    const void* coeffs = nullptr;
    const void* p2p_own = nullptr;
    const void* act = nullptr;
    const void* cnt = nullptr;
    bool operator==(const GraphKey& o) const {
      return init_step == o.init_step && conv == o.conv && maxiter == o.maxiter && ncand == o.ncand &&
             p2p == o.p2p && cnt_w == o.cnt_w && sample == o.sample && y_smem == o.y_smem &&
             coeffs == o.coeffs && p2p_own == o.p2p_own && act == o.act && cnt == o.cnt;
    }
  } gkey;
  DevBuf<double> partbuf;                // [2][K*N + K] (single rank) / [1][...] summed partials (peers)
  DevBuf<int> done;                      // [K] cand CTAs of a column that have finished
  DevBuf<double> lam_store;              // [K][kCand] projection multipliers of the previous iteration
  DevBuf<double> gbuf, gobj;             // [K][N] gradient; [K][ceil(N/256)][3] objective partials
  DevBuf<unsigned long long> epoch_dev;  // [1] epoch of the peer exchange at the start of a graph launch
  int* h_poll = nullptr;                 // pinned: [K] act | [K] niter | [K] halv
  ~Mle();
};

// One cudaMalloc'd, IPC-shared block per rank = two (KN + K)-double buffers + one epoch
// word per peer, with the peers' mappings.  Blocks are cached for the life of the process,
// keyed by size: the reference's API builds a new M-step object per fit, and mapping seven
// peers' handles costs milliseconds each.  The epoch counter lives with the block, so it
// stays monotonic across the workspaces that use it one after the other.
struct P2PBlock {
  size_t bytes = 0;
  void* own = nullptr;
  void* peer[kMaxPeers] = {nullptr};
  int world = 0, rank = 0;
  unsigned long long epoch = 0;
  int* err = nullptr;
  int id = 0;           // index in g_p2p_blocks: the ranks of a job must agree on it
  bool opened = false, in_use = false;
  bool dirty = false;   // a barrier timed out on this block: flags / epoch must be reset
};
static std::vector<P2PBlock*> g_p2p_blocks;

Mle::~Mle() {
  if (p2p) p2p->in_use = false;
  if (gexec) cudaGraphExecDestroy(gexec);
  if (ev_in) cudaEventDestroy(ev_in);
  if (ev_out) cudaEventDestroy(ev_out);
  if (gstream) cudaStreamDestroy(gstream);
  if (h_poll) cudaFreeHost(h_poll);
}

constexpr int kGraphIters = 16;  // outer iterations per graph launch (even: buffer parity is static)

constexpr int kCand = 16;  // trial step sizes of backtracking evaluated concurrently

constexpr int kColThreads = 256;

// small unsigned integer -> double without the (quarter-rate) I2F.F64: put it in the low
// mantissa word of 2^52 and subtract 2^52 (one DADD)
__device__ __forceinline__ double small_uint_to_double(unsigned int c) {
  return __hiloint2double(0x43300000, (int)c) - 4503599627370496.0;
}

// Counters are read 8 bytes at a time: C = 4 (uint16) or 8 (uint8) adjacent genes.
template <class CT>
struct CntVec {
  static constexpr int C = 8 / (int)sizeof(CT);
  static __device__ __forceinline__ void unpack(const uint2 v, double (&e)[C]) {
    if constexpr (sizeof(CT) == 2) {
      e[0] = small_uint_to_double(__byte_perm(v.x, 0u, 0x4410));
      e[1] = small_uint_to_double(__byte_perm(v.x, 0u, 0x4432));
      e[2] = small_uint_to_double(__byte_perm(v.y, 0u, 0x4410));
      e[3] = small_uint_to_double(__byte_perm(v.y, 0u, 0x4432));
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        e[q] = small_uint_to_double(__byte_perm(v.x, 0u, 0x4440 + q));
        e[4 + q] = small_uint_to_double(__byte_perm(v.y, 0u, 0x4440 + q));
      }
    }
  }
};

// MODE 0: out[type][i] += Y_li * E[S_li]            (m_step.py:88-89)
// MODE 1: out[type][i] += E[S_li] * R_l / u_l       (m_step.py:123-125), active types only
// A thread owns C adjacent genes (one 64-bit load of counters per cell) and walks a
// chunk of the type-sorted cells.  The per-cell metadata of the chunk (row index, weight
// R_l/u_l or 0 for a skipped row, type) is staged in shared memory once per CTA, so the
// inner loop is: broadcast LDS of (row, weight), one 8-byte load, C x (byte permute, DADD,
// DFMA); kColDepth independent loads are in flight per thread.  MODE 1 finishes opt_u
// from mle_u_kernel's partial sums while staging.  HBM: 1 or 2 B per cell x gene (MODE 0 adds the 4 B count).
constexpr int kColDepth = 8;
constexpr int kColChunkMax = 512;  // cells per CTA (shared-memory staging)

template <class CT, int MODE>
__global__ void __launch_bounds__(kColThreads)
mle_colsum_kernel(const float* Y, const CT* cnt, const int* G, const int* order,
                  const double* u_acc, const double* R, double* u, const int* active, long long L,
                  int N, int chunk, double inv_sample, double* out) {
  constexpr int C = CntVec<CT>::C;
  __shared__ int row_s[kColChunkMax];
  __shared__ int type_s[kColChunkMax];
  __shared__ double w_s[kColChunkMax];
  const long long c0 = (long long)blockIdx.y * chunk;
  const int n = (int)(min(L, c0 + chunk) - c0);
  for (int t = threadIdx.x; t < n; t += kColThreads) {
    const int l = order[c0 + t];
    const int ty = G[l];
    const bool use = MODE == 0 || active[ty];
    row_s[t] = l;
    type_s[t] = ty;
    double w = use ? inv_sample : 0.0;
    if (MODE == 1 && use) {  // opt_u (:108-115): u_l, and the weight R_l / u_l of the row
      const double ul = u_acc[l] * inv_sample;
      if (blockIdx.x == 0) u[l] = ul;
      w = (R[l] / ul) * inv_sample;
    }
    w_s[t] = w;
  }
  __syncthreads();
  const int i = C * (blockIdx.x * kColThreads + threadIdx.x);
  if (i >= N) return;
  const bool vec = (i + C - 1 < N) && ((N % C) == 0);  // 64-bit loads need 8-byte aligned rows
  double acc[C];
#pragma unroll
  for (int q = 0; q < C; ++q) acc[q] = 0.0;
  int cur = n > 0 ? type_s[0] : -1;
  auto flush = [&](int next) {
#pragma unroll
    for (int q = 0; q < C; ++q) {
      if (acc[q] != 0.0 && i + q < N) atomicAdd(&out[(size_t)cur * N + i + q], acc[q]);
      acc[q] = 0.0;
    }
    cur = next;
  };
  auto loadv = [&](int l) {
    const CT* row = cnt + (size_t)l * N + i;
    if (vec) return *reinterpret_cast<const uint2*>(row);
    unsigned long long bits = 0ull;  // ragged tail: assemble the same 8 bytes from scalars
#pragma unroll
    for (int q = 0; q < C; ++q)
      if (i + q < N) bits |= (unsigned long long)row[q] << (8 * (int)sizeof(CT) * q);
    return make_uint2((unsigned int)bits, (unsigned int)(bits >> 32));
  };
  for (int t0 = 0; t0 < n; t0 += kColDepth) {
    uint2 v[kColDepth];
#pragma unroll
    for (int q = 0; q < kColDepth; ++q) {
      const int t = t0 + q;
      v[q] = (t < n && w_s[t] != 0.0) ? loadv(row_s[t]) : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int q = 0; q < kColDepth; ++q) {
      const int t = t0 + q;
      if (t >= n) break;
      if (type_s[t] != cur) flush(type_s[t]);
      const double w = w_s[t];
      if (w == 0.0) continue;
      double e[C];
      CntVec<CT>::unpack(v[q], e);
      if (MODE == 0) {
        const float* y = Y + (size_t)row_s[t] * N + i;
#pragma unroll
        for (int r = 0; r < C; ++r)
          if (i + r < N) acc[r] = fma((double)y[r] * e[r], w, acc[r]);
      } else {
#pragma unroll
        for (int r = 0; r < C; ++r) acc[r] = fma(e[r], w, acc[r]);
      }
    }
  }
  flush(-1);
}

// u_l = sum_i A[i, G_l] E[S_li] for the cells of active types (opt_u, :108-115), as
// partial dot products summed into u_acc (zeroed by the caller); mle_colsum_kernel<., 1>
// turns them into u_l and the weights R_l / u_l when it stages its cells.
// A CTA owns a block of consecutive cells of the TYPE-SORTED order and one segment of the
// gene axis.  For every run of same-type cells in its block it stages that type's segment
// of A in shared memory once (split into the (a0,a1) and (a2,a3) halves of each group of
// four genes, so that the 16-byte reads are conflict free) and its warps then stream the
// counters: a warp takes kURows cells at a time, a lane owns four adjacent genes per step
// (one 4- or 8-byte counter load per cell), kUDepth steps in flight.  A is read from L2
// once per CTA and run instead of once per cell; counter -> double is a byte permute and
// one DADD (small_uint_to_double).
constexpr int kURows = 4, kUDepth = 4, kUThreads = 256;
constexpr int kUSegGenes = 4096;  // genes per segment: 32 KB of A in shared memory

// four adjacent counters as one raw 4- or 8-byte load (kept packed while in flight)
template <class CT>
struct Raw4 {
  uint2 v;
  __device__ __forceinline__ void load(const CT* p) {
    if constexpr (sizeof(CT) == 1) {
      v.x = *reinterpret_cast<const unsigned int*>(p);
      v.y = 0u;
    } else {
      v = *reinterpret_cast<const uint2*>(p);
    }
  }
  __device__ __forceinline__ double get(int r) const {
    if constexpr (sizeof(CT) == 1) return small_uint_to_double(__byte_perm(v.x, 0u, 0x4440 + r));
    return small_uint_to_double(__byte_perm(r < 2 ? v.x : v.y, 0u, (r & 1) ? 0x4432 : 0x4410));
  }
};

template <class CT>
__global__ void __launch_bounds__(kUThreads)
mle_u_kernel(const CT* cnt, const int* order, const int* type_start, const double* A,
             const int* active, long long L, int N, int K, int rows_per_cta, int seg_genes,
             double* u_acc) {
  __shared__ double2 A01[kUSegGenes / 4], A23[kUSegGenes / 4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = kUThreads >> 5;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(L, r0 + rows_per_cta);
  const int g_lo = blockIdx.y * seg_genes, g_hi = min(N, g_lo + seg_genes);
  const bool vec = (N & 3) == 0;  // rows keep the alignment of the vector loads
  const int nv = (g_hi - g_lo) >> 2;  // groups of four genes in this segment (vec only)
  for (int k = 0; k < K; ++k) {
    const long long lo = max(r0, (long long)type_start[k]);
    const long long hi = min(r1, (long long)type_start[k + 1]);
    if (lo >= hi || !active[k]) continue;
    const double* Ak = A + (size_t)k * N;
    if (vec) {
      __syncthreads();  // previous run's readers are done
      const double2* src = reinterpret_cast<const double2*>(Ak + g_lo);
      for (int v = tid; v < nv; v += kUThreads) {
        A01[v] = src[2 * v];
        A23[v] = src[2 * v + 1];
      }
      __syncthreads();
    }
    for (long long c0 = lo + (long long)warp * kURows; c0 < hi; c0 += (long long)nwarp * kURows) {
      int l[kURows];
      double s[kURows];
#pragma unroll
      for (int q = 0; q < kURows; ++q) {
        l[q] = c0 + q < hi ? order[c0 + q] : -1;
        s[q] = 0.0;
      }
      if (vec) {
        const CT* row[kURows];
#pragma unroll
        for (int q = 0; q < kURows; ++q) row[q] = cnt + (size_t)(l[q] >= 0 ? l[q] : l[0]) * N + g_lo;
        for (int v0 = lane; v0 < nv; v0 += 32 * kUDepth) {
          Raw4<CT> c[kUDepth][kURows];
#pragma unroll
          for (int t = 0; t < kUDepth; ++t) {
            const int v = min(v0 + 32 * t, nv - 1);  // clamped; out-of-range steps are skipped
#pragma unroll
            for (int q = 0; q < kURows; ++q) c[t][q].load(row[q] + 4 * (size_t)v);
          }
#pragma unroll
          for (int t = 0; t < kUDepth; ++t) {
            const int v = v0 + 32 * t;
            if (v >= nv) break;
            const double2 a01 = A01[v], a23 = A23[v];
#pragma unroll
            for (int q = 0; q < kURows; ++q) {
              s[q] = fma(a01.x, c[t][q].get(0), s[q]);
              s[q] = fma(a01.y, c[t][q].get(1), s[q]);
              s[q] = fma(a23.x, c[t][q].get(2), s[q]);
              s[q] = fma(a23.y, c[t][q].get(3), s[q]);
            }
          }
        }
      } else {  // ragged rows: scalar loads
#pragma unroll
        for (int q = 0; q < kURows; ++q) {
          if (l[q] < 0) continue;
          const CT* row = cnt + (size_t)l[q] * N;
          double acc = 0.0;
          for (int i = g_lo + lane; i < g_hi; i += 32)
            acc = fma(Ak[i], small_uint_to_double((unsigned int)row[i]), acc);
          s[q] = acc;
        }
      }
#pragma unroll
      for (int q = 0; q < kURows; ++q) {
        const double tot = warp_sum(s[q]);
        if (lane == 0 && l[q] >= 0) atomicAdd(&u_acc[l[q]], tot);
      }
    }
  }
}

// opt_u (:108-115) + update_gd_coeffConst (:118-125) in ONE pass for gene vectors short enough
// that a block of cells' counter rows fits in shared memory (N <= 8192: c2, c5): a CTA takes
// `nc` consecutive cells of the type-sorted order; per run of same-type cells, (1) its warps
// stream the rows once from HBM / L2 -- 8 bytes per lane and load -- into shared memory while
// accumulating the row dot with A[:, type] (u_l), (2) the threads then own groups of 8 bytes of
// adjacent counters and sum the staged rows, weighted by R_l / u_l, into one atomic per gene and
// CTA.  The counters cross the memory system once per outer iteration (the two-kernel path reads
// them twice) and there is one launch instead of two.
constexpr int kFusedThreads = 512;

template <class CT>
__global__ void __launch_bounds__(kFusedThreads, 2)
mle_pass_fused_kernel(const CT* cnt, const int* order, const int* type_start, const double* A,
                      const int* active, const double* R, double* u, long long L, int N, int K,
                      int nc, double inv_sample, double* out) {
  constexpr int C8 = 8 / (int)sizeof(CT);    // counters per 8-byte load / group
  extern __shared__ __align__(16) unsigned char fused_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = kFusedThreads >> 5;
  const size_t row_bytes = (size_t)N * sizeof(CT);  // a multiple of 8, N even (host checks)
  double* w_s = reinterpret_cast<double*>(fused_smem);           // [nc] R_l / u_l / sample
  double* dot_s = w_s + nc;                                       // [nc][2] half-row dot products
  double2* A_s = reinterpret_cast<double2*>(fused_smem + (((size_t)nc * 24 + 15) & ~(size_t)15));  // [C8/2][nv]
  unsigned char* rows = reinterpret_cast<unsigned char*>(A_s + N / 2);  // [nc][row_bytes]
  const long long c_lo = (long long)blockIdx.x * nc, c_hi = min(L, c_lo + nc);
  const int nvA = (int)(row_bytes >> 3);
  for (int k = 0; k < K; ++k) {
    const long long lo = max(c_lo, (long long)type_start[k]);
    const long long hi = min(c_hi, (long long)type_start[k + 1]);
    if (lo >= hi || !active[k]) continue;
    {
      // the column A[:, k] as C8/2 arrays of pairs (pair q of group v at [q][v]): consecutive
      // lanes read consecutive 16 bytes in the dot below -- conflict free, and no global load
      // sits between the counter loads and their use
      const double2* src = reinterpret_cast<const double2*>(A + (size_t)k * N);
      for (int j = tid; j < N / 2; j += kFusedThreads) {
        const int v = j / (C8 / 2), q = j - v * (C8 / 2);
        A_s[q * nvA + v] = src[j];
      }
      __syncthreads();
    }
    // ---- phase 1: stage the rows, row dots.  Two warps per row (interleaved 8-byte loads,
    // eight in flight per lane: the pass is a chain of memory round trips, not arithmetic)
    for (long long cb = lo; cb < hi; cb += nwarp / 2) {
      const long long c = cb + (warp >> 1);
      const int half = warp & 1;
      if (c < hi) {
        const int l = order[c];
        const uint2* src = reinterpret_cast<const uint2*>(cnt + (size_t)l * N);
        uint2* dst = reinterpret_cast<uint2*>(rows + (size_t)(c - c_lo) * row_bytes);
        const int nv = (int)(row_bytes >> 3);
        double acc = 0.0, acc2 = 0.0;
        for (int v0 = half * 32 + lane; v0 < nv; v0 += 64 * 8) {
          uint2 x[8];
#pragma unroll
          for (int h = 0; h < 8; ++h) {
            const int v = v0 + 64 * h;
            x[h] = v < nv ? src[v] : make_uint2(0u, 0u);
          }
#pragma unroll
          for (int h = 0; h < 8; ++h) {
            const int v = v0 + 64 * h;
            if (v >= nv) break;
            dst[v] = x[h];
            double e[C8];
            CntVec<CT>::unpack(x[h], e);
#pragma unroll
            for (int q = 0; q < C8 / 2; ++q) {
              const double2 p = A_s[q * nvA + v];
              acc = fma(p.x, e[2 * q], acc);
              acc2 = fma(p.y, e[2 * q + 1], acc2);
            }
          }
        }
        acc = warp_sum(acc + acc2);
        if (lane == 0) dot_s[2 * (c - c_lo) + half] = acc;
      }
    }
    __syncthreads();
    for (long long c = lo + tid; c < hi; c += kFusedThreads) {
      const int l = order[c];
      const double ul = (dot_s[2 * (c - c_lo)] + dot_s[2 * (c - c_lo) + 1]) * inv_sample;
      u[l] = ul;
      w_s[c - c_lo] = (R[l] / ul) * inv_sample;
    }
    __syncthreads();
    // ---- phase 2: weighted column sums of the staged rows
    const int ngrp = (int)(row_bytes >> 3);
    for (int gq = tid; gq < ngrp; gq += kFusedThreads) {
      double acc[C8];
#pragma unroll
      for (int q = 0; q < C8; ++q) acc[q] = 0.0;
      for (long long c = lo; c < hi; ++c) {
        const uint2 v = *reinterpret_cast<const uint2*>(rows + (size_t)(c - c_lo) * row_bytes +
                                                        (size_t)gq * 8);
        const double w = w_s[c - c_lo];
        double e[C8];
        CntVec<CT>::unpack(v, e);
#pragma unroll
        for (int q = 0; q < C8; ++q) acc[q] = fma(e[q], w, acc[q]);
      }
      double* o = out + (size_t)k * N + (size_t)gq * C8;
#pragma unroll
      for (int q = 0; q < C8; ++q)
        if (acc[q] != 0.0) atomicAdd(o + q, acc[q]);
    }
    __syncthreads();
  }
}

// The same pass for SHORT gene vectors (at most 256 eight-byte groups per row: N <= 2048 with
// byte counters, the c5 shape), where the kernel above leaves half of its 512 threads without a
// group.  256 threads, one warp per row, thread t owns group t; PERSISTENT CTAs: a CTA walks a
// contiguous range of the type-sorted cells in blocks of `nc` rows and keeps (a) the column
// A[:, type] staged in shared memory and (b) its column sums in registers for as long as the type
// does not change -- one fp64 atomic per gene, CTA and run instead of one per block (at c5 the
// atomics of 2,451 blocks x 2,000 genes were a cost of their own).
constexpr int kSmallThreads = 256;

template <class CT>
__global__ void __launch_bounds__(kSmallThreads, 4)
mle_pass_fused_small_kernel(const CT* cnt, const int* order, const int* type_start, const double* A,
                            const int* active, const double* R, double* u, long long L, int N, int K,
                            int nc, long long cells_per_cta, double inv_sample, double* out) {
  constexpr int C8 = 8 / (int)sizeof(CT);
  constexpr int NW = kSmallThreads / 32;
  extern __shared__ __align__(16) unsigned char fused_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t row_bytes = (size_t)N * sizeof(CT);
  const int nv = (int)(row_bytes >> 3);  // <= 256 (host checks)
  double* w_s = reinterpret_cast<double*>(fused_smem);   // [nc] R_l / u_l / sample
  double* dot_s = w_s + nc;                               // [nc] row dot products
  double* A_s = reinterpret_cast<double*>(fused_smem + (((size_t)nc * 16 + 15) & ~(size_t)15));  // [N]
  unsigned char* rows = reinterpret_cast<unsigned char*>(A_s + N);  // [nc][row_bytes]
  const long long cta_lo = (long long)blockIdx.x * cells_per_cta;
  const long long cta_hi = min(L, cta_lo + cells_per_cta);
  const bool own = tid < nv;
  double acc[C8];
#pragma unroll
  for (int q = 0; q < C8; ++q) acc[q] = 0.0;
  int cur = -1;  // type whose column is staged and whose sums are in `acc`
  int k0 = 0;    // first type that can still have cells at or after the block in hand
  for (long long blk_lo = cta_lo; blk_lo < cta_hi; blk_lo += nc) {
    const long long blk_hi = min(cta_hi, blk_lo + nc);
    while (k0 < K - 1 && (long long)type_start[k0 + 1] <= blk_lo) ++k0;
    for (int k = k0; k < K && (long long)type_start[k] < blk_hi; ++k) {
      const long long lo = max(blk_lo, (long long)type_start[k]);
      const long long hi = min(blk_hi, (long long)type_start[k + 1]);
      if (lo >= hi || !active[k]) continue;
      if (k != cur) {
        if (cur >= 0 && own) {
          double* o = out + (size_t)cur * N + (size_t)tid * C8;
#pragma unroll
          for (int q = 0; q < C8; ++q) {
            if (acc[q] != 0.0) atomicAdd(o + q, acc[q]);
            acc[q] = 0.0;
          }
        }
        cur = k;
        __syncthreads();  // readers of the previous column are done
        // the column as C8/2 arrays of pairs, pair q of group v at [q][v]: the lanes of a warp
        // (consecutive v) then read consecutive 16 bytes -- conflict free
        const double2* src = reinterpret_cast<const double2*>(A + (size_t)k * N);
        double2* dst = reinterpret_cast<double2*>(A_s);
        for (int j = tid; j < N / 2; j += kSmallThreads) {
          const int v = j / (C8 / 2), q = j - v * (C8 / 2);
          dst[q * nv + v] = src[j];
        }
      }
      __syncthreads();  // column staged; the previous run's rows / weights are no longer read
      // ---- (1) stage the rows, row dots: one warp per row, eight 8-byte loads in flight per lane
      for (long long c = lo + warp; c < hi; c += NW) {
        const int l = order[c];
        const uint2* src = reinterpret_cast<const uint2*>(cnt + (size_t)l * N);
        uint2* dst = reinterpret_cast<uint2*>(rows + (size_t)(c - blk_lo) * row_bytes);
        uint2 x[8];
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const int v = lane + 32 * h;
          x[h] = v < nv ? src[v] : make_uint2(0u, 0u);
        }
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const int v = lane + 32 * h;
          if (v < nv) {
            dst[v] = x[h];
            double e[C8];
            CntVec<CT>::unpack(x[h], e);
            const double2* a2 = reinterpret_cast<const double2*>(A_s) + v;
#pragma unroll
            for (int q = 0; q < C8 / 2; ++q) {
              const double2 p = a2[q * nv];
              d0 = fma(p.x, e[2 * q], d0);
              d1 = fma(p.y, e[2 * q + 1], d1);
            }
          }
        }
        const double d = warp_sum(d0 + d1);
        if (lane == 0) dot_s[c - blk_lo] = d;
      }
      __syncthreads();
      for (long long c = lo + tid; c < hi; c += kSmallThreads) {
        const int l = order[c];
        const double ul = dot_s[c - blk_lo] * inv_sample;
        u[l] = ul;
        w_s[c - blk_lo] = (R[l] / ul) * inv_sample;
      }
      __syncthreads();
      // ---- (2) weighted column sums of the staged rows, into the thread's registers
      if (own) {
        const unsigned char* col = rows + (size_t)tid * 8;
        for (long long c = lo; c < hi; ++c) {
          const uint2 v = *reinterpret_cast<const uint2*>(col + (size_t)(c - blk_lo) * row_bytes);
          const double w = w_s[c - blk_lo];
          double e[C8];
          CntVec<CT>::unpack(v, e);
#pragma unroll
          for (int q = 0; q < C8; ++q) acc[q] = fma(e[q], w, acc[q]);
        }
      }
    }
  }
  if (cur >= 0 && own) {
    double* o = out + (size_t)cur * N + (size_t)tid * C8;
#pragma unroll
    for (int q = 0; q < C8; ++q)
      if (acc[q] != 0.0) atomicAdd(o + q, acc[q]);
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

// lambda such that sum_i max(y_i + lambda, m) = 1  (utils.py:8-31 without the sort).
// Michelot's active-set iteration = Newton's method on the convex piecewise-linear
// f(lambda) = sum_i max(y_i + lambda, m) - 1: from ANY start the first step lands at or above
// the root and the iterates then decrease to it in finitely many steps; the active sets
// {y_i + lambda > m} are threshold sets, hence nested, so an unchanged count means an unchanged
// set and the lambda computed from it is final.  `warm` (may be null / NaN) is a guess -- the
// lambda of the same trial in the previous outer iteration -- that replaces the cold start
// (every entry active) and saves most of the passes; the result is the same lambda, summed in
// the same order.
__device__ double proj_lambda(const double* y, int N, double m, double* red,
                              const double* warm = nullptr) {
  double s = 0.0, c = 0.0;
  double lam = warm ? *warm : nan("");
  double prev = -1.0;
  bool cold = !(lam == lam);
  if (cold) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) s += y[i];
    s = block_sum(s, red);
    lam = (1.0 - s) / N;
    prev = N;
  }
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
    if (c == 0.0) {
      if (cold) break;  // degenerate (N*m >= 1): everything sits on the floor
      // a guess so low that nothing is active: start over from the cold start
      s = 0.0;
      for (int i = threadIdx.x; i < N; i += blockDim.x) s += y[i];
      s = block_sum(s, red);
      lam = (1.0 - s) / N;
      prev = N;
      cold = true;
      continue;
    }
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
  int y_smem;          // trial column lives in dynamic shared memory (N doubles)
  // fused loop (ursm_mle_opt_run; all null / 0 in the step-by-step entry points):
  double* Aw;          // [K][N] writable alias of A: the last CTA of a column installs the winner
  double* zero_out;    // [K*N + K] partials buffer of the pass that follows: zeroed here
  double* u_acc;       // [L] partial dot products of the pass that follows: zeroed here
  const int* order;    // [L] cells sorted by type
  const int* type_start;  // [K+1]
  int* done;           // [K] arrival counters
  int* act_w;          // [K] loop state, advanced in place by the last CTA of a column
  int *niter, *halv, *base_w;
  double* delta;
  double* lam_store;   // [K][kCand] lambda of each trial in the previous outer iteration (warm start)
  const double* gbuf;  // [K][N] gradient of the current columns (mle_grad_kernel), or null
  const double* gobj;  // [K][gslots][3] block partials of the current objective
  int gslots;
  double conv;
  int maxiter, ncand, K;
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

// What every trial of a column shares (fused loop): the gradient -get_grad_A (:262-272) and the
// three sums of the current objective -get_obj_A (:275-287).  One thread per gene over the whole
// machine instead of once per trial CTA (16 x the fp64 log / divide work on 1/16 of the SMs);
// the block partials go to fixed slots that the trial CTAs add up in a fixed order.
constexpr int kGradThreads = 256;

__global__ void __launch_bounds__(kGradThreads)
mle_grad_kernel(const double* A, const double* cAinv, const double* cA, const double* coeffA,
                const double* part, const int* act, int N, double* gbuf, double* gobj) {
  __shared__ double red[6 * 32];
  const int k = blockIdx.y;
  if (!act[k]) return;
  const int i = blockIdx.x * kGradThreads + threadIdx.x;
  double v[6] = {0, 0, 0, 0, 0, 0};
  if (i < N) {
    const size_t e = (size_t)k * N + i;
    const double x = A[e], cc = coeffA[e] - part[e];
    gbuf[e] = -((cAinv[e] / x + cA[e] * x + cc) * (1.0 / N));
    v[0] = cAinv[e] * log(x);
    v[1] = cA[e] * x * x;
    v[2] = cc * x;
  }
  block_sum6(v, red);
  if (threadIdx.x == 0) {
    double* o = gobj + ((size_t)k * gridDim.x + blockIdx.x) * 3;
    o[0] = v[0];
    o[1] = v[1];
    o[2] = v[2];
  }
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
  __shared__ int last_s;
  extern __shared__ double cand_ysm[];
  const int k = blockIdx.y, c = blockIdx.x, N = a.N, tid = threadIdx.x, T = blockDim.x;
  if (a.zero_out && c == 0 && tid == 0) a.zero_out[(size_t)a.K * N + k] = 0.0;  // sum_l R_l log u_l
  if (!a.act[k]) return;
  if (a.zero_out) {
    // what the pass after this step accumulates into, for this column: slice c of the
    // partials row and of the per-cell dot products of the column's cells
    const int nc = gridDim.x;
    const int i0 = (int)((long long)N * c / nc), i1 = (int)((long long)N * (c + 1) / nc);
    for (int i = i0 + tid; i < i1; i += T) a.zero_out[(size_t)k * N + i] = 0.0;
    const long long t0 = a.type_start[k], t1 = a.type_start[k + 1];
    const long long l0 = t0 + (t1 - t0) * c / nc, l1 = t0 + (t1 - t0) * (c + 1) / nc;
    for (long long l = l0 + tid; l < l1; l += T) a.u_acc[a.order[l]] = 0.0;
  }
  const double* Ak = a.A + (size_t)k * N;
  double* yout = a.cand + ((size_t)k * kCand + c) * N;
  // the trial column is re-read by every pass of the projection: keep it on chip
  double* y = a.y_smem ? cand_ysm : yout;
  const double* ci = a.cAinv + (size_t)k * N;
  const double* ca = a.cA + (size_t)k * N;
  const double* c0 = a.coeffA + (size_t)k * N;
  const double* pt = a.part + (size_t)k * N;
  const double invN = 1.0 / N;
  const double step = ldexp(a.init_step, -(a.base[k] + c));
  const double* gk = a.gbuf ? a.gbuf + (size_t)k * N : nullptr;
  double v[6] = {0, 0, 0, 0, 0, 0};
  if (gk) {  // gradient and current objective precomputed once per column (mle_grad_kernel)
    for (int i = tid; i < N; i += T) y[i] = Ak[i] - step * gk[i];
    if (tid < 32) {
      const double* o = a.gobj + (size_t)k * a.gslots * 3;
      for (int q = tid; q < a.gslots; q += 32) {  // fixed order: lane q takes slots q, q+32, ...
        v[0] += o[3 * q];
        v[1] += o[3 * q + 1];
        v[2] += o[3 * q + 2];
      }
    }
  } else {
    for (int i = tid; i < N; i += T) {
      const double x = Ak[i], cc = c0[i] - pt[i];
      const double g = -((ci[i] / x + ca[i] * x + cc) * invN);
      y[i] = x - step * g;
      v[0] += ci[i] * log(x);
      v[1] += ca[i] * x * x;
      v[2] += cc * x;
    }
  }
  block_sum6(v, red);
  const double obj_old = -((v[0] + v[1] / 2.0 + v[2]) * invN);
  __syncthreads();
  double* lam_slot = a.lam_store ? a.lam_store + (size_t)k * kCand + c : nullptr;
  const double lam = proj_lambda(y, N, a.min_A, red, lam_slot);
  if (lam_slot && tid == 0) *lam_slot = lam;
  for (int q = 0; q < 6; ++q) v[q] = 0.0;
  for (int i = tid; i < N; i += T) {
    const double x = Ak[i], cc = c0[i] - pt[i];
    const double g = gk ? gk[i] : -((ci[i] / x + ca[i] * x + cc) * invN);
    const double nv = fmax(y[i] + lam, a.min_A);
    const double Gt = (x - nv) / step;
    yout[i] = nv;
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
  if (!a.done) return;
  // ---- fused loop: the LAST trial CTA of the column to finish decides (the first passing
  // trial = the step the sequential halving accepts, :209-233), advances the opt_Ak loop
  // state (:164-181) and installs the accepted column
  __threadfence();
  __syncthreads();
  if (tid == 0) last_s = (atomicAdd(&a.done[k], 1) == a.ncand - 1) ? 1 : 0;
  __syncthreads();
  if (!last_s) return;
  __threadfence();
  if (tid == 0) {
    a.done[k] = 0;
    int w = -1;
    for (int q = 0; q < a.ncand; ++q)
      if (a.cres[((size_t)k * kCand + q) * 4] != 0.0) {
        w = q;
        break;
      }
    if (w < 0 && a.base_w[k] + a.ncand >= 1024) w = a.ncand - 1;  // step underflow guard
    if (w < 0) {  // every trial of this batch was rejected: keep halving
      a.base_w[k] += a.ncand;
    } else {
      const double d = a.cres[((size_t)k * kCand + w) * 4 + 2];
      a.delta[k] = d;
      a.halv[k] += a.base_w[k] + w;
      a.base_w[k] = 0;
      const int n = ++a.niter[k];
      a.act_w[k] = (d > a.conv && n < a.maxiter) ? 1 : 0;
    }
    last_s = w;
  }
  __syncthreads();
  const int w = last_s;
  if (w < 0) return;
  const double* src = a.cand + ((size_t)k * kCand + w) * N;
  for (int i = tid; i < N; i += T) a.Aw[(size_t)k * N + i] = src[i];
}

// One CTA per column: thread 0 accepts the first passing trial of the batch (or schedules
// the next batch of halvings) and advances the opt_Ak loop state (:164-181); the CTA
// then installs the accepted trial column, A[k] <- cand[k][w].
__global__ void __launch_bounds__(1024)
mle_decide_copy_kernel(const double* cres, const int* act, int* act_next, int* niter, int* base,
                       int* halv, double* delta, double* A, const double* cand, int N,
                       double conv, int maxiter, int ncand) {
  __shared__ int win_s;
  const int k = blockIdx.x;
  if (!act[k]) {
    if (threadIdx.x == 0) act_next[k] = 0;
    return;
  }
  if (threadIdx.x == 0) {
    int w = -1;
    for (int c = 0; c < ncand; ++c)
      if (cres[((size_t)k * kCand + c) * 4] != 0.0) {
        w = c;
        break;
      }
    if (w < 0 && base[k] + ncand >= 1024) w = ncand - 1;  // step underflow guard
    if (w < 0) {  // every trial of this batch was rejected: keep halving
      base[k] += ncand;
      act_next[k] = 1;
    } else {
      const double d = cres[((size_t)k * kCand + w) * 4 + 2];
      delta[k] = d;
      halv[k] += base[k] + w;
      base[k] = 0;
      const int n = ++niter[k];
      act_next[k] = (d > conv && n < maxiter) ? 1 : 0;
    }
    win_s = w;
  }
  __syncthreads();
  const int w = win_s;
  if (w < 0) return;
  const double* src = cand + ((size_t)k * kCand + w) * N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) A[(size_t)k * N + i] = src[i];
}

// part[k][:] <- part_local[k][:] for the columns that were stepped; the per-type
// sum_l R_l log u_l tail is refreshed for every type.  (The loop state advances by the
// host swapping the act_cur / act_next buffers: no kernel.)
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

// The same installation fused with the cross-rank sum (one all-reduce of
// N*K + K doubles per outer iteration), over NVLink peer memory instead of a collective
// call.  Every rank's partials sit in an IPC-shared buffer (double-buffered by epoch);
// the first CTA of this kernel publishes "my buffer of this epoch is complete" to every
// peer's flag word, every CTA waits until all peers have done so, then each element is
// the sum of the W peer buffers read in rank order -- bit-identical on every rank, so the
// replicated loop state cannot diverge.  No second barrier: a buffer is reused two epochs
// later, by which time every peer has passed the next epoch's barrier, i.e. finished this one.
struct P2PArgs {
  const double* peer[kMaxPeers];        // peer r's buffer of this epoch
  unsigned long long* peer_flags[kMaxPeers];  // peer r's flag array (we write entry `rank`)
  const unsigned long long* my_flags;   // [world] written by the peers
  int world, rank;
  unsigned long long epoch;                 // epoch of this exchange, or (fused loop) its offset ...
  const unsigned long long* epoch_base;     // ... from the device-resident base (null: absolute)
  int* err;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__global__ void mle_p2p_merge_kernel(P2PArgs a, double* part, const int* act, int N, int K) {
  const unsigned long long epoch = a.epoch + (a.epoch_base ? *a.epoch_base : 0ull);
  if (blockIdx.x == 0 && blockIdx.y == 0 && (int)threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(a.peer_flags[threadIdx.x] + a.rank, epoch);
  }
  if ((int)threadIdx.x < a.world) {
    const long long t0 = clock64();
    // a barrier that timed out once poisons the block: later launches do not wait again
    while (ld_acquire_sys(a.my_flags + threadIdx.x) < epoch && *((volatile int*)a.err) == 0) {
      if (clock64() - t0 > 40000000000ll) {  // ~20 s: a peer is gone; fail loudly instead of hanging
        *a.err = 1;
        break;
      }
    }
  }
  __syncthreads();
  const int k = blockIdx.y;
  if (k == K) {
    if (blockIdx.x == 0 && (int)threadIdx.x < K) {
      double s = 0.0;
      for (int r = 0; r < a.world; ++r) s += ld_relaxed_sys(a.peer[r] + (size_t)K * N + threadIdx.x);
      part[(size_t)K * N + threadIdx.x] = s;
    }
    return;
  }
  if (!act[k]) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
    double s = 0.0;
    for (int r = 0; r < a.world; ++r) s += ld_relaxed_sys(a.peer[r] + (size_t)k * N + i);
    part[(size_t)k * N + i] = s;
  }
}

__global__ void mle_epoch_bump_kernel(unsigned long long* epoch, unsigned long long by) { *epoch += by; }

// A_k = simplex_proj(v / sum v)   (m_step.py:149-153, utils.py:8-31)
__global__ void __launch_bounds__(1024)
mle_project_kernel(const double* v, double* y, double* out, int N, double m, int normalise) {
  __shared__ double red[96];
  v += (size_t)blockIdx.x * N;  // one CTA per vector
  y += (size_t)blockIdx.x * N;
  out += (size_t)blockIdx.x * N;
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
  ny = std::max<long long>(ny, (L + kColChunkMax - 1) / kColChunkMax);  // smem staging per CTA
  URSM_REQUIRE(ny <= 65535, "M-step: too many cells in one shard for the column-sum grid");
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
    m->y.alloc(KN);
    m->act_cur.alloc(K);
    m->act_next.alloc(K);
    m->niter.alloc(K);
    m->base.alloc(K);
    m->halv.alloc(K);
    m->delta.alloc(K);
    m->act_cur.zero();
    m->act_next.zero();
    m->cand.alloc(KN * kCand);
    m->cres.alloc((size_t)K * kCand * 4);
    m->terms.alloc(4);
    if (m->sc) {
      std::vector<int> ts(K + 1);
      for (int k = 0; k <= K; ++k) ts[k] = (int)m->sc->type_start[k];
      m->type_start.alloc(K + 1);
      URSM_CUDA(cudaMemcpy(m->type_start.p, ts.data(), (K + 1) * sizeof(int), cudaMemcpyHostToDevice));
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
    URSM_REQUIRE(s->cnt_w > 0, "ursm_mle_begin: the single-cell E-step has not run yet");
    const int C = 8 / s->cnt_w;
    const int gx = (m->N + C * kColThreads - 1) / (C * kColThreads);
    int ny;
    const int chunk = colsum_chunk(m, s->L, gx, &ny);
    if (s->cnt_w == 1)
      mle_colsum_kernel<unsigned char, 0><<<dim3(gx, ny), kColThreads, 0, st>>>(
          s->Y, s->cnt.p, s->G.p, s->order.p, nullptr, nullptr, nullptr, nullptr, s->L, m->N,
          chunk, 1.0 / s->sample, d_part);
    else
      mle_colsum_kernel<unsigned short, 0><<<dim3(gx, ny), kColThreads, 0, st>>>(
          s->Y, reinterpret_cast<const unsigned short*>(s->cnt.p), s->G.p, s->order.p, nullptr,
          nullptr, nullptr, nullptr, s->L, m->N, chunk, 1.0 / s->sample, d_part);
    URSM_LAUNCH_CHECK();
  }
  URSM_API_END
}

// The one-pass kernel applies when rows keep 8-byte alignment and a useful block of cells fits
// in shared memory next to two resident CTAs per SM (URSM_MLE_FUSED=0 forces the two-kernel path).
static bool mle_fused_ok(const Mle* m) {
  if (const char* e = getenv("URSM_MLE_FUSED"))
    if (!atoi(e)) return false;
  const size_t row_bytes = (size_t)m->N * m->sc->cnt_w;
  return (row_bytes % 8) == 0 && (m->N % 2) == 0 && m->N <= 8192;
}

// short rows (at most 256 eight-byte groups): the persistent 256-thread kernel
// (URSM_MLE_FUSED_SMALL=0 keeps the general one: tests)
static bool mle_fused_small_ok(const Mle* m) {
  if (const char* e = getenv("URSM_MLE_FUSED_SMALL"))
    if (!atoi(e)) return false;
  const size_t row_bytes = (size_t)m->N * m->sc->cnt_w;
  return row_bytes / 8 <= 256 && (size_t)m->N * sizeof(double) + 4 * (row_bytes + 16) + 64 <= 54 * 1024;
}

static void mle_fused_geometry(const Mle* m, int* nc_out, size_t* smem_out) {
  const size_t row_bytes = (size_t)m->N * m->sc->cnt_w;
  const size_t a_bytes = (size_t)m->N * sizeof(double);  // the staged column A[:, type]
  const long long cap = std::max<long long>(
      1, ((long long)(108 * 1024) - (long long)a_bytes) / (long long)row_bytes);
  int per_sm = 2;  // CTAs per SM the grid aims at: their phases overlap
  if (const char* e = getenv("URSM_MLE_FUSED_CTAS_PER_SM")) per_sm = std::max(1, atoi(e));
  const long long slots = (long long)per_sm * m->num_sms;
  long long nc = (m->sc->L + slots - 1) / slots;
  nc = std::max<long long>(std::min<long long>(nc, cap), std::min<long long>(8, cap));
  const size_t smem = (((size_t)nc * 24 + 15) & ~(size_t)15) + a_bytes + (size_t)nc * row_bytes;
  static size_t configured[2] = {0, 0};
  const int which = m->sc->cnt_w == 1 ? 0 : 1;
  if (smem > 48 * 1024 && smem > configured[which]) {
    if (which == 0)
      URSM_CUDA(cudaFuncSetAttribute(mle_pass_fused_kernel<unsigned char>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
      URSM_CUDA(cudaFuncSetAttribute(mle_pass_fused_kernel<unsigned short>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[which] = smem;
  }
  *nc_out = (int)nc;
  *smem_out = smem;
}

// opt_u + update_gd_coeffConst for the columns flagged in m->act_cur (device flags).
// `lean`: the fused loop -- the buffers were zeroed by the step kernel and sum_l R_l log u_l is
// left to the full pass that follows the loop.
static void mle_pass(Mle* m, double* d_part, cudaStream_t st, bool lean = false) {
  ScShard* s = m->sc;
  const size_t KN = (size_t)m->K * m->N;
  if (!lean) URSM_CUDA(cudaMemsetAsync(d_part, 0, (KN + m->K) * sizeof(double), st));
  const double inv = 1.0 / s->sample;
  URSM_REQUIRE(s->cnt_w > 0, "M-step: the single-cell E-step has not run yet");
  if (mle_fused_ok(m) && mle_fused_small_ok(m)) {
    // short gene vectors: persistent CTAs, four per SM; A column + a block of rows in shared memory
    const size_t row_bytes = (size_t)m->N * s->cnt_w;
    const size_t budget = 54 * 1024;
    const size_t a_bytes = (size_t)m->N * sizeof(double);
    long long nc = std::max<long long>(1, ((long long)budget - (long long)a_bytes - 64) / (long long)(row_bytes + 16));
    nc = std::min<long long>(nc, 64);
    long long slots = 4LL * m->num_sms;
    if (const char* e = getenv("URSM_MLE_FUSED_SMALL_CTAS"))  // tests: many blocks and runs per CTA
      slots = std::max(1, atoi(e));
    long long per = (s->L + slots - 1) / slots;
    per = std::max<long long>(nc, (per + nc - 1) / nc * nc);  // whole blocks per CTA
    const unsigned nb = (unsigned)((s->L + per - 1) / per);
    const size_t smem = ((((size_t)nc * 16 + 15) & ~(size_t)15)) + a_bytes + (size_t)nc * row_bytes;
    static bool configured = false;
    if (!configured) {
      URSM_CUDA(cudaFuncSetAttribute(mle_pass_fused_small_kernel<unsigned char>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      URSM_CUDA(cudaFuncSetAttribute(mle_pass_fused_small_kernel<unsigned short>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      configured = true;
    }
    if (s->cnt_w == 1)
      mle_pass_fused_small_kernel<unsigned char><<<nb, kSmallThreads, smem, st>>>(
          s->cnt.p, s->order.p, m->type_start.p, m->A.p, m->act_cur.p, s->R.p, m->u.p, s->L, m->N,
          m->K, (int)nc, per, inv, d_part);
    else
      mle_pass_fused_small_kernel<unsigned short><<<nb, kSmallThreads, smem, st>>>(
          reinterpret_cast<const unsigned short*>(s->cnt.p), s->order.p, m->type_start.p, m->A.p,
          m->act_cur.p, s->R.p, m->u.p, s->L, m->N, m->K, (int)nc, per, inv, d_part);
    URSM_LAUNCH_CHECK();
    if (lean) return;
    mle_rlogu_kernel<<<(unsigned)((s->L + 255) / 256), 256, 0, st>>>(s->G.p, s->R.p, m->u.p, s->L,
                                                                    d_part + KN);
    URSM_LAUNCH_CHECK();
    return;
  }
  if (mle_fused_ok(m)) {
    int nc;
    size_t smem;
    mle_fused_geometry(m, &nc, &smem);
    const unsigned nb = (unsigned)((s->L + nc - 1) / nc);
    if (s->cnt_w == 1)
      mle_pass_fused_kernel<unsigned char><<<nb, kFusedThreads, smem, st>>>(
          s->cnt.p, s->order.p, m->type_start.p, m->A.p, m->act_cur.p, s->R.p, m->u.p, s->L, m->N,
          m->K, nc, inv, d_part);
    else
      mle_pass_fused_kernel<unsigned short><<<nb, kFusedThreads, smem, st>>>(
          reinterpret_cast<const unsigned short*>(s->cnt.p), s->order.p, m->type_start.p, m->A.p,
          m->act_cur.p, s->R.p, m->u.p, s->L, m->N, m->K, nc, inv, d_part);
    URSM_LAUNCH_CHECK();
    if (lean) return;
    mle_rlogu_kernel<<<(unsigned)((s->L + 255) / 256), 256, 0, st>>>(s->G.p, s->R.p, m->u.p, s->L,
                                                                    d_part + KN);
    URSM_LAUNCH_CHECK();
    return;
  }
  const int C = 8 / s->cnt_w;
  const int gx = (m->N + C * kColThreads - 1) / (C * kColThreads);
  int ny;
  const int chunk = colsum_chunk(m, s->L, gx, &ny);
  // u pass geometry: blocks of cells x gene segments (at most kUSegGenes genes, a multiple
  // of the 128 genes a warp covers per step), sized for ~4 CTAs per SM over the whole grid
  const int group = (kUThreads / 32) * kURows;
  int nseg = (m->N + kUSegGenes - 1) / kUSegGenes;
  long long rows = (s->L * nseg + 4LL * m->num_sms - 1) / (4LL * m->num_sms);
  rows = std::max<long long>(group, std::min<long long>(8 * group, (rows + group - 1) / group * group));
  const unsigned ub = (unsigned)((s->L + rows - 1) / rows);
  if ((long long)ub * nseg < 4LL * m->num_sms)
    nseg = (int)std::min<long long>((m->N + 127) / 128, (4LL * m->num_sms + ub - 1) / ub);
  const int seg_genes = std::min(kUSegGenes, ((m->N + nseg - 1) / nseg + 127) / 128 * 128);
  nseg = (m->N + seg_genes - 1) / seg_genes;
  if (!lean) m->ratio.zero(st);  // u_acc: partial dot products
  if (s->cnt_w == 1) {
    mle_u_kernel<unsigned char><<<dim3(ub, nseg), kUThreads, 0, st>>>(
        s->cnt.p, s->order.p, m->type_start.p, m->A.p, m->act_cur.p, s->L, m->N, m->K, (int)rows,
        seg_genes, m->ratio.p);
    URSM_LAUNCH_CHECK();
    mle_colsum_kernel<unsigned char, 1><<<dim3(gx, ny), kColThreads, 0, st>>>(
        s->Y, s->cnt.p, s->G.p, s->order.p, m->ratio.p, s->R.p, m->u.p, m->act_cur.p, s->L, m->N,
        chunk, inv, d_part);
    URSM_LAUNCH_CHECK();
  } else {
    const unsigned short* c16 = reinterpret_cast<const unsigned short*>(s->cnt.p);
    mle_u_kernel<unsigned short><<<dim3(ub, nseg), kUThreads, 0, st>>>(
        c16, s->order.p, m->type_start.p, m->A.p, m->act_cur.p, s->L, m->N, m->K, (int)rows,
        seg_genes, m->ratio.p);
    URSM_LAUNCH_CHECK();
    mle_colsum_kernel<unsigned short, 1><<<dim3(gx, ny), kColThreads, 0, st>>>(
        s->Y, c16, s->G.p, s->order.p, m->ratio.p, s->R.p, m->u.p, m->act_cur.p, s->L, m->N, chunk,
        inv, d_part);
    URSM_LAUNCH_CHECK();
  }
  if (lean) return;
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

// a batch is ncand <= kCand trial steps per column, sized so that all ncand x K CTAs are
// resident in one wave (the accepted step does not depend on the batch size)
static int mle_ncand(const Mle* m) { return std::max(1, std::min(kCand, m->num_sms / m->K)); }

// trial column in dynamic shared memory (N doubles) when it fits; opt in once, outside captures
static int mle_cand_y_smem(const Mle* m) {
  const size_t ybytes = (size_t)m->N * sizeof(double);
  int y_smem = ybytes <= (size_t)160 * 1024 ? 1 : 0;
  if (const char* e = getenv("URSM_MLE_Y_SMEM"))  // tests: the global-memory trial column
    y_smem = (y_smem && atoi(e)) ? 1 : 0;
  static size_t configured = 0;
  if (y_smem && ybytes > 48 * 1024 && ybytes > configured) {
    URSM_CUDA(cudaFuncSetAttribute(mle_cand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)ybytes));
    configured = ybytes;
  }
  return y_smem;
}

// fills the fields every launch of the step kernel shares and launches it
static void mle_launch_cand(Mle* m, StepArgs& a, const double* part_in, const int* act,
                            cudaStream_t st) {
  a.A = m->A.p;
  a.cand = m->cand.p;
  a.cres = m->cres.p;
  a.cAinv = m->cAinv;
  a.cA = m->cA;
  a.coeffA = m->coeffA;
  a.part = part_in;
  a.act = act;
  a.base = m->base.p;
  a.N = m->N;
  a.K = m->K;
  a.min_A = m->min_A;
  a.init_step = m->init_step;
  a.ncand = mle_ncand(m);
  const int T = std::max(64, std::min(1024, (m->N / 4 + 31) / 32 * 32));
  a.y_smem = mle_cand_y_smem(m);
  const size_t dyn = a.y_smem ? (size_t)m->N * sizeof(double) : 0;
  mle_cand_kernel<<<dim3(a.ncand, m->K), T, dyn, st>>>(a);
  URSM_LAUNCH_CHECK();
}

int ursm_mle_opt_step(ursm_mle* mm, const double* d_cconst_part, void* stream) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && d_cconst_part && m->cAinv && m->maxiter > 0, "ursm_mle_opt_step: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  StepArgs a;
  memset(&a, 0, sizeof a);
  mle_launch_cand(m, a, d_cconst_part, m->act_cur.p, st);
  const int Tc = std::max(64, std::min(1024, (m->N / 4 + 31) / 32 * 32));
  mle_decide_copy_kernel<<<m->K, Tc, 0, st>>>(m->cres.p, m->act_cur.p, m->act_next.p, m->niter.p,
                                             m->base.p, m->halv.p, m->delta.p, m->A.p, m->cand.p,
                                             m->N, m->conv, m->maxiter, a.ncand);
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
  std::swap(m->act_cur.p, m->act_next.p);  // advance: the flags decided in this iteration
  URSM_API_END
}

static size_t p2p_buf_doubles(const Mle* m) { return (size_t)m->K * m->N + m->K; }
static size_t p2p_bytes(const Mle* m) {
  return 2 * p2p_buf_doubles(m) * sizeof(double) + kMaxPeers * sizeof(unsigned long long);
}
static double* p2p_buf(const Mle* m, void* base, unsigned long long epoch) {
  return static_cast<double*>(base) + (epoch & 1ull) * p2p_buf_doubles(m);
}
static unsigned long long* p2p_flags(const Mle* m, void* base) {
  return reinterpret_cast<unsigned long long*>(static_cast<double*>(base) + 2 * p2p_buf_doubles(m));
}

int ursm_mle_p2p_alloc(ursm_mle* mm, void* out_handle64, int32_t* out_cached) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && out_handle64 && out_cached, "ursm_mle_p2p_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!m->p2p) {
    const size_t bytes = p2p_bytes(m);
    for (P2PBlock* b : g_p2p_blocks)
      if (b->bytes == bytes && !b->in_use) {
        m->p2p = b;
        break;
      }
    if (!m->p2p) {
      P2PBlock* b = new P2PBlock();
      b->bytes = bytes;
      URSM_CUDA(cudaMalloc(&b->own, bytes));  // IPC needs a cudaMalloc block, not the pool
      URSM_CUDA(cudaMemset(b->own, 0, bytes));
      URSM_CUDA(cudaMalloc((void**)&b->err, sizeof(int)));
      URSM_CUDA(cudaMemset(b->err, 0, sizeof(int)));
      URSM_CUDA(cudaDeviceSynchronize());
      b->id = (int)g_p2p_blocks.size();
      g_p2p_blocks.push_back(b);
      m->p2p = b;
    }
    m->p2p->in_use = true;
  }
  *out_cached = (m->p2p->opened && !m->p2p->dirty) ? 1 : 0;
  cudaIpcMemHandle_t h;
  URSM_CUDA(cudaIpcGetMemHandle(&h, m->p2p->own));
  memcpy(out_handle64, &h, sizeof h);
  URSM_API_END
}

int ursm_mle_p2p_open(ursm_mle* mm, const void* handles, int32_t world, int32_t rank) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && m->p2p, "ursm_mle_p2p_open: call ursm_mle_p2p_alloc first");
  URSM_REQUIRE(world >= 2 && world <= kMaxPeers && rank >= 0 && rank < world,
               "ursm_mle_p2p_open: bad world/rank");
  P2PBlock* b = m->p2p;
  if (b->opened) {  // cached mapping: every rank reuses the block it used before
    URSM_REQUIRE(b->world == world && b->rank == rank, "ursm_mle_p2p_open: cached block of another job");
  } else {
    URSM_REQUIRE(handles, "ursm_mle_p2p_open: no handles");
    const cudaIpcMemHandle_t* hs = static_cast<const cudaIpcMemHandle_t*>(handles);
    for (int r = 0; r < world; ++r) {
      if (r == rank) {
        b->peer[r] = b->own;
      } else {
        URSM_CUDA(cudaIpcOpenMemHandle(&b->peer[r], hs[r], cudaIpcMemLazyEnablePeerAccess));
      }
    }
    b->world = world;
    b->rank = rank;
    b->opened = true;
  }
  URSM_API_END
}

int ursm_mle_p2p_state(ursm_mle* mm, int64_t* out4) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && m->p2p && out4, "ursm_mle_p2p_state: call ursm_mle_p2p_alloc first");
  out4[0] = (m->p2p->opened && !m->p2p->dirty) ? 1 : 0;
  out4[1] = m->p2p->id;
  out4[2] = (int64_t)m->p2p->epoch;
  out4[3] = m->p2p->opened ? 1 : 0;
  URSM_API_END
}

int ursm_mle_p2p_reset(ursm_mle* mm) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && m->p2p, "ursm_mle_p2p_reset: call ursm_mle_p2p_alloc first");
  P2PBlock* b = m->p2p;
  URSM_CUDA(cudaDeviceSynchronize());
  URSM_CUDA(cudaMemset(p2p_flags(m, b->own), 0, kMaxPeers * sizeof(unsigned long long)));
  URSM_CUDA(cudaMemset(b->err, 0, sizeof(int)));
  URSM_CUDA(cudaDeviceSynchronize());
  b->epoch = 0;
  b->dirty = false;
  URSM_API_END
}

int ursm_mle_p2p_disable(ursm_mle* mm) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  if (m && m->p2p) {
    m->p2p->in_use = false;
    m->p2p = nullptr;
  }
  URSM_API_END
}

int ursm_mle_opt_pass_p2p(ursm_mle* mm, void* stream) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && m->sc && m->p2p && m->p2p->opened, "ursm_mle_opt_pass_p2p: peer exchange not set up");
  ++m->p2p->epoch;
  mle_pass(m, p2p_buf(m, m->p2p->own, m->p2p->epoch), (cudaStream_t)stream);
  URSM_API_END
}

int ursm_mle_opt_merge_p2p(ursm_mle* mm, double* d_part, void* stream) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && d_part && m->p2p && m->p2p->opened, "ursm_mle_opt_merge_p2p: peer exchange not set up");
  cudaStream_t st = (cudaStream_t)stream;
  P2PBlock* b = m->p2p;
  P2PArgs a;
  memset(&a, 0, sizeof a);
  for (int r = 0; r < b->world; ++r) {
    a.peer[r] = p2p_buf(m, b->peer[r], b->epoch);
    a.peer_flags[r] = p2p_flags(m, b->peer[r]);
  }
  a.my_flags = p2p_flags(m, b->own);
  a.world = b->world;
  a.rank = b->rank;
  a.epoch = b->epoch;
  a.err = b->err;
  mle_p2p_merge_kernel<<<dim3((m->N + 255) / 256, m->K + 1), 256, 0, st>>>(a, d_part, m->act_cur.p,
                                                                           m->N, m->K);
  URSM_LAUNCH_CHECK();
  std::swap(m->act_cur.p, m->act_next.p);  // advance: the flags decided in this iteration
  URSM_API_END
}

int ursm_mle_p2p_error(ursm_mle* mm, int32_t* h_err) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && h_err, "ursm_mle_p2p_error: bad argument");
  *h_err = 0;
  if (m->p2p && m->p2p->err) {
    URSM_CUDA(cudaMemcpy(h_err, m->p2p->err, sizeof(int), cudaMemcpyDeviceToHost));
    if (*h_err) m->p2p->dirty = true;  // the next fit resets the block behind a barrier
  }
  URSM_API_END
}

static size_t p2p_buf_doubles(const Mle* m);
static double* p2p_buf(const Mle* m, void* base, unsigned long long epoch);
static unsigned long long* p2p_flags(const Mle* m, void* base);

// One outer iteration of opt_Ak for all columns, enqueued on `st`:
//   step   (trial steps + decision + installation, and the zeroing the pass needs)
//   pass   (opt_u + update_gd_coeffConst of the stepped columns -> part_out)
//   merge  (peer exchange only: part_in <- sum over ranks of the peers' part_out)
static void mle_enqueue_iteration(Mle* m, int it, cudaStream_t st) {
  const size_t stride = (size_t)m->K * m->N + m->K;
  const bool peers = m->p2p != nullptr;
  const double* part_in;
  double* part_out;
  if (peers) {  // epochs advance by kGraphIters (even) per launch from an even base: static parity
    part_in = m->partbuf.p;
    part_out = p2p_buf(m, m->p2p->own, (unsigned long long)(it + 1));
  } else {
    part_in = m->partbuf.p + (size_t)(it & 1) * stride;
    part_out = m->partbuf.p + (size_t)((it + 1) & 1) * stride;
  }
  const int gslots = (m->N + kGradThreads - 1) / kGradThreads;
  mle_grad_kernel<<<dim3(gslots, m->K), kGradThreads, 0, st>>>(m->A.p, m->cAinv, m->cA, m->coeffA,
                                                              part_in, m->act_cur.p, m->N,
                                                              m->gbuf.p, m->gobj.p);
  URSM_LAUNCH_CHECK();
  StepArgs a;
  memset(&a, 0, sizeof a);
  a.gbuf = m->gbuf.p;
  a.gobj = m->gobj.p;
  a.gslots = gslots;
  a.Aw = m->A.p;
  a.zero_out = part_out;
  a.u_acc = m->ratio.p;
  a.order = m->sc->order.p;
  a.type_start = m->type_start.p;
  a.done = m->done.p;
  a.act_w = m->act_cur.p;
  a.niter = m->niter.p;
  a.halv = m->halv.p;
  a.base_w = m->base.p;
  a.delta = m->delta.p;
  a.lam_store = m->lam_store.p;
  a.conv = m->conv;
  a.maxiter = m->maxiter;
  mle_launch_cand(m, a, part_in, m->act_cur.p, st);
  mle_pass(m, part_out, st, /*lean=*/true);
  if (peers) {
    P2PBlock* b = m->p2p;
    P2PArgs pa;
    memset(&pa, 0, sizeof pa);
    for (int r = 0; r < b->world; ++r) {
      pa.peer[r] = p2p_buf(m, b->peer[r], (unsigned long long)(it + 1));
      pa.peer_flags[r] = p2p_flags(m, b->peer[r]);
    }
    pa.my_flags = p2p_flags(m, b->own);
    pa.world = b->world;
    pa.rank = b->rank;
    pa.epoch = (unsigned long long)(it + 1);
    pa.epoch_base = m->epoch_dev.p;
    pa.err = b->err;
    mle_p2p_merge_kernel<<<dim3((m->N + 255) / 256, m->K + 1), 256, 0, st>>>(pa, m->partbuf.p,
                                                                             m->act_cur.p, m->N, m->K);
    URSM_LAUNCH_CHECK();
  }
}

// opt_A_u (:128-154) for the columns flagged in h_active, to completion: the loop of opt_Ak
// (:157-182) runs on the device -- kGraphIters outer iterations per launch of a captured CUDA
// graph on the workspace's own stream, three kernels per iteration and no memset; the host only
// reads the K loop flags back between launches.  With the peer exchange set up
// (ursm_mle_p2p_open) every rank runs the same graph and the per-iteration sum crosses NVLink
// inside it.  d_part_init: the all-reduced partials of the pass that preceded the loop.
// The partials of the last accepted steps are NOT returned: the caller follows up with a full
// pass (ursm_mle_pass_u), which also refreshes sum_l R_l log u_l.
int ursm_mle_opt_run(ursm_mle* mm, const int32_t* h_active, double init_step, double conv,
                     int32_t maxiter, const double* d_part_init, int32_t* h_niter,
                     int32_t* h_halvings, int32_t* h_launches, void* stream) {
  URSM_API_BEGIN
  Mle* m = reinterpret_cast<Mle*>(mm);
  URSM_REQUIRE(m && m->sc && h_active && d_part_init && m->cAinv && init_step > 0 && maxiter >= 1,
               "ursm_mle_opt_run: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int K = m->K;
  const size_t stride = (size_t)K * m->N + K;
  const bool peers = m->p2p != nullptr && m->p2p->opened;
  URSM_REQUIRE(m->p2p == nullptr || peers, "ursm_mle_opt_run: peer exchange allocated but not opened");
  if (!m->gstream) {
    URSM_CUDA(cudaStreamCreateWithFlags(&m->gstream, cudaStreamNonBlocking));
    URSM_CUDA(cudaEventCreateWithFlags(&m->ev_in, cudaEventDisableTiming));
    URSM_CUDA(cudaEventCreateWithFlags(&m->ev_out, cudaEventDisableTiming));
    URSM_CUDA(cudaHostAlloc((void**)&m->h_poll, 3 * K * sizeof(int), cudaHostAllocPortable));
    m->done.alloc(K);
    m->lam_store.alloc((size_t)K * kCand);
    m->gbuf.alloc((size_t)K * m->N);
    m->gobj.alloc((size_t)K * ((m->N + 255) / 256) * 3);
    m->epoch_dev.alloc(1);
    URSM_CUDA(cudaDeviceSynchronize());  // pool allocations are ordered on the default stream
  }
  if (m->partbuf.n != 2 * stride) {
    m->partbuf.alloc(2 * stride);
    URSM_CUDA(cudaDeviceSynchronize());
  }
  m->init_step = init_step;
  m->conv = conv;
  m->maxiter = maxiter;
  // everything below runs on the workspace's stream, after what the caller has enqueued
  URSM_CUDA(cudaEventRecord(m->ev_in, st));
  URSM_CUDA(cudaStreamWaitEvent(m->gstream, m->ev_in, 0));
  cudaStream_t g = m->gstream;
  URSM_CUDA(cudaMemcpyAsync(m->act_cur.p, h_active, K * sizeof(int), cudaMemcpyHostToDevice, g));
  URSM_CUDA(cudaMemsetAsync(m->niter.p, 0, K * sizeof(int), g));
  URSM_CUDA(cudaMemsetAsync(m->base.p, 0, K * sizeof(int), g));
  URSM_CUDA(cudaMemsetAsync(m->halv.p, 0, K * sizeof(int), g));
  URSM_CUDA(cudaMemsetAsync(m->delta.p, 0, K * sizeof(double), g));
  URSM_CUDA(cudaMemsetAsync(m->done.p, 0, K * sizeof(int), g));
  URSM_CUDA(cudaMemsetAsync(m->lam_store.p, 0xFF, (size_t)K * kCand * sizeof(double), g));  // NaN: cold
  URSM_CUDA(cudaMemcpyAsync(m->partbuf.p, d_part_init, stride * sizeof(double),
                            cudaMemcpyDeviceToDevice, g));
  if (peers) {
    URSM_REQUIRE((m->p2p->epoch & 1ull) == 0, "ursm_mle_opt_run: peer-exchange epoch is odd");
    URSM_CUDA(cudaMemcpyAsync(m->epoch_dev.p, &m->p2p->epoch, sizeof(unsigned long long),
                              cudaMemcpyHostToDevice, g));
  }
  Mle::GraphKey key;
  key.init_step = init_step;
  key.conv = conv;
  key.maxiter = maxiter;
  key.ncand = mle_ncand(m);
  key.p2p = peers ? 1 : 0;
  key.coeffs = m->cAinv;
  key.p2p_own = peers ? m->p2p->own : nullptr;
  key.act = m->act_cur.p;
  key.cnt = m->sc->cnt.p;
  key.cnt_w = m->sc->cnt_w;
  key.sample = m->sc->sample;
  key.y_smem = mle_cand_y_smem(m);  // (also opts the kernel in to its shared memory, before any capture)
  if (mle_fused_ok(m)) {  // same for the one-pass kernel
    int nc;
    size_t smem;
    mle_fused_geometry(m, &nc, &smem);
  }
  if (!m->gexec || !(m->gkey == key)) {
    if (m->gexec) {
      cudaGraphExecDestroy(m->gexec);
      m->gexec = nullptr;
    }
    URSM_CUDA(cudaStreamSynchronize(g));
    cudaGraph_t graph = nullptr;
    URSM_CUDA(cudaStreamBeginCapture(g, cudaStreamCaptureModeThreadLocal));
    try {
      for (int it = 0; it < kGraphIters; ++it) mle_enqueue_iteration(m, it, g);
      if (peers) {
        mle_epoch_bump_kernel<<<1, 1, 0, g>>>(m->epoch_dev.p, (unsigned long long)kGraphIters);
        URSM_LAUNCH_CHECK();
      }
    } catch (...) {
      cudaStreamEndCapture(g, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    URSM_CUDA(cudaStreamEndCapture(g, &graph));
    cudaError_t e = cudaGraphInstantiate(&m->gexec, graph, 0);
    cudaGraphDestroy(graph);
    URSM_CUDA(e);
    m->gkey = key;
  }
  int launches = 0;
  const int max_launches = (maxiter + kGraphIters - 1) / kGraphIters + 64 + 1024 / mle_ncand(m);
  while (true) {
    URSM_CUDA(cudaGraphLaunch(m->gexec, g));
    count_launch(kGraphIters * (peers ? 5 : 4));
    ++launches;
    if (peers) m->p2p->epoch += kGraphIters;
    URSM_CUDA(cudaMemcpyAsync(m->h_poll, m->act_cur.p, K * sizeof(int), cudaMemcpyDeviceToHost, g));
    URSM_CUDA(cudaStreamSynchronize(g));
    bool any = false;
    for (int k = 0; k < K; ++k) any = any || m->h_poll[k] != 0;
    if (!any) break;
    URSM_REQUIRE(launches < max_launches, "ursm_mle_opt_run: the opt_Ak loop did not terminate");
  }
  URSM_CUDA(cudaMemcpyAsync(m->h_poll + K, m->niter.p, K * sizeof(int), cudaMemcpyDeviceToHost, g));
  URSM_CUDA(cudaMemcpyAsync(m->h_poll + 2 * K, m->halv.p, K * sizeof(int), cudaMemcpyDeviceToHost, g));
  URSM_CUDA(cudaStreamSynchronize(g));
  for (int k = 0; k < K; ++k) {
    if (h_niter) h_niter[k] = m->h_poll[K + k];
    if (h_halvings) h_halvings[k] = m->h_poll[2 * K + k];
  }
  if (h_launches) *h_launches = launches;
  URSM_CUDA(cudaEventRecord(m->ev_out, g));
  URSM_CUDA(cudaStreamWaitEvent(st, m->ev_out, 0));
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

int ursm_simplex_proj_rows(const double* h_y, int32_t rows, int32_t n, double min_y, double* h_out) {
  URSM_API_BEGIN
  URSM_REQUIRE(h_y && h_out && n > 0 && rows > 0, "ursm_simplex_proj_rows: bad argument");
  const size_t tot = (size_t)rows * n;
  DevBuf<double> v, y, o;
  v.alloc(tot);
  y.alloc(tot);
  o.alloc(tot);
  URSM_CUDA(cudaMemcpy(v.p, h_y, tot * sizeof(double), cudaMemcpyHostToDevice));
  const int T = std::max(64, std::min(1024, (n + 31) / 32 * 32));
  mle_project_kernel<<<rows, T>>>(v.p, y.p, o.p, n, min_y, 0);
  URSM_LAUNCH_CHECK();
  URSM_CUDA(cudaMemcpy(h_out, o.p, tot * sizeof(double), cudaMemcpyDeviceToHost));
  URSM_API_END
}

int ursm_simplex_proj(const double* h_y, int32_t n, double min_y, double* h_out) {
  return ursm_simplex_proj_rows(h_y, 1, n, min_y, h_out);
}

}  // extern "C"
