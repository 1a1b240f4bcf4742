// draw_Z (e_step_gibbs.py:132-138) for one Gibbs sweep: Z_ji. ~ Mult(X_ji, W_.j A_i. / AW_ij) as
// a chain of K-1 conditional binomials (the decomposition numpy.random.multinomial uses).
//
// The 32 lanes of a warp are 32 bulk SAMPLES and the warp walks genes: for a given (gene, type)
// all lanes share A_ik and see similar counts.  Counts are read from a gene-major copy of X
// (samples contiguous: one 128-byte line per warp and gene).  sum_i Z_jik is private to a lane (a
// shared-memory column per thread, flushed once per CTA); sum_j Z_jik is a warp reduction + one
// shared-memory atomic per (gene, type), flushed once per CTA.  Z (M x N x K) and AW (N x M)
// never exist.  The RNG is keyed on the gene id and the GLOBAL sample id, so results do not
// depend on the sharding or on which warp takes which gene.
//
// bk_split_kernel (the production path) draws the binomials of one (gene, step) with the WHOLE
// WARP IN LOCKSTEP -- no lane runs its own data-dependent loop:
//   * lanes with n r <= 40: inversion from 0 in straight-line code.  All lanes stand on the same
//     atom x at the same step, so 1/x is a compile-time immediate and a step is five
//     instructions (FSET, FADD, FFMA, FMUL, FADD: no MUFU, no cursor bookkeeping); the warp
//     leaves at the first chunk boundary (4, 8, 16, 24, ... atoms) where its last lane has landed;
//   * lanes with n r > 40 and n < 1024: inversion from the MODE, both directions per step
//     (~1 + 1.6 sigma atoms instead of 1 + n r; log2 k! table in shared memory, one MUFU
//     reciprocal per pair of atoms), again in lockstep;
//   * a lane whose item fits neither (n r > 40 with n >= 1024: BTRS territory) or whose uniform
//     falls beyond the fp32 mass sum (~1e-7 of the draws; redrawn, which keeps the draw
//     unbiased) hands the rest of its chain to a list that bk_split_cold_kernel finishes, one
//     thread per chain, so the lockstep loop carries neither the registers nor the divergence of
//     the rejection sampler.
// Conditional probabilities are ratios of positive masses: the masses of the steps still to come
// are SUFFIX SUMS accumulated backwards into a private shared-memory column, so that
// P(this type | not yet assigned) = q / S_t and its complement S_{t+1} / S_t involve no
// subtraction.  The chain visits the types in their order except that the type with the largest
// A_ik trades places with the last one: it takes the remainder and is never drawn (any order
// fixed before the draws gives the same multinomial; this one removes the largest n r).
// Genes are dealt to CTAs round-robin over the order of decreasing total count (every CTA gets
// the same mix of deep and shallow genes) and the warps of a CTA that share a sample group take
// them from a shared-memory counter.
//
// bk_split_lane_kernel is the per-lane sampler (modal inversion / BTRS state machine per lane):
// slower on shallow counts, but it takes any count; the host runs it instead when a sizeable
// part of X is 1024 or more (deep bulk libraries).
#pragma once
#include <algorithm>

#include "samplers.cuh"

namespace ursm {

constexpr int kSplitThreads = 256;

struct BkSplitArgs {
  // gene-indexed inputs of the lockstep kernel are stored BY POSITION in the order of decreasing
  // total count (`order[pos]` = gene id); the per-lane kernel takes them in the natural order
  const float* Xt;      // [N][Mpad] counts, gene-major
  const float* A32;     // [K][N] fp32 A^T
  const float* W32g;    // [ngroups][K][32] fp32 W of this sweep
  const double* lf2;    // [URSM_BINOM_TABLE] log2 k!
  const int* order;     // [N] gene id of every position (null: natural order)
  const unsigned char* kmax;  // [N] by position: type with the largest A_ik, drawn last (null: type K-1)
  double* n_cur;        // [M][K] accumulates this sweep's sum_i Z (zero on entry)
  double* sumI;         // [K][N] exp_Zik accumulator
  long long M, sample_offset;
  int Mpad, N, K, ngroups, gpc, genes_per_cta;
  double inv_sample;
  int retained;
  unsigned long long key;
  PhiloxKeys rk;
  unsigned int sweep;
  // chains the lockstep kernel could not finish (see bk_split_cold_kernel): entries
  // {position, local sample, step | fresh << 8, float bits of the count left}, their number (may
  // exceed the capacity: the excess is dropped and reported by the host), the capacity
  uint4* cold_list;
  unsigned int* cold_count;
  unsigned int cold_cap;
  int* fail;            // [1] error counter of the E-step (a list that did not fit adds 1 << 20)
  unsigned long long* probe;  // optional event counters (tools/bk_probe.cu), null in production
};

#if defined(__CUDACC__)

#ifndef URSM_BK_MINBLOCKS
#define URSM_BK_MINBLOCKS 4
#endif
#ifndef URSM_LS_ZERO_MAX
#define URSM_LS_ZERO_MAX 40.0f     // warp-max n r up to which the warp inverts from 0
#endif
constexpr int kLsZeroAtoms = 128;  // atoms of the straight-line search (n r <= 40: P(X > 128) ~ 0)
constexpr int kLsModePairs = 128;  // pairs of atoms of the modal search (n < 1024: sigma <= 16)

// log2(1 - r), 0 <= r <= 1/2: series below 1/16 (relative error < 3e-8), MUFU on the complement
__device__ __forceinline__ float log2_1m(float r, float rc) {
  const float ser = -r * fmaf(r, fmaf(r, fmaf(r, fmaf(r, fmaf(r, 1.0f / 6, 0.2f), 0.25f), 1.0f / 3),
                                      0.5f), 1.0f);
  return (r < 0.0625f) ? 1.4426950408889634f * ser : URSM_LG2(rc);
}

// Bin(n, r), 0 <= r <= 1/2, n r <= URSM_LS_ZERO_MAX, for all 32 lanes at once:
// x = #{t : u > F(t)} with F the running sum of pmf(t) = pmf(t-1) ((n+1) s / t - s), s = r/(1-r).
// A lane without an item passes u = 0 (it lands at once whatever n and r are); n = 0 or r = 0
// gives pmf(0) = 1.  Beyond t = n the recurrence leaves rounding residue of relative size 1e-7,
// which is harmless: the caller clamps x to n, and a uniform above the final sum is reported as a
// miss and redrawn.
__device__ __forceinline__ float binom_zero_lockstep(float nf, float r, float rc, float u, bool& miss) {
  float px = URSM_EX2(nf * log2_1m(r, rc));
  const float s = r * URSM_RCP(fmaxf(rc, 1e-30f));  // (r = rc = 0: a lane without masses, pmf(0) = 1)
  const float c = fmaf(nf, s, s);
  float cdf = px;
  float x = 0.0f;
#define URSM_LS_ATOM(T)                         \
  {                                             \
    x += (u > cdf) ? 1.0f : 0.0f;               \
    px *= fmaf(c, 1.0f / (float)((T) + 1), -s); \
    cdf += px;                                  \
  }
  bool open = !__all_sync(0xffffffffu, u <= cdf);
  if (open) {
#pragma unroll
    for (int q = 0; q < 4; ++q) URSM_LS_ATOM(q)
    open = !__all_sync(0xffffffffu, u <= cdf);
  }
  if (open) {
#pragma unroll
    for (int q = 4; q < 8; ++q) URSM_LS_ATOM(q)
#pragma unroll
    for (int ch = 1; ch < kLsZeroAtoms / 8; ++ch) {
      if (__all_sync(0xffffffffu, u <= cdf)) break;
#pragma unroll
      for (int q = 0; q < 8; ++q) URSM_LS_ATOM(ch * 8 + q)
    }
  }
#undef URSM_LS_ATOM
  miss = u > cdf;
  return x;
}

// Bin(n, r), 0 < r <= 1/2, n < URSM_BINOM_TABLE, by inversion from the mode m = floor((n+1) r)
// (Kemp 1986): atoms in the fixed order m, m+1, m-1, m+2, m-2, ... for all 32 lanes at once.
// Lanes without an item pass n = 0, r = rc = 1/2, u = 0.
__device__ __forceinline__ float binom_mode_lockstep(int n, float nf, float r, float rc, float u,
                                                     const double* lf2, bool& miss) {
  int m = (int)((nf + 1.0f) * r);
  m = m < n ? m : n;
  const float mf = (float)m;
  const float logc = (float)(lf2[n] - lf2[m] - lf2[n - m]);  // log2 C(n, m)
  const float pm = URSM_EX2(fmaf(nf - mf, log2_1m(r, rc), fmaf(mf, URSM_LG2(r), logc)));
  const float t = URSM_RCP(r * rc);
  const float s = r * r * t, is = rc * rc * t;  // r/(1-r) and its reciprocal
  const float c = (nf + 1.0f) * s;
  float cdf = pm, res = mf;
  bool found = u <= cdf;
  float pu = pm, pd = pm, fu = mf, fd = mf, gd = nf - mf + 1.0f;
  for (int it = 0; it < kLsModePairs / 2; ++it) {
    if (__all_sync(0xffffffffu, found || fmaxf(pu, pd) <= 0.0f)) break;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      fu += 1.0f;
      const float R = URSM_RCP(fu * gd);  // 1/fu = R gd, 1/gd = R fu
      pu = (fu <= nf) ? pu * fmaf(c, R * gd, -s) : 0.0f;
      cdf += pu;
      if (!found && u <= cdf) {
        res = fu;
        found = true;
      }
      pd *= fd * is * (R * fu);
      fd -= 1.0f;
      gd += 1.0f;
      cdf += pd;
      if (!found && u <= cdf) {
        res = fd;
        found = true;
      }
    }
  }
  miss = !found;
  return res;
}

// Shared memory of bk_split_kernel, in 4-byte words (everything is addressed by word index from
// the start of the dynamic window, so that an access is one LDS/STS at [index * 4 + constant]):
//   [0, 2 TABLE)                   log2 k! table (doubles)
//   [accI0, accI0 + G K)           sum_j Z tile of the CTA's genes
//   per warp, 3 K + 1 rows of 32:  W_jk of the warp's sample group (row = type), the lanes'
//                                  sum_i Z columns (row = type), the suffix sums (row = step),
//                                  and one row holding A_ik of the current gene in step order
__host__ __device__ inline int bk_split_warp_words(int K) { return (3 * K + 1) * 32; }

template <bool PROBE>
__global__ void __launch_bounds__(kSplitThreads, URSM_BK_MINBLOCKS) bk_split_kernel(BkSplitArgs a) {
  extern __shared__ __align__(16) unsigned char split_smem[];
  __shared__ unsigned int qctr[kSplitThreads / 32];
  float* smf = reinterpret_cast<float*>(split_smem);
  unsigned int* smu = reinterpret_cast<unsigned int*>(split_smem);
  const double* lf2 = reinterpret_cast<const double*>(split_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = a.K, N = a.N, gpc = a.gpc;
  const int nphase = (kSplitThreads / 32) / gpc;
  const int grp_local = warp % gpc, phase = warp / gpc;
  const int grp = blockIdx.y * gpc + grp_local;
  const bool live = grp < a.ngroups && phase < nphase;
  // this CTA's genes: positions blockIdx.x, blockIdx.x + gridDim.x, ... of the cost-sorted order
  const int gx = gridDim.x;
  const int ng = (N - (int)blockIdx.x + gx - 1) / gx;
  const int accI0 = 2 * URSM_BINOM_TABLE;
  const int wb = accI0 + a.genes_per_cta * K + warp * bk_split_warp_words(K) + lane;
  const int K32 = K * 32;
  const int ab = wb - lane + 3 * K32;  // A row of the warp
  for (int t = tid; t < URSM_BINOM_TABLE; t += kSplitThreads) reinterpret_cast<double*>(split_smem)[t] = a.lf2[t];
  for (int k = 0; k < K; ++k) {
    smf[wb + k * 32] = live ? a.W32g[((size_t)grp * K + k) * 32 + lane] : 0.0f;
    smu[wb + K32 + k * 32] = 0u;
  }
  const bool keepI = a.retained != 0;
  if (keepI)
    for (int t = tid; t < ng * K; t += kSplitThreads) smu[accI0 + t] = 0u;
  if (tid < kSplitThreads / 32) qctr[tid] = 0u;
  __syncthreads();
  const uint32_t ctr2 = kPurposeBulkZ | a.sweep;
  const uint32_t jl = (uint32_t)grp * 32u + (uint32_t)lane;
  const uint32_t jg = (uint32_t)a.sample_offset + jl;
  unsigned long long pr_zero = 0, pr_mode = 0, pr_deep = 0, pr_miss = 0;

  if (live) {
    // The warps that share a sample group take the CTA's genes from a common counter.  A gene's
    // inputs -- its id, the counts of the warp's 32 samples, the type drawn last and A_ik (lane k
    // loads type k) -- depend on the position alone, and are loaded one gene ahead so that their
    // latency hides behind the chain of the gene in hand.
    unsigned int gl = 0, gl_n = 0;
    int gene = 0, km = 0, gene_n = 0, km_n = 0;
    float rem = 0.0f, aval = 0.0f, rem_n = 0.0f, aval_n = 0.0f;
#define URSM_SPLIT_FETCH(GL, GENE, REM, KM, AVAL)                                            \
  {                                                                                          \
    unsigned int g_ = 0;                                                                     \
    if (lane == 0) g_ = atomicAdd(&qctr[grp_local], 1u);                                     \
    GL = __shfl_sync(0xffffffffu, g_, 0);                                                    \
    if (GL < (unsigned int)ng) {                                                             \
      const int pos_ = (int)blockIdx.x + (int)GL * gx;                                       \
      GENE = a.order ? __ldg(a.order + pos_) : pos_;                                         \
      REM = __ldg(a.Xt + (size_t)pos_ * a.Mpad + jl);                                        \
      KM = a.kmax ? (int)__ldg(a.kmax + pos_) : K - 1;                                       \
      AVAL = lane < K ? __ldg(a.A32 + (size_t)lane * N + pos_) : 0.0f;                       \
    }                                                                                        \
  }
    URSM_SPLIT_FETCH(gl, gene, rem, km, aval)
    for (; gl < (unsigned int)ng; gl = gl_n, gene = gene_n, rem = rem_n, km = km_n, aval = aval_n) {
      URSM_SPLIT_FETCH(gl_n, gene_n, rem_n, km_n, aval_n)
      if (__all_sync(0xffffffffu, rem == 0.0f)) continue;  // nothing to split for these samples
      // chain order: the types in their order, except that the one with the largest A_ik trades
      // places with the last one; A_ik of the gene goes to the warp's A row in that order
      __syncwarp();
      if (lane < K) smf[ab + (lane == km ? K - 1 : (lane == K - 1 ? km : lane))] = aval;
      __syncwarp();
      // suffix sums of the masses along the chain, last step first
      float S = smf[wb + km * 32] * smf[ab + K - 1];
      smf[wb + 2 * K32 + (K - 1) * 32] = S;
      for (int t = K - 2; t >= 0; --t) {
        const int kk = (t == km) ? K - 1 : t;
        S += smf[wb + kk * 32] * smf[ab + t];
        smf[wb + 2 * K32 + t * 32] = S;
      }
      uint32_t w0, w1, w2, w3;
      philox_block(a.rk, (uint32_t)gene, jg, ctr2, 0u, w0, w1, w2, w3);
      const int ip = accI0 + (int)gl * K;
      bool gone = false;  // this lane has handed the rest of its chain to the cold list
#pragma unroll 1
      for (int t = 0; t < K - 1; ++t) {
        if (t != 0 && (t & 3) == 0)
          philox_block(a.rk, (uint32_t)gene, jg, ctr2, (uint32_t)(t >> 2), w0, w1, w2, w3);
        const float u = u01(w0);
        w0 = w1;
        w1 = w2;
        w2 = w3;
        const int kk = (t == km) ? K - 1 : t;
        const float q = smf[wb + kk * 32] * smf[ab + t];
        const float Sn = smf[wb + 2 * K32 + (t + 1) * 32];
        const float inv = URSM_RCP(fmaxf(S, 1e-37f));
        // P(this type | not yet assigned) and its complement, each formed from its own mass
        const float p = q * inv, pc = Sn * inv;
        S = Sn;
        const bool flip = p > 0.5f;
        const float r = flip ? pc : p, rc = flip ? p : pc;
        const float nr = rem * r;
        const bool big = nr > URSM_LS_ZERO_MAX;
        const bool table = rem < (float)URSM_BINOM_TABLE;
        // which search a lane's draw comes from depends on its own (n, r) alone -- never on the
        // other lanes of the warp -- so that a draw does not depend on how samples are grouped
        // (sharding invariance); the warp runs a search when any of its lanes needs it
        const bool inMode = big && table;
        const bool inZero = !big;
        const bool anyMode = __any_sync(0xffffffffu, inMode);
        float x = 0.0f;
        bool miss = false;
        if (anyMode) {
          const int ni = inMode ? (int)rem : 0;
          x = binom_mode_lockstep(ni, (float)ni, inMode ? r : 0.5f, inMode ? rc : 0.5f,
                                  inMode ? u : 0.0f, lf2, miss);
          if (PROBE) pr_mode += inMode;
        }
        if (__any_sync(0xffffffffu, inZero && rem > 0.0f)) {
          bool ms;
          const float xz = binom_zero_lockstep(rem, r, rc, inZero ? u : 0.0f, ms);
          if (inZero) {
            x = xz;
            miss = ms;
          }
          if (PROBE) pr_zero += (inZero && rem > 0.0f);
        }
        x = fminf(x, rem);  // (rounding residue beyond atom n can never yield more than n)
        float z = flip ? rem - x : x;
        if ((big && !table) || miss) {  // this lane's chain is finished by bk_split_cold_kernel
          const unsigned int slot = atomicAdd(a.cold_count, 1u);
          if (slot < a.cold_cap)
            a.cold_list[slot] = make_uint4(blockIdx.x + gl * (unsigned int)gx, jl,
                                           (unsigned int)t | (miss ? 0x100u : 0u), __float_as_uint(rem));
          gone = true;
          z = rem;  // (nothing left for the lockstep steps to come; not accumulated here)
          if (PROBE) {
            pr_deep += !miss;
            pr_miss += miss;
          }
        }
        rem -= z;
        const unsigned int zi = gone ? 0u : __float2uint_rn(z);
        smu[wb + K32 + kk * 32] += zi;
        if (keepI) {
          const unsigned int tot = __reduce_add_sync(0xffffffffu, zi);
          if (lane == 0 && tot != 0u) atomicAdd(&smu[ip + kk], tot);
        }
      }
      {  // the last step takes what is left
        const unsigned int zi = __float2uint_rn(rem);
        smu[wb + K32 + km * 32] += zi;
        if (keepI) {
          const unsigned int tot = __reduce_add_sync(0xffffffffu, zi);
          if (lane == 0 && tot != 0u) atomicAdd(&smu[ip + km], tot);
        }
      }
    }
#undef URSM_SPLIT_FETCH
    if ((long long)jl < a.M)
      for (int k = 0; k < K; ++k) {
        const unsigned int v = smu[wb + K32 + k * 32];
        if (v != 0u) atomicAdd(&a.n_cur[(size_t)jl * K + k], (double)v);
      }
  }
  if (PROBE && a.probe) {
    atomicAdd(a.probe + 0, pr_zero);
    atomicAdd(a.probe + 1, pr_mode);
    atomicAdd(a.probe + 2, pr_deep);
    atomicAdd(a.probe + 3, pr_miss);
  }
  if (keepI) {
    __syncthreads();
    for (int t = tid; t < ng * K; t += kSplitThreads) {
      const unsigned int v = smu[accI0 + t];
      const int gl = t / K, k = t - gl * K;
      if (v != 0u) {
        const int pos = (int)blockIdx.x + gl * gx;
        const int gene = a.order ? __ldg(a.order + pos) : pos;
        atomicAdd(&a.sumI[(size_t)k * N + gene], (double)v * a.inv_sample);
      }
    }
  }
}

// one resumable binomial run to completion by its lane (samplers.cuh): inversion from 0 for
// n r <= 20, BTRS above; further Philox blocks 0x100 + (step << 8) + attempt
template <class KeyT>  // PhiloxKeys (round keys from the constant bank) or the 64-bit key
__device__ __forceinline__ int binom_lane(int n, float p, float pc, float u, const KeyT& rk,
                                          uint32_t gene, uint32_t jg, uint32_t ctr2, int step,
                                          const double* lf2) {
  BinomState b;
  binom_begin(b, n, p, pc, u);
  int extra = 0;
  while (b.mode != kBinDone) {
    if (b.mode == kBinInv) {
      binom_inv_chunk4(b);
    } else {
      uint32_t e0, e1, e2, e3;
      philox_block(rk, gene, jg, ctr2, 0x100u + (uint32_t)(step << 8) + (uint32_t)extra, e0, e1, e2, e3);
      ++extra;
      if (b.mode == kBinBtrs) {
        binom_btrs_attempt(b, u01(e0), u01(e1), lf2);
        if (b.mode != kBinDone) binom_btrs_attempt(b, u01(e2), u01(e3), lf2);
      } else {  // kBinRetry
        binom_inv_start(b, u01(e0));
      }
    }
  }
  return b.result;
}

// The rest of the chains the lockstep kernel handed over: one thread per chain, from the step
// the chain was left at (its first draw from a fresh Philox block when it was a redraw).
__global__ void __launch_bounds__(128) bk_split_cold_kernel(BkSplitArgs a) {
  const unsigned int n = min(*a.cold_count, a.cold_cap);
  const unsigned int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx == 0 && *a.cold_count > a.cold_cap && a.fail) atomicAdd(a.fail, 1 << 20);
  if (idx >= n) return;
  const uint4 e = a.cold_list[idx];
  const int pos = (int)e.x, K = a.K, N = a.N;
  const int gene = a.order ? a.order[pos] : pos;
  const long long jl = e.y;
  const int t0 = (int)(e.z & 0xffu);
  const bool first_fresh = (e.z >> 8) != 0u;
  int rem = (int)__uint_as_float(e.w);
  const uint32_t jg = (uint32_t)(a.sample_offset + jl);
  const uint32_t ctr2 = kPurposeBulkZ | a.sweep;
  const int km = a.kmax ? (int)a.kmax[pos] : K - 1;
  const float* Ag = a.A32 + pos;
  const float* Wj = a.W32g + (size_t)(jl >> 5) * K * 32 + (jl & 31);
  double rest = (double)(Wj[km * 32] * Ag[(size_t)km * N]);  // mass of the steps t0 .. K-1
  for (int t = K - 2; t >= t0; --t) {
    const int kk = (t == km) ? K - 1 : t;
    rest += (double)(Wj[kk * 32] * Ag[(size_t)kk * N]);
  }
  for (int t = t0; t < K - 1 && rem > 0; ++t) {
    uint32_t w0, w1, w2, w3;
    const bool fresh = first_fresh && t == t0;
    philox_block(a.rk, (uint32_t)gene, jg, ctr2, fresh ? 0x80u + (uint32_t)(t << 8) : (uint32_t)(t >> 2),
                 w0, w1, w2, w3);
    const uint32_t wt = (fresh || (t & 3) == 0) ? w0 : ((t & 3) == 1 ? w1 : ((t & 3) == 2 ? w2 : w3));
    const int kk = (t == km) ? K - 1 : t;
    const float q = Wj[kk * 32] * Ag[(size_t)kk * N];
    const double after = fmax(rest - (double)q, 0.0);
    const float inv = URSM_RCP(fmaxf((float)rest, 1e-37f));
    const float p = q * inv, pc = (float)after * inv;
    rest = after;
    int z;
    if (!(p > 0.0f))
      z = 0;
    else if (!(pc > 0.0f))
      z = rem;
    else
      z = binom_lane(rem, p, pc, u01(wt), a.rk, (uint32_t)gene, jg, ctr2, t, a.lf2);
    rem -= z;
    if (z != 0) {
      atomicAdd(&a.n_cur[jl * K + kk], (double)z);
      if (a.retained) atomicAdd(&a.sumI[(size_t)kk * N + gene], (double)z * a.inv_sample);
    }
  }
  if (rem != 0) {  // the last step takes what is left
    atomicAdd(&a.n_cur[jl * K + km], (double)rem);
    if (a.retained) atomicAdd(&a.sumI[(size_t)km * N + gene], (double)rem * a.inv_sample);
  }
}

// ---------------------------------------------------------------------------
// The per-lane kernel (deep counts).  Same layout and accumulation as above, contiguous gene
// ranges per CTA, types in their natural order; each lane runs its own sampler: inversion from
// the mode for n < 1024 with n r <= 32 (samplers.cuh::binom_mode_try), the resumable
// inversion / BTRS state machine otherwise.
// ---------------------------------------------------------------------------
__device__ __noinline__ int binom_deep(int n, float p, float pc, float u, unsigned long long key,
                                       uint32_t gene, uint32_t jg, uint32_t ctr2, int k,
                                       const double* lf2) {
  return binom_lane(n, p, pc, u, key, gene, jg, ctr2, k, lf2);
}

__global__ void __launch_bounds__(kSplitThreads, URSM_BK_MINBLOCKS) bk_split_lane_kernel(BkSplitArgs a) {
  extern __shared__ __align__(16) unsigned char split_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = a.K, N = a.N, gpc = a.gpc;
  const int nphase = (kSplitThreads / 32) / gpc;
  const int grp_local = warp % gpc, phase = warp / gpc;
  const int grp = blockIdx.y * gpc + grp_local;
  const bool live = grp < a.ngroups && phase < nphase;
  const int g0 = blockIdx.x * a.genes_per_cta;
  const int ng = min(a.genes_per_cta, N - g0);
  double* lf2 = reinterpret_cast<double*>(split_smem);                       // [URSM_BINOM_TABLE]
  float* Wsm = reinterpret_cast<float*>(lf2 + URSM_BINOM_TABLE);              // [gpc][K][32]
  unsigned int* accJ = reinterpret_cast<unsigned int*>(Wsm + gpc * K * 32) + tid;  // [K][T] own column
  unsigned int* accI = accJ - tid + K * kSplitThreads;                         // [genes_per_cta][K]
  for (int t = tid; t < URSM_BINOM_TABLE; t += kSplitThreads) lf2[t] = a.lf2[t];
  for (int t = tid; t < gpc * K * 32; t += kSplitThreads) {
    const int gl = t / (K * 32);
    const int gg = blockIdx.y * gpc + gl;
    Wsm[t] = gg < a.ngroups ? a.W32g[(size_t)gg * K * 32 + (t - gl * K * 32)] : 0.f;
  }
  for (int k = 0; k < K; ++k) accJ[k * kSplitThreads] = 0u;
  const bool keepI = a.retained != 0;
  if (keepI)
    for (int t = tid; t < ng * K; t += kSplitThreads) accI[t] = 0u;
  __syncthreads();
  const uint32_t ctr2 = kPurposeBulkZ | a.sweep;
  const float* Wj = Wsm + grp_local * K * 32 + lane;  // W_jk at Wj[k * 32]
  const uint32_t jg = (uint32_t)(a.sample_offset + (long long)grp * 32 + lane);

  if (live) {
    for (int gl = phase; gl < ng; gl += nphase) {
      const int gene = g0 + gl;
      const float xf = a.Xt[(size_t)gene * a.Mpad + (size_t)grp * 32 + lane];
      int rem = (int)xf;
      if (__all_sync(0xffffffffu, rem == 0)) continue;  // nothing to split for these samples
      const float* Ai = a.A32 + gene;                    // A_ik at Ai[k * N] (warp-uniform)
      // AW_ij and the mass of the types still to come (fp64: exact enough down to the last type)
      double rest = 0.0;
      for (int k = 0; k < K; ++k) rest += (double)(Wj[k * 32] * __ldg(Ai + (size_t)k * N));
      uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma unroll 1
      for (int k = 0; k < K; ++k) {
        int z;
        if (k == K - 1) {
          z = rem;  // the last type takes what is left
        } else {
          if ((k & 3) == 0)
            philox_block(a.rk, (uint32_t)gene, jg, ctr2, (uint32_t)(k >> 2), w0, w1, w2, w3);
          const uint32_t wk = (k & 3) == 0 ? w0 : ((k & 3) == 1 ? w1 : ((k & 3) == 2 ? w2 : w3));
          const float qk = Wj[k * 32] * __ldg(Ai + (size_t)k * N);
          const double after = fmax(rest - (double)qk, 0.0);
          const float inv = URSM_RCP((float)rest);
          // P(type k | not yet assigned) and its complement, each formed from its own mass
          const float p = qk * inv, pc = (float)after * inv;
          rest = after;
          const bool flip = p > 0.5f;
          const float r = flip ? pc : p, rc = flip ? p : pc;
          if (rem <= 0 || !(r > 0.0f)) {
            z = flip ? rem : 0;
          } else if (binom_use_mode(rem, (float)rem * r)) {
            int x = binom_mode_try(rem, r, rc, u01(wk), lf2);
            for (int extra = 0; x < 0; ++extra) {  // uniform beyond the fp32 mass sum (~1e-6): redraw
              uint32_t e0, e1, e2, e3;
              philox_block(a.rk, (uint32_t)gene, jg, ctr2, 0x100u + (uint32_t)(k << 8) + (uint32_t)extra,
                           e0, e1, e2, e3);
              x = binom_mode_try(rem, r, rc, u01(e0), lf2);
            }
            z = flip ? rem - x : x;
          } else {
            z = binom_deep(rem, p, pc, u01(wk), a.key, (uint32_t)gene, jg, ctr2, k, lf2);
          }
          rem -= z;
        }
        accJ[k * kSplitThreads] += (unsigned int)z;
        if (keepI) {
          const unsigned int tot = __reduce_add_sync(0xffffffffu, (unsigned int)z);
          if (lane == 0 && tot != 0u) atomicAdd(&accI[gl * K + k], tot);
        }
      }
    }
    const long long j = (long long)grp * 32 + lane;
    if (j < a.M)
      for (int k = 0; k < K; ++k) {
        const unsigned int v = accJ[k * kSplitThreads];
        if (v != 0u) atomicAdd(&a.n_cur[j * K + k], (double)v);
      }
  }
  if (keepI) {
    __syncthreads();
    for (int t = tid; t < ng * K; t += kSplitThreads) {
      const unsigned int v = accI[t];
      const int gl = t / K, k = t - gl * K;
      if (v != 0u) atomicAdd(&a.sumI[(size_t)k * N + g0 + gl], (double)v * a.inv_sample);
    }
  }
}

#endif  // __CUDACC__

// Launch geometry of the split kernels.  gpc sample groups share a CTA (their per-gene sums
// meet in shared memory before one global atomic per gene and type); the genes are divided so
// that the grid is a whole number of waves of `per_sm` CTAs per SM.
struct BkSplitGeom {
  int ngroups, gpc, rows, genes_per_cta, gx;
  size_t smem, smem_lane;
};

inline BkSplitGeom bk_split_geometry(long long M, int N, int K, int num_sms) {
  BkSplitGeom g;
  g.ngroups = (int)((M + 31) / 32);
  g.gpc = 1;
  for (int p = 8; p >= 2; p >>= 1)
    if (g.ngroups % p == 0 || (p == 8 && g.ngroups >= 40)) {
      g.gpc = p;
      break;
    }
  if (const char* e = getenv("URSM_BK_GROUPS_PER_CTA")) {
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4 || v == 8) g.gpc = v;
  }
  g.rows = (g.ngroups + g.gpc - 1) / g.gpc;
  int per_sm = 4;
  if (const char* e = getenv("URSM_BK_CTAS_PER_SM")) per_sm = std::max(1, atoi(e));
  const long long wave = (long long)per_sm * num_sms;
  const int max_genes = std::max(8, (16 * 1024) / (4 * K));  // accI tile <= 16 KB
  long long gx = std::max<long long>(1, wave / g.rows);
  gx = std::max<long long>(gx, (N + max_genes - 1) / max_genes);
  gx = std::min<long long>(gx, N);
  // whole waves: round the CTA count up to a multiple of `wave` when several waves are needed
  if (gx * g.rows > wave) {
    const long long waves = (gx * g.rows + wave - 1) / wave;
    gx = std::min<long long>(N, std::max<long long>(gx, waves * wave / g.rows));
  }
  g.genes_per_cta = (int)((N + gx - 1) / gx);
  g.gx = (N + g.genes_per_cta - 1) / g.genes_per_cta;
  // the per-lane kernel: table, W of the CTA's groups, sum_i Z columns, sum_j Z tile
  g.smem_lane = (size_t)URSM_BINOM_TABLE * 8 + (size_t)g.gpc * K * 32 * 4 + (size_t)K * kSplitThreads * 4 +
                (size_t)g.genes_per_cta * K * 4;
  // the lockstep kernel: table, sum_j Z tile, per-warp blocks (bk_split_warp_words)
  g.smem = (size_t)URSM_BINOM_TABLE * 8 + (size_t)g.genes_per_cta * K * 4 +
           (size_t)(kSplitThreads / 32) * bk_split_warp_words(K) * 4;
  return g;
}

}  // namespace ursm
