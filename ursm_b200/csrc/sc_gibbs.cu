// Single-cell Gibbs E-step for sm_100a: e_step_gibbs.py:171-377 as ONE persistent
// kernel.
//
// Design:
//  * cells are independent inside an E-step (parameters frozen), so one CTA runs
//    ALL burnin + sample*thin sweeps of a cell with the chain state on chip:
//    per gene 4 B (w, fp32) + 4 B (draw_S threshold) + 2 B (E[S] counter) + 1 B (flags)
//    of shared memory.  HBM is touched once per cell: read the count row (which
//    genes are zero), write the E[S] counters.  psi, w, S never exist in HBM (the
//    reference keeps five L x N fp64/int64 arrays, :216-220,232).
//  * phase 1 of a sweep, draw_w (:318-323): a rejection sampler.  A warp owns the
//    32 g genes of its threads as a queue; every trip of the warp loop is ONE attempt
//    per lane (one Philox block and a fixed amount of arithmetic,
//    samplers.cuh::pg_attempt) and a lane whose attempt is accepted takes the next
//    gene of the queue, so neither rejections nor uneven runs leave lanes idle.  The
//    accepting lane also folds the PREVIOUS sweep's (w, S) into the coeffA/coeffAsq
//    partials (update_suffStats :245-270 needs w(t) with the kappa/tau drawn after
//    it) and stores the threshold form of the gene's S draw.
//  * phase 2, draw_S (:326-342): a sequential scan over the zero genes of a cell
//    through the running sum_AS.  It is kept EXACT: each update is rewritten as
//    "s_other > T_i" (samplers.cuh::s_threshold; T_i depends on the uniform only), a
//    thread scans its own run of g genes assuming the sweep-start sum as its incoming
//    value and records the smallest |s_other - T_i| it saw (its margin); a block-wide
//    exclusive scan of the per-thread net changes gives every thread its true
//    incoming sum; only threads whose incoming sum moved by more than their margin
//    re-scan (thresholds are in shared memory: three flops per zero gene); repeat to
//    the fixed point, which is the unique sequential solution.
//  * draw_kappa_tau (:345-377): five block reductions in fp64, closed-form 2x2
//    Cholesky draw by one thread, broadcast through shared memory.
//  * per-gene accumulators for coeffA/coeffAsq are CTA-private fp32 (shared memory for
//    small N, an L2-resident scratch row otherwise) and flushed with fp64 atomics
//    every 16 cells or when the CTA changes cell type.
#include "common.cuh"
#include "samplers.cuh"
#include "sc_shard.h"
#include "../../include/ursm_b200.h"

#include <algorithm>
#include <numeric>
#include <type_traits>

namespace ursm {

struct ScGibbsArgs {
  const float* Y;       // [L][N]
  const int* G;         // [L]
  const double* R;      // [L]
  const int* order;     // [L] cells sorted by type
  const float* A32;     // [K][slots] fp32, slot order (gene t g + j at j T + t)
  long long L;
  int N, K;
  long long cell_offset;
  int g, slots;
  double mu_k, prec_k, mu_t, prec_t;
  int burnin, nsweeps, thin, sample;
  int sweep0;           // first sweep index (debug: run sweep `sweep0` only)
  unsigned long long key;
  PhiloxKeys rk;        // round keys of `key` (constant-bank operands of the Philox rounds)
  void* cnt;            // [L][N] counters, cnt_w bytes each
  int cnt_w;
  double* cellmom;      // [4][L]
  double* stats;        // coeffA[K][N] | coeffAsq[K][N] | elbo | sk | skk | st | stt | scan failures
  float* scratch;       // [grid][slots][2]
  int flush_every;
  int* counter;
  int acc_red;          // accumulate the scratch row with reductions (no L1 left)
  // debug (all optional)
  const unsigned char* S0;  // [L][N]
  const double* kt0;        // [2][L]
  float* w_out;             // [L][N]
  unsigned char* S_out;     // [L][N]
  double* kt_out;           // [2][L]
  int* rounds_out;          // [L]
  int dump_all;             // w_out / S_out / kt_out hold EVERY sweep ([nsweeps][...]), not the last
  unsigned char* state;     // BIG variant: [grid][slots] x (w fp32 | threshold fp32 | counter u16)
};

// bit masks over a thread's run of genes: 32 bits, or 64 in the BIG variant
__device__ __forceinline__ int mask_ffs(unsigned int m) { return __ffs((int)m); }
__device__ __forceinline__ int mask_ffs(unsigned long long m) { return __ffsll((long long)m); }
__device__ __forceinline__ int mask_popc(unsigned int m) { return __popc(m); }
__device__ __forceinline__ int mask_popc(unsigned long long m) { return __popcll(m); }

// acc[slot] += (dA, dQ).  Shared memory or an L1-cached scratch row: plain 8-byte
// read-modify-write (one lane per slot at a time).  L2-resident scratch row and no L1 to
// speak of (long gene vectors): ONE fire-and-forget vector reduction instead of a
// load-add-store round trip.
template <bool ACC_SMEM>
__device__ __forceinline__ void acc_add(float2* p, float dA, float dQ, int use_red) {
  if (ACC_SMEM || !use_red) {
    float2 v = *p;
    v.x += dA;
    v.y += dQ;
    *p = v;
  } else {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(dA), "f"(dQ) : "memory");
  }
}

// Shared-memory layout of the per-gene chain state: thread t OWNS the run of genes
// [t g, (t+1) g) (the unit of the sequential draw_S scan) and gene (t, j) lives in slot
// j T + t.  Both access patterns are then conflict free without padding: phase 1 walks a
// warp's genes with consecutive lanes on consecutive owner threads, phase 2 walks a
// thread's own run with stride T.  The fp32 copy of A^T is stored in the same slot order
// (ursm_sc_set_params), so its loads are coalesced in both phases.
__device__ __forceinline__ int slot_of_gene(int i, int g, int T, float inv_g) {
  int t = (int)((float)i * inv_g);  // i / g for i < 2^22, fixed up below
  int j = i - t * g;
  if (j < 0) {
    --t;
    j += g;
  } else if (j >= g) {
    ++t;
    j -= g;
  }
  return j * T + t;
}

constexpr int kNormChunk = 64;  // (kappa, tau) normal pairs precomputed per refill

// BIG: gene vectors whose chain state (10 B per gene) does not fit in shared memory (N > ~22,700)
// keep w, the thresholds and the counters in a CTA-private, L2-resident scratch row instead, and
// a thread owns up to 64 genes (64-bit masks): any N <= 65,536.  Same code otherwise.
template <bool ACC_SMEM, bool BIG>
__global__ void __launch_bounds__(1024, 1) sc_gibbs_kernel(ScGibbsArgs a) {
  using MT = typename std::conditional<BIG, unsigned long long, unsigned int>::type;
  constexpr int kMaskBits = BIG ? 64 : 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = a.g, N = a.N, slots = a.slots, nw = T >> 5;
  const float inv_g = 1.0f / (float)g;

  double* red = reinterpret_cast<double*>(smem_raw);         // [32]    block_sum scratch
  double* redA = red + 32;                                    // [2][32] scan: warp totals
  double* red1 = redA + 64;                                   // [3][32] sum w, w a, w a^2 per warp
  double* nrm = red1 + 96;                                    // [kNormChunk][2] normal pairs
  double* bc = nrm + 2 * kNormChunk;                          // [4] broadcast
  int* redN = reinterpret_cast<int*>(bc + 4);                 // [2][32] scan: warp counts of S
  int* ibc = redN + 64;                                       // [4]
  unsigned char* st_base = BIG ? a.state + (size_t)blockIdx.x * slots * 10
                               : reinterpret_cast<unsigned char*>(ibc + 4);
  float* w_s = reinterpret_cast<float*>(st_base);             // [slots] w of the current sweep
  float* T_s = w_s + slots;                                   // [slots] draw_S thresholds
  // coeffA / coeffAsq partials of gene slot s: acc[s] = (dA, dQ), one 8-byte access
  float2* acc = reinterpret_cast<float2*>(ACC_SMEM ? (T_s + slots)
                                                   : (a.scratch + (size_t)blockIdx.x * 2 * slots));
  MT* Sm_s = reinterpret_cast<MT*>(
      BIG ? reinterpret_cast<unsigned char*>(ibc + 4)
          : reinterpret_cast<unsigned char*>(ACC_SMEM ? (T_s + 3 * (size_t)slots) : (T_s + slots)));  // [T]
  unsigned short* cnt_s = BIG ? reinterpret_cast<unsigned short*>(T_s + slots)
                              : reinterpret_cast<unsigned short*>(Sm_s + T);                 // [slots]

  const double inv_sample = 1.0 / (double)a.sample;
  double* coeffA = a.stats;
  double* coeffQ = a.stats + (size_t)a.K * N;
  double* scal = a.stats + 2 * (size_t)a.K * N;

  // this thread's run (phase 2): genes i0 .. i0 + gcount - 1, bit j of the masks below
  const int i0 = tid * g;
  const int gcount = max(0, min(g, N - i0));
  const MT validm = gcount >= kMaskBits ? ~(MT)0 : (((MT)1 << gcount) - (MT)1);
  // this warp's queue (phase 1): entry e -> owner thread wfirst + (e & 31), gene j = e >> 5
  const int wfirst = tid & ~31;
  const int qtotal = (wfirst * g < N) ? 32 * g : 0;
  unsigned int lt_mask;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));

  int cur_type = -1, cells_in_acc = 0;
  for (int s = tid; s < slots; s += T) acc[s] = make_float2(0.f, 0.f);

  while (true) {
    __syncthreads();
    if (tid == 0) ibc[0] = atomicAdd(a.counter, 1);
    __syncthreads();
    const long long c = ibc[0];
    const bool finished = c >= a.L;
    const int l = finished ? 0 : a.order[c];
    const int type = finished ? -1 : a.G[l];
    if (cur_type >= 0 && (finished || type != cur_type || cells_in_acc >= a.flush_every)) {
      // flush the CTA-private accumulators into coeffA/coeffAsq[:, cur_type]
      for (int i = tid; i < N; i += T) {
        const int slot = slot_of_gene(i, g, T, inv_g);
        const float2 v = acc[slot];
        if (v.x != 0.f) atomicAdd(&coeffA[(size_t)cur_type * N + i], (double)v.x * inv_sample);
        if (v.y != 0.f) atomicAdd(&coeffQ[(size_t)cur_type * N + i], (double)v.y * inv_sample);
        acc[slot] = make_float2(0.f, 0.f);
      }
      cells_in_acc = 0;
      __syncthreads();
    }
    if (finished) break;
    cur_type = type;
    ++cells_in_acc;

    const unsigned int gl = (unsigned int)(a.cell_offset + l);  // global cell id keys the RNG
    const float Rf = (float)a.R[l];
    const float invR = Rf > 0.f ? 1.0f / Rf : 0.f;
    const float* Acol = a.A32 + (size_t)type * slots;  // A[:, type] in SLOT order

    // ---- init_gibbs (:212-225): S ~ Bern(1/2), S[Y>0] = 1, kappa/tau = prior means.
    // The count row is read coalesced and staged through w_s; the owner threads then
    // build their masks: posm (Y > 0) and Sm (the chain's S, one bit per gene).
    for (int i = tid; i < N; i += T) {
      const int slot = slot_of_gene(i, g, T, inv_g);
      w_s[slot] = a.Y[(size_t)l * N + i];
      cnt_s[slot] = 0;
    }
    __syncthreads();
    MT posm = 0, Sm = 0;
    for (int j = 0; j < gcount; ++j) {
      const int i = i0 + j;
      const MT pos = w_s[j * T + tid] > 0.f ? 1 : 0;
      MT S;
      if (a.S0) {
        S = a.S0[(size_t)l * N + i] ? 1 : 0;
      } else {
        Philox r0(a.key, (uint32_t)i, gl, kPurposeInit);
        S = r0.next() >> 31;
      }
      posm |= pos << j;
      Sm |= (S | pos) << j;
    }
    const MT zerom = validm & ~posm;  // genes the draw_S scan visits
    Sm_s[tid] = Sm;
    double kappa = a.kt0 ? a.kt0[l] : a.mu_k;
    double tau = a.kt0 ? a.kt0[a.L + l] : a.mu_t;
    double s_cur;
    {
      double part = 0.0;
      for (MT m = Sm; m; m &= m - 1) part += (double)Acol[(mask_ffs(m) - 1) * T + tid];
      s_cur = block_sum(part, red);  // (its barriers also order the staging reads above)
    }

    double m_k = 0, m_kk = 0, m_t = 0, m_tt = 0, elbo = 0;  // meaningful in thread 0
    bool prev_ret = false;
    float pk = 0.f, pt = 0.f;
    int rounds_total = 0;

    for (int sweep = a.sweep0; sweep < a.sweep0 + a.nsweeps; ++sweep) {
      const bool retained = sweep >= a.burnin && ((sweep - a.burnin) % a.thin == 0);
      const float fk = (float)kappa, ft = (float)tau;
      const float cA1 = pt, cA2 = pt * pk, cQ = -0.5f * pt * pt;
      const uint32_t ctr2 = kPurposeGene | (uint32_t)sweep;
      // the normal pairs of the (kappa, tau) draws depend on (cell, sweep) only: threads
      // 0 .. kNormChunk-1 precompute a chunk of them off the critical path (the barriers
      // of this sweep's scan order the writes before thread 0 reads them)
      const int nrm_idx = (sweep - a.sweep0) % kNormChunk;
      if (nrm_idx == 0 && tid < kNormChunk && sweep + tid < a.sweep0 + a.nsweeps) {
        Philox rk(a.key, 0xffffffffu, gl, kPurposeCell | (uint32_t)(sweep + tid));
        double n0, n1;
        normal_pair(rk, n0, n1);
        nrm[2 * tid] = n0;
        nrm[2 * tid + 1] = n1;
      }
      double sw = 0, swa = 0, swaa = 0;
      // ---- phase 1: draw_w (:318-323), the deferred update_suffStats fold and the
      // draw_S thresholds, for the 32 g genes of this warp in ANY order.  Every trip
      // of the loop is one sampling attempt (one Philox block, fixed work) per lane; a
      // lane whose attempt is accepted takes the next gene of the warp's queue
      // (ballot + popc), so neither rejections nor uneven runs leave lanes idle.
      {
        int cursor = 0, i = 0, slot = 0, attempt = 0;
        bool need = true, active = false;
        float av = 0.f, psi = 0.f;
        uint32_t rS = 0u;
        PgTilt pg;
        while (true) {
          const unsigned int free_mask = __ballot_sync(0xffffffffu, need);
          if (need) {
            const int e = cursor + __popc(free_mask & lt_mask);
            if (e < qtotal) {
              const int to = wfirst + (e & 31), j = e >> 5;
              i = to * g + j;
              if (i < N) {
                need = false;
                slot = j * T + to;
                av = Acol[slot];
                if (prev_ret) {  // update_suffStats (:245-270) of the previous sweep
                  const unsigned int Sp = (unsigned int)((Sm_s[to] >> j) & 1);
                  const float wold = w_s[slot];
                  const float dA = ((float)Sp - 0.5f) * cA1 - wold * cA2, dQ = cQ * wold;
                  acc_add<ACC_SMEM>(acc + slot, dA, dQ, a.acc_red);
                  cnt_s[slot] += (unsigned short)Sp;
                }
                psi = fmaf(ft, av, fk);
                pg = pg_tilt(psi);
                attempt = 0;
                active = true;
              }
            } else {
              need = false;  // queue exhausted: this lane is done for the sweep
            }
          }
          cursor += __popc(free_mask);
          if (!__any_sync(0xffffffffu, active || need)) break;
          if (active) {
            uint32_t r0, r1, r2, r3;
            philox_block(a.rk, (uint32_t)i, gl, ctr2, (uint32_t)attempt, r0, r1, r2, r3);
            if (attempt == 0) rS = r3;  // word 3 of the first attempt: the S uniform
            float x;
            if (pg_attempt(pg, u01(r0), u01(r1), u01(r2), x)) {
              const float wn = 0.25f * x;
              w_s[slot] = wn;
              T_s[slot] = s_threshold(psi, u01(rS), av, Rf, invR);
              const double wd = wn, ad = av;
              sw += wd;
              swa += wd * ad;
              swaa += wd * ad * ad;
              active = false;
              need = true;
            } else {
              ++attempt;
            }
          }
        }
      }
      sw = warp_sum(sw);
      swa = warp_sum(swa);
      swaa = warp_sum(swaa);
      if (lane == 0) {
        red1[warp] = sw;
        red1[32 + warp] = swa;
        red1[64 + warp] = swaa;
      }
      __syncwarp();  // phase 2 reads what other lanes of this warp wrote
      // ---- phase 2: draw_S (:326-342), the in-cell sequential scan, kept EXACT:
      // every thread scans the zero genes of its own run from the sweep-start sum
      // (speculation), a block scan of the net changes gives the true incoming sums,
      // threads whose incoming sum moved by more than their margin re-scan (thresholds
      // are in shared memory, so a re-scan is a handful of flops per zero gene), to the
      // fixed point.  S lives in a register bit mask per owner thread.
      double s_run = s_cur, used_in = s_cur, total = 0.0;
      MT Snew_m = Sm;
      float margin = INFINITY;
      bool scan = true;
      int rounds = 0;
      while (true) {
        if (scan) {
          margin = INFINITY;
          s_run = used_in;
          Snew_m = Sm;
          for (MT m = zerom; m; m &= m - 1) {
            const int j = mask_ffs(m) - 1;
            const double ad = (double)Acol[j * T + tid];
            const double Tth = (double)T_s[j * T + tid];
            const double so = s_run - (((Sm >> j) & 1) ? ad : 0.0);
            const bool on = so > Tth;
            margin = fminf(margin, fabsf((float)(so - Tth)));
            s_run = so + (on ? ad : 0.0);
            Snew_m = on ? (Snew_m | ((MT)1 << j)) : (Snew_m & ~((MT)1 << j));
          }
        }
        // block-wide exclusive scan of the net changes (double-buffered warp totals)
        const double delta = s_run - used_in;
        double inc = delta;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        const int buf = (rounds & 1) * 32;
        const int nSw = __reduce_add_sync(0xffffffffu, mask_popc(Snew_m & validm));
        if (lane == 31) redA[buf + warp] = inc;
        if (lane == 0) redN[buf + warp] = nSw;
        __syncthreads();
        const double x = (lane < nw) ? redA[buf + lane] : 0.0;
        const double pre = warp_sum(lane < warp ? x : 0.0);
        total = warp_sum(x);
        const double incoming = s_cur + pre + inc - delta;
        const bool viol = !(fabs(incoming - used_in) < (double)margin) && incoming != used_in &&
                          margin != INFINITY;
        const bool more = __syncthreads_or(viol);
        // stats tail slot 5 counts (cell, sweep) pairs whose scan hit the round limit: never
        // seen, but the host raises if it ever is (it rides the statistics all-reduce)
        if (more && rounds > T && tid == 0) atomicAdd(&scal[5], 1.0);
        if (!more || rounds > T) {
          // ---- draw_kappa_tau (:345-377): warp 0 sums the per-warp partials, one thread
          // solves the 2x2 system and draws from the precomputed normal pair
          if (warp == 0) {
            double x0 = (lane < nw) ? red1[lane] : 0.0;
            double x1 = (lane < nw) ? red1[32 + lane] : 0.0;
            double x2 = (lane < nw) ? red1[64 + lane] : 0.0;
            int x4 = (lane < nw) ? redN[buf + lane] : 0;
            x0 = warp_sum(x0);
            x1 = warp_sum(x1);
            x2 = warp_sum(x2);
            x4 = warp_sum(x4);
            if (lane == 0) {
              const double x3 = s_cur + total;  // sum_AS after this sweep's S draws
              const double d0 = x0 + a.prec_k, d1 = x2 + a.prec_t, off = x1;
              const double det = d0 * d1 - off * off;
              const double c00 = d1 / det, c01 = -off / det, c11 = d0 / det;
              const double b0 = (double)x4 - 0.5 * N + a.mu_k * a.prec_k;
              const double b1 = x3 - 0.5 + a.mu_t * a.prec_t;
              const double m0 = c00 * b0 + c01 * b1, m1 = c01 * b0 + c11 * b1;
              const double l00 = sqrt(c00), l10 = c01 / l00;
              const double l11 = sqrt(fmax(c11 - l10 * l10, 0.0));
              const double n0 = nrm[2 * nrm_idx], n1 = nrm[2 * nrm_idx + 1];
              bc[0] = m0 + l00 * n0;
              bc[1] = m1 + l10 * n0 + l11 * n1;
              bc[2] = x3;
              bc[3] = x0;  // sum_i w
              ibc[1] = x4;
            }
          }
          break;
        }
        ++rounds;
        scan = viol;
        if (viol) used_in = incoming;
      }
      rounds_total += rounds;
      Sm = Snew_m;
      Sm_s[tid] = Sm;  // read by this warp's lanes in the next sweep's fold
      __syncthreads();
      kappa = bc[0];
      tau = bc[1];
      s_cur = bc[2];
      if (retained && tid == 0) {
        m_k += kappa;
        m_kk += kappa * kappa;
        m_t += tau;
        m_tt += tau * tau;
        // sum_i [-w kappa^2/2 + (S - 1/2) kappa]   (:254-257)
        elbo += -0.5 * kappa * kappa * bc[3] + kappa * ((double)ibc[1] - 0.5 * N);
      }
      prev_ret = retained;
      pk = (float)kappa;
      pt = (float)tau;
      if (a.dump_all) {  // parity hook (ursm_sc_chain_debug): the state after EVERY sweep
        const size_t sw_i = (size_t)(sweep - a.sweep0);
        const size_t base = (sw_i * a.L + l) * N;
        for (int i = tid; i < N; i += T) {
          const int slot = slot_of_gene(i, g, T, inv_g);
          a.w_out[base + i] = w_s[slot];
          a.S_out[base + i] = (unsigned char)((Sm_s[slot % T] >> (slot / T)) & 1);
        }
        if (tid == 0) {
          a.kt_out[sw_i * 2 * a.L + l] = kappa;
          a.kt_out[sw_i * 2 * a.L + a.L + l] = tau;
        }
        __syncthreads();  // the next sweep overwrites w_s
      }
    }
    // ---- fold of the last sweep (owner threads, own runs), write-back
    if (prev_ret) {
      const float cA1 = pt, cA2 = pt * pk, cQ = -0.5f * pt * pt;
      for (int j = 0; j < gcount; ++j) {
        const int slot = j * T + tid;
        const unsigned int Sold = (unsigned int)((Sm >> j) & 1);
        const float wold = w_s[slot];
        const float dA = ((float)Sold - 0.5f) * cA1 - wold * cA2, dQ = cQ * wold;
        acc_add<ACC_SMEM>(acc + slot, dA, dQ, a.acc_red);
        cnt_s[slot] += (unsigned short)Sold;
      }
    }
    __syncthreads();
    for (int i = tid; i < N; i += T) {
      const int slot = slot_of_gene(i, g, T, inv_g);
      if (a.cnt_w == 1)
        static_cast<unsigned char*>(a.cnt)[(size_t)l * N + i] = (unsigned char)cnt_s[slot];
      else
        static_cast<unsigned short*>(a.cnt)[(size_t)l * N + i] = cnt_s[slot];
      if (a.w_out && !a.dump_all) a.w_out[(size_t)l * N + i] = w_s[slot];
      if (a.S_out && !a.dump_all) {
        const int to = slot % T, j = slot / T;
        a.S_out[(size_t)l * N + i] = (unsigned char)((Sm_s[to] >> j) & 1);
      }
    }
    if (tid == 0) {
      a.cellmom[l] = m_k * inv_sample;
      a.cellmom[a.L + l] = m_t * inv_sample;
      a.cellmom[2 * a.L + l] = m_kk * inv_sample;
      a.cellmom[3 * a.L + l] = m_tt * inv_sample;
      atomicAdd(&scal[0], elbo * inv_sample);
      atomicAdd(&scal[1], m_k * inv_sample);
      atomicAdd(&scal[2], m_kk * inv_sample);
      atomicAdd(&scal[3], m_t * inv_sample);
      atomicAdd(&scal[4], m_tt * inv_sample);
      if (a.kt_out && !a.dump_all) {
        a.kt_out[l] = kappa;
        a.kt_out[a.L + l] = tau;
      }
      if (a.rounds_out) a.rounds_out[l] = rounds_total;
    }
  }
}

// ---------------------------------------------------------------------------
// update_suffStats (:245-270) on an injected state: parity check #1.
// One CTA per cell; fp64 throughout; same formulas as the fold above.
// ---------------------------------------------------------------------------
__global__ void sc_inject_kernel(const double* w, const unsigned char* S, const double* kappa,
                                 const double* tau, const int* G, long long L, int N, int K,
                                 double inv_sample, unsigned short* cnt, double* cellmom,
                                 double* stats) {
  __shared__ double red[32];
  const long long l = blockIdx.x;
  const int type = G[l];
  const double kp = kappa[l], tu = tau[l];
  double* coeffA = stats + (size_t)type * N;
  double* coeffQ = stats + (size_t)K * N + (size_t)type * N;
  double* scal = stats + 2 * (size_t)K * N;
  double e = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double wv = w[(size_t)l * N + i];
    const double Sv = S[(size_t)l * N + i] ? 1.0 : 0.0;
    e += -0.5 * wv * kp * kp + (Sv - 0.5) * kp;
    atomicAdd(&coeffA[i], ((Sv - 0.5) * tu - wv * tu * kp) * inv_sample);
    atomicAdd(&coeffQ[i], (-0.5 * wv * tu * tu) * inv_sample);
    cnt[(size_t)l * N + i] += (unsigned short)(Sv > 0.5);
  }
  e = block_sum(e, red);
  if (threadIdx.x == 0) {
    cellmom[l] += kp * inv_sample;
    cellmom[L + l] += tu * inv_sample;
    cellmom[2 * L + l] += kp * kp * inv_sample;
    cellmom[3 * L + l] += tu * tu * inv_sample;
    atomicAdd(&scal[0], e * inv_sample);
    atomicAdd(&scal[1], kp * inv_sample);
    atomicAdd(&scal[2], kp * kp * inv_sample);
    atomicAdd(&scal[3], tu * inv_sample);
    atomicAdd(&scal[4], tu * tu * inv_sample);
  }
}

__global__ void sc_rowsum_kernel(const float* Y, long long L, int N, double* R) {
  __shared__ double red[32];
  const long long l = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s += (double)Y[(size_t)l * N + i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) R[l] = s;
}

// fp64 -> fp32 storage of a count matrix; *bad is set when an entry is not a count that
// fp32 holds exactly (negative, non-finite or above 2^24)
__global__ void cast_f64_f32_kernel(const double* in, float* out, size_t n, int* bad) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  bool b = false;
  for (; i < n; i += stride) {
    const double v = in[i];
    b = b || !(v >= 0.0 && v <= 16777216.0);
    out[i] = (float)v;
  }
  if (b) *bad = 1;
}

// out[type][i] += sum over the chunk's rows of Y[row][i] / R[row]: one streaming pass over
// the counts (4 B per entry), a thread owns four adjacent genes (16-byte loads), four rows
// in flight, per-type partials flushed with fp64 atomics (std_row + the per-type means of
// gem.py:225-255, utils.py:34-38; a read-less row gives NaN as it does there).
__global__ void __launch_bounds__(256)
normalised_colsum_kernel(const float* Y, const int* order, const int* G, const double* R,
                         long long L, int N, int chunk, double* out) {
  const int i = 4 * (blockIdx.x * 256 + threadIdx.x);
  if (i >= N) return;
  const long long c0 = (long long)blockIdx.y * chunk, c1 = min(L, c0 + (long long)chunk);
  const bool vec = (i + 3 < N) && ((N & 3) == 0);
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  int cur = -1;
  for (long long c = c0; c < c1; c += 4) {
    float4 v[4];
    double w[4];
    int ty[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const bool ok = c + q < c1;
      const long long l = ok ? (order ? (long long)order[c + q] : c + q) : 0;
      ty[q] = ok ? (G ? G[l] : 0) : -1;
      w[q] = ok ? 1.0 / R[l] : 0.0;
      const float* row = Y + (size_t)l * N + i;
      if (!ok) {
        v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      } else if (vec) {
        v[q] = *reinterpret_cast<const float4*>(row);
      } else {
        v[q] = make_float4(row[0], i + 1 < N ? row[1] : 0.f, i + 2 < N ? row[2] : 0.f,
                           i + 3 < N ? row[3] : 0.f);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (ty[q] < 0) continue;
      if (ty[q] != cur) {
        if (cur >= 0)
          for (int r = 0; r < 4; ++r)
            if (i + r < N && acc[r] != 0.0) atomicAdd(&out[(size_t)cur * N + i + r], acc[r]);
        acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
        cur = ty[q];
      }
      acc[0] += (double)v[q].x * w[q];
      acc[1] += (double)v[q].y * w[q];
      acc[2] += (double)v[q].z * w[q];
      acc[3] += (double)v[q].w * w[q];
    }
  }
  if (cur >= 0)
    for (int r = 0; r < 4; ++r)
      if (i + r < N && !(acc[r] == 0.0)) atomicAdd(&out[(size_t)cur * N + i + r], acc[r]);
}

template <class CT>
__global__ void cnt_to_f64_kernel(const CT* cnt, double* out, size_t n, double inv) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = (double)cnt[i] * inv;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
void launch_rowsum(const float* Y, long long L, int N, double* R, cudaStream_t st) {
  sc_rowsum_kernel<<<(unsigned)L, 256, 0, st>>>(Y, L, N, R);
  URSM_LAUNCH_CHECK();
}

void launch_normalised_colsum(const float* Y, const int* order, const int* G, const double* R,
                              long long L, int N, int num_sms, double* out, cudaStream_t st) {
  const int gx = (N + 4 * 256 - 1) / (4 * 256);
  long long ny = std::max<long long>(1, (8LL * num_sms + gx - 1) / gx);
  ny = std::min<long long>(std::min<long long>(ny, L), 65535);
  const int chunk = (int)((L + ny - 1) / ny);
  ny = (L + chunk - 1) / chunk;
  normalised_colsum_kernel<<<dim3(gx, (unsigned)ny), 256, 0, st>>>(Y, order, G, R, L, N, chunk, out);
  URSM_LAUNCH_CHECK();
}

void upload_f64_as_f32(const double* h, float* d, size_t n) {
  // chunked: stage fp64 on the device, cast there
  const size_t chunk = (size_t)1 << 24;
  DevBuf<double> stage;
  DevBuf<int> bad;
  stage.alloc(std::min(n, chunk));
  bad.alloc(1);
  bad.zero();
  for (size_t off = 0; off < n; off += chunk) {
    size_t m = std::min(chunk, n - off);
    URSM_CUDA(cudaMemcpy(stage.p, h + off, m * sizeof(double), cudaMemcpyHostToDevice));
    cast_f64_f32_kernel<<<1184, 256>>>(stage.p, d + off, m, bad.p);
    URSM_LAUNCH_CHECK();
  }
  int h_bad = 0;
  URSM_CUDA(cudaMemcpy(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost));
  // the reference keeps counts in float64 (e_step_gibbs.py:137 splits them exactly); here they
  // are stored as fp32, exact up to 2^24 -- refuse anything else instead of rounding it
  URSM_REQUIRE(!h_bad, "count matrix holds an entry that is negative, not finite or above 2^24 "
                       "(counts are stored as fp32 on the device)");
}

// Shared memory of one CTA with T threads x g genes per thread: reductions / broadcast /
// normal pairs, then per gene slot 10 bytes (w, threshold, E[S] counter) + 8 for in-smem
// accumulators, and one S bit mask per thread.
static size_t sc_smem_bytes(int T, int g, bool acc_smem) {
  const size_t fixed = (32 + 64 + 96 + 2 * kNormChunk + 4) * 8 + (64 + 4) * 4 + (size_t)T * 4;
  return fixed + (size_t)T * g * (acc_smem ? 18 : 10) + 16;
}
// BIG variant: the per-gene state lives in global scratch; shared memory holds the fixed part
// and one 64-bit S mask per thread
static size_t sc_smem_bytes_big(int T) {
  return (32 + 64 + 96 + 2 * kNormChunk + 4) * 8 + (64 + 4) * 4 + (size_t)T * 8 + 16;
}

// Geometry: the kernel is bound by dependent-instruction latency, not by instruction
// count, so what matters is resident warps per SM (64 registers per thread and the
// cell's 10 B per gene of shared memory set the limit) times the fraction of lanes
// that own genes.  Candidates: g = 8..32 genes per thread, T = ceil(N / g) rounded
// up to a warp; ties go to the smaller CTA (cheaper barriers).
static void sc_pick_geometry(ScShard* s) {
  s->num_sms = device_sm_count();
  const size_t smem_optin = device_smem_optin();
  int maxT = 1024;
  if (const char* e = getenv("URSM_SC_MAX_THREADS")) maxT = std::max(64, std::min(1024, atoi(e)));
  maxT = maxT / 32 * 32;
  int acc_limit = 32 * 1024;
  if (const char* e = getenv("URSM_SC_ACC_SMEM_LIMIT")) acc_limit = atoi(e);
  int g_lo = 8, g_hi = 32;
  if (const char* e = getenv("URSM_SC_GENES_PER_THREAD")) g_lo = g_hi = std::max(1, std::min(32, atoi(e)));
  URSM_CUDA(cudaFuncSetAttribute(sc_gibbs_kernel<true, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin));
  URSM_CUDA(cudaFuncSetAttribute(sc_gibbs_kernel<false, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin));
  double best = -1.0;
  bool force_big = false;
  if (const char* e = getenv("URSM_SC_STATE_GLOBAL")) force_big = atoi(e) != 0;  // tests
  for (int gpt = g_lo; gpt <= g_hi && !force_big; ++gpt) {
    int T = ((s->N + gpt - 1) / gpt + 31) / 32 * 32;
    T = std::max(64, std::min(maxT, T));
    const int g = (s->N + T - 1) / T;
    if (g > 32) continue;
    const bool acc = sc_smem_bytes(T, g, true) <= (size_t)acc_limit;
    const size_t smem = sc_smem_bytes(T, g, acc);
    if (smem > smem_optin) continue;
    int occ = 0;
    if (acc) {
      URSM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sc_gibbs_kernel<true, false>, T,
                                                              smem));
    } else {
      URSM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sc_gibbs_kernel<false, false>, T,
                                                              smem));
    }
    if (occ < 1) continue;
    const double score = (double)occ * (T / 32) * ((double)s->N / ((double)T * g));
    if (score > best * (1.0 + 1e-9) || (score > best * (1.0 - 1e-9) && T < s->T)) {
      best = score;
      s->T = T;
      s->g = g;
      s->acc_smem = acc;
      s->smem = (int)smem;
      s->occ = occ;
    }
  }
  s->big = false;
  if (best <= 0.0) {
    // the chain state of a cell (10 B per gene) does not fit in shared memory: keep it in a
    // CTA-private scratch row in global memory (L2-resident) and let a thread own up to 64 genes
    int T = ((s->N + g_hi - 1) / g_hi + 31) / 32 * 32;  // same genes-per-thread target as above
    T = std::max(64, std::min(maxT, T));
    const int g = (s->N + T - 1) / T;
    URSM_REQUIRE(g <= 64, "gene count too large for the single-cell kernel (at most 65,536 genes)");
    s->T = T;
    s->g = g;
    s->acc_smem = false;
    s->smem = (int)sc_smem_bytes_big(T);
    s->big = true;
    int occ = 0;
    URSM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sc_gibbs_kernel<false, true>, T,
                                                            (size_t)s->smem));
    URSM_REQUIRE(occ >= 1, "single-cell kernel (large gene vectors): no resident CTA fits");
    s->occ = occ;
  }
  s->slots = s->T * s->g;
  long long want = (long long)s->num_sms * s->occ;
  if (const char* e = getenv("URSM_SC_GRID"))  // tests: many cells (and type changes) per CTA
    want = std::max(1, atoi(e));
  s->grid = (int)std::max<long long>(1, std::min<long long>(want, s->L));
  if (!s->acc_smem) s->scratch.alloc((size_t)s->grid * 2 * s->slots);
  if (s->big) s->state.alloc((size_t)s->grid * s->slots * 10);
}

static void sc_finish_create(ScShard* s, const int64_t* h_G) {
  const long long L = s->L;
  std::vector<int> G(L), order(L);
  s->type_count.assign(s->K, 0);
  for (long long l = 0; l < L; ++l) {
    URSM_REQUIRE(h_G[l] >= 0 && h_G[l] < s->K, "cell type outside {0..K-1}");
    G[l] = (int)h_G[l];
    s->type_count[G[l]]++;
  }
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return G[x] < G[y]; });
  s->type_start.assign(s->K + 1, 0);
  for (int k = 0; k < s->K; ++k) s->type_start[k + 1] = s->type_start[k] + s->type_count[k];
  s->G.alloc(L);
  s->order.alloc(L);
  URSM_CUDA(cudaMemcpy(s->G.p, G.data(), L * sizeof(int), cudaMemcpyHostToDevice));
  URSM_CUDA(cudaMemcpy(s->order.p, order.data(), L * sizeof(int), cudaMemcpyHostToDevice));
  s->R.alloc(L);
  sc_rowsum_kernel<<<(unsigned)L, 256>>>(s->Y, L, s->N, s->R.p);
  URSM_LAUNCH_CHECK();
  s->cellmom.alloc(4 * (size_t)L);
  s->cellmom.zero();
  s->counter.alloc(1);
  s->A.alloc((size_t)s->N * s->K);
  sc_pick_geometry(s);
  s->Aslot.alloc((size_t)s->K * s->slots);
  s->sample = 1;
  URSM_CUDA(cudaDeviceSynchronize());
}

static void sc_launch(ScShard* s, ScGibbsArgs& a, cudaStream_t st) {
  a.Y = s->Y;
  a.G = s->G.p;
  a.R = s->R.p;
  a.order = s->order.p;
  a.A32 = s->Aslot.p;
  a.L = s->L;
  a.N = s->N;
  a.K = s->K;
  a.cell_offset = s->cell_offset;
  a.g = s->g;
  a.slots = s->slots;
  a.mu_k = s->pkappa[0];
  a.prec_k = 1.0 / s->pkappa[1];
  a.mu_t = s->ptau[0];
  a.prec_t = 1.0 / s->ptau[1];
  a.cnt = s->cnt.p;
  a.cnt_w = s->cnt_w;
  a.cellmom = s->cellmom.p;
  a.scratch = s->scratch.p;
  a.flush_every = 16;
  a.counter = s->counter.p;
  URSM_CUDA(cudaMemsetAsync(s->counter.p, 0, sizeof(int), st));
  a.acc_red = (!s->acc_smem && (s->smem > 128 * 1024 || s->big)) ? 1 : 0;
  if (const char* e = getenv("URSM_SC_ACC_RED"))  // tests: the reduction path on small inputs
    a.acc_red = (!s->acc_smem && atoi(e)) ? 1 : 0;
  if (!s->ev0) {
    URSM_CUDA(cudaEventCreate(&s->ev0));
    URSM_CUDA(cudaEventCreate(&s->ev1));
  }
  URSM_CUDA(cudaEventRecord(s->ev0, st));
  a.state = s->state.p;
  if (s->big)
    sc_gibbs_kernel<false, true><<<s->grid, s->T, s->smem, st>>>(a);
  else if (s->acc_smem)
    sc_gibbs_kernel<true, false><<<s->grid, s->T, s->smem, st>>>(a);
  else
    sc_gibbs_kernel<false, false><<<s->grid, s->T, s->smem, st>>>(a);
  URSM_LAUNCH_CHECK();
  URSM_CUDA(cudaEventRecord(s->ev1, st));
  s->timed = true;
}

}  // namespace ursm

using namespace ursm;

extern "C" {

int ursm_sc_create(ursm_sc** out, const double* h_Y, const int64_t* h_G, int64_t L, int32_t N,
                   int32_t K, int64_t cell_offset) {
  URSM_API_BEGIN
  URSM_REQUIRE(out && h_Y && h_G && L > 0 && N > 0 && K > 0, "ursm_sc_create: bad argument");
  URSM_REQUIRE(L < (1ll << 31), "ursm_sc_create: more than 2^31 cells in one shard");
  ScShard* s = new ScShard();
  s->L = L;
  s->N = N;
  s->K = K;
  s->cell_offset = cell_offset;
  try {
    s->Yown.alloc((size_t)L * N);
    s->Y = s->Yown.p;
    upload_f64_as_f32(h_Y, s->Yown.p, (size_t)L * N);
    sc_finish_create(s, h_G);
  } catch (...) {
    delete s;
    throw;
  }
  *out = reinterpret_cast<ursm_sc*>(s);
  URSM_API_END
}

int ursm_sc_create_dev(ursm_sc** out, const float* d_Y, const int64_t* h_G, int64_t L, int32_t N,
                       int32_t K, int64_t cell_offset) {
  URSM_API_BEGIN
  URSM_REQUIRE(out && d_Y && h_G && L > 0 && N > 0 && K > 0, "ursm_sc_create_dev: bad argument");
  URSM_REQUIRE(L < (1ll << 31), "ursm_sc_create_dev: more than 2^31 cells in one shard");
  ScShard* s = new ScShard();
  s->L = L;
  s->N = N;
  s->K = K;
  s->cell_offset = cell_offset;
  s->Y = d_Y;  // borrowed: caller keeps it alive
  try {
    sc_finish_create(s, h_G);
  } catch (...) {
    delete s;
    throw;
  }
  *out = reinterpret_cast<ursm_sc*>(s);
  URSM_API_END
}

int ursm_sc_destroy(ursm_sc* sc) {
  URSM_API_BEGIN
  cudaDeviceSynchronize();  // the pool frees below are ordered on the default stream only
  delete reinterpret_cast<ScShard*>(sc);
  URSM_API_END
}

int64_t ursm_sc_stats_len(const ursm_sc* sc) {
  const ScShard* s = reinterpret_cast<const ScShard*>(sc);
  return 2 * (int64_t)s->N * s->K + 6;
}

int ursm_sc_geometry(const ursm_sc* sc, int32_t* threads, int32_t* genes_per_thread, int32_t* grid,
                     int32_t* smem_bytes, int32_t* acc_in_smem) {
  const ScShard* s = reinterpret_cast<const ScShard*>(sc);
  if (threads) *threads = s->T;
  if (genes_per_thread) *genes_per_thread = s->g;
  if (grid) *grid = s->grid;
  if (smem_bytes) *smem_bytes = s->smem;
  if (acc_in_smem) *acc_in_smem = s->big ? -1 : (s->acc_smem ? 1 : 0);  // -1: state in global scratch
  return 0;
}

int ursm_sc_kernel_ms(ursm_sc* sc, float* ms) {
  URSM_API_BEGIN
  ScShard* s = reinterpret_cast<ScShard*>(sc);
  URSM_REQUIRE(s && ms && s->timed, "ursm_sc_kernel_ms: no Gibbs kernel has been launched");
  URSM_CUDA(cudaEventSynchronize(s->ev1));
  URSM_CUDA(cudaEventElapsedTime(ms, s->ev0, s->ev1));
  URSM_API_END
}

int ursm_sc_set_params(ursm_sc* sc, const double* h_A, const double* h_pkappa,
                       const double* h_ptau) {
  URSM_API_BEGIN
  ScShard* s = reinterpret_cast<ScShard*>(sc);
  URSM_REQUIRE(s && h_A && h_pkappa && h_ptau, "ursm_sc_set_params: bad argument");
  URSM_REQUIRE(h_pkappa[1] > 0 && h_ptau[1] > 0, "ursm_sc_set_params: prior variance must be > 0");
  const int N = s->N, K = s->K;
  URSM_CUDA(cudaMemcpy(s->A.p, h_A, (size_t)N * K * sizeof(double), cudaMemcpyHostToDevice));
  // A^T in fp32, in the kernel's SLOT order: gene i = t g + j sits at j T + t
  std::vector<float> slot((size_t)K * s->slots, 0.f);
  for (int k = 0; k < K; ++k)
    for (int i = 0; i < N; ++i)
      slot[(size_t)k * s->slots + (size_t)(i % s->g) * s->T + i / s->g] =
          (float)h_A[(size_t)i * K + k];
  URSM_CUDA(cudaMemcpy(s->Aslot.p, slot.data(), slot.size() * sizeof(float),
                       cudaMemcpyHostToDevice));
  s->pkappa[0] = h_pkappa[0];
  s->pkappa[1] = h_pkappa[1];
  s->ptau[0] = h_ptau[0];
  s->ptau[1] = h_ptau[1];
  s->has_params = true;
  URSM_API_END
}

int ursm_sc_gibbs(ursm_sc* sc, int32_t burnin, int32_t sample, int32_t thin, uint64_t seed,
                  uint64_t em_iter, double* d_stats, void* stream) {
  URSM_API_BEGIN
  ScShard* s = reinterpret_cast<ScShard*>(sc);
  URSM_REQUIRE(s && d_stats, "ursm_sc_gibbs: bad argument");
  URSM_REQUIRE(s->has_params, "ursm_sc_gibbs: call ursm_sc_set_params first");
  URSM_REQUIRE(burnin >= 0 && sample >= 1 && thin >= 1, "ursm_sc_gibbs: bad sweep counts");
  URSM_REQUIRE(sample <= 65535, "ursm_sc_gibbs: sample must be <= 65535 (uint16 E[S] counters)");
  URSM_REQUIRE((long long)burnin + (long long)sample * thin < (1 << 28),
               "ursm_sc_gibbs: too many sweeps");
  cudaStream_t st = (cudaStream_t)stream;
  URSM_CUDA(cudaMemsetAsync(d_stats, 0, (size_t)ursm_sc_stats_len(sc) * sizeof(double), st));
  ScGibbsArgs a;
  memset(&a, 0, sizeof a);
  a.burnin = burnin;
  a.nsweeps = burnin + sample * thin;
  a.thin = thin;
  a.sample = sample;
  a.sweep0 = 0;
  a.key = seed + 0x9E3779B97F4A7C15ull * (em_iter + 1);
  a.rk = philox_keys(a.key);
  a.stats = d_stats;
  s->sample = sample;
  s->cnt_ensure((sample <= 255 && !getenv("URSM_SC_CNT16")) ? 1 : 2);
  sc_launch(s, a, st);
  URSM_API_END
}

int ursm_sc_type_sums(ursm_sc* sc, double* d_out, void* stream) {
  URSM_API_BEGIN
  ScShard* s = reinterpret_cast<ScShard*>(sc);
  URSM_REQUIRE(s && d_out, "ursm_sc_type_sums: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  URSM_CUDA(cudaMemsetAsync(d_out, 0, (size_t)s->K * s->N * sizeof(double), st));
  launch_normalised_colsum(s->Y, s->order.p, s->G.p, s->R.p, s->L, s->N, s->num_sms, d_out, st);
  URSM_API_END
}

int ursm_sc_get_cell_moments(ursm_sc* sc, double* h_exp_kappa, double* h_exp_tau,
                             double* h_exp_kappasq, double* h_exp_tausq) {
  URSM_API_BEGIN
  ScShard* s = reinterpret_cast<ScShard*>(sc);
  const size_t L = s->L, b = L * sizeof(double);
  URSM_CUDA(cudaDeviceSynchronize());
  if (h_exp_kappa) URSM_CUDA(cudaMemcpy(h_exp_kappa, s->cellmom.p, b, cudaMemcpyDeviceToHost));
  if (h_exp_tau) URSM_CUDA(cudaMemcpy(h_exp_tau, s->cellmom.p + L, b, cudaMemcpyDeviceToHost));
  if (h_exp_kappasq)
    URSM_CUDA(cudaMemcpy(h_exp_kappasq, s->cellmom.p + 2 * L, b, cudaMemcpyDeviceToHost));
  if (h_exp_tausq)
    URSM_CUDA(cudaMemcpy(h_exp_tausq, s->cellmom.p + 3 * L, b, cudaMemcpyDeviceToHost));
  URSM_API_END
}

int ursm_sc_get_exp_S(ursm_sc* sc, double* h_out, int64_t row0, int64_t nrows) {
  URSM_API_BEGIN
  ScShard* s = reinterpret_cast<ScShard*>(sc);
  URSM_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= s->L, "ursm_sc_get_exp_S: bad rows");
  URSM_REQUIRE(s->cnt_w > 0, "ursm_sc_get_exp_S: no E-step has run yet");
  const size_t rows_per = std::max<size_t>(1, ((size_t)1 << 24) / s->N);
  DevBuf<double> stage;
  stage.alloc(std::min<size_t>(rows_per, (size_t)nrows) * s->N);
  URSM_CUDA(cudaDeviceSynchronize());
  for (int64_t r = 0; r < nrows; r += rows_per) {
    size_t m = std::min<size_t>(rows_per, (size_t)(nrows - r)) * s->N;
    const size_t first = (size_t)(row0 + r) * s->N;
    if (s->cnt_w == 1)
      cnt_to_f64_kernel<<<1184, 256>>>(s->cnt.p + first, stage.p, m, 1.0 / s->sample);
    else
      cnt_to_f64_kernel<<<1184, 256>>>(reinterpret_cast<const unsigned short*>(s->cnt.p) + first,
                                       stage.p, m, 1.0 / s->sample);
    URSM_LAUNCH_CHECK();
    URSM_CUDA(cudaMemcpy(h_out + (size_t)r * s->N, stage.p, m * sizeof(double),
                         cudaMemcpyDeviceToHost));
  }
  URSM_API_END
}

int ursm_sc_inject(ursm_sc* sc, const double* h_w, const int64_t* h_S, const double* h_kappa,
                   const double* h_tau, int32_t sample, int32_t reset, double* d_stats,
                   void* stream) {
  URSM_API_BEGIN
  ScShard* s = reinterpret_cast<ScShard*>(sc);
  URSM_REQUIRE(s && h_w && h_S && h_kappa && h_tau && d_stats && sample >= 1,
               "ursm_sc_inject: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t LN = (size_t)s->L * s->N;
  if (reset || s->cnt_w != 2) {
    URSM_REQUIRE(reset, "ursm_sc_inject: accumulate onto a Gibbs run is not supported; reset first");
    s->cnt_ensure(2);  // the parity hook always counts in 16 bits
    URSM_CUDA(cudaMemsetAsync(d_stats, 0, (size_t)ursm_sc_stats_len(sc) * sizeof(double), st));
    s->cnt.zero(st);
    s->cellmom.zero(st);
  }
  s->sample = sample;
  DevBuf<double> w, kt;
  DevBuf<unsigned char> S;
  w.alloc(LN);
  S.alloc(LN);
  kt.alloc(2 * (size_t)s->L);
  std::vector<unsigned char> S8(LN);
  for (size_t i = 0; i < LN; ++i) S8[i] = h_S[i] != 0;
  URSM_CUDA(cudaMemcpyAsync(w.p, h_w, LN * sizeof(double), cudaMemcpyHostToDevice, st));
  URSM_CUDA(cudaMemcpyAsync(S.p, S8.data(), LN, cudaMemcpyHostToDevice, st));
  URSM_CUDA(cudaMemcpyAsync(kt.p, h_kappa, s->L * sizeof(double), cudaMemcpyHostToDevice, st));
  URSM_CUDA(cudaMemcpyAsync(kt.p + s->L, h_tau, s->L * sizeof(double), cudaMemcpyHostToDevice, st));
  sc_inject_kernel<<<(unsigned)s->L, 256, 0, st>>>(w.p, S.p, kt.p, kt.p + s->L, s->G.p, s->L, s->N,
                                                   s->K, 1.0 / sample,
                                                   reinterpret_cast<unsigned short*>(s->cnt.p), s->cellmom.p,
                                                   d_stats);
  URSM_LAUNCH_CHECK();
  URSM_CUDA(cudaStreamSynchronize(st));
  URSM_API_END
}

// Parity hook: `nsweeps` production gibbs_cycles from a supplied (S, kappa, tau), starting at
// sweep index `sweep0`, with the retention schedule (burnin, thin, sample) of ursm_sc_gibbs.
// Returns the state after EVERY sweep and what the production update_suffStats fold made of
// it: the packed statistics vector, the E[S] counters and the per-cell moments.
static void sc_chain_debug(ScShard* s, const int64_t* h_S_in, const double* h_kappa_in,
                           const double* h_tau_in, uint64_t seed, uint64_t em_iter, int sweep0,
                           int nsweeps, int burnin, int thin, int sample, bool dump_all,
                           float* h_w_out, uint8_t* h_S_out, double* h_kt_out,
                           int32_t* h_scan_rounds, double* h_stats, uint16_t* h_cnt,
                           double* h_cellmom) {
  const size_t LN = (size_t)s->L * s->N, L = s->L;
  const size_t reps = dump_all ? (size_t)nsweeps : 1;
  DevBuf<unsigned char> S0, Sout;
  DevBuf<double> kt0, ktout, stats;
  DevBuf<float> wout;
  DevBuf<int> rounds;
  S0.alloc(LN);
  Sout.alloc(LN * reps);
  wout.alloc(LN * reps);
  kt0.alloc(2 * L);
  ktout.alloc(2 * L * reps);
  rounds.alloc(L);
  const size_t nstats = (size_t)(2 * (int64_t)s->N * s->K + 6);
  stats.alloc(nstats);
  stats.zero();
  std::vector<unsigned char> S8(LN);
  for (size_t i = 0; i < LN; ++i) S8[i] = h_S_in[i] != 0;
  URSM_CUDA(cudaMemcpy(S0.p, S8.data(), LN, cudaMemcpyHostToDevice));
  URSM_CUDA(cudaMemcpy(kt0.p, h_kappa_in, L * sizeof(double), cudaMemcpyHostToDevice));
  URSM_CUDA(cudaMemcpy(kt0.p + L, h_tau_in, L * sizeof(double), cudaMemcpyHostToDevice));
  ScGibbsArgs a;
  memset(&a, 0, sizeof a);
  a.burnin = burnin;
  a.nsweeps = nsweeps;
  a.thin = thin;
  a.sample = sample;
  a.sweep0 = sweep0;
  a.key = seed + 0x9E3779B97F4A7C15ull * (em_iter + 1);
  a.rk = philox_keys(a.key);
  a.stats = stats.p;
  a.S0 = S0.p;
  a.kt0 = kt0.p;
  a.w_out = wout.p;
  a.S_out = Sout.p;
  a.kt_out = ktout.p;
  a.rounds_out = rounds.p;
  a.dump_all = dump_all ? 1 : 0;
  s->sample = sample;
  s->cnt_ensure(2);
  sc_launch(s, a, 0);
  URSM_CUDA(cudaDeviceSynchronize());
  if (h_w_out)
    URSM_CUDA(cudaMemcpy(h_w_out, wout.p, LN * reps * sizeof(float), cudaMemcpyDeviceToHost));
  if (h_S_out) URSM_CUDA(cudaMemcpy(h_S_out, Sout.p, LN * reps, cudaMemcpyDeviceToHost));
  if (h_kt_out)
    URSM_CUDA(cudaMemcpy(h_kt_out, ktout.p, 2 * L * reps * sizeof(double), cudaMemcpyDeviceToHost));
  if (h_scan_rounds)
    URSM_CUDA(cudaMemcpy(h_scan_rounds, rounds.p, L * sizeof(int), cudaMemcpyDeviceToHost));
  if (h_stats)
    URSM_CUDA(cudaMemcpy(h_stats, stats.p, nstats * sizeof(double), cudaMemcpyDeviceToHost));
  if (h_cnt) URSM_CUDA(cudaMemcpy(h_cnt, s->cnt.p, LN * sizeof(uint16_t), cudaMemcpyDeviceToHost));
  if (h_cellmom)
    URSM_CUDA(cudaMemcpy(h_cellmom, s->cellmom.p, 4 * L * sizeof(double), cudaMemcpyDeviceToHost));
}

int ursm_sc_set_counts(ursm_sc* sc, const uint16_t* h_cnt, int32_t sample, const double* h_cellmom) {
  URSM_API_BEGIN
  ScShard* s = reinterpret_cast<ScShard*>(sc);
  URSM_REQUIRE(s && h_cnt && sample >= 1 && sample <= 65535, "ursm_sc_set_counts: bad argument");
  s->cnt_ensure(2);
  s->sample = sample;
  URSM_CUDA(cudaMemcpy(s->cnt.p, h_cnt, (size_t)s->L * s->N * sizeof(uint16_t),
                       cudaMemcpyHostToDevice));
  if (h_cellmom)
    URSM_CUDA(cudaMemcpy(s->cellmom.p, h_cellmom, 4 * (size_t)s->L * sizeof(double),
                         cudaMemcpyHostToDevice));
  URSM_API_END
}

int ursm_sc_sweep_debug(ursm_sc* sc, const int64_t* h_S_in, const double* h_kappa_in,
                        const double* h_tau_in, uint64_t seed, uint64_t em_iter, int32_t sweep,
                        float* h_w_out, uint8_t* h_S_out, double* h_kappa_out, double* h_tau_out,
                        int32_t* h_scan_rounds) {
  URSM_API_BEGIN
  ScShard* s = reinterpret_cast<ScShard*>(sc);
  URSM_REQUIRE(s && h_S_in && h_kappa_in && h_tau_in, "ursm_sc_sweep_debug: bad argument");
  URSM_REQUIRE(s->has_params, "ursm_sc_sweep_debug: call ursm_sc_set_params first");
  std::vector<double> kt(2 * (size_t)s->L);
  // the single sweep counts as retained (burnin = sweep, sample = 1)
  sc_chain_debug(s, h_S_in, h_kappa_in, h_tau_in, seed, em_iter, sweep, 1, sweep, 1, 1, false,
                 h_w_out, h_S_out, kt.data(), h_scan_rounds, nullptr, nullptr, nullptr);
  if (h_kappa_out) memcpy(h_kappa_out, kt.data(), s->L * sizeof(double));
  if (h_tau_out) memcpy(h_tau_out, kt.data() + s->L, s->L * sizeof(double));
  URSM_API_END
}

int ursm_sc_chain_debug(ursm_sc* sc, const int64_t* h_S_in, const double* h_kappa_in,
                        const double* h_tau_in, uint64_t seed, uint64_t em_iter, int32_t burnin,
                        int32_t sample, int32_t thin, float* h_w_out, uint8_t* h_S_out,
                        double* h_kt_out, double* h_stats, uint16_t* h_cnt, double* h_cellmom) {
  URSM_API_BEGIN
  ScShard* s = reinterpret_cast<ScShard*>(sc);
  URSM_REQUIRE(s && h_S_in && h_kappa_in && h_tau_in, "ursm_sc_chain_debug: bad argument");
  URSM_REQUIRE(s->has_params, "ursm_sc_chain_debug: call ursm_sc_set_params first");
  URSM_REQUIRE(burnin >= 0 && sample >= 1 && thin >= 1 && burnin + sample * thin <= 4096,
               "ursm_sc_chain_debug: bad sweep counts");
  sc_chain_debug(s, h_S_in, h_kappa_in, h_tau_in, seed, em_iter, 0, burnin + sample * thin, burnin,
                 thin, sample, true, h_w_out, h_S_out, h_kt_out, nullptr, h_stats, h_cnt, h_cellmom);
  URSM_API_END
}

}  // extern "C"
