// Device-resident state of one single-cell shard (LogitNormalGibbs_SC's data and
// chain outputs), shared between sc_gibbs.cu and mle.cu.
#pragma once
#include <vector>
#include "common.cuh"

namespace ursm {

struct ScShard {
  long long L = 0;
  int N = 0, K = 0;
  long long cell_offset = 0;
  DevBuf<float> Yown;            // owned copy of the counts (fp32), or empty
  const float* Y = nullptr;      // [L][N] counts (owned or borrowed)
  DevBuf<int> G, order;          // types; local cell ids sorted by type
  DevBuf<double> R;              // read depth per cell
  // [L][N] number of retained sweeps with S = 1 (E[S] = cnt / sample, exact).  One byte
  // per entry when sample <= 255, two otherwise: the M-step streams this matrix twice per
  // outer iteration, so its width is the M-step's HBM traffic.
  DevBuf<unsigned char> cnt;
  int cnt_w = 0;                 // bytes per counter (0: not allocated yet)
  void cnt_ensure(int width) {
    if (cnt_w == width) return;
    cnt.alloc((size_t)L * N * width);
    cnt.zero();
    cudaStreamSynchronize(0);  // allocation and zeroing are ordered on the default stream only
    cnt_w = width;
  }
  DevBuf<double> cellmom;        // [4][L] E kappa, E tau, E kappa^2, E tau^2
  DevBuf<int> counter;           // work counter of the persistent kernel
  DevBuf<double> A;              // [N][K] fp64 parameters (row-major, as given)
  DevBuf<float> Aslot;           // [K][slots] fp32 copy of A^T in the Gibbs kernel's slot order
  DevBuf<float> scratch;         // [grid][slots][2] CTA-private accumulators (L2-resident)
  int T = 0, g = 0, slots = 0, smem = 0, occ = 0, grid = 0, num_sms = 0;
  DevBuf<unsigned char> state;   // BIG variant: [grid][slots] x 10 B chain state (w, threshold, counter)
  bool acc_smem = false, has_params = false, big = false;
  double pkappa[2] = {0, 1}, ptau[2] = {0, 1};
  int sample = 1;                // divisor that turns cnt into E[S]
  std::vector<long long> type_count, type_start;  // cells per type; offsets into `order`
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;       // bracket the last Gibbs kernel launch
  bool timed = false;
  ~ScShard() {
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
  }
};

struct BkShard {
  long long M = 0;
  int N = 0, K = 0;
  long long sample_offset = 0;
  DevBuf<float> Xown;
  const float* X = nullptr;      // [M][N] counts
  DevBuf<float> Xt;              // [N][Mpad] gene-major copy (Gibbs split kernel: lanes = samples)
  int Mpad = 0;                  // M rounded up to a multiple of 32
  DevBuf<float> XtN;             // [N][MpadN] gene-major copy in the natural gene order (mean-mode row pass)
  int MpadN = 0;
  DevBuf<float> A32;             // [K][N] fp32 A^T
  DevBuf<float> W32g;            // [ceil(M/32)][K][32] fp32 W of the current sweep
  DevBuf<double> A;              // [K][N] type-major fp64
  DevBuf<double> alpha;          // [K]
  DevBuf<double> W;              // [M][K] (sample-major)
  DevBuf<double> nA, nB;         // [M][K] sum_i Z_jik: previous sweep / being built
  DevBuf<double> lf2;            // [1024] log2 k! (modal binomial inversion)
  DevBuf<double> acc;            // [M][K] scratch (nmf numerators)
  DevBuf<double> eZjk, eLogW, eW;  // [M][K] each
  DevBuf<int> order;             // [N] lockstep split kernel: gene id of every position (genes by
                                 // decreasing total count); Xt / A32 / kmax are stored by position
  DevBuf<unsigned char> kmax;    // [N] type with the largest A_ik (drawn last), by position
  DevBuf<uint4> cold_list;       // chains the lockstep kernel hands to bk_split_cold_kernel
  DevBuf<unsigned int> cold_count;
  unsigned int cold_cap = 0;
  bool lane_kernel = false;      // deep counts: the per-lane split kernel does every chain
  DevBuf<int> fail;              // [1] sampler bail-outs of the last Gibbs E-step (0 in a correct run)
  bool has_params = false, has_state = false;
  int num_sms = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;       // bracket the last E-step's kernels
  bool timed = false;
  ~BkShard() {
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
  }
};

}  // namespace ursm
