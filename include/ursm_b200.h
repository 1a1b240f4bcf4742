/* ursm_b200.h -- C ABI of the B200-native URSM Monte-Carlo EM hot path.
 *
 * The reference (lingxuez/URSM) is pure Python: there is no FFI layer to mirror.
 * The seam this library sits behind is the class contract between gem.py and
 * e_step_gibbs.py / m_step.py.  Each entry point below names the
 * reference method(s) it replaces (file:line under /root/reference); the Python
 * shims in ursm_b200/{e_step_gibbs,m_step}.py bind them with ctypes and keep the
 * reference's class / method / attribute names, so gem.py and scUnif.py run
 * unchanged on top (INTEGRATION.md shows the three-line binding).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types in any signature;
 *   - `h_` pointers are HOST memory, `d_` pointers are DEVICE memory owned by the
 *     caller (the Python host allocates them as torch tensors so that NCCL can
 *     all-reduce them in place), `stream` is a cudaStream_t passed as void*;
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     available from ursm_last_error() (thread-local);
 *   - matrices are row-major; expression matrices are cells|samples x genes as in
 *     the reference's CSVs; A is genes x K;
 *   - cells / bulk samples may be a SHARD of the global problem: `*_offset` is
 *     the global index of the first local row and keys the counter-based RNG, so
 *     results do not depend on the number of shards.
 *   - there is NO CPU fallback: every compute entry point launches sm_100a kernels.
 */
#ifndef URSM_B200_H
#define URSM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ursm_sc ursm_sc; /* single-cell shard: LogitNormalGibbs_SC state  */
typedef struct ursm_bk ursm_bk; /* bulk shard:        LogitNormalGibbs_BK state  */
typedef struct ursm_mle ursm_mle; /* M-step workspace:  LogitNormalMLE state       */

/* ---- library ---------------------------------------------------------- */
const char* ursm_last_error(void);
int ursm_version(void);
/* number of kernel launches issued by this library in this process (bench.py's
 * `gpu_launches`), and reset */
int64_t ursm_launch_count(void);
void ursm_launch_count_reset(void);
/* page-locked host blocks for results read back whole (scUnif.py:71 writes exp_S, an
 * L x N float64 matrix, straight after the fit): cached by size inside the library so a
 * read-back is one DMA into memory the host never has to fault in.  ursm_host_trim
 * releases the cache. */
int ursm_host_alloc(void** out, int64_t bytes);
int ursm_host_free(void* p);
int ursm_host_trim(void);

/* ---- single cells: e_step_gibbs.py:171-377 ---------------------------------
 * ursm_sc_create        __init__ :174-197  (caches Y, G, read depths; uploads once)
 *   h_Y: L x N float64 counts (host), h_G: L int64 types in [0,K)
 * ursm_sc_create_dev    same, from an fp32 L x N matrix already on the device
 *                        (used by bench.py so c3-c5 inputs never exist on the host)
 * ursm_sc_set_params    update_parameters :273-280  (A: N x K, pkappa/ptau: [mean,var])
 * ursm_sc_gibbs         gibbs :286-304 = init_gibbs :212-225 + (burnin+sample*thin) x
 *                        gibbs_cycle :307-312 {draw_w :318, draw_S :326, draw_kappa_tau
 *                        :345, update_psi :314} + update_suffStats :245-270 on retained
 *                        sweeps.  ONE fused persistent kernel; chain state stays on chip.
 *   d_stats (device, caller-owned, length ursm_sc_stats_len()) receives, already
 *   divided by `sample`:  coeffA[N*K] | coeffAsq[N*K] | exp_elbo_const |
 *   sum_l E[kappa] | sum_l E[kappa^2] | sum_l E[tau] | sum_l E[tau^2] | failure count
 *   (failure count: (cell, sweep) pairs whose draw_S scan :326-342 did not reach its fixed
 *   point -- 0 in a correct run; the Python mirror raises otherwise)
 *   -- the vector a multi-GPU caller all-reduces.
 * ursm_sc_get_cell_moments  suff_stats exp_kappa/exp_tau/exp_kappasq/exp_tausq (:233-236)
 * ursm_sc_get_exp_S         suff_stats["exp_S"] (:232) rows [row0,row0+nrows) as float64 (on the
 *                            device it is a count per entry, one byte when sample <= 255, two
 *                            otherwise: exp_S * sample is an integer)
 * ursm_sc_inject            parity hook: one update_suffStats(sample) (:245-270) on a
 *                            caller-supplied state (w, S, kappa, tau) instead of a drawn one
 * ursm_sc_sweep_debug       parity hook: ONE gibbs_cycle of the production kernel from
 *                            a caller-supplied (S, kappa, tau); returns the w and S it
 *                            drew and the new (kappa, tau)
 */
int ursm_sc_create(ursm_sc** out, const double* h_Y, const int64_t* h_G, int64_t L, int32_t N,
                   int32_t K, int64_t cell_offset);
int ursm_sc_create_dev(ursm_sc** out, const float* d_Y, const int64_t* h_G, int64_t L, int32_t N,
                       int32_t K, int64_t cell_offset);
int ursm_sc_destroy(ursm_sc* sc);
int64_t ursm_sc_stats_len(const ursm_sc* sc);
int ursm_sc_set_params(ursm_sc* sc, const double* h_A, const double* h_pkappa,
                       const double* h_ptau);
int ursm_sc_gibbs(ursm_sc* sc, int32_t burnin, int32_t sample, int32_t thin, uint64_t seed,
                  uint64_t em_iter, double* d_stats, void* stream);
int ursm_sc_get_cell_moments(ursm_sc* sc, double* h_exp_kappa, double* h_exp_tau,
                             double* h_exp_kappasq, double* h_exp_tausq);
int ursm_sc_get_exp_S(ursm_sc* sc, double* h_out, int64_t row0, int64_t nrows);
int ursm_sc_inject(ursm_sc* sc, const double* h_w, const int64_t* h_S, const double* h_kappa,
                   const double* h_tau, int32_t sample, int32_t reset, double* d_stats,
                   void* stream);
int ursm_sc_sweep_debug(ursm_sc* sc, const int64_t* h_S_in, const double* h_kappa_in,
                        const double* h_tau_in, uint64_t seed, uint64_t em_iter, int32_t sweep,
                        float* h_w_out, uint8_t* h_S_out, double* h_kappa_out,
                        double* h_tau_out, int32_t* h_scan_rounds);
/* statistics that were NOT produced by this library (a suff_stats dict of the reference's own
 * sampler, e_step_gibbs.py:228-242, handed to the M-step): installs E[S] as counts per entry
 * (h_cnt[l][i] = exp_S * sample, an integer by construction :247) and the per-cell moments
 * [4][L] (E kappa, E tau, E kappa^2, E tau^2), so that update_suff_stats m_step.py:50-89 and the
 * opt_u / update_gd_coeffConst passes :108-125 can run on them */
int ursm_sc_set_counts(ursm_sc* sc, const uint16_t* h_cnt, int32_t sample, const double* h_cellmom);
/* parity hook for update_suffStats :245-270 as the PRODUCTION kernel folds it: burnin +
 * sample*thin gibbs_cycles (:307-312) from a caller-supplied (S, kappa, tau) with the retention
 * schedule of gibbs :299-304; returns the state after EVERY sweep -- h_w_out / h_S_out
 * [nsweeps][L][N], h_kt_out [nsweeps][2][L] (kappa row, tau row) -- and what the kernel
 * accumulated from it: h_stats (ursm_sc_stats_len doubles, layout of ursm_sc_gibbs),
 * h_cnt [L][N] (retained sweeps with S = 1), h_cellmom [4][L] (E kappa, E tau, E kappa^2, E tau^2) */
int ursm_sc_chain_debug(ursm_sc* sc, const int64_t* h_S_in, const double* h_kappa_in,
                        const double* h_tau_in, uint64_t seed, uint64_t em_iter, int32_t burnin,
                        int32_t sample, int32_t thin, float* h_w_out, uint8_t* h_S_out,
                        double* h_kt_out, double* h_stats, uint16_t* h_cnt, double* h_cellmom);
/* device time (ms, CUDA events on the launch stream) of the last Gibbs kernel */
int ursm_sc_kernel_ms(ursm_sc* sc, float* ms);
/* launch geometry the shard picked (threads per CTA, genes per thread, CTAs) */
int ursm_sc_geometry(const ursm_sc* sc, int32_t* threads, int32_t* genes_per_thread,
                     int32_t* grid, int32_t* smem_bytes, int32_t* acc_in_smem);

/* ---- bulk samples: e_step_gibbs.py:19-164 -----------------------------------
 * ursm_bk_create       __init__ :22-36
 * ursm_bk_set_params   update_parameters :81-84
 * ursm_bk_set_state    init_gibbs :39-55 result: W (K x M) and, for the Gibbs mode,
 *                       the initial sum_i Z_jik (M x K; zeros without markers)
 * ursm_bk_mean_step    gibbs(mean_approx=True) :95-103 = get_nmf_W :157-164 +
 *                       draw_Z_mean :149-155 + update_suffStats(1) :69-78
 * ursm_bk_gibbs        gibbs(mean_approx=False) :105-117: per sweep draw_W :140-146 then
 *                       draw_Z :132-138, update_suffStats on retained sweeps
 *   d_stats (length ursm_bk_stats_len()):  exp_Zik[N*K] | sum_j exp_logW[K] |
 *                                          sum_jk exp_Zjk*exp_logW | sampler failure count (0)
 * ursm_bk_get_sample_stats  exp_Zjk (M x K), exp_logW (K x M), exp_W (K x M)
 * freeze flags (parity hooks for the moment tests): bit0 = keep W fixed, bit1 = keep Z sums fixed
 */
int ursm_bk_create(ursm_bk** out, const double* h_X, int64_t M, int32_t N, int32_t K,
                   int64_t sample_offset);
int ursm_bk_create_dev(ursm_bk** out, const float* d_X, int64_t M, int32_t N, int32_t K,
                       int64_t sample_offset);
int ursm_bk_destroy(ursm_bk* bk);
int64_t ursm_bk_stats_len(const ursm_bk* bk);
int ursm_bk_set_params(ursm_bk* bk, const double* h_A, const double* h_alpha);
int ursm_bk_set_state(ursm_bk* bk, const double* h_W, const double* h_Zjk);
int ursm_bk_mean_step(ursm_bk* bk, double* d_stats, void* stream);
int ursm_bk_gibbs(ursm_bk* bk, int32_t burnin, int32_t sample, int32_t thin, uint64_t seed,
                  uint64_t em_iter, int32_t freeze, double* d_stats, void* stream);
int ursm_bk_get_sample_stats(ursm_bk* bk, double* h_exp_Zjk, double* h_exp_logW,
                             double* h_exp_W);
int ursm_bk_get_W(ursm_bk* bk, double* h_W);
/* device time (ms) of the last bulk E-step's kernels (mean step or all Gibbs sweeps) */
int ursm_bk_kernel_ms(ursm_bk* bk, float* ms);

/* ---- M-step: m_step.py:13-320 ------------------------------------------------
 * The workspace borrows the shards (it reads Y, the E[S] counters, read depths).
 * Either shard may be NULL (bulk-only / single-cell-only model).
 * ursm_mle_begin        update_suff_stats :50-89, local part: sum_{l in k} Y_li E[S_li]
 *                        (-> d_part[0 : K*N], type-major) -- caller all-reduces d_part
 * ursm_mle_pass_u       opt_u :108-115 + update_gd_coeffConst :118-125, local part for the
 *                        columns flagged active: d_part[k*N+i] = sum_{l in k} E[S_li] R_l/u_l,
 *                        d_part[K*N] = sum_l R_l log u_l   (length K*N+1)
 * ursm_mle_set_coeffs   installs the (all-reduced) coefficient matrices, all K x N type-major
 * ursm_mle_opt_begin/step/pass/merge/poll   opt_A_u :128-154 for every column that has cells,
 *                        all columns at once and without a host round trip per iteration:
 *     begin  arms the per-column loop state of opt_Ak :157-182 (active flag, iteration count)
 *     step   backtracking :209-233 -- up to 16 trial step sizes init_step/2^c of every active
 *            column are projected (simplex_proj, utils.py:8-31) and tested CONCURRENTLY, then
 *            the first passing one is committed: the step the reference's sequential halving
 *            loop accepts; get_grad_A :262-272, get_obj_A :275-287.  d_cconst_part is the
 *            all-reduced output of the previous pass.
 *     pass   opt_u :108-115 + update_gd_coeffConst :118-125 for the stepped columns -> d_part_local
 *            (the caller all-reduces it across ranks)
 *     merge  installs the reduced partials of the stepped columns, advances the loop state
 *            (pass_p2p / merge_p2p below: the same with the cross-rank sum inside the kernel)
 *     poll   device -> host: which columns are still iterating, iteration / halving counts
 * ursm_mle_closed_form  the bulk-only column rule :149-153
 * ursm_mle_elbo_terms   the A-dependent sums of compute_elbo :290-320
 */
/* ---- parameter initialisation: gem.py:225-255 with utils.std_row utils.py:34-38 -------------
 * ursm_sc_type_sums    d_out[k][i] = sum over this shard's cells of type k of Y_li / sum_i Y_li
 *                      (the caller all-reduces and divides by the global cell count of the type)
 * ursm_bk_profile_sum  d_out[i]    = sum over this shard's bulk samples of X_ji / sum_i X_ji
 * One streaming pass over the device-resident counts each (4 B per entry). */
int ursm_sc_type_sums(ursm_sc* sc, double* d_out, void* stream);
int ursm_bk_profile_sum(ursm_bk* bk, double* d_out, void* stream);

int ursm_mle_create(ursm_mle** out, ursm_sc* sc, ursm_bk* bk, int32_t N, int32_t K, double min_A);
int ursm_mle_destroy(ursm_mle* m);
int ursm_mle_set_A(ursm_mle* m, const double* h_A);
int ursm_mle_get_A(ursm_mle* m, double* h_A);
int ursm_mle_begin(ursm_mle* m, double* d_part, void* stream);
int ursm_mle_pass_u(ursm_mle* m, const int32_t* h_active, double* d_part, void* stream);
int ursm_mle_set_coeffs(ursm_mle* m, const double* d_cAinv, const double* d_cA,
                        const double* d_coeffA);
int ursm_mle_opt_begin(ursm_mle* m, const int32_t* h_active, double init_step, double conv,
                       int32_t maxiter, void* stream);
int ursm_mle_opt_step(ursm_mle* m, const double* d_cconst_part, void* stream);
int ursm_mle_opt_pass(ursm_mle* m, double* d_part_local, void* stream);
int ursm_mle_opt_merge(ursm_mle* m, const double* d_part_local, double* d_part, void* stream);
/* The per-iteration cross-rank sum of `pass` + `merge` (N*K + K doubles per outer
 * iteration of opt_Ak) without a collective call: each rank's partials live in an IPC-shared
 * block (ursm_mle_p2p_alloc returns its 64-byte cudaIpcMemHandle_t; the caller exchanges the
 * handles, ursm_mle_p2p_open maps the peers'), `pass_p2p` writes them there and `merge_p2p`
 * is ONE kernel that signals the peers, waits for them and installs the rank-ordered sum read
 * over NVLink.  Blocks and their peer mappings are cached per process by size (`out_cached`:
 * this rank already holds the mappings; when every rank says so `open` needs no handles).
 * ursm_mle_p2p_error reports a peer that never showed up (the kernel gives up after ~20 s
 * instead of hanging).
 * ursm_mle_p2p_state -> {usable cached mapping?, block id, epoch, peers mapped?}: every rank
 * of a job must report the same first three before a cached block is reused without a handshake;
 * ursm_mle_p2p_reset zeroes this rank's flag words, the error word and the epoch (the caller
 * puts a barrier on both sides). */
int ursm_mle_p2p_alloc(ursm_mle* m, void* out_handle64, int32_t* out_cached);
int ursm_mle_p2p_state(ursm_mle* m, int64_t* out4);
int ursm_mle_p2p_reset(ursm_mle* m);
int ursm_mle_p2p_disable(ursm_mle* m);
int ursm_mle_p2p_open(ursm_mle* m, const void* handles, int32_t world, int32_t rank);
int ursm_mle_opt_pass_p2p(ursm_mle* m, void* stream);
int ursm_mle_opt_merge_p2p(ursm_mle* m, double* d_part, void* stream);
int ursm_mle_p2p_error(ursm_mle* m, int32_t* h_err);
int ursm_mle_opt_poll(ursm_mle* m, int32_t* h_active, int32_t* h_niter, int32_t* h_halvings,
                      double* h_delta, void* stream);
/* opt_A_u :128-154 for the columns flagged in h_active, run to completion in one call: the
 * loop of opt_Ak :157-182 lives on the device (16 outer iterations per launch of a captured
 * CUDA graph on the workspace's own stream: step + decision + installation in one kernel, then
 * the opt_u / update_gd_coeffConst pass; with the peer exchange opened, the cross-rank sum of
 * every iteration inside the same graph).  d_part_init = the all-reduced partials of the pass
 * before the loop (ursm_mle_pass_u); the caller runs that full pass again afterwards (it also
 * refreshes sum_l R_l log u_l).  h_niter / h_halvings: per column, as ursm_mle_opt_poll. */
int ursm_mle_opt_run(ursm_mle* m, const int32_t* h_active, double init_step, double conv,
                     int32_t maxiter, const double* d_part_init, int32_t* h_niter,
                     int32_t* h_halvings, int32_t* h_graph_launches, void* stream);
int ursm_mle_closed_form(ursm_mle* m, int32_t k, const double* d_exp_Zik, void* stream);
int ursm_mle_elbo_terms(ursm_mle* m, const double* d_cconst_part, double* h_terms3,
                        void* stream);
int ursm_mle_get_u(ursm_mle* m, double* h_u);
/* projection onto {y >= min_y, sum y = 1} of a host vector, on the device
 * (utils.py:8-31) -- parity hook for the M-step's inner kernel */
int ursm_simplex_proj(const double* h_y, int32_t n, double min_y, double* h_out);
/* the same projection for `rows` vectors of length n stored one after the other (one launch;
 * gem.py:252-254 projects the K columns of the initial A) */
int ursm_simplex_proj_rows(const double* h_y, int32_t rows, int32_t n, double min_y, double* h_out);

/* ---- file contract of scUnif.py (:85-94), SURVEY 8f row 1 ---------------------------
 * Headerless comma-separated matrices, byte-compatible with np.savetxt(delimiter=",")
 * ('%.18e' per value) and np.loadtxt(delimiter=","); multi-threaded host code.
 * ursm_csv_write_f64   mtx2csv :85-87
 * ursm_csv_shape + ursm_csv_read_f64   load_from_file :90-94 (float matrices)
 */
int ursm_csv_write_f64(const char* path, const double* data, int64_t rows, int64_t cols,
                       int32_t nthreads);
int ursm_csv_shape(const char* path, int64_t* rows, int64_t* cols);
int ursm_csv_read_f64(const char* path, double* out, int64_t rows, int64_t cols, int32_t nthreads);

/* ---- synthetic inputs on the device --------------------------------------
 * The distributions of the reference's demo generator, demo/demo_simulate_data.py:
 * ursm_synth_bulk   simulate_bulk :45-69 with depths ~ Poisson(depth_per_gene * N) (:150-152):
 *                   W_j ~ Dir(alpha), X_j ~ Mult(depth_j, A W_j); d_X [M][N] fp32 on the device,
 *                   h_W [M][K] (may be null)
 * ursm_synth_cells  simulate_sc :72-114 with depths ~ Poisson(depth_per_gene * N): kappa_l ~
 *                   N(kappa, kappa_sd), tau_l ~ N(tau_scale N, tau_sd_frac tau_scale N),
 *                   S_li ~ Bern(expit(kappa_l + tau_l A[i, G_l])), Y_l ~ Mult(depth_l,
 *                   normalise(A[:, G_l] S_l)); d_Y [L][N] fp32, d_S [L][N] bytes on the device
 * A multinomial with a Poisson total is drawn as independent Poissons (the same joint law).
 * Streams are keyed on (seed, gene, row_offset + row): a shard of rows does not depend on who
 * draws it.  h_A is N x K row-major. */
int ursm_synth_bulk(const double* h_A, int32_t N, int32_t K, int64_t M, const double* h_alpha,
                    double depth_per_gene, uint64_t seed, int64_t row_offset, float* d_X, double* h_W);
int ursm_synth_cells(const double* h_A, int32_t N, int32_t K, const int64_t* h_G, int64_t L,
                     uint64_t seed, int64_t row_offset, double depth_per_gene, double kappa,
                     double kappa_sd, double tau_scale, double tau_sd_frac, float* d_Y, uint8_t* d_S,
                     double* h_kappa, double* h_tau);

#ifdef __cplusplus
}
#endif
#endif /* URSM_B200_H */
