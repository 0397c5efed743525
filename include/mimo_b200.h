/* mimo_b200 -- C-ABI of the B200-native mixture-inference sweep.
 *
 * The reference (hanyas/mimo) is pure Python/NumPy and has no FFI: the
 * "interface" this library replaces is the set of NumPy call sites on the
 * sweep.  Each entry point below names the reference
 * function(s) it stands in for (paths relative to /root/reference/mimo).  The
 * reference-side binding a maintainer would add is the ctypes stub shown in
 * INTEGRATION.md; mimo_b200/_lib.py is that stub as shipped here.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns all memory, including workspaces (sizes are queried);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*)
 *     unless stated otherwise, re-entrant, and returns an int status;
 *   - dtype selects the arithmetic type of the per-point path:
 *       MIMO_F32: FP32 compute, FP64 accumulation of statistics / lower bound
 *       MIMO_F64: FP64 everywhere (1e-9 parity mode)
 *     posterior (per-component) math is always FP64;
 *   - layouts follow the reference: data (N, D) row-major; per-point x
 *     per-component arrays (K, N) component-major (mixtures/gmm.py:67-75);
 *     labels (N,) int32 (utils/stats.py:8).
 *
 * Packed operand form.  Every full-covariance-family
 * E-step (Gaussian, linear-Gaussian, their NW / MNW expectations, and the ILR
 * sum basis + models) is
 *       a[k][n] = cst[k] - 0.5 * sum_{i<Rp} ( sum_{j<D} W[k][i][j] z[n][j] + W[k][i][D] )^2
 *   W  : (K, Rp, Dpp)   Rp = rows padded to a power of two in [8,128] (zero rows),
 *                       Dpp = D+1 padded to a multiple of 4 (zero columns),
 *                       column D multiplies the implicit constant 1.
 * Diagonal family:
 *       a[k][n] = cst[k] - 0.5 * sum_j ( S[k][j] z[n][j] - T[k][j] )^2
 *
 * Packed statistics.  Features of zt = [z ; 1] (length Dp = D+1):
 *       stat[k][f] = sum_n r[k][n] * zt[n][fi[f]] * zt[n][fj[f]]        (FP64)
 *   full family : all pairs j <= i, f = i(i+1)/2 + j    (F = Dp(Dp+1)/2)
 *   diag family : (j,D) j<D ; (j,j) j<D ; (D,D)         (F = 2D+1)
 * This K x F buffer is exactly what the data-sharded driver all-reduces.
 * Any (fi, fj) table is accepted.  The list / tensor-core statistics kernels of mimo_sweep produce the "full family"
 * order above; the sweep reads the table back once per call (F * 8 bytes) and takes the generic kernel for any other
 * table, so the result never depends on which kernel ran.
 */
#ifndef MIMO_B200_H
#define MIMO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIMO_OK            0
#define MIMO_EINVAL        1   /* bad argument (shape, alignment, null pointer) */
#define MIMO_ECUDA         2   /* CUDA runtime error; see mimo_last_error_string */
#define MIMO_ENOTPD        3   /* a matrix was not positive definite (np.linalg.LinAlgError) */
#define MIMO_EUNSUPPORTED  4   /* shape outside what this build supports */

#define MIMO_F32 0
#define MIMO_F64 1

/* flags of mimo_softmax / mimo_sweep_* */
#define MIMO_WRITE_RESP    1   /* overwrite the log-joint with responsibilities */
#define MIMO_WRITE_LSE     2   /* write the per-point log-normaliser */
#define MIMO_DRAW_LABELS   4   /* inverse-CDF categorical draw (utils/stats.py:8-21) */
#define MIMO_ACC_LSE       8   /* add sum_n lse[n] into *lse_sum (FP64) */

const char* mimo_last_error_string(void);
int mimo_version(void);
/* 1 when the tcgen05 (tensor-core) kernels are compiled in and the shape is
 * handled by them; informational. */
int mimo_device_ok(void);

/* ---- per-point log-likelihoods ---------------------------------------- */

/* replaces: StackedGaussiansWithPrecision.log_likelihood  distributions/gaussian.py:510-523
 *           StackedLinearGaussiansWithPrecision.log_likelihood  lingauss.py:330-347
 *           Stacked*With{NormalWisharts,MatrixNormalWisharts}.expected_log_likelihood
 *           bayesian.py:287-301, 933-947 (same kernel, different W / cst)      */
int mimo_loglik_quad(int dtype, const void* Z, int64_t N, int D, int64_t ldz,
                     const void* W, const void* cst, int K, int Rp, int Dpp,
                     void* out, int64_t ldo, void* stream);

/* replaces: StackedGaussiansWithDiagonalPrecision.log_likelihood gaussian.py:837-850
 *           StackedGaussiansWithNormalGammas.expected_log_likelihood bayesian.py:446-460 */
int mimo_loglik_diag(int dtype, const void* Z, int64_t N, int D, int64_t ldz,
                     const void* S, const void* T, const void* cst, int K,
                     void* out, int64_t ldo, void* stream);

/* replaces: logsumexp + exp  mixtures/gmm.py:72-75, 256-259
 *           sample_discrete_from_log  utils/stats.py:8-21
 * a: (K, ldo) log-joint for n points.  uniforms: FP64 (n,) regardless of dtype
 * (what npr.random returns at stats.py:14), or NULL: then labels use
 * Philox4x32-10 keyed by (seed, point_offset + n).                           */
int mimo_softmax(int dtype, void* a, int K, int64_t n, int64_t ldo, int flags,
                 void* lse, const void* uniforms, uint64_t seed, uint64_t point_offset,
                 int32_t* labels, double* lse_sum, void* stream);

/* ---- tensor-core (tcgen05 / TMEM) variants: FP32 data, quad family, D <= 128 ----
 * Operands are split into two FP16 halves after a power-of-two pre-scale and the product is
 * accumulated in FP32 as Ah*Bh + Ah*Bl + Al*Bh (22 significand bits, FP32-class results; see
 * DESIGN.md section 5).  mimo_sweep picks these kernels by itself when
 * mimo_sweep_uses_tensor_cores() says so; the stand-alone entry points exist for parity tests.
 * Same reference call sites as mimo_loglik_quad / mimo_stats_soft.                          */
/* mode 0: CUDA cores only; 1 (default): tcgen05 kernels on CTA pairs, E-step screened (one FP16 pass over all
 * pairs + exact recomputation of the pairs within 40 nats of a point's best component, or -- device-side choice
 * when more than 4 % of the pairs qualify -- the dense 3-pass kernel) and, for mean-field sweeps, the statistics
 * summed over those candidate pairs only (pair-list kernel; the dense tensor-core statistics on fallback chunks);
 * 2: single-CTA dense kernels; 3: CTA pairs, dense; 4: as 1 with dense statistics; 5: as 1 with the screening
 * starting on its second tier (all operand rows in one FP16 pass instead of the 32-row projection -- the tier a sweep
 * moves to by itself when the projection leaves too many candidates).  Returns the old mode. */
int mimo_set_tensor_cores(int mode);
/* {candidate pairs, dense-fallback flag} of the most recent screened point chunk (synchronises; diagnostics) */
int mimo_tc_screen_last(uint32_t* out_host2);
/* totals of the most recent screened sweep over ALL its point chunks: {candidate pairs of the refined chunks,
 * points of the refined chunks, chunks that took the dense pass, chunks, tier the sweep ended on} (synchronises) */
int mimo_tc_screen_totals(uint64_t* out_host5);
/* screening tier the most recent screened sweep ended on: 0 projection, 1 all operand rows, 2 none (dense); -1 unknown */
int mimo_tc_screen_level(void);
int mimo_sweep_uses_tensor_cores(int dtype, int family, int D, int Rp);
/* ---- data-sharded sweeps: the one exchange step -------------------------------------------------------
 * Sum over shards of the packed FP64 statistics (and the lower-bound scalar packed behind them): the reference's
 * list-of-arrays semantics, distributions/gaussian.py:503-505, utils/abstraction.py:12-14.  NCCL is loaded at run time
 * (dlopen "libnccl.so.2"); MIMO_EUNSUPPORTED when it is not installed.  Rank 0 calls mimo_comm_unique_id and hands the
 * mimo_comm_unique_id_bytes() = 128 bytes to every rank out of band; every rank (one process per GPU, its device
 * current) then calls mimo_comm_init; mimo_comm_allreduce_stats sums `count` doubles in place on `stream`. */
size_t mimo_comm_unique_id_bytes(void);
int mimo_comm_unique_id(void* out);
int mimo_comm_init(int world, int rank, const void* unique_id, void** comm_out);
int mimo_comm_allreduce_stats(void* comm, double* stat, int64_t count, void* stream);
int mimo_comm_destroy(void* comm);

/* ---- prediction path (mixtures/ilr.py:325-430) ----------------------------------------------------------------
 * mimo_studentt_from_quad: a[k][n] <- add[k] + log1p(2 (c0[k] - a[k][n]) / df[k]), the reference's Student-t form
 * (utils/stats.py:53-79) of a Gaussian-form log-joint a = c0 - delta/2 produced by mimo_loglik_quad.
 * mimo_predict_lingauss: per point the posterior-predictive moments of the K linear-Gaussian experts
 * (distributions/bayesian.py:876-912, 949-985: mu = M x~, c = 1 + x~^T K^-1 x~, Lambda = Psi df / c), combined with the
 * predictive weights W (K, ldw) into the mixture mean / covariance (mode 0, ilr.py:375-383) or taken from the expert
 * with the largest weight (mode 1); studentt scales the covariances by df / (df - 2); with Y the negative log
 * predictive density (ilr.py:407-411) is written to nlpd_out.  M (K, o, c), Kinv (K, c, c), Sigma = Psi^-1 and Psi
 * (K, o, o), logdet_psi and df (K) are FP64 device arrays (one entry instead of K when tied); c = din + affine;
 * o <= 4.  Outputs mu (N, o), cov (N, o, o), nlpd (N) in dtype. */
int mimo_studentt_from_quad(int dtype, void* a, int K, int64_t N, int64_t lda, const double* c0, const double* add, const double* df,
                            void* stream);
int mimo_predict_lingauss(int dtype, const void* X, int64_t N, int64_t ldx, int din, int affine, const void* W, int64_t ldw, int K,
                          const double* M, const double* Kinv, const double* Sigma, const double* Psi, const double* logdet_psi,
                          const double* df, int o, int tied, int mode, int studentt, const void* Y, int64_t ldy, double eps,
                          void* mu_out, void* cov_out, void* nlpd_out, void* stream);

/* Diagonal family on the tensor pipe (FP32, D <= 64, K <= 256): the log-density is linear in [z', z'^2] of the centred
 * point, i.e. one tcgen05 GEMM; labels (inverse CDF of `uniforms`, or Philox(seed, point_offset + n) when NULL) and the
 * log-normalisers come out of its epilogue.  Replaces distributions/gaussian.py:837-850, bayesian.py:446-460 and
 * mixtures/gmm.py:72-75 on that shape; mimo_sweep uses it by itself.  out / labels / lse / lse_sum are optional.
 * *guard_host (optional, synchronises) = 1 when the operands failed the cancellation guard and NOTHING was computed
 * (mimo_sweep then runs the CUDA-core kernels on the device's own decision). */
size_t mimo_loglik_diag_tc_workspace(void);
int mimo_loglik_diag_tc(const void* Z, int64_t N, int D, int64_t ldz, const void* S, const void* T, const void* cst, int K,
                        void* out, int64_t ldo, int32_t* labels, const void* uniforms, uint64_t seed, uint64_t point_offset,
                        void* lse, double* lse_sum, uint32_t* guard_host, void* workspace, size_t workspace_bytes, void* stream);
int mimo_tc_diag_enable(int on);            /* A/B: 0 keeps diagonal sweeps on the CUDA cores; returns the old setting */
/* One-shot hint for the NEXT mimo_sweep / mimo_sweep_timed of the calling thread: max |Z| over the data it will be given
 * (>= the true maximum).  The tensor-core paths then skip their own pass over Z for the common power-of-two data scale
 * -- for callers that keep the data resident and unchanged across sweeps (the Python Session does).  The hint is consumed
 * by that sweep whichever kernels it runs (a sweep that stays on the CUDA cores discards it) and is never seen by the
 * stand-alone entry points. */
int mimo_sweep_absmax_hint(double absmax);
/* One-shot promise for the NEXT mimo_sweep of the calling thread: the (fi, fj) tables are (1) / are not (0) the canonical
 * packed triangle f = i (i + 1) / 2 + j.  Without it a soft quad-family sweep reads the tables back once to decide whether
 * its list / tensor-core statistics kernels apply (a stream synchronisation, which a CUDA-graph capture cannot contain). */
int mimo_sweep_tables_hint(int canonical);
/* diagnostics: clocks the MMA issuers of tc_fstats_kernel waited, summed over clusters since the last call: {A tile, peer's A
 * tile, B stage, peer's B stage, accumulator drain, total issuer clocks, stages issued, 0} (synchronises, resets) */
int mimo_tc_fstats_stall_clocks(uint64_t* out_host8);
int mimo_tc_set_min_dim(int d);             /* A/B: smallest D a quad-family sweep takes to the tensor pipe (default 8; 24 = round-1 behaviour); returns the old value */
int mimo_tc_set_quad_generations(int on);   /* A/B: dense E-step, 64 < D <= 128: 1 = four components per accumulator generation with the zero block of the Cholesky factors skipped (tc_estep4.cu; measured slower), 0 (default) = the plain CTA-pair kernel (tc_estep2.cu); returns the old setting */
int mimo_tc_set_triangular(int rows);       /* dense E-step, 64 < D <= 128: rows per step of the triangular skip (16 default, 32; 0 = kernel with both operands in shared memory); returns the old value */
int mimo_tc_set_flush_tiles(int tiles);     /* 128-point tiles accumulated in TMEM (FP32) between FP64 drains */
size_t mimo_loglik_quad_tc_workspace(int K, int Rp, int D);
int mimo_loglik_quad_tc(const void* Z, int64_t N, int D, int64_t ldz,
                        const void* W, const void* cst, int K, int Rp, int Dpp,
                        void* out, int64_t ldo, void* workspace, size_t workspace_bytes, void* stream);
/* full-triangle packed statistics only (F = (D+1)(D+2)/2, the order documented above) */
size_t mimo_stats_soft_tc_workspace(int64_t N, int K);
int mimo_stats_soft_tc(const void* Z, int64_t N, int D, int64_t ldz,
                       const void* resp, int64_t ldr, int K, int F,
                       double* stat, void* workspace, size_t workspace_bytes, void* stream);

/* ---- weighted sufficient statistics ----------------------------------- */

/* replaces: *.weighted_statistics  gaussian.py:491-505, 819-832; lingauss.py:306-325;
 *           categorical.py:41-46.  stat (K,F) FP64 is ACCUMULATED into.      */
int mimo_stats_soft(int dtype, const void* Z, int64_t N, int D, int64_t ldz,
                    const void* resp, int64_t ldr, int K,
                    const int32_t* fi, const int32_t* fj, int F,
                    double* stat, void* stream);

/* replaces: one_hot + weighted_statistics  utils/data.py:160-169 + the above;
 *           categorical.py:35-39 (bincount = the (D,D) feature).
 * labels outside [0,K) give MIMO_EINVAL (the assert of data.py:162) after a
 * device-side check; this call synchronises `stream` once for that check.   */
size_t mimo_stats_hard_workspace(int64_t N, int K);
int mimo_stats_hard(int dtype, const void* Z, int64_t N, int D, int64_t ldz,
                    const int32_t* labels, int K,
                    const int32_t* fi, const int32_t* fj, int F,
                    double* stat, void* workspace, size_t workspace_bytes, void* stream);

/* ---- one sweep over resident data ------------------------------------- */

/* One E-step + statistics pass over N resident points, processed in point
 * chunks through an L2-sized (K, chunk) scratch so (K,N) is never built:
 *   family 0 (quad): op_a = W, op_b unused;  family 1 (diag): op_a = S, op_b = T
 *   hard = 0: mean-field   -- stat += sum r phi, *lse_sum += sum lse
 *                              (mixtures/gmm.py:275-279 fused; the data + label
 *                              lower-bound terms of :338-356 equal sum lse)
 *   hard = 1: Gibbs        -- labels drawn (uniforms or Philox), stat += phi[label]
 *                              (mixtures/gmm.py:220-223 fused)
 * labels_out / lse_out / ll_out may be NULL.  ll_out, if given, is (K, N) and
 * receives the log-joint (hard) or the responsibilities (soft).             */
size_t mimo_sweep_workspace(int dtype, int family, int hard, int64_t N, int D, int K, int Rp);
int mimo_sweep(int dtype, int family, int hard,
               const void* Z, int64_t N, int D, int64_t ldz,
               const void* op_a, const void* op_b, const void* cst, int K, int Rp, int Dpp,
               const int32_t* fi, const int32_t* fj, int F,
               const void* uniforms, uint64_t seed, uint64_t point_offset,
               double* stat, double* lse_sum, int32_t* labels_out, void* lse_out,
               void* ll_out, int64_t ldo,
               void* workspace, size_t workspace_bytes, void* stream);

/* mimo_sweep with per-phase device timing for benchmarks: CUDA events are recorded on
 * `stream` around every launch; on return (this variant synchronises) phase_ms_host[0..2]
 * have been incremented by the milliseconds spent in the E-step, softmax / label and
 * statistics kernels, phase_ms_host[3] by the number of kernel launches and
 * phase_ms_host[4] by the number of point chunks (= launches of each per-chunk kernel) and
 * phase_ms_host[5] by the device time of the dominant E-step kernel alone (the screening pass on
 * the screened path; equal to phase 0 otherwise).  phase_ms_host holds 6 doubles.               */
int mimo_sweep_timed(int dtype, int family, int hard,
                     const void* Z, int64_t N, int D, int64_t ldz,
                     const void* op_a, const void* op_b, const void* cst, int K, int Rp, int Dpp,
                     const int32_t* fi, const int32_t* fj, int F,
                     const void* uniforms, uint64_t seed, uint64_t point_offset,
                     double* stat, double* lse_sum, int32_t* labels_out, void* lse_out,
                     void* ll_out, int64_t ldo,
                     void* workspace, size_t workspace_bytes, void* stream, double* phase_ms_host);

/* ---- batched per-component posterior kernels (always FP64 math) -------- */

/* Operand placement: a posterior kernel writes its whitening rows into the
 * shared W (K, Rp, Dpp) at row_off, mapping its own variable j to column
 * col_map[j] (device int32 array; the constant-1 column is D), and ADDS its
 * per-component constant into cst.  op_dtype is the dtype of W / cst / S / T. */

/* Normal-Wishart.  replaces composite.py:50-72 (std<->nat), :106-118, :77-86;
 * wishart.py:59-92, 129-143; gaussian.py:295-313, 352-354; bayesian.py:209-243;
 * tied: composite.py:275-283.
 *  prior (std form): m0 (K,d) kappa0 (K) psi0 (K,d,d) nu0 (K)
 *  stat: packed (K, F) over zt with `d` variables mapped through stat_idx
 *        (device int32 (d+1): position of each variable and of the constant in zt)
 *  mode: 0 mean-field (expected operands), 1 Gibbs (draw from variates),
 *        2 MAP (posterior mode), 3 none (posterior parameters only)
 *  variates (mode 1): (K, d(d-1)/2 + d + d) doubles per component in the
 *        reference order: normals (tril row-major), chi-square draws, normals
 *  outputs (any may be NULL): post m (K,d) kappa (K) psi (K,d,d) nu (K);
 *        lik_mu (K,d) lik_lmbda (K,d,d) (sampled / mode parameters);
 *        vlb (K): entropy - cross-entropy (bayesian.py:240-243)
 *  info: device int32[2] = {status, first failing component}                 */
size_t mimo_nw_workspace(int K, int d);
int mimo_nw_posterior(int K, int d, int tied, int mode,
                      const double* m0, const double* kappa0, const double* psi0, const double* nu0,
                      const double* stat, int F, const int32_t* stat_idx, int Dp,
                      const double* variates,
                      double* post_m, double* post_kappa, double* post_psi, double* post_nu,
                      double* lik_mu, double* lik_lmbda, double* vlb,
                      int op_dtype, void* W, void* cst, int Rp, int Dpp, int row_off, const int32_t* col_map,
                      void* workspace, size_t workspace_bytes, int32_t* info, void* stream);

/* Normal-Gamma (diagonal family).  replaces composite.py:313-398, gamma.py:54-113,
 * bayesian.py:370-404, 446-460; tied: composite.py:539-547.
 *  stat: diag-family packed layout [sum r x (d) | sum r x^2 (d) | sum r]
 *  bug_compat != 0 reproduces the reference's StackedNormalGammas setters
 *  (composite.py:472-484): alphas/betas stay at the prior values.
 *  variates (mode 1): (K, 2d): gamma draws then normals.                     */
size_t mimo_ng_workspace(int K, int d);
int mimo_ng_posterior(int K, int d, int tied, int mode, int bug_compat,
                      const double* m0, const double* kappa0, const double* alpha0, const double* beta0,
                      const double* stat, int F,
                      const double* variates,
                      double* post_m, double* post_kappa, double* post_alpha, double* post_beta,
                      double* lik_mu, double* lik_lmbda_diag, double* vlb,
                      int op_dtype, void* S, void* T, void* cst,
                      void* workspace, size_t workspace_bytes, int32_t* info, void* stream);

/* Matrix-Normal-Wishart (linear-Gaussian experts).  replaces composite.py:577-663,
 * 800-808; matrix.py:98-125; bayesian.py:823-857, 933-947.
 *  c = column_dim (input_dim + 1 if affine), o = row_dim
 *  stat_idx: device int32 (c + o + 1): positions in zt of xt[0..c), of y[0..o) and
 *            of the constant
 *  mean-field rows written: o (residual) + c (input) ; Gibbs / MAP rows: o
 *  variates (mode 1): (K, o(o-1)/2 + o + o*c).                               */
size_t mimo_mnw_workspace(int K, int c, int o);
int mimo_mnw_posterior(int K, int c, int o, int tied, int mode,
                       const double* M0, const double* K0, const double* psi0, const double* nu0,
                       const double* stat, int F, const int32_t* stat_idx, int Dp,
                       const double* variates,
                       double* post_M, double* post_K, double* post_psi, double* post_nu,
                       double* lik_A, double* lik_lmbda, double* vlb,
                       int op_dtype, void* W, void* cst, int Rp, int Dpp, int row_off, const int32_t* col_map,
                       void* workspace, size_t workspace_bytes, int32_t* info, void* stream);

/* Gating.  kind 0 Dirichlet (dirichlet.py:8-97, bayesian.py:62-99),
 *          kind 1 truncated stick-breaking (dirichlet.py:100-214, bayesian.py:128-179).
 *  counts are read from stat[k*F + count_feature].
 *  prior_a = alphas | gammas, prior_b = NULL | deltas.
 *  mode 0: expected log-weights; 1: Gibbs from variates (K gamma draws | K-1 beta
 *  draws), clipped as bayesian.py:75; 2: log of the posterior mode; 4: log of the
 *  posterior mean (ilr.py:333).  cst[k] is SET to the log-weight.            */
size_t mimo_gating_workspace(int K);
int mimo_gating_posterior(int K, int kind, int mode,
                          const double* prior_a, const double* prior_b,
                          const double* stat, int F, int count_feature,
                          const double* variates,
                          double* post_a, double* post_b, double* probs, double* vlb,
                          int op_dtype, void* cst, void* workspace, size_t workspace_bytes,
                          int32_t* info, void* stream);

/* Likelihood parameters -> operands (Gibbs / EM E-step with explicit
 * parameters; gaussian.py:507-523, lingauss.py:327-347).
 *  quad: rows = U_k (upper Cholesky of lmbda_k), offset -U_k mu_k             */
size_t mimo_operands_workspace(int K, int d);
int mimo_operands_gauss(int K, int d, const double* mu, const double* lmbda,
                        int op_dtype, void* W, void* cst, int Rp, int Dpp, int row_off, const int32_t* col_map,
                        void* workspace, size_t workspace_bytes, int32_t* info, void* stream);
int mimo_operands_gauss_diag(int K, int d, const double* mu, const double* lmbda_diag,
                             int op_dtype, void* S, void* T, void* cst, void* stream);
int mimo_operands_lingauss(int K, int c, int o, const double* A, const double* lmbda,
                           int op_dtype, void* W, void* cst, int Rp, int Dpp, int row_off, const int32_t* col_map,
                           void* workspace, size_t workspace_bytes, int32_t* info, void* stream);

/* EM M-steps from packed statistics (gaussian.py:525-542, 852-862; lingauss.py:350-367;
 * tied variants gaussian.py:550-572, 870-887).                                */
size_t mimo_mstep_workspace(int K, int d);
int mimo_mstep_gauss(int K, int d, int tied, const double* stat, int F, const int32_t* stat_idx, int Dp,
                     double* mu, double* lmbda, void* workspace, size_t workspace_bytes,
                     int32_t* info, void* stream);
int mimo_mstep_gauss_diag(int K, int d, int tied, const double* stat, int F,
                          double* mu, double* lmbda_diag,
                          void* workspace, size_t workspace_bytes, void* stream);
size_t mimo_mstep_lingauss_workspace(int K, int c, int o);
int mimo_mstep_lingauss(int K, int c, int o, int tied, const double* stat, int F, const int32_t* stat_idx, int Dp,
                        double* A, double* lmbda, void* workspace, size_t workspace_bytes,
                        int32_t* info, void* stream);

/* ---- host-buffer convenience (the end-to-end path bench.py's `e2e` times) */

/* One mean-field / Gibbs sweep of the quad or diag family with HOST pointers:
 * copies Z (N, D) to the device, runs mimo_sweep, copies stat / lse_sum (and
 * labels if requested) back.  Synchronous.  Allocates its own device buffers
 * (the only entry point that does) and keeps them for the next call;
 * mimo_sweep_host_release() frees them.  Z is uploaded in
 * segments of whole point chunks on a copy stream while the previous segment
 * is swept, so the call costs max(upload, compute) rather than their sum; the
 * result is the statistics of utils/abstraction.py:12-14 summed over segments
 * (the reference's list-of-arrays semantics, gaussian.py:503-505).
 * mimo_sweep_host_set_segment: points per segment (rounded up to whole chunks;
 * 0 = automatic, about 0.5 GB of data).                                        */
int mimo_sweep_host_set_segment(int64_t points);
int mimo_sweep_host_release(void);
int mimo_sweep_host(int dtype, int family, int hard,
                    const void* Z_host, int64_t N, int D,
                    const void* op_a_host, const void* op_b_host, const void* cst_host,
                    int K, int Rp, int Dpp,
                    const int32_t* fi_host, const int32_t* fj_host, int F,
                    const void* uniforms_host, uint64_t seed,
                    double* stat_host, double* lse_sum_host, int32_t* labels_host);

#ifdef __cplusplus
}
#endif
#endif /* MIMO_B200_H */
