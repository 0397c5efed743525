// Internal (C++) entry points behind the extern "C" wrappers of abi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace mimo {

int loglik_quad(int dtype, const void* Z, int64_t N, int D, int64_t ldz, const void* W, const void* cst,
                int K, int Rp, int Dpp, void* out, int64_t ldo, cudaStream_t st);
int loglik_diag(int dtype, const void* Z, int64_t N, int D, int64_t ldz, const void* S, const void* Tm,
                const void* cst, int K, void* out, int64_t ldo, cudaStream_t st,
                const unsigned int* gate = nullptr, unsigned int gate_value = 0u);
// diagonal family on the tensor pipe (tc_diag.cu): FP32, D <= 64, K <= 256; E-step as one feature GEMM, fused label draw
bool tc_diag_supported(int dtype, int D, int K);
int tc_diag_enable(int on);
size_t tc_diag_workspace();
const unsigned int* tc_diag_gate(void* ws);      // 1: the operands failed the cancellation guard, the CUDA-core kernels must run
int tc_diag_prepare(const float* Z, int64_t N, int D, int64_t ldz, const float* S, const float* T, const float* cst, int K,
                    void* ws, cudaStream_t st, float absmax_hint = 0.f);
int tc_diag_chunk(const float* Z, int64_t N, int D, int64_t ldz, int K, float* out, int64_t ldo,
                  int32_t* labels, const double* uniforms, uint64_t seed, uint64_t point_offset,
                  float* lse_out, double* lse_sum, void* ws, cudaStream_t st);
int softmax(int dtype, void* a, int K, int64_t n, int64_t ldo, int flags, void* lse, const void* uniforms,
            uint64_t seed, uint64_t point_offset, int32_t* labels, double* lse_sum, cudaStream_t st,
            const unsigned int* gate = nullptr, unsigned int gate_value = 0u);

int stats_soft(int dtype, const void* Z, int64_t N, int D, int64_t ldz, const void* resp, int64_t ldr, int K,
               const int32_t* fi, const int32_t* fj, int F, double* stat, cudaStream_t st,
               const unsigned int* gate = nullptr, unsigned int gate_value = 0u);
size_t stats_hard_workspace(int64_t N, int K);
int stats_hard(int dtype, const void* Z, int64_t N, int D, int64_t ldz, const int32_t* labels, int K,
               const int32_t* fi, const int32_t* fj, int F, double* stat,
               void* workspace, size_t workspace_bytes, bool check, cudaStream_t st);

// tensor-core (tcgen05) path, FP32 quad family, D <= 128  (tc_estep.cu, tc_stats.cu)
int tc_mode();                       // 0 CUDA cores only; 1 tensor cores: CTA pairs + screened E-step + pair-list statistics (default); 2 single-CTA dense; 3 CTA pairs dense; 4 as 1 with dense statistics; 5 as 1, screening starts on the all-rows tier
int tc_set_mode(int mode);
int tc_set_min_dim(int d);
bool tc_estep_supported(int dtype, int D, int Rp);
size_t tc_operand_workspace(int K, int Rp, int D);
int tc_data_scale(const float* Z, int64_t N, int D, int64_t ldz, void* ws, cudaStream_t st, float absmax_hint = 0.f);   // hint > 0: max |Z| is known, no pass over Z
const unsigned int* tc_maxbits(void* ws);
void tc_set_absmax_hint(float v);     // one-shot: the next sweep() of this thread takes v instead of scanning Z
float tc_take_absmax_hint();          // read and clear
int tc_prepare_operands(const float* W, const float* cst, int K, int Rp, int Dpp, int D, void* ws, cudaStream_t st);
unsigned int* tc_flags(void* ws);    // [0] max |z| bits, [2] max_k ||W'_k||_F (screening operands), [3] max_n ||z_n||_2, [4] max_k ||W_k||_F (all columns)
int tc_estep_pass(const float* Z, int64_t N, int D, int64_t ldz, int K, int Rp, float* out, int64_t ldo, void* ws,
                  int passes, const unsigned int* gate, unsigned int gate_value, float* lower, int* guess, int64_t ldl, cudaStream_t st,
                  float* lse_vals = nullptr, double* lse_sum = nullptr);
// screened E-step (tc_screen.cu): projected single-pass screening + exact refinement of the candidates / gated dense pass
bool tc_screen_supported(int D, int Rp);
size_t tc_screen_workspace(int64_t chunk_points, int K);
size_t tc_screen_operand_workspace(int K, int Rp, int Dpp, int D);
int tc_screen_rows(int Rp);
int tc_screen_prepare(const float* Z, int64_t N, int D, int64_t ldz, const float* W, const float* cst, int K, int Rp, int Dpp,
                      void* ops_ws, void* sops_ws, cudaStream_t st);
int tc_screen_pass(const float* Z, int64_t n, int D, int64_t ldz, int K, int Rp, int Dpp, float* out, int64_t ldo,
                   void* ops_ws, void* sops_ws, int64_t plan_points, void* ws, cudaStream_t st);
int tc_screen_begin(int64_t plan_points, int K, int Rp, int start_level, void* ws, cudaStream_t st);
const unsigned int* tc_screen_gate(void* ws, int64_t plan_points, int K);
int tc_screen_last(unsigned int* out_host2);
int tc_screen_totals(unsigned long long* out_host5);
void tc_screen_forget();
int tc_screen_level();
int tc_screen_select(const float* Z, int D, int64_t ldz, const float* W, const float* cst, int K, int Rp, int Dpp,
                     float* a, int64_t n, int64_t ldo, void* ops_ws, void* sops_ws, int64_t plan_points, void* ws, cudaStream_t st);
int tc_screen_refine(const float* Z, int D, int64_t ldz, const float* W, int K, int Rp, int Dpp, const float* cst,
                     float* a, int64_t ldo, int64_t plan_points, void* ws, cudaStream_t st);
int tc_screen_lse(const float* a, int K, int64_t n, int64_t ldo, double* lse_sum, int64_t plan_points, void* ws, cudaStream_t st);
const float* tc_screen_lse_values(void* ws, int64_t plan_points, int K);
void tc_screen_lists(void* ws, int64_t plan_points, int K, int which, const int32_t** perm, const int32_t** offsets, const int32_t** slabs);
// lse_vals / lse_sum (optional, CTA-pair kernel): per-point log-normalisers of the log-joints written to `out`, and their sum
int tc_estep(const float* Z, int64_t N, int D, int64_t ldz, const float* cst, int K, int Rp,
             float* out, int64_t ldo, void* ws, cudaStream_t st, float* lse_vals = nullptr, double* lse_sum = nullptr);
// CTA-pair (cta_group::2) E-step, tc_estep2.cu
size_t tc2_offsets_bytes(int K, int Rp);
int tc2_prepare_offsets(const float* rowoff, const float* invS2, const float* cst, int K, int Rp, float* offs2, cudaStream_t st);
int tc_estep2(const float* Z, int64_t N, int D, int64_t ldz, int K, int Rp, int KB, const void* Bimg, const float* offs2,
              const unsigned int* maxbits, float* out, int64_t ldo, int passes, const unsigned int* gate, unsigned int gate_value,
              float* lower, int* guess, int64_t ldl, cudaStream_t st, float* lse_vals = nullptr, double* lse_sum = nullptr);
// CTA-pair E-step with the points operand in tensor memory + triangular skip (tc_estep3.cu): 64 < D <= 128, Rp = 128
bool tc3_supported(int D, int Rp);
int tc3_set_granularity(int g);      // rows per step of the triangular skip: 16 (default) or 32; 0 turns the kernel off
size_t tc3_workspace(int K);
int tc3_prepare(const float* W, const float* cst, int K, int Dpp, int D, unsigned int* flags, void* ws3, cudaStream_t st);
int tc_estep3(const float* Z, int64_t N, int D, int64_t ldz, int K, const void* ws3, const unsigned int* flags,
              float* out, int64_t ldo, const unsigned int* gate, unsigned int gate_value,
              float* lse_vals, double* lse_sum, cudaStream_t st);
// CTA-pair E-step, four components per accumulator generation, zero block skipped at full MMA width (tc_estep4.cu)
bool tc4_supported(int D, int Rp);
int tc4_enable(int on);
size_t tc4_workspace(int K);
int tc4_prepare(const float* W, const float* cst, int K, int Dpp, int D, unsigned int* flags, void* ws4, cudaStream_t st);
int tc_estep4(const float* Z, int64_t N, int D, int64_t ldz, int K, const void* ws4, const unsigned int* flags,
              float* out, int64_t ldo, const unsigned int* gate, unsigned int gate_value,
              float* lse_vals, double* lse_sum, cudaStream_t st);
int loglik_quad_tc(const void* Z, int64_t N, int D, int64_t ldz, const void* W, const void* cst,
                   int K, int Rp, int Dpp, void* out, int64_t ldo, void* ws, size_t ws_bytes, cudaStream_t st);
bool tc_stats_supported(int dtype, int D, int F);
size_t tc_stats_workspace(int64_t chunk_points, int K);
int tc_stats_begin(int64_t chunk_points, int K, void* ws, cudaStream_t st);
int tc_stats_chunk(const float* Z, int64_t N, int D, int64_t ldz, const float* R, int64_t ldr, int K, int F,
                   const unsigned int* maxbits, double* stat, int64_t plan_points, void* ws, cudaStream_t st,
                   const unsigned int* gate = nullptr, unsigned int gate_value = 0u, const float* lse = nullptr);
int tc_stats_end(int64_t plan_points, int K, int D, int F, const unsigned int* maxbits, double* stat, void* ws, cudaStream_t st);
void tc_set_flush_tiles(int tiles);
// feature-form statistics (tc_fstats.cu): folded lower triangle, 64 < D <= 128
bool tc_fstats_supported(int dtype, int D, int F);
size_t tc_fstats_workspace(int64_t chunk_points, int K);
int tc_fstats_begin(int64_t chunk_points, int K, void* ws, cudaStream_t st);
int tc_fstats_chunk(const float* Z, int64_t N, int D, int64_t ldz, const float* R, int64_t ldr, int K,
                    const unsigned int* maxbits, int64_t plan_points, void* ws, cudaStream_t st,
                    const unsigned int* gate = nullptr, unsigned int gate_value = 0u, const float* lse = nullptr);
int tc_fstats_end(int64_t plan_points, int K, int D, int F, const unsigned int* maxbits, double* stat, void* ws, cudaStream_t st);
void tc_fstats_set_flush_tiles(int tiles);
int tc_fstats_stall_clocks(unsigned long long* out_host8);
// feature-form statistics for small dimensions (tc_sstats.cu): D <= 21, accumulates straight into stat (K, F)
bool tc_sstats_supported(int dtype, int D, int F);
int tc_sstats_enable(int on);
int tc_sstats_chunk(const float* Z, int64_t N, int D, int64_t ldz, const float* R, int64_t ldr, const float* lse, int K, int F,
                    const unsigned int* maxbits, double* stat, cudaStream_t st,
                    const unsigned int* gate = nullptr, unsigned int gate_value = 0u);
size_t stats_soft_tc_workspace(int64_t N, int K);
int stats_soft_tc(const void* Z, int64_t N, int D, int64_t ldz, const void* resp, int64_t ldr, int K, int F,
                  double* stat, void* ws, size_t ws_bytes, cudaStream_t st);

// statistics all-reduce over NCCL, loaded at run time (comm.cu)
size_t comm_unique_id_bytes();
int comm_unique_id(void* out);
int comm_init(int world, int rank, const void* unique_id, void** comm_out);
int comm_allreduce_stats(void* comm, double* stat, int64_t count, cudaStream_t st);
int comm_destroy(void* comm);

// prediction path of the linear-Gaussian mixtures (predict.cu)
int studentt_from_quad(int dtype, void* a, int K, int64_t N, int64_t lda, const double* c0, const double* add, const double* df,
                       cudaStream_t st);
int predict_lingauss(int dtype, const void* X, int64_t N, int64_t ldx, int din, int affine, const void* W, int64_t ldw, int K,
                     const double* M, const double* Kinv, const double* Sig, const double* Psi, const double* logdet, const double* df,
                     int o, int tied, int mode, int studentt, const void* Y, int64_t ldy, double eps,
                     void* mu_out, void* cov_out, void* nlpd_out, cudaStream_t st);

// statistics over (component, point) lists grouped by component (pair_stats.cu): hard labels / screened candidates
constexpr int PS_SLAB = 1024;        // listed points of one component per work item
bool pair_stats_supported(int dtype, int D, int F);
int pair_stats(const float* Z, int D, int64_t ldz, const int32_t* perm, const int32_t* offsets, const int32_t* slabs, int K,
               const float* R, int64_t ldr, const float* lse, const unsigned int* gate, unsigned int gate_value,
               double* stat, int F, cudaStream_t st);

// responsibility lists of the CUDA-core path (tc_screen.cu)
size_t resp_list_workspace(int64_t chunk_points, int K);
int resp_list_build(const float* R, int K, int64_t n, int64_t ldr, int D, int64_t plan_points, void* ws, cudaStream_t st);
const unsigned int* resp_list_gate(void* ws, int64_t plan_points, int K);
void resp_list_get(void* ws, int64_t plan_points, int K, const int32_t** perm, const int32_t** offsets, const int32_t** slabs);

void sweep_set_tables_hint(int canonical);
bool sweep_uses_tc(int dtype, int family, int D, int Rp);
int64_t sweep_chunk_points(int dtype, int family, int64_t N, int D, int K, int Rp);
size_t sweep_workspace(int dtype, int family, int hard, int64_t N, int D, int K, int Rp);
int sweep(int dtype, int family, int hard, const void* Z, int64_t N, int D, int64_t ldz,
          const void* op_a, const void* op_b, const void* cst, int K, int Rp, int Dpp,
          const int32_t* fi, const int32_t* fj, int F,
          const void* uniforms, uint64_t seed, uint64_t point_offset,
          double* stat, double* lse_sum, int32_t* labels_out, void* lse_out, void* ll_out, int64_t ldo,
          void* workspace, size_t workspace_bytes, cudaStream_t st, double* phase_ms = nullptr);
int64_t sweep_host_set_segment(int64_t points);
void sweep_host_release();
int sweep_host(int dtype, int family, int hard, const void* Z_host, int64_t N, int D,
               const void* op_a_host, const void* op_b_host, const void* cst_host, int K, int Rp, int Dpp,
               const int32_t* fi_host, const int32_t* fj_host, int F,
               const void* uniforms_host, uint64_t seed,
               double* stat_host, double* lse_sum_host, int32_t* labels_host);

size_t nw_workspace(int K, int d);
int nw_posterior(int K, int d, int tied, int mode,
                 const double* m0, const double* kappa0, const double* psi0, const double* nu0,
                 const double* stat, int F, const int32_t* stat_idx, int Dp, const double* variates,
                 double* post_m, double* post_kappa, double* post_psi, double* post_nu,
                 double* lik_mu, double* lik_lmbda, double* vlb,
                 int op_dtype, void* W, void* cst, int Rp, int Dpp, int row_off, const int32_t* col_map,
                 void* workspace, size_t workspace_bytes, int32_t* info, cudaStream_t st);
size_t ng_workspace(int K, int d);
int ng_posterior(int K, int d, int tied, int mode, int bug_compat,
                 const double* m0, const double* kappa0, const double* alpha0, const double* beta0,
                 const double* stat, int F, const double* variates,
                 double* post_m, double* post_kappa, double* post_alpha, double* post_beta,
                 double* lik_mu, double* lik_l, double* vlb,
                 int op_dtype, void* S, void* T, void* cst, void* workspace, size_t workspace_bytes,
                 int32_t* info, cudaStream_t st);
size_t mnw_workspace(int K, int c, int o);
int mnw_posterior(int K, int c, int o, int tied, int mode,
                  const double* M0, const double* K0, const double* psi0, const double* nu0,
                  const double* stat, int F, const int32_t* stat_idx, int Dp, const double* variates,
                  double* post_M, double* post_K, double* post_psi, double* post_nu,
                  double* lik_A, double* lik_lmbda, double* vlb,
                  int op_dtype, void* W, void* cst, int Rp, int Dpp, int row_off, const int32_t* col_map,
                  void* workspace, size_t workspace_bytes, int32_t* info, cudaStream_t st);
size_t gating_workspace(int K);
int gating_posterior(int K, int kind, int mode, const double* prior_a, const double* prior_b,
                     const double* stat, int F, int count_feature, const double* variates,
                     double* post_a, double* post_b, double* probs, double* vlb,
                     int op_dtype, void* cst, void* workspace, size_t workspace_bytes,
                     int32_t* info, cudaStream_t st);
int operands_gauss(int K, int d, int c, const double* mu_or_A, const double* lmbda,
                   int op_dtype, void* W, void* cst, int Rp, int Dpp, int row_off, const int32_t* col_map,
                   void* workspace, size_t workspace_bytes, int32_t* info, cudaStream_t st);
int operands_gauss_diag(int K, int d, const double* mu, const double* lam, int op_dtype,
                        void* S, void* T, void* cst, cudaStream_t st);
size_t mstep_workspace(int K, int d);
int mstep_gauss(int K, int d, int tied, const double* stat, int F, const int32_t* stat_idx, int Dp,
                double* mu, double* lmbda, void* workspace, size_t workspace_bytes, int32_t* info, cudaStream_t st);
int mstep_gauss_diag(int K, int d, int tied, const double* stat, int F, double* mu, double* lam,
                     void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t mstep_lingauss_workspace(int K, int c, int o);
int mstep_lingauss(int K, int c, int o, int tied, const double* stat, int F, const int32_t* stat_idx, int Dp,
                   double* A, double* lmbda, void* workspace, size_t workspace_bytes, int32_t* info, cudaStream_t st);

}  // namespace mimo
