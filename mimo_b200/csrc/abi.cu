// extern "C" boundary of libmimo_b200.so (see include/mimo_b200.h).
#include "common.cuh"
#include "internal.h"
#include <stdarg.h>
#include <string.h>

namespace mimo {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaDeviceProp p;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&p, dev) == cudaSuccess) n = p.multiProcessorCount;
        else n = 148;
    }
    return n;
}

}  // namespace mimo

using namespace mimo;
#define ST(s) ((cudaStream_t)(s))

extern "C" {

const char* mimo_last_error_string(void) { return g_err; }
int mimo_version(void) { return 100; }

int mimo_device_ok(void) {
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
        set_error("no CUDA device");
        return 0;
    }
    if (p.major != 10) { set_error("device is sm_%d%d; this library is built for sm_100a only", p.major, p.minor); return 0; }
    return 1;
}

int mimo_loglik_quad(int dtype, const void* Z, int64_t N, int D, int64_t ldz, const void* W, const void* cst,
                     int K, int Rp, int Dpp, void* out, int64_t ldo, void* stream) {
    return loglik_quad(dtype, Z, N, D, ldz, W, cst, K, Rp, Dpp, out, ldo, ST(stream));
}
int mimo_loglik_diag(int dtype, const void* Z, int64_t N, int D, int64_t ldz, const void* S, const void* T,
                     const void* cst, int K, void* out, int64_t ldo, void* stream) {
    return loglik_diag(dtype, Z, N, D, ldz, S, T, cst, K, out, ldo, ST(stream));
}
size_t mimo_loglik_diag_tc_workspace(void) { return tc_diag_workspace(); }
int mimo_loglik_diag_tc(const void* Z, int64_t N, int D, int64_t ldz, const void* S, const void* T, const void* cst, int K,
                        void* out, int64_t ldo, int32_t* labels, const void* uniforms, uint64_t seed, uint64_t point_offset,
                        void* lse, double* lse_sum, uint32_t* guard_host, void* workspace, size_t workspace_bytes, void* stream) {
    MIMO_CHECK_ARG(Z && S && T && cst && workspace, "null pointer");
    MIMO_CHECK_ARG(N >= 0 && ldz >= D && (!out || ldo >= N), "shape");
    if (!tc_diag_supported(MIMO_F32, D, K)) { set_error("tensor-core diagonal E-step: unsupported shape D=%d K=%d", D, K); return MIMO_EUNSUPPORTED; }
    MIMO_CHECK_ARG(workspace_bytes >= tc_diag_workspace(), "workspace too small");
    int rc = tc_diag_prepare((const float*)Z, N, D, ldz, (const float*)S, (const float*)T, (const float*)cst, K, workspace, ST(stream));
    if (rc) return rc;
    rc = tc_diag_chunk((const float*)Z, N, D, ldz, K, (float*)out, ldo, labels, (const double*)uniforms, seed, point_offset,
                       (float*)lse, lse_sum, workspace, ST(stream));
    if (rc) return rc;
    if (guard_host) {
        MIMO_CUDA(cudaMemcpyAsync(guard_host, tc_diag_gate(workspace), 4, cudaMemcpyDeviceToHost, ST(stream)));
        MIMO_CUDA(cudaStreamSynchronize(ST(stream)));
    }
    return MIMO_OK;
}
int mimo_studentt_from_quad(int dtype, void* a, int K, int64_t N, int64_t lda, const double* c0, const double* add, const double* df,
                            void* stream) {
    return studentt_from_quad(dtype, a, K, N, lda, c0, add, df, ST(stream));
}
int mimo_predict_lingauss(int dtype, const void* X, int64_t N, int64_t ldx, int din, int affine, const void* W, int64_t ldw, int K,
                          const double* M, const double* Kinv, const double* Sigma, const double* Psi, const double* logdet_psi,
                          const double* df, int o, int tied, int mode, int studentt, const void* Y, int64_t ldy, double eps,
                          void* mu_out, void* cov_out, void* nlpd_out, void* stream) {
    return predict_lingauss(dtype, X, N, ldx, din, affine, W, ldw, K, M, Kinv, Sigma, Psi, logdet_psi, df, o, tied, mode, studentt,
                            Y, ldy, eps, mu_out, cov_out, nlpd_out, ST(stream));
}
size_t mimo_comm_unique_id_bytes(void) { return comm_unique_id_bytes(); }
int mimo_comm_unique_id(void* out) { return comm_unique_id(out); }
int mimo_comm_init(int world, int rank, const void* unique_id, void** comm_out) { return comm_init(world, rank, unique_id, comm_out); }
int mimo_comm_allreduce_stats(void* comm, double* stat, int64_t count, void* stream) { return comm_allreduce_stats(comm, stat, count, ST(stream)); }
int mimo_comm_destroy(void* comm) { return comm_destroy(comm); }
int mimo_sweep_tables_hint(int canonical) { sweep_set_tables_hint(canonical); return MIMO_OK; }
int mimo_sweep_absmax_hint(double absmax) { tc_set_absmax_hint((float)absmax); return MIMO_OK; }
int mimo_tc_fstats_stall_clocks(uint64_t* out_host8) { return tc_fstats_stall_clocks((unsigned long long*)out_host8); }
int mimo_tc_diag_enable(int on) { return tc_diag_enable(on); }
int mimo_tc_set_min_dim(int d) { return tc_set_min_dim(d); }
int mimo_softmax(int dtype, void* a, int K, int64_t n, int64_t ldo, int flags, void* lse, const void* uniforms,
                 uint64_t seed, uint64_t point_offset, int32_t* labels, double* lse_sum, void* stream) {
    return softmax(dtype, a, K, n, ldo, flags, lse, uniforms, seed, point_offset, labels, lse_sum, ST(stream));
}
int mimo_stats_soft(int dtype, const void* Z, int64_t N, int D, int64_t ldz, const void* resp, int64_t ldr, int K,
                    const int32_t* fi, const int32_t* fj, int F, double* stat, void* stream) {
    return stats_soft(dtype, Z, N, D, ldz, resp, ldr, K, fi, fj, F, stat, ST(stream));
}
size_t mimo_stats_hard_workspace(int64_t N, int K) { return stats_hard_workspace(N, K); }
int mimo_stats_hard(int dtype, const void* Z, int64_t N, int D, int64_t ldz, const int32_t* labels, int K,
                    const int32_t* fi, const int32_t* fj, int F, double* stat,
                    void* workspace, size_t workspace_bytes, void* stream) {
    return stats_hard(dtype, Z, N, D, ldz, labels, K, fi, fj, F, stat, workspace, workspace_bytes, true, ST(stream));
}
size_t mimo_sweep_workspace(int dtype, int family, int hard, int64_t N, int D, int K, int Rp) {
    return sweep_workspace(dtype, family, hard, N, D, K, Rp);
}
int mimo_set_tensor_cores(int mode) { return tc_set_mode(mode); }
int mimo_tc_screen_last(uint32_t* out_host2) { return tc_screen_last(out_host2); }
int mimo_tc_screen_totals(uint64_t* out_host5) { return tc_screen_totals((unsigned long long*)out_host5); }
int mimo_sweep_uses_tensor_cores(int dtype, int family, int D, int Rp) { return sweep_uses_tc(dtype, family, D, Rp) ? 1 : 0; }
int mimo_tc_set_triangular(int rows) { return tc3_set_granularity(rows); }
int mimo_tc_set_quad_generations(int on) { return tc4_enable(on); }
int mimo_tc_set_flush_tiles(int tiles) { tc_set_flush_tiles(tiles); tc_fstats_set_flush_tiles(tiles); return MIMO_OK; }
size_t mimo_loglik_quad_tc_workspace(int K, int Rp, int D) { return tc_operand_workspace(K, Rp, D); }
int mimo_loglik_quad_tc(const void* Z, int64_t N, int D, int64_t ldz, const void* W, const void* cst,
                        int K, int Rp, int Dpp, void* out, int64_t ldo, void* workspace, size_t workspace_bytes, void* stream) {
    return loglik_quad_tc(Z, N, D, ldz, W, cst, K, Rp, Dpp, out, ldo, workspace, workspace_bytes, ST(stream));
}
size_t mimo_stats_soft_tc_workspace(int64_t N, int K) { return stats_soft_tc_workspace(N, K); }
int mimo_stats_soft_tc(const void* Z, int64_t N, int D, int64_t ldz, const void* resp, int64_t ldr, int K, int F,
                       double* stat, void* workspace, size_t workspace_bytes, void* stream) {
    return stats_soft_tc(Z, N, D, ldz, resp, ldr, K, F, stat, workspace, workspace_bytes, ST(stream));
}
int mimo_sweep(int dtype, int family, int hard, const void* Z, int64_t N, int D, int64_t ldz,
               const void* op_a, const void* op_b, const void* cst, int K, int Rp, int Dpp,
               const int32_t* fi, const int32_t* fj, int F,
               const void* uniforms, uint64_t seed, uint64_t point_offset,
               double* stat, double* lse_sum, int32_t* labels_out, void* lse_out, void* ll_out, int64_t ldo,
               void* workspace, size_t workspace_bytes, void* stream) {
    return sweep(dtype, family, hard, Z, N, D, ldz, op_a, op_b, cst, K, Rp, Dpp, fi, fj, F, uniforms, seed,
                 point_offset, stat, lse_sum, labels_out, lse_out, ll_out, ldo, workspace, workspace_bytes, ST(stream));
}
int mimo_sweep_timed(int dtype, int family, int hard, const void* Z, int64_t N, int D, int64_t ldz,
                     const void* op_a, const void* op_b, const void* cst, int K, int Rp, int Dpp,
                     const int32_t* fi, const int32_t* fj, int F,
                     const void* uniforms, uint64_t seed, uint64_t point_offset,
                     double* stat, double* lse_sum, int32_t* labels_out, void* lse_out, void* ll_out, int64_t ldo,
                     void* workspace, size_t workspace_bytes, void* stream, double* phase_ms_host) {
    return sweep(dtype, family, hard, Z, N, D, ldz, op_a, op_b, cst, K, Rp, Dpp, fi, fj, F, uniforms, seed,
                 point_offset, stat, lse_sum, labels_out, lse_out, ll_out, ldo, workspace, workspace_bytes, ST(stream),
                 phase_ms_host);
}
int mimo_tc_screen_level(void) { return tc_screen_level(); }
int mimo_sweep_host_release(void) { sweep_host_release(); return MIMO_OK; }
int mimo_sweep_host_set_segment(int64_t points) { sweep_host_set_segment(points); return MIMO_OK; }
int mimo_sweep_host(int dtype, int family, int hard, const void* Z_host, int64_t N, int D,
                    const void* op_a_host, const void* op_b_host, const void* cst_host, int K, int Rp, int Dpp,
                    const int32_t* fi_host, const int32_t* fj_host, int F, const void* uniforms_host, uint64_t seed,
                    double* stat_host, double* lse_sum_host, int32_t* labels_host) {
    return sweep_host(dtype, family, hard, Z_host, N, D, op_a_host, op_b_host, cst_host, K, Rp, Dpp,
                      fi_host, fj_host, F, uniforms_host, seed, stat_host, lse_sum_host, labels_host);
}

size_t mimo_nw_workspace(int K, int d) { return nw_workspace(K, d); }
int mimo_nw_posterior(int K, int d, int tied, int mode,
                      const double* m0, const double* kappa0, const double* psi0, const double* nu0,
                      const double* stat, int F, const int32_t* stat_idx, int Dp, const double* variates,
                      double* post_m, double* post_kappa, double* post_psi, double* post_nu,
                      double* lik_mu, double* lik_lmbda, double* vlb,
                      int op_dtype, void* W, void* cst, int Rp, int Dpp, int row_off, const int32_t* col_map,
                      void* workspace, size_t workspace_bytes, int32_t* info, void* stream) {
    return nw_posterior(K, d, tied, mode, m0, kappa0, psi0, nu0, stat, F, stat_idx, Dp, variates,
                        post_m, post_kappa, post_psi, post_nu, lik_mu, lik_lmbda, vlb,
                        op_dtype, W, cst, Rp, Dpp, row_off, col_map, workspace, workspace_bytes, info, ST(stream));
}
size_t mimo_ng_workspace(int K, int d) { return ng_workspace(K, d); }
int mimo_ng_posterior(int K, int d, int tied, int mode, int bug_compat,
                      const double* m0, const double* kappa0, const double* alpha0, const double* beta0,
                      const double* stat, int F, const double* variates,
                      double* post_m, double* post_kappa, double* post_alpha, double* post_beta,
                      double* lik_mu, double* lik_lmbda_diag, double* vlb,
                      int op_dtype, void* S, void* T, void* cst,
                      void* workspace, size_t workspace_bytes, int32_t* info, void* stream) {
    return ng_posterior(K, d, tied, mode, bug_compat, m0, kappa0, alpha0, beta0, stat, F, variates,
                        post_m, post_kappa, post_alpha, post_beta, lik_mu, lik_lmbda_diag, vlb,
                        op_dtype, S, T, cst, workspace, workspace_bytes, info, ST(stream));
}
size_t mimo_mnw_workspace(int K, int c, int o) { return mnw_workspace(K, c, o); }
int mimo_mnw_posterior(int K, int c, int o, int tied, int mode,
                       const double* M0, const double* K0, const double* psi0, const double* nu0,
                       const double* stat, int F, const int32_t* stat_idx, int Dp, const double* variates,
                       double* post_M, double* post_K, double* post_psi, double* post_nu,
                       double* lik_A, double* lik_lmbda, double* vlb,
                       int op_dtype, void* W, void* cst, int Rp, int Dpp, int row_off, const int32_t* col_map,
                       void* workspace, size_t workspace_bytes, int32_t* info, void* stream) {
    return mnw_posterior(K, c, o, tied, mode, M0, K0, psi0, nu0, stat, F, stat_idx, Dp, variates,
                         post_M, post_K, post_psi, post_nu, lik_A, lik_lmbda, vlb,
                         op_dtype, W, cst, Rp, Dpp, row_off, col_map, workspace, workspace_bytes, info, ST(stream));
}
size_t mimo_gating_workspace(int K) { return gating_workspace(K); }
int mimo_gating_posterior(int K, int kind, int mode, const double* prior_a, const double* prior_b,
                          const double* stat, int F, int count_feature, const double* variates,
                          double* post_a, double* post_b, double* probs, double* vlb,
                          int op_dtype, void* cst, void* workspace, size_t workspace_bytes,
                          int32_t* info, void* stream) {
    return gating_posterior(K, kind, mode, prior_a, prior_b, stat, F, count_feature, variates,
                            post_a, post_b, probs, vlb, op_dtype, cst, workspace, workspace_bytes, info, ST(stream));
}

size_t mimo_operands_workspace(int K, int d) { return 8 * (size_t)K * d * d; }
int mimo_operands_gauss(int K, int d, const double* mu, const double* lmbda,
                        int op_dtype, void* W, void* cst, int Rp, int Dpp, int row_off, const int32_t* col_map,
                        void* workspace, size_t workspace_bytes, int32_t* info, void* stream) {
    return operands_gauss(K, d, 0, mu, lmbda, op_dtype, W, cst, Rp, Dpp, row_off, col_map,
                          workspace, workspace_bytes, info, ST(stream));
}
int mimo_operands_lingauss(int K, int c, int o, const double* A, const double* lmbda,
                           int op_dtype, void* W, void* cst, int Rp, int Dpp, int row_off, const int32_t* col_map,
                           void* workspace, size_t workspace_bytes, int32_t* info, void* stream) {
    return operands_gauss(K, o, c, A, lmbda, op_dtype, W, cst, Rp, Dpp, row_off, col_map,
                          workspace, workspace_bytes, info, ST(stream));
}
int mimo_operands_gauss_diag(int K, int d, const double* mu, const double* lmbda_diag,
                             int op_dtype, void* S, void* T, void* cst, void* stream) {
    return operands_gauss_diag(K, d, mu, lmbda_diag, op_dtype, S, T, cst, ST(stream));
}

size_t mimo_mstep_workspace(int K, int d) { return mstep_workspace(K, d); }
int mimo_mstep_gauss(int K, int d, int tied, const double* stat, int F, const int32_t* stat_idx, int Dp,
                     double* mu, double* lmbda, void* workspace, size_t workspace_bytes, int32_t* info, void* stream) {
    return mstep_gauss(K, d, tied, stat, F, stat_idx, Dp, mu, lmbda, workspace, workspace_bytes, info, ST(stream));
}
int mimo_mstep_gauss_diag(int K, int d, int tied, const double* stat, int F, double* mu, double* lmbda_diag,
                          void* workspace, size_t workspace_bytes, void* stream) {
    return mstep_gauss_diag(K, d, tied, stat, F, mu, lmbda_diag, workspace, workspace_bytes, ST(stream));
}
size_t mimo_mstep_lingauss_workspace(int K, int c, int o) { return mstep_lingauss_workspace(K, c, o); }
int mimo_mstep_lingauss(int K, int c, int o, int tied, const double* stat, int F, const int32_t* stat_idx, int Dp,
                        double* A, double* lmbda, void* workspace, size_t workspace_bytes, int32_t* info, void* stream) {
    return mstep_lingauss(K, c, o, tied, stat, F, stat_idx, Dp, A, lmbda, workspace, workspace_bytes, info, ST(stream));
}

}  // extern "C"
