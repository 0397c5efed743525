// One sweep over resident data: E-step -> softmax / label draw -> statistics, walked in
// point chunks through an L2-sized (K, chunk) scratch so the (K, N) arrays of the
// reference (mixtures/gmm.py:67-75, utils/data.py:160-169) are never built.
// CUDA-core path; the tcgen05 path (tc_*.cu) replaces the inner three launches where
// it supports the shape.
#include "common.cuh"
#include "internal.h"
#include <nvtx3/nvToolsExt.h>      // header-only; ranges cost nothing unless a profiler is attached (tracing)
#include <vector>
#include <chrono>
#include <stdlib.h>
#include <algorithm>

namespace mimo {

static size_t a256(size_t x) { return (x + 255) / 256 * 256; }

static int g_tc_mode = 1;
int tc_mode() { return g_tc_mode; }
int tc_set_mode(int mode) { int old = g_tc_mode; g_tc_mode = (mode < 0 || mode > 5) ? 1 : mode; return old; }

// the tensor-core path takes FP32 quad-family sweeps from D = 8 up: below D = 24 the E-step is bound by the read-back of
// its K * Rp accumulator columns per point (one 16-wide K step feeds 256 columns) and the statistics run in feature form
// (tc_sstats.cu); from D = 24 the contraction is wide enough for the per-component / folded-triangle statistics kernels
static int g_tc_min_d = 8;
int tc_set_min_dim(int d) { int old = g_tc_min_d; g_tc_min_d = d < 1 ? 1 : d; return old; }
bool sweep_uses_tc(int dtype, int family, int D, int Rp) {
    return g_tc_mode && family == 0 && D >= g_tc_min_d && tc_estep_supported(dtype, D, Rp);
}
// screened E-step: worth it when a point's candidates (>= 1) can stay below 4 % of the K components
static bool sweep_uses_screen(int dtype, int family, int D, int K, int Rp) {
    return (g_tc_mode == 1 || g_tc_mode == 4 || g_tc_mode == 5) && K >= 32 && sweep_uses_tc(dtype, family, D, Rp) && tc_screen_supported(D, Rp);
}

// CUDA-core quad path in FP32, 16 <= D < 24: statistics over the list of pairs with a non-negligible responsibility
// (pair_stats.cu).  Below D = 16 the triangle has 3 register tiles or fewer and the list kernel loses to the dense one
// at any list length (measured on cfg2, D = 9).
static bool sweep_uses_resp_list(int dtype, int family, int hard, int D, int K, int Rp) {
    return (g_tc_mode == 1 || g_tc_mode == 5) && !hard && family == 0 && dtype == MIMO_F32 && D >= 16 && K >= 8 && !sweep_uses_tc(dtype, family, D, Rp);
}

// points per chunk.
//  CUDA-core path: keep the (K, chunk) scratch around 64 MB (half of the 126 MB L2) so the softmax /
//    statistics passes over it are served from L2.
//  tensor-core path: the kernels are persistent (one CTA per SM walking 256-point pairs / 4-component
//    units), so a chunk is many waves of 256 x #SM points; the scratch (<= 4 GB) streams through HBM,
//    which costs ~12 B per pair against ~66 kflop at d = 128.
// The list / tensor-core statistics kernels write the CANONICAL packed triangle (f = i (i + 1) / 2 + j over zt = [z ; 1]);
// a caller's (fi, fj) table of the same length in another order must take the generic kernel.  The table is read back
// once per sweep (F * 8 bytes) -- mimo_sweep_host checks its host copy instead and leaves the answer in g_tables_hint.
static thread_local int g_tables_hint = -1;              // -1 unknown, 0 not canonical, 1 canonical
static bool canonical_host(const int32_t* fi, const int32_t* fj, int F, int D) {
    if (F != (D + 1) * (D + 2) / 2) return false;
    int f = 0;
    for (int i = 0; i <= D; ++i)
        for (int j = 0; j <= i; ++j, ++f)
            if (fi[f] != i || fj[f] != j) return false;
    return true;
}
// A caller that built the tables itself may say so for its NEXT sweep (mimo_sweep_tables_hint): the sweep then does not
// synchronise to read them back -- which is also what lets a whole iteration be captured in a CUDA graph.  One-shot.
static thread_local int g_tables_promise = -1;
void sweep_set_tables_hint(int canonical) { g_tables_promise = canonical < 0 ? -1 : (canonical ? 1 : 0); }

static int tables_canonical(const int32_t* fi, const int32_t* fj, int F, int D, cudaStream_t st, int promise, bool* out) {
    *out = false;
    if (F != (D + 1) * (D + 2) / 2) return MIMO_OK;
    if (g_tables_hint >= 0) { *out = g_tables_hint == 1; return MIMO_OK; }
    if (promise >= 0) { *out = promise == 1; return MIMO_OK; }
    std::vector<int32_t> h((size_t)2 * F);
    MIMO_CUDA(cudaMemcpyAsync(h.data(), fi, (size_t)F * 4, cudaMemcpyDeviceToHost, st));
    MIMO_CUDA(cudaMemcpyAsync(h.data() + F, fj, (size_t)F * 4, cudaMemcpyDeviceToHost, st));
    MIMO_CUDA(cudaStreamSynchronize(st));
    *out = canonical_host(h.data(), h.data() + F, F, D);
    return MIMO_OK;
}

// diagonal family on the tensor pipe (tc_diag.cu)
static bool sweep_uses_diag_tc(int dtype, int family, int D, int K) {
    return g_tc_mode && family == 1 && tc_diag_supported(dtype, D, K);
}

int64_t sweep_chunk_points(int dtype, int family, int64_t N, int D, int K, int Rp) {
    size_t es = dtype == MIMO_F32 ? 4 : 8;
    int64_t npad = (N + 255) / 256 * 256;
    if (npad <= 0) npad = 256;
    if (sweep_uses_tc(dtype, family, D, Rp) || sweep_uses_diag_tc(dtype, family, D, K)) {
        const int64_t wave = (int64_t)256 * sm_count();
        int64_t c = (int64_t)(((size_t)4 << 30) / ((size_t)K * es));
        c = c / wave * wave;
        if (c < wave) c = wave;
        return c < npad ? c : npad;
    }
    int64_t c = (int64_t)((64u << 20) / ((size_t)K * es));
    c = c / 256 * 256;
    if (c < 1024) c = 1024;
    if (c > (1 << 20)) c = 1 << 20;
    return c < npad ? c : npad;
}

size_t sweep_workspace(int dtype, int family, int hard, int64_t N, int D, int K, int Rp) {
    size_t es = dtype == MIMO_F32 ? 4 : 8;
    int64_t c = sweep_chunk_points(dtype, family, N, D, K, Rp);
    size_t b = a256((size_t)K * c * es);
    // Gibbs: labels of ALL points (when the caller does not keep them) + one counting sort over N
    if (hard) b += a256((size_t)N * 4) + a256(stats_hard_workspace(N, K));
    if (sweep_uses_resp_list(dtype, family, hard, D, K, Rp)) b += a256(resp_list_workspace(c, K));
    if (sweep_uses_diag_tc(dtype, family, D, K)) b += a256(tc_diag_workspace());
    if (sweep_uses_tc(dtype, family, D, Rp)) {
        b += a256(tc_operand_workspace(K, Rp, D));
        if (sweep_uses_screen(dtype, family, D, K, Rp)) b += a256(tc_screen_workspace(c, K)) + a256(tc_screen_operand_workspace(K, Rp, D + 4, D));
        if (!hard) {                                                    // log-normalisers of a chunk + statistics partials
            b += a256((size_t)c * 4);
            if (!tc_sstats_supported(dtype, D, (D + 1) * (D + 2) / 2)) b += a256(std::max(tc_stats_workspace(c, K), tc_fstats_workspace(c, K)));
        }
    }
    return b;
}

int sweep(int dtype, int family, int hard, const void* Z, int64_t N, int D, int64_t ldz,
          const void* op_a, const void* op_b, const void* cst, int K, int Rp, int Dpp,
          const int32_t* fi, const int32_t* fj, int F,
          const void* uniforms, uint64_t seed, uint64_t point_offset,
          double* stat, double* lse_sum, int32_t* labels_out, void* lse_out, void* ll_out, int64_t ldo,
          void* workspace, size_t workspace_bytes, cudaStream_t st, double* phase_ms) {
    const int promise = g_tables_promise;              // one-shot: whatever this sweep does with it, the next one starts clean
    g_tables_promise = -1;
    const float absmax_hint = tc_take_absmax_hint();   // likewise: taken here, whichever kernels the sweep ends up running
    MIMO_CHECK_ARG(dtype == MIMO_F32 || dtype == MIMO_F64, "dtype");
    MIMO_CHECK_ARG(family == 0 || family == 1, "family");
    MIMO_CHECK_ARG(Z && op_a && cst && workspace && (family == 0 || op_b), "null pointer");
    MIMO_CHECK_ARG(!stat || (fi && fj && F >= 1), "feature tables");
    MIMO_CHECK_ARG(workspace_bytes >= sweep_workspace(dtype, family, hard, N, D, K, Rp), "workspace too small");
    MIMO_CHECK_ARG(!ll_out || ldo >= N, "ldo");
    const size_t es = dtype == MIMO_F32 ? 4 : 8;
    const bool use_tc = sweep_uses_tc(dtype, family, D, Rp);
    const int64_t C = sweep_chunk_points(dtype, family, N, D, K, Rp);
    char* ws = (char*)workspace;
    void* scratch = ws; ws += a256((size_t)K * C * es);
    int32_t* lab_all = labels_out;
    if (hard && !lab_all) { lab_all = (int32_t*)ws; ws += a256((size_t)N * 4); }
    void* hard_ws = ws;
    const size_t hard_ws_bytes = hard ? stats_hard_workspace(N, K) : 0;
    if (hard) ws += a256(hard_ws_bytes);
    void* resp_ws = nullptr;
    // soft statistics through the list / tensor-core kernels need the canonical packed triangle (checked, not assumed)
    bool canon = false;
    if (stat && !hard && family == 0 && dtype == MIMO_F32 && (use_tc || sweep_uses_resp_list(dtype, family, hard, D, K, Rp))) {
        int rc = tables_canonical(fi, fj, F, D, st, promise, &canon);
        if (rc) return rc;
    }
    const bool resp_list = stat && canon && sweep_uses_resp_list(dtype, family, hard, D, K, Rp) && pair_stats_supported(dtype, D, F);
    if (sweep_uses_resp_list(dtype, family, hard, D, K, Rp)) { resp_ws = ws; ws += a256(resp_list_workspace(C, K)); }
    // diagonal family: E-step (+ label draw of a Gibbs sweep) on the tensor pipe; the CUDA-core kernels stay behind a
    // device-side gate for operands that fail the cancellation guard
    const bool diag_tc = sweep_uses_diag_tc(dtype, family, D, K);
    const bool diag_fused = diag_tc && hard && !ll_out;          // labels and log-normalisers out of the kernel's epilogue
    void* diag_ws = nullptr;
    if (diag_tc) {
        diag_ws = ws; ws += a256(tc_diag_workspace());
        int rc = tc_diag_prepare((const float*)Z, N, D, ldz, (const float*)op_a, (const float*)op_b, (const float*)cst, K, diag_ws, st, absmax_hint);
        if (rc) return rc;
    }
    void* tc_ops_ws = nullptr;
    void* tc_stat_ws = nullptr;
    void* screen_ws = nullptr;
    void* screen_ops_ws = nullptr;
    // the (K, N) log-joint output must hold FP32-class values for every pair: no screening then
    const bool use_screen = sweep_uses_screen(dtype, family, D, K, Rp) && !ll_out && Dpp <= D + 4;   // (workspace is sized for Dpp <= D + 4)
    // the packed full-triangle statistics are what the tensor-core statistics kernel produces
    const bool tc_stats = use_tc && stat && !hard && canon && tc_stats_supported(dtype, D, F);
    const bool tc_fstats = tc_stats && tc_fstats_supported(dtype, D, F);     // feature form (folded triangle) for D > 64
    const bool tc_sstats = tc_stats && tc_sstats_supported(dtype, D, F);     // feature form, whole triangle in one accumulator, D <= 21
    const bool pair_stats_list = tc_stats && use_screen && g_tc_mode != 4 && pair_stats_supported(dtype, D, F);
    // ... and the log-normalisers too: no pass over the (K, chunk) scratch after the refinement on such chunks
    const bool list_softmax = pair_stats_list && !lse_out;
    // dense tensor-core chunks: the E-step kernel forms the log-normalisers itself (online log-sum-exp in its epilogue) and
    // the statistics pre-pass takes r = exp(a - lse_n) from the log-joints -- no softmax pass over the (K, chunk) scratch.
    // Not when the caller wants the (K, N) responsibilities, labels, or the screened path's refined scratch is normalised
    // by the softmax kernel.
    const bool fused_lse = use_tc && !hard && !ll_out && (!stat || tc_stats) && (!use_screen || list_softmax);
    float* lse_chunk = nullptr;
    if (use_tc) {
        tc_ops_ws = ws; ws += a256(tc_operand_workspace(K, Rp, D));
        if (sweep_uses_screen(dtype, family, D, K, Rp)) {
            screen_ws = ws; ws += a256(tc_screen_workspace(C, K));
            screen_ops_ws = ws; ws += a256(tc_screen_operand_workspace(K, Rp, D + 4, D));
        }
        if (!hard) { lse_chunk = (float*)ws; ws += a256((size_t)C * 4); tc_stat_ws = ws; }
        int rc = tc_data_scale((const float*)Z, N, D, ldz, tc_ops_ws, st, absmax_hint);
        if (rc) return rc;
        rc = tc_prepare_operands((const float*)op_a, (const float*)cst, K, Rp, Dpp, D, tc_ops_ws, st);
        if (rc) return rc;
        if (use_screen) {
            rc = tc_screen_prepare((const float*)Z, N, D, ldz, (const float*)op_a, (const float*)cst, K, Rp, Dpp, tc_ops_ws, screen_ops_ws, st);
            if (rc) return rc;
            rc = tc_screen_begin(C, K, Rp, g_tc_mode == 5 ? 1 : 0, screen_ws, st);
            if (rc) return rc;
        }
        if (tc_stats && !tc_sstats) { rc = tc_fstats ? tc_fstats_begin(C, K, tc_stat_ws, st) : tc_stats_begin(C, K, tc_stat_ws, st); if (rc) return rc; }
    }

    // optional per-phase device timing (bench.py's roofline leg): events on the launching stream
    std::vector<cudaEvent_t> ev, evk;
    auto mark = [&]() { if (phase_ms) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); ev.push_back(e); } };
    struct Range { explicit Range(const char* n) { nvtxRangePushA(n); } ~Range() { nvtxRangePop(); } };
    Range sweep_range(hard ? "mimo_sweep (Gibbs)" : "mimo_sweep (mean field)");
    for (int64_t n0 = 0; n0 < N; n0 += C) {
        const int64_t nc = (N - n0 < C) ? (N - n0) : C;
        mark();
        const char* Zc = (const char*)Z + (size_t)n0 * ldz * es;
        int rc;
        nvtxRangePushA("E-step");
        if (use_screen) {
            // screening pass over all pairs (projected operands, one FP16 pass), exact guesses, candidate lists; then
            // either the exact refinement of the candidates or (device-side flag, when > 4 % of the pairs are
            // candidates) the dense 3-pass kernel
            rc = tc_screen_pass((const float*)Zc, nc, D, ldz, K, Rp, Dpp, (float*)scratch, C, tc_ops_ws, screen_ops_ws, C, screen_ws, st);
            if (rc) return rc;
            if (phase_ms) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); evk.push_back(e); }   // end of the screening kernel
            rc = tc_screen_select((const float*)Zc, D, ldz, (const float*)op_a, (const float*)cst, K, Rp, Dpp, (float*)scratch, nc, C,
                                  tc_ops_ws, screen_ops_ws, C, screen_ws, st);
            if (rc) return rc;
            rc = tc_estep_pass((const float*)Zc, nc, D, ldz, K, Rp, (float*)scratch, C, tc_ops_ws, 3, tc_screen_gate(screen_ws, C, K), 1u, nullptr, nullptr, 0, st,
                               fused_lse ? lse_chunk : nullptr, fused_lse ? lse_sum : nullptr);
            if (rc) return rc;
            rc = tc_screen_refine((const float*)Zc, D, ldz, (const float*)op_a, K, Rp, Dpp, (const float*)cst, (float*)scratch, C, C, screen_ws, st);
        }
        else if (use_tc)      rc = tc_estep((const float*)Zc, nc, D, ldz, (const float*)cst, K, Rp, (float*)scratch, C, tc_ops_ws, st,
                                            fused_lse ? (lse_out ? (float*)lse_out + n0 : lse_chunk) : nullptr, fused_lse ? lse_sum : nullptr);
        else if (family == 0) rc = loglik_quad(dtype, Zc, nc, D, ldz, op_a, cst, K, Rp, Dpp, scratch, C, st);
        else if (diag_tc) {
            rc = tc_diag_chunk((const float*)Zc, nc, D, ldz, K, diag_fused ? nullptr : (float*)scratch, C,
                               diag_fused ? lab_all + n0 : nullptr, uniforms ? (const double*)uniforms + n0 : nullptr, seed,
                               point_offset + (uint64_t)n0, (diag_fused && lse_out) ? (float*)lse_out + n0 : nullptr,
                               diag_fused ? lse_sum : nullptr, diag_ws, st);
            if (rc) return rc;
            rc = loglik_diag(dtype, Zc, nc, D, ldz, op_a, op_b, cst, K, scratch, C, st, tc_diag_gate(diag_ws), 1u);
        }
        else             rc = loglik_diag(dtype, Zc, nc, D, ldz, op_a, op_b, cst, K, scratch, C, st);
        nvtxRangePop();
        if (rc) return rc;
        mark();
        int flags = (lse_sum ? MIMO_ACC_LSE : 0) | (lse_out ? MIMO_WRITE_LSE : 0)
                  | (hard ? MIMO_DRAW_LABELS : MIMO_WRITE_RESP);
        int32_t* lab = hard ? lab_all + n0 : nullptr;
        const double* uni = uniforms ? (const double*)uniforms + n0 : nullptr;
        void* lse_c = lse_out ? (char*)lse_out + (size_t)n0 * es : nullptr;
        if (list_softmax) {
            rc = tc_screen_lse((const float*)scratch, K, nc, C, lse_sum, C, screen_ws, st);
            if (rc) return rc;
        }
        if (!fused_lse) {
            rc = softmax(dtype, scratch, K, nc, C, flags, lse_c, uni, seed, point_offset + (uint64_t)n0, lab, lse_sum, st,
                         list_softmax ? tc_screen_gate(screen_ws, C, K) : (diag_fused ? tc_diag_gate(diag_ws) : nullptr), 1u);
            if (rc) return rc;
        }
        const float* lse_f = fused_lse ? (lse_out ? (const float*)lse_out + n0 : lse_chunk) : nullptr;   // scratch holds log-joints
        mark();
        if (ll_out)
            MIMO_CUDA(cudaMemcpy2DAsync((char*)ll_out + (size_t)n0 * es, (size_t)ldo * es, scratch, (size_t)C * es,
                                        (size_t)nc * es, K, cudaMemcpyDeviceToDevice, st));
        if (tc_stats) {
            // behind the screened E-step the candidate lists carry the whole statistic (every other pair has r < e^-40):
            // the pair-list kernel runs when the chunk was refined, the dense tensor-core kernels when it fell back
            const unsigned int* sgate = pair_stats_list ? tc_screen_gate(screen_ws, C, K) : nullptr;
            if (sgate) {
                for (int which = 0; which < 2; ++which) {                  // the guesses, then the other candidates
                    const int32_t *perm, *offsets, *slabs;
                    tc_screen_lists(screen_ws, C, K, which, &perm, &offsets, &slabs);
                    rc = pair_stats((const float*)Zc, D, ldz, perm, offsets, slabs, K, (const float*)scratch, C,
                                    list_softmax ? tc_screen_lse_values(screen_ws, C, K) : nullptr, sgate, 0u, stat, F, st);
                    if (rc) return rc;
                }
            }
            rc = tc_sstats ? tc_sstats_chunk((const float*)Zc, nc, D, ldz, (const float*)scratch, C, lse_f, K, F, tc_maxbits(tc_ops_ws), stat, st, sgate, 1u)
               : tc_fstats ? tc_fstats_chunk((const float*)Zc, nc, D, ldz, (const float*)scratch, C, K, tc_maxbits(tc_ops_ws), C, tc_stat_ws, st, sgate, 1u, lse_f)
                           : tc_stats_chunk((const float*)Zc, nc, D, ldz, (const float*)scratch, C, K, F, tc_maxbits(tc_ops_ws), stat, C, tc_stat_ws, st, sgate, 1u, lse_f);
            if (rc) return rc;
        } else if (stat && !hard) {
            const unsigned int* rgate = nullptr;
            if (resp_list) {                                   // pairs with r >= e^-40, grouped by component
                rc = resp_list_build((const float*)scratch, K, nc, C, D, C, resp_ws, st);
                if (rc) return rc;
                rgate = resp_list_gate(resp_ws, C, K);
                const int32_t *perm, *offsets, *slabs;
                resp_list_get(resp_ws, C, K, &perm, &offsets, &slabs);
                rc = pair_stats((const float*)Zc, D, ldz, perm, offsets, slabs, K, (const float*)scratch, C, nullptr, rgate, 0u, stat, F, st);
                if (rc) return rc;
            }
            rc = stats_soft(dtype, Zc, nc, D, ldz, scratch, C, K, fi, fj, F, stat, st, rgate, 1u);
            if (rc) return rc;
        }
        mark();
    }
    if (tc_stats && !tc_sstats) {
        int rc = tc_fstats ? tc_fstats_end(C, K, D, F, tc_maxbits(tc_ops_ws), stat, tc_stat_ws, st)
                           : tc_stats_end(C, K, D, F, tc_maxbits(tc_ops_ws), stat, tc_stat_ws, st);
        if (rc) return rc;
    }
    // Gibbs: ONE counting sort + segmented FP64 reduction over all N labels (re-reads Z once)
    cudaEvent_t h0 = nullptr, h1 = nullptr;
    if (stat && hard) {
        if (phase_ms) { cudaEventCreate(&h0); cudaEventCreate(&h1); cudaEventRecord(h0, st); }
        int rc = stats_hard(dtype, Z, N, D, ldz, lab_all, K, fi, fj, F, stat, hard_ws, hard_ws_bytes, false, st);
        if (rc) return rc;
        if (phase_ms) cudaEventRecord(h1, st);
    }
    if (phase_ms) {
        MIMO_CUDA(cudaStreamSynchronize(st));
        for (size_t i = 0; i + 3 < ev.size(); i += 4) {
            for (int ph = 0; ph < 3; ++ph) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ev[i + ph], ev[i + ph + 1]);
                phase_ms[ph] += ms;
                if (ph == 0 && evk.empty()) phase_ms[5] += ms;            // the E-step phase is one kernel
            }
            if (!evk.empty()) {                                           // the screening kernel alone
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ev[i], evk[i / 4]);
                phase_ms[5] += ms;
            }
            // kernel launches of this chunk: E-step (+ offsets blocks for the CTA-pair kernel), softmax,
            // statistics (feature form: data image + responsibility image + GEMM)
            phase_ms[3] += 2.0 + (use_screen ? 10.0 : 0.0)
                         + ((hard || !stat) ? 0.0 : (tc_fstats ? 3.0 : 1.0)) + (pair_stats_list ? 2.0 : 0.0) + (list_softmax ? 4.0 : 0.0) + (resp_list ? 4.0 : 0.0);
        }
        if (h0) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, h0, h1);
            phase_ms[2] += ms;
            phase_ms[3] += 4.0;
            cudaEventDestroy(h0); cudaEventDestroy(h1);
        }
        if (use_tc) phase_ms[3] += 3.0 + (tc_stats ? 1.0 : 0.0) + (use_screen ? 6.0 : 0.0);   // data scale, operand image + offsets, norms, statistics fold
        phase_ms[4] += (double)((N + C - 1) / C);                      // point chunks
        for (auto e : ev) cudaEventDestroy(e);
        for (auto e : evk) cudaEventDestroy(e);
    }
    return MIMO_OK;
}

// Device buffers of mimo_sweep_host, kept between calls (a fresh cudaMalloc / cudaFree of tens of GB per call costs
// 30 ms at best and hundreds when the driver has to reclaim memory); mimo_sweep_host_release() frees them.
struct HostSweepCache {
    static constexpr int SLOTS = 11;
    void* ptr[SLOTS] = {};
    size_t bytes[SLOTS] = {};
    int device = -1;
    void release() {
        for (int i = 0; i < SLOTS; ++i) { if (ptr[i]) cudaFree(ptr[i]); ptr[i] = nullptr; bytes[i] = 0; }
    }
    cudaError_t get(int slot, size_t need, void** out) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev != device) { release(); device = dev; }
        need = std::max<size_t>(need, 16);
        if (bytes[slot] < need) {
            if (ptr[slot]) cudaFree(ptr[slot]);
            ptr[slot] = nullptr; bytes[slot] = 0;
            cudaError_t e = cudaMalloc(&ptr[slot], need);
            if (e != cudaSuccess) return e;
            bytes[slot] = need;
        }
        *out = ptr[slot];
        return cudaSuccess;
    }
};
static HostSweepCache g_host_cache;
void sweep_host_release() { g_host_cache.release(); }

static int64_t g_host_segment = 0;
int64_t sweep_host_set_segment(int64_t points) { int64_t old = g_host_segment; g_host_segment = points < 0 ? 0 : points; return old; }

// Host-buffer variant: what bench.py's `e2e` leg and a caller without device memory use.
// The data is uploaded in segments of whole point chunks on a copy stream while the sweep of the previous segment
// runs on the compute stream (statistics, the lower-bound term and the labels are additive / per point, so a sweep
// over segments is the same sweep): end to end the call costs max(upload, compute), not their sum.
int sweep_host(int dtype, int family, int hard, const void* Z_host, int64_t N, int D,
               const void* op_a_host, const void* op_b_host, const void* cst_host, int K, int Rp, int Dpp,
               const int32_t* fi_host, const int32_t* fj_host, int F,
               const void* uniforms_host, uint64_t seed,
               double* stat_host, double* lse_sum_host, int32_t* labels_host) {
    MIMO_CHECK_ARG(Z_host && op_a_host && cst_host && fi_host && fj_host && stat_host, "null pointer");
    MIMO_CHECK_ARG(N >= 0 && D >= 1 && K >= 1 && F >= 1, "shape");
    const size_t es = dtype == MIMO_F32 ? 4 : 8;
    struct HintGuard { HintGuard(int v) { g_tables_hint = v; } ~HintGuard() { g_tables_hint = -1; } }
        hint(canonical_host(fi_host, fj_host, F, D) ? 1 : 0);
    const size_t zb = (size_t)N * D * es;
    const size_t ab = (family == 0 ? (size_t)K * Rp * Dpp : (size_t)K * D) * es;
    const int64_t C = sweep_chunk_points(dtype, family, N, D, K, Rp);
    const int64_t want = g_host_segment > 0 ? g_host_segment : (int64_t)(((size_t)1 << 29) / ((size_t)D * es));   // ~0.5 GB of data
    const int64_t seg = (want + C - 1) / C * C;                                                               // whole chunks
    const int n_seg = N > 0 ? (int)((N + seg - 1) / seg) : 0;
    const size_t wsb = sweep_workspace(dtype, family, hard, std::min<int64_t>(N, seg), D, K, Rp);
    char *dZ = nullptr, *dA = nullptr, *dB = nullptr, *dC = nullptr, *dws = nullptr;
    int32_t *dfi = nullptr, *dfj = nullptr, *dlab = nullptr;
    double *dstat = nullptr, *dlse = nullptr, *duni = nullptr;
    cudaStream_t st = nullptr, sc = nullptr;
    std::vector<cudaEvent_t> landed;
    int rc = MIMO_OK;
    const bool dbg = getenv("MIMO_HOST_DEBUG") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = now();
    auto body = [&]() -> int {
        MIMO_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        MIMO_CUDA(cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking));
        MIMO_CUDA(g_host_cache.get(0, zb, (void**)&dZ));
        MIMO_CUDA(g_host_cache.get(1, ab, (void**)&dA));
        MIMO_CUDA(g_host_cache.get(2, (size_t)K * es, (void**)&dC));
        MIMO_CUDA(g_host_cache.get(3, (size_t)F * 4, (void**)&dfi));
        MIMO_CUDA(g_host_cache.get(4, (size_t)F * 4, (void**)&dfj));
        MIMO_CUDA(g_host_cache.get(5, (size_t)K * F * 8, (void**)&dstat));
        MIMO_CUDA(g_host_cache.get(6, 8, (void**)&dlse));
        MIMO_CUDA(g_host_cache.get(7, wsb, (void**)&dws));
        if (family == 1) { MIMO_CUDA(g_host_cache.get(8, ab, (void**)&dB)); MIMO_CUDA(cudaMemcpyAsync(dB, op_b_host, ab, cudaMemcpyHostToDevice, st)); }
        if (hard) MIMO_CUDA(g_host_cache.get(9, (size_t)N * 4, (void**)&dlab));
        if (uniforms_host) MIMO_CUDA(g_host_cache.get(10, (size_t)N * 8, (void**)&duni));
        MIMO_CUDA(cudaMemcpyAsync(dA, op_a_host, ab, cudaMemcpyHostToDevice, st));
        MIMO_CUDA(cudaMemcpyAsync(dC, cst_host, (size_t)K * es, cudaMemcpyHostToDevice, st));
        MIMO_CUDA(cudaMemcpyAsync(dfi, fi_host, (size_t)F * 4, cudaMemcpyHostToDevice, st));
        MIMO_CUDA(cudaMemcpyAsync(dfj, fj_host, (size_t)F * 4, cudaMemcpyHostToDevice, st));
        MIMO_CUDA(cudaMemsetAsync(dstat, 0, (size_t)K * F * 8, st));
        MIMO_CUDA(cudaMemsetAsync(dlse, 0, 8, st));
        if (dbg) fprintf(stderr, "mimo_sweep_host: %d segment(s) of %lld points, workspace %.2f GB, allocations %.1f ms\n", n_seg, (long long)seg, wsb / 1e9, now() - t_start);
        // uploads: one segment after the other on the copy stream, an event per segment
        landed.resize(n_seg);
        for (int s = 0; s < n_seg; ++s) {
            const int64_t n0 = (int64_t)s * seg, ns = std::min<int64_t>(seg, N - n0);
            MIMO_CUDA(cudaMemcpyAsync(dZ + (size_t)n0 * D * es, (const char*)Z_host + (size_t)n0 * D * es, (size_t)ns * D * es,
                                      cudaMemcpyHostToDevice, sc));
            if (uniforms_host)
                MIMO_CUDA(cudaMemcpyAsync(duni + n0, (const double*)uniforms_host + n0, (size_t)ns * 8, cudaMemcpyHostToDevice, sc));
            MIMO_CUDA(cudaEventCreateWithFlags(&landed[s], cudaEventDisableTiming));
            MIMO_CUDA(cudaEventRecord(landed[s], sc));
        }
        for (int s = 0; s < n_seg; ++s) {
            const int64_t n0 = (int64_t)s * seg, ns = std::min<int64_t>(seg, N - n0);
            MIMO_CUDA(cudaStreamWaitEvent(st, landed[s], 0));
            int r = sweep(dtype, family, hard, dZ + (size_t)n0 * D * es, ns, D, D, dA, dB, dC, K, Rp, Dpp, dfi, dfj, F,
                          duni ? duni + n0 : nullptr, seed, (uint64_t)n0, dstat, dlse, dlab ? dlab + n0 : nullptr,
                          nullptr, nullptr, 0, dws, wsb, st, nullptr);
            if (r) return r;
        }
        MIMO_CUDA(cudaMemcpyAsync(stat_host, dstat, (size_t)K * F * 8, cudaMemcpyDeviceToHost, st));
        if (lse_sum_host) MIMO_CUDA(cudaMemcpyAsync(lse_sum_host, dlse, 8, cudaMemcpyDeviceToHost, st));
        if (hard && labels_host && N > 0) MIMO_CUDA(cudaMemcpyAsync(labels_host, dlab, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
        if (dbg) fprintf(stderr, "mimo_sweep_host: everything enqueued at %.1f ms\n", now() - t_start);
        MIMO_CUDA(cudaStreamSynchronize(st));
        if (dbg) fprintf(stderr, "mimo_sweep_host: device done at %.1f ms\n", now() - t_start);
        return MIMO_OK;
    };
    rc = body();
    if (rc != MIMO_OK) cudaDeviceSynchronize();
    tc_screen_forget();                                   // its counters live in a workspace the next call may replace
    for (auto e : landed) if (e) cudaEventDestroy(e);
    if (rc != MIMO_OK) g_host_cache.release();
    if (st) cudaStreamDestroy(st);
    if (sc) cudaStreamDestroy(sc);
    if (dbg) fprintf(stderr, "mimo_sweep_host: buffers freed at %.1f ms\n", now() - t_start);
    return rc;
}

}  // namespace mimo

namespace mimo {

// stand-alone tensor-core statistics (mimo_stats_soft_tc): data scale + one chunk + reduce
size_t stats_soft_tc_workspace(int64_t N, int K) { return 2048 + std::max(tc_stats_workspace(N, K), tc_fstats_workspace(N, K)); }

int stats_soft_tc(const void* Z, int64_t N, int D, int64_t ldz, const void* resp, int64_t ldr, int K, int F,
                  double* stat, void* ws, size_t ws_bytes, cudaStream_t st) {
    MIMO_CHECK_ARG(Z && resp && stat && ws, "null pointer");
    MIMO_CHECK_ARG(N >= 0 && K >= 1 && ldz >= D && ldr >= N, "shape");
    if (!tc_stats_supported(MIMO_F32, D, F)) { set_error("tensor-core statistics: unsupported shape D=%d F=%d", D, F); return MIMO_EUNSUPPORTED; }
    MIMO_CHECK_ARG(ws_bytes >= stats_soft_tc_workspace(N, K), "workspace too small");
    if (N == 0) return MIMO_OK;
    int rc = tc_data_scale((const float*)Z, N, D, ldz, ws, st);
    if (rc) return rc;
    void* pws = (char*)ws + 2048;
    if (tc_sstats_supported(MIMO_F32, D, F))
        return tc_sstats_chunk((const float*)Z, N, D, ldz, (const float*)resp, ldr, nullptr, K, F, tc_maxbits(ws), stat, st);
    if (tc_fstats_supported(MIMO_F32, D, F)) {
        rc = tc_fstats_begin(N, K, pws, st);
        if (rc) return rc;
        rc = tc_fstats_chunk((const float*)Z, N, D, ldz, (const float*)resp, ldr, K, tc_maxbits(ws), N, pws, st);
        if (rc) return rc;
        return tc_fstats_end(N, K, D, F, tc_maxbits(ws), stat, pws, st);
    }
    rc = tc_stats_begin(N, K, pws, st);
    if (rc) return rc;
    rc = tc_stats_chunk((const float*)Z, N, D, ldz, (const float*)resp, ldr, K, F, tc_maxbits(ws), stat, N, pws, st);
    if (rc) return rc;
    return tc_stats_end(N, K, D, F, tc_maxbits(ws), stat, pws, st);
}

}  // namespace mimo
