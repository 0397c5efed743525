// Tensor-core E-step + label draw for the diagonal-covariance family, sm_100a (tcgen05, CTA pairs).
//
//   a[k][n] = cst[k] - 0.5 * sum_j (S[k][j] z[n][j] - T[k][j])^2           (include/mimo_b200.h, diagonal operand form)
//
// replaces distributions/gaussian.py:837-850 (GaussianWithDiagonalPrecision.log_likelihood), bayesian.py:446-460
// (the Normal-Gamma expectation) and, for Gibbs sweeps, mixtures/gmm.py:72-75 + utils/stats.py:8-21 (label draw) for
// FP32 data with D <= 64 and K <= 256 (cfg3 of BASELINE.json: N = 100M, d = 64, K = 256).
//
// The log-density is linear in the features [z', z'^2] of the centred point z' = z - mu0:
//   a = c_k + sum_j L_kj z'_j + sum_j Q_kj z'_j^2,   L = S T',  Q = -S^2 / 2,  T' = T - S mu0,  c_k = cst_k - |T'_k|^2 / 2
// i.e. ONE GEMM (points x 2D features) . (2D features x K components) with a per-point epilogue.  Centring at the mean
// of the component centres keeps the cancelling constant |T'_k|^2 / 2 small; the prepare step measures it and raises a
// device-side flag when it is too large for FP32 accumulation (max_k |T'_k|^2 > TD_GUARD): the CUDA-core kernels of
// estep.cu, which form (S z - T)^2 term by term, then run instead (every kernel of both branches is launched and returns
// at once when it is not its turn -- no host round trip).
//
// Structure (the CTA-pair protocol of tc_estep2.cu): clusters of 2 CTAs walk tiles of 2 x 128 points.  The 8 converter /
// epilogue warps write the tile's features in the 3 x FP16 split as the K-major swizzled A operand (double-buffered:
// the next tile is converted while this one is in the tensor pipe), the weights of all K <= 256 components stay resident
// in shared memory (each CTA holds its 128 rows, loaded once), the leader's issuing thread runs 8 K steps x 3 passes of
// tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16) into one of two 256-column accumulators.  Epilogue: thread =
// point (TMEM lane) x column quarter: the thread pulls its 64 accumulator columns into registers, releases the
// accumulator, and keeps e = exp(a - block max) of every component in those registers together with (max, sum) per
// 32-component block; the four quarters of a point meet in shared memory, which gives the log-normaliser and the
// position u * sum of the point's uniform in the cumulative sum; the quarter that holds the crossing walks its
// registers to the label.  One read of the accumulator per pair, no (K, chunk) scratch, no softmax kernel.
#include <algorithm>
#include "tc_common.cuh"
#include "internal.h"

namespace mimo {

using namespace tc;

constexpr int TD_EPI = 512;                   // 16 converter / epilogue warps
constexpr int TD_THREADS = TD_EPI;             // no extra warps: 16 warps = 4 per scheduler leave 128 registers per thread; thread 0 also loads the
                                              // weights and issues the MMAs (leader CTA) / relays the peer's barriers
constexpr uint32_t TD_TILE = 16384;           // 128 rows x 64 FP16
constexpr uint32_t TD_ABUF = 4 * TD_TILE;     // [hi linear | hi quadratic | lo linear | lo quadratic]
constexpr uint32_t TD_STAGE = 2 * TD_TILE;    // one K block of this CTA's 128 components: hi | lo
constexpr int TD_KMAX = 256;
constexpr float TD_GUARD = 4096.f;            // largest |T'_k|^2 the GEMM form is trusted with (error ~ 1e-7 x this)
// parameter block (floats): [0] s1, [1] s2, [2] max_k |T'_k|^2 (bits, atomicMax), [3] gate (uint: 1 = take the CUDA-core kernels), [4 .. 67] mu0

struct TdBars {
    uint64_t b_full[2], peer_b_full[2], c_full;
    uint64_t a_full[2], peer_a_full[2];
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};
constexpr uint32_t TD_OFF_B = TD_ABUF;                         // A0 | B (kb0, kb1) | A1
constexpr uint32_t TD_OFF_A1 = TD_ABUF + 2 * TD_STAGE;
constexpr uint32_t TD_OFF_C = TD_OFF_A1 + TD_ABUF;             // constants: 256 x (c_k, 1 / g_k)
constexpr uint32_t TD_OFF_X = TD_OFF_C + TD_KMAX * 8;          // block statistics: [tile parity][half][block 4][point 128] (max, sum)
constexpr uint32_t TD_OFF_U = TD_OFF_X + 2 * 2 * 4 * 128 * 8;   // the points' uniforms: [tile parity][point 128] floats
constexpr uint32_t TD_OFF_BARS = TD_OFF_U + 2 * 128 * 4;
constexpr uint32_t TD_SMEM = TD_OFF_BARS + sizeof(TdBars);

// ---- prepare: centre, scales, weight image, constants -----------------------------------------------------------
__global__ void td_center_kernel(const float* __restrict__ S, const float* __restrict__ T, int K, int D,
                                 const unsigned int* __restrict__ zmaxbits, float* __restrict__ prm) {
    __shared__ float mmax[64];
    const int j = threadIdx.x;                 // 64 threads
    double s = 0.0;
    if (j < D)
        for (int k = 0; k < K; ++k) { const float sk = S[(size_t)k * D + j]; s += sk != 0.f ? (double)T[(size_t)k * D + j] / sk : 0.0; }
    const float mu = (float)(s / K);
    prm[4 + j] = j < D ? mu : 0.f;
    mmax[j] = j < D ? fabsf(mu) : 0.f;
    __syncthreads();
    if (j == 0) {
        float m = 0.f;
        for (int i = 0; i < 64; ++i) m = fmaxf(m, mmax[i]);
        const float zm = __uint_as_float(*zmaxbits) + m;                  // >= max |z - mu0|
        prm[0] = pow2_scale_for(zm);                                       // |z'| s1 < 2^14
        int e = 0;
        if (zm > 0.f && zm < 3.0e38f) frexpf(zm, &e);
        e = 7 - e; e = e > 40 ? 40 : (e < -40 ? -40 : e);
        prm[1] = ldexpf(1.f, e);                                           // (|z'| s2)^2 < 2^14
        prm[2] = 0.f;
        reinterpret_cast<unsigned int*>(prm)[3] = 0u;
    }
}

// grid = 256 component slots, block = 64 (thread = dimension).  img [kb][rank][hi|lo][128][64] FP16 (SW128),
// consts [256] (c_k, 1/g_k) x log2(e): the epilogue works on log2-domain values (one FADD + one MUFU.EX2 per exp)
__global__ void __launch_bounds__(64)
td_prep_kernel(const float* __restrict__ S, const float* __restrict__ T, const float* __restrict__ cst, int K, int D,
               float* __restrict__ prm, unsigned char* __restrict__ img, float2* __restrict__ consts) {
    __shared__ double red[3][64];
    const int k = blockIdx.x, j = threadIdx.x;
    const int rank = k >> 7, row = k & 127;
    const bool on = k < K && j < D;
    double L = 0.0, Q = 0.0, t2 = 0.0;
    if (on) {
        const double s = S[(size_t)k * D + j], tp = (double)T[(size_t)k * D + j] - s * (double)prm[4 + j];
        L = s * tp; Q = -0.5 * s * s; t2 = tp * tp;
    }
    red[0][j] = fabs(L); red[1][j] = fabs(Q); red[2][j] = t2;
    __syncthreads();
    for (int o = 32; o > 0; o >>= 1) {
        if (j < o) { red[0][j] = fmax(red[0][j], red[0][j + o]); red[1][j] = fmax(red[1][j], red[1][j + o]); red[2][j] += red[2][j + o]; }
        __syncthreads();
    }
    const float s1 = prm[0], s2 = prm[1];
    // one power-of-two scale g per component: weights L g / s1 and Q g / s2^2 both below 2^14
    const float gl = pow2_scale_for((float)red[0][0] / s1), gq = pow2_scale_for((float)red[1][0] / (s2 * s2));
    const float g = fminf(gl, gq);
    __half lh, ll, qh, ql;
    split_f16(on ? (float)(L * (double)(g / s1)) : 0.f, lh, ll);
    split_f16(on ? (float)(Q * (double)(g / (s2 * s2))) : 0.f, qh, ql);
    const uint32_t o = sw128_chunk_off(row, j >> 3) + (j & 7) * 2;
    unsigned char* lin = img + (size_t)(0 * 2 + rank) * TD_STAGE + o;
    unsigned char* quad = img + (size_t)(1 * 2 + rank) * TD_STAGE + o;
    *reinterpret_cast<__half*>(lin) = lh;  *reinterpret_cast<__half*>(lin + TD_TILE) = ll;
    *reinterpret_cast<__half*>(quad) = qh; *reinterpret_cast<__half*>(quad + TD_TILE) = ql;
    if (j == 0) {
        if (k < K) {
            consts[k] = make_float2((float)(((double)cst[k] - 0.5 * red[2][0]) * 1.4426950408889634), (float)(1.4426950408889634 / (double)g));   // log2 domain
            atomicMax(reinterpret_cast<unsigned int*>(prm) + 2, __float_as_uint((float)red[2][0]));
        } else {
            consts[k] = make_float2(-INFINITY, 0.f);
        }
    }
}

__global__ void td_gate_kernel(float* __restrict__ prm) {
    reinterpret_cast<unsigned int*>(prm)[3] = (prm[2] > TD_GUARD) ? 1u : 0u;
}

// ---- main kernel -------------------------------------------------------------------------------------------------
// 16 converter / epilogue warps: warp w reads TMEM lanes 32 (w % 4) .. + 31 (its points) and the column quarter w / 4
// (64 components = two 32-component blocks).  With 8 warps the epilogue ran at 1.2 warp instructions per clock (two
// warps per scheduler, dependent chains): profiles/r02_tc_diag_kernel.md.  The 24 MMAs of a tile are issued by thread 0
// right after the tile's features are stored, one iteration before its epilogue.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TD_THREADS, 1)
tc_diag_kernel(const float* __restrict__ Z, int64_t N, int D, int64_t ldz, int vec4,
               const unsigned char* __restrict__ img, const float2* __restrict__ consts, const float* __restrict__ prm, int K,
               float* __restrict__ out, int64_t ldo, int32_t* __restrict__ labels, const double* __restrict__ uniforms,
               uint64_t seed, uint64_t point_offset, float* __restrict__ lse_out, double* __restrict__ lse_sum) {
    if (reinterpret_cast<const unsigned int*>(prm)[3] != 0u) return;      // cancellation guard: the CUDA-core kernels run instead
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem + TD_OFF_B;
    const float2* sC = reinterpret_cast<const float2*>(smem + TD_OFF_C);
    float2* sX = reinterpret_cast<float2*>(smem + TD_OFF_X);
    float* sU = reinterpret_cast<float*>(smem + TD_OFF_U);
    TdBars* bars = reinterpret_cast<TdBars*>(smem + TD_OFF_BARS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int64_t n_tiles = (N + 255) / 256;
    const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->b_full[b], 1); mbar_init(&bars->peer_b_full[b], 1);
            mbar_init(&bars->a_full[b], TD_EPI); mbar_init(&bars->peer_a_full[b], 1);
            mbar_init(&bars->tmem_full[b], 1); mbar_init(&bars->tmem_empty[b], 2 * TD_EPI / 32);   // leader's: one arrival per epilogue warp of BOTH CTAs
        }
        mbar_init(&bars->c_full, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (warp == 0) tmem_alloc2(&bars->tmem_base, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    {
        // ================= converter + epilogue warps =================
        const float s1 = __ldg(prm), s2 = __ldg(prm + 1);
        const int j4 = (lane & 15) * 4, sub = lane >> 4;                 // this lane's 4 dimensions, row parity
        float mu[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) mu[e] = __ldg(prm + 4 + j4 + e);
        const int qtr = warp >> 2, qd = warp & 3;
        const int prow = qd * 32 + lane;                                 // point row inside the tile = TMEM lane
        const uint32_t lane_base = tmem_base + ((uint32_t)(qd * 32) << 16);

        // a tile's rows reach the A operand in two steps, a whole tile apart: the global loads are issued one iteration
        // ahead (their latency hides behind that iteration's epilogue), the feature tiles are written the iteration after
        float4 z[4];
        auto load_rows = [&](int64_t tile) {
            const int64_t n0 = tile * 256 + rank * 128;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = warp * 8 + 2 * i + sub;
                const int64_t n = n0 + r;
                z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tile < n_tiles && n < N && j4 < D) {
                    const float* src = Z + n * ldz + j4;
                    if (vec4 && j4 + 3 < D) z[i] = __ldg(reinterpret_cast<const float4*>(src));
                    else {
                        z[i].x = __ldg(src);
                        if (j4 + 1 < D) z[i].y = __ldg(src + 1);
                        if (j4 + 2 < D) z[i].z = __ldg(src + 2);
                        if (j4 + 3 < D) z[i].w = __ldg(src + 3);
                    }
                }
            }
        };
        auto store_features = [&](uint32_t buf) {
            unsigned char* sA = smem + (buf ? TD_OFF_A1 : 0u);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = warp * 8 + 2 * i + sub;
                // (rows beyond N carry -mu: their results are never stored)
                const bool on = j4 < D;
                const float zz[4] = {on ? z[i].x - mu[0] : 0.f, (on && j4 + 1 < D) ? z[i].y - mu[1] : 0.f,
                                     (on && j4 + 2 < D) ? z[i].z - mu[2] : 0.f, (on && j4 + 3 < D) ? z[i].w - mu[3] : 0.f};
                float f1[4], f2[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) { f1[e] = zz[e] * s1; const float t = zz[e] * s2; f2[e] = t * t; }
                uint2 h1, l1, h2, l2;
                split4(f1, h1, l1);
                split4(f2, h2, l2);
                unsigned char* base = sA + sw128_chunk_off(r, j4 >> 3) + (j4 & 7) * 2;
                *reinterpret_cast<uint2*>(base) = h1;
                *reinterpret_cast<uint2*>(base + TD_TILE) = h2;
                *reinterpret_cast<uint2*>(base + 2 * TD_TILE) = l1;
                *reinterpret_cast<uint2*>(base + 3 * TD_TILE) = l2;
            }
            fence_proxy_async();
            mbar_arrive(&bars->a_full[buf]);
        };

        // thread 0: the MMAs of local tile i (leader CTA), or the peer's "features stored" event forwarded to the leader
        auto issue = [&](uint32_t i) {
            const uint32_t buf = i & 1, par = (i >> 1) & 1;
            mbar_wait(&bars->a_full[buf], par);
            if (rank != 0) { mbar_arrive_remote(map_to_rank(smem_u32(&bars->peer_a_full[buf]), 0)); return; }
            mbar_wait_cluster(&bars->peer_a_full[buf], par);
            mbar_wait_cluster(&bars->tmem_empty[buf], par ^ 1);
            tc_fence_after();
            const uint32_t idesc = make_idesc_f16(256, 256);
            const int S = (D + 15) >> 4;                                  // 16-wide K steps that hold data, per feature block
            const uint32_t a0 = smem_u32(smem + (buf ? TD_OFF_A1 : 0u)), b0 = smem_u32(sB);
            const uint32_t d = tmem_base + buf * 256;
            bool first = true;
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
                const uint64_t ah = make_desc_sw128(a0 + kb * TD_TILE), al = make_desc_sw128(a0 + (2 + kb) * TD_TILE);
                const uint64_t bh = make_desc_sw128(b0 + kb * TD_STAGE), bl = make_desc_sw128(b0 + kb * TD_STAGE + TD_TILE);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    if (kk >= S) continue;
                    umma2_f16(d, al + 2 * kk, bh + 2 * kk, idesc, first ? 0u : 1u);
                    umma2_f16(d, ah + 2 * kk, bl + 2 * kk, idesc, 1);
                    umma2_f16(d, ah + 2 * kk, bh + 2 * kk, idesc, 1);
                    first = false;
                }
            }
            umma2_commit(&bars->tmem_full[buf]);
        };

        uint32_t it = 0;
        if (tid == 0) {
            // this CTA's 128 weight rows + the constants, once
            for (int kb = 0; kb < 2; ++kb) {
                mbar_arrive_expect_tx(&bars->b_full[kb], TD_STAGE);
                bulk_g2s(sB + (size_t)kb * TD_STAGE, img + (size_t)(kb * 2 + rank) * TD_STAGE, TD_STAGE, &bars->b_full[kb]);
            }
            mbar_arrive_expect_tx(&bars->c_full, TD_KMAX * 8);
            bulk_g2s(smem + TD_OFF_C, consts, TD_KMAX * 8, &bars->c_full);
        }
        if (cluster_id < n_tiles) { load_rows(cluster_id); store_features(0); }
        load_rows(cluster_id + n_clusters);
        if (tid == 0) {
            mbar_wait(&bars->b_full[0], 0); mbar_wait(&bars->b_full[1], 0);
            if (rank == 0) { mbar_wait_cluster(&bars->peer_b_full[0], 0); mbar_wait_cluster(&bars->peer_b_full[1], 0); }
            else { mbar_arrive_remote(map_to_rank(smem_u32(&bars->peer_b_full[0]), 0)); mbar_arrive_remote(map_to_rank(smem_u32(&bars->peer_b_full[1]), 0)); }
            if (cluster_id < n_tiles) issue(0);
        }
        __syncwarp();
        mbar_wait(&bars->c_full, 0);
        for (int64_t tile = cluster_id; tile < n_tiles; tile += n_clusters, ++it) {
            const uint32_t buf = it & 1;
            // the MMAs that read A[buf ^ 1] (tile it - 1) have completed: this thread waited for their accumulator
            if (tile + n_clusters < n_tiles) {
                store_features(buf ^ 1);
                if (tid == 0) issue(it + 1);
                __syncwarp();
            }
            load_rows(tile + 2 * n_clusters);

            const int64_t n = tile * 256 + rank * 128 + prow;
            const bool pvalid = n < N;
            mbar_wait(&bars->tmem_full[buf], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = lane_base + buf * 256 + qtr * 64;
            const float2* cc = sC + qtr * 64;
            float2* xw = sX + ((size_t)buf * 8 + 2 * qtr) * 128 + prow;
            // ---- the thread's 64 columns into registers, then the accumulator is free again ----
            float e[2][32], gs[2][4];
            tmem_ld32(taddr, e[0]);
            tmem_ld32(taddr + 32, e[1]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {                                               // the leader's issuer may reuse the buffer
                if (rank == 0) mbar_arrive(&bars->tmem_empty[buf]);
                else mbar_arrive_remote_nofence(map_to_rank(smem_u32(&bars->tmem_empty[buf]), 0));
            }
            // (leaving the second half of the read in flight while the first is worked on was measured slower: 27.8 vs
            //  26.6 ms per cfg3 sweep -- the accumulator is released later and the issuer waits for it)
            // ---- per 32-component block: log2-domain log-joints, max, e = 2^(a - max) kept in the registers, sum ----
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float4 c = *reinterpret_cast<const float4*>(cc + 32 * b + i);   // (c, 1/g) of two components, broadcast read
                    e[b][i] = fmaf(e[b][i], c.y, c.x);
                    e[b][i + 1] = fmaf(e[b][i + 1], c.w, c.z);
                    m0 = fmaxf(m0, e[b][i]); m1 = fmaxf(m1, e[b][i + 1]);
                }
                const float mb = fmaxf(m0, m1);
                if (out != nullptr && pvalid) {
                    float* outp = out + n;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int k = qtr * 64 + 32 * b + i;
                        if (k < K) outp[(int64_t)k * ldo] = e[b][i] * 0.6931471805599453f;
                    }
                }
                const float ms = (mb == -INFINITY) ? 0.f : mb;             // an all-padding block: every e is 2^-inf = 0
                // e = 2^(a - max) stays in the registers; sums of the four groups of 8 consecutive components are kept
                // for the two-level search of the label below (two add chains per group)
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
                    float s0 = 0.f, s1_ = 0.f;
#pragma unroll
                    for (int i = 8 * g4; i < 8 * g4 + 8; i += 2) {
                        e[b][i] = ex2_ftz(e[b][i] - ms);         s0 += e[b][i];
                        e[b][i + 1] = ex2_ftz(e[b][i + 1] - ms); s1_ += e[b][i + 1];
                    }
                    gs[b][g4] = s0 + s1_;
                }
                xw[b * 128] = make_float2(mb, (gs[b][0] + gs[b][1]) + (gs[b][2] + gs[b][3]));
            }
            // the point's uniform: drawn once, by the first quarter, and handed over with the block statistics
            if (qtr == 0 && labels != nullptr)
                sU[buf * 128 + prow] = pvalid ? (float)(uniforms ? uniforms[n] : philox_uniform(seed, point_offset + (uint64_t)n)) : 0.f;
            // only the four warps that share this TMEM lane quarter exchange anything: one named barrier per quarter
            asm volatile("bar.sync %0, 128;" ::"r"(1 + qd) : "memory");
            // ---- the four quarters of a point meet: log-normaliser, position of the uniform in the cumulative sum ----
            const float2* xr = sX + ((size_t)buf * 8) * 128 + prow;
            float sbs[8];
            float m = -INFINITY;
#pragma unroll
            for (int b = 0; b < 8; ++b) m = fmaxf(m, xr[b * 128].x);
            float Ssum = 0.f;
#pragma unroll
            for (int b = 0; b < 8; ++b) { const float2 t = xr[b * 128]; sbs[b] = t.y * ex2_ftz(t.x - m); Ssum += sbs[b]; }   // 2^-inf = 0: empty blocks
            if (labels != nullptr) {
                const float thr = sU[buf * 128 + prow] * Ssum;
                float before = 0.f;                                        // cumulative sum in front of this quarter
#pragma unroll
                for (int b = 0; b < 6; ++b) if (b < 2 * qtr) before += sbs[b];
                const float after = before + sbs[2 * qtr] + sbs[2 * qtr + 1];
                // exactly one quarter owns the crossing (the block sums are the same numbers in all four threads)
                if ((qtr == 0 || before < thr) && (qtr == 3 || thr <= after)) {
                    // two-level search over this quarter's 64 components (the cumulative sum is non-decreasing): first the
                    // group of 8 whose end passes the threshold, from the group sums kept above, then the 8 components of
                    // that group, picked from the registers with a select tree (register arrays have no dynamic index)
                    float scb[2];
#pragma unroll
                    for (int b = 0; b < 2; ++b) scb[b] = sbs[2 * qtr + b] > 0.f ? ex2_ftz(xr[(2 * qtr + b) * 128].x - m) : 0.f;
                    float cum = before, base = before;                     // base: cumulative sum in front of the crossing group
                    int gi = 0;                                            // groups that end below the threshold
#pragma unroll
                    for (int g8 = 0; g8 < 8; ++g8) {
                        cum = fmaf(gs[g8 >> 2][g8 & 3], scb[g8 >> 2], cum);
                        const bool below = cum < thr;
                        gi += below ? 1 : 0;
                        base = below ? cum : base;
                    }
                    int lab = qtr * 64 + 63;                               // rounding put the threshold past the last component
                    if (gi < 8) {
                        const bool b0 = (gi & 1) != 0, b1 = (gi & 2) != 0, b2 = (gi & 4) != 0;
                        const float sc = b2 ? scb[1] : scb[0];
                        float c = base;
                        int lt = 0;                                        // components of the group whose cumulative sum stays below thr
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float t0 = b0 ? e[0][8 + j] : e[0][j], t1 = b0 ? e[0][24 + j] : e[0][16 + j];
                            const float t2 = b0 ? e[1][8 + j] : e[1][j], t3 = b0 ? e[1][24 + j] : e[1][16 + j];
                            const float u0 = b1 ? t1 : t0, u1 = b1 ? t3 : t2;
                            c = fmaf(b2 ? u1 : u0, sc, c);
                            lt += (c < thr) ? 1 : 0;
                        }
                        lab = qtr * 64 + gi * 8 + min(lt, 7);
                    }
                    if (lab >= K) lab = K - 1;
                    if (pvalid) labels[n] = lab;
                }
            }
            if (qtr == 0 && (lse_out != nullptr || lse_sum != nullptr)) {
                const float lse = (m + __log2f(Ssum)) * 0.6931471805599453f;
                double part = pvalid ? (double)lse : 0.0;
                if (lse_out != nullptr && pvalid) lse_out[n] = lse;
                if (lse_sum != nullptr) {
#pragma unroll
                    for (int o2 = 16; o2 > 0; o2 >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o2);
                    if (lane == 0 && part != 0.0) atomicAdd(lse_sum, part);
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();                                   // no CTA leaves (or frees TMEM) while its partner may still signal it
    if (warp == 0) tmem_dealloc2(tmem_base, 512);
}

// ---- host side ---------------------------------------------------------------------------------------------------

static bool g_td_enabled = true;
int tc_diag_enable(int on) { int old = g_td_enabled; g_td_enabled = on != 0; return old; }

bool tc_diag_supported(int dtype, int D, int K) { return g_td_enabled && dtype == MIMO_F32 && D >= 8 && D <= 64 && K >= 1 && K <= TD_KMAX; }

static char* align1k_d(void* p) { return (char*)(((uintptr_t)p + 1023) / 1024 * 1024); }
// [flags 256 B (max |z| bits) | parameter block | constants | weight image], 1 KB aligned inside
size_t tc_diag_workspace() { return 2048 + 1024 + TD_KMAX * 8 + 4 * (size_t)TD_STAGE; }

struct TdLayout { unsigned int* zmax; float* prm; float2* consts; unsigned char* img; };
static TdLayout td_layout(void* ws) {
    char* base = align1k_d(ws);
    TdLayout L;
    L.zmax = (unsigned int*)base;
    L.prm = (float*)(base + 256);
    L.consts = (float2*)(base + 1024);
    L.img = (unsigned char*)(base + 1024 + 2048);
    return L;
}

const unsigned int* tc_diag_gate(void* ws) { return reinterpret_cast<const unsigned int*>(td_layout(ws).prm) + 3; }

// once per sweep: data scale, centre, weight image
int tc_diag_prepare(const float* Z, int64_t N, int D, int64_t ldz, const float* S, const float* T, const float* cst, int K,
                    void* ws, cudaStream_t st, float absmax_hint) {
    TdLayout L = td_layout(ws);
    int rc = tc_data_scale(Z, N, D, ldz, ws, st, absmax_hint);    // max |z| -> L.zmax[0]
    if (rc) return rc;
    td_center_kernel<<<1, 64, 0, st>>>(S, T, K, D, L.zmax, L.prm);
    MIMO_LAUNCH_CHECK();
    td_prep_kernel<<<TD_KMAX, 64, 0, st>>>(S, T, cst, K, D, L.prm, L.img, L.consts);
    MIMO_LAUNCH_CHECK();
    td_gate_kernel<<<1, 1, 0, st>>>(L.prm);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

// one chunk of points: log-joints (out, optional), labels (optional), log-normalisers (optional)
int tc_diag_chunk(const float* Z, int64_t N, int D, int64_t ldz, int K, float* out, int64_t ldo,
                  int32_t* labels, const double* uniforms, uint64_t seed, uint64_t point_offset,
                  float* lse_out, double* lse_sum, void* ws, cudaStream_t st) {
    if (N == 0) return MIMO_OK;
    TdLayout L = td_layout(ws);
    MIMO_CUDA(cudaFuncSetAttribute(tc_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TD_SMEM));
    const int64_t tiles = (N + 255) / 256;
    const int clusters = (int)std::min<int64_t>(tiles, sm_count() / 2);
    const int vec4 = (ldz % 4 == 0) && (((uintptr_t)Z & 15) == 0);
    tc_diag_kernel<<<2 * clusters, TD_THREADS, TD_SMEM, st>>>(Z, N, D, ldz, vec4, L.img, L.consts, L.prm, K, out, ldo, labels, uniforms,
                                                              seed, point_offset, lse_out, lse_sum);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

}  // namespace mimo
