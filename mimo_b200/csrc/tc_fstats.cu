// Tensor-core weighted sufficient statistics in FEATURE form, sm_100a (tcgen05 / TMEM).
//
//   stat[k][(i,j)] = sum_n r[k][n] * zt[n][i] * zt[n][j]        zt = [z ; 1],  j <= i
//
// replaces distributions/gaussian.py:491-505 and lingauss.py:306-325 (the einsums
// 'nd,kn,nl->kdl', 'kn,nd->kd', 'kn->k') for FP32 data with 64 < D <= 128.
//
// One GEMM over the points:  S (components x features) = R (components x points) . Phi (points x features)
// where Phi holds the products z_i z_j of the LOWER TRIANGLE only (half the flops of the
// per-component X^T diag(r_k) X form) and is generated on the fly in shared memory, once per
// 64-point block for 128 components, never stored in HBM.
//
// Folded triangle.  The triangle rows i = 0..127 are paired (i, 126 - i) so that every "folded
// row" has exactly 128 slots; thread t of a row computes (i_hi, t) when t <= i_hi and
// (i_lo, 127 - t) otherwise.  Each thread therefore only ever needs its own two columns of the
// point block, z_t and z_{127-t}, which it keeps in registers; the other factor z_i is a
// warp-wide broadcast read of shared memory.  66 folded rows x 128 slots = 8448 slots hold the
// 8256 + 128 + 1 features of [z ; 1] (rows 63 and 127 stand alone, row 65 is z_t * 1, and the
// spare slot of row 64 holds 1 * 1).
//
// The GEMM runs on CTA PAIRS (tcgen05 cta_group::2: one 256 x 256 x 16 MMA per pair; see tc_estep2.cu for
// the pair protocol): a single-CTA 128 x 128 x 16 version was bound by the issue rate of its one MMA thread
// and by the products its 512 producer threads could form.  In a pair each CTA owns 128 components (its half
// of M) and forms every second folded row (its half of N), so per MMA flop the producers do half the work.
// Work unit = (pair of 128-component blocks, 4 folded rows = 512 TMEM columns[, slab of points]) on a static
// lock-step schedule.  Two small pre-passes write the operands in their shared-memory image
// (K-major, 128-byte swizzle, 3xFP16 split of tc_common.cuh): the responsibilities
// [component block][64-point block][hi|lo][128][64] and the transposed data
// [64-point block][hi|lo][128 columns][64], so the main kernel receives both with one bulk copy
// (TMA engine) per block.  Per 64-point block: 512 producer threads form, per folded row, one B
// stage of 128 slots x 64 points of products directly in split FP16 (Dekker product on the FP16
// FMA pipe); the leader's issuing thread runs 4 x 3 tcgen05.mma (M=256, N=256, K=16) per pair of rows.  Accumulators
// are FP32 in TMEM; every `flush` blocks they are drained with tcgen05.ld and added in FP64
// (red.global) to the partial buffer [component block][slot][component lane], which
// tc_fstats_end folds into the packed (K, F) statistics once per sweep.
#include <algorithm>
#include <cmath>
#include "tc_common.cuh"
#include "internal.h"

namespace mimo {

using namespace tc;

constexpr int TF_PRODUCERS = 512;                // 16 producer / drain warps (4 warpgroups)
constexpr int TF_THREADS = 640;                  // + 1 warpgroup: MMA warp, bulk-copy warp, 2 idle warps
constexpr int TF_KB = 64;                        // points per block (one 128-byte operand row)
constexpr int TF_DC = 128;                       // columns of the folded triangle
constexpr int TF_H = TF_DC / 2;
constexpr int TF_ROWS = TF_H + 2;                // folded rows
constexpr int TF_FBROWS = 4;                     // folded rows per unit (4 x 128 = 512 TMEM columns)
constexpr int TF_BSTAGES = 3;
constexpr uint32_t TF_TILE = 16384;              // [128 rows][64 points] FP16
constexpr uint32_t TF_PAIR = 2 * TF_TILE;        // hi | lo
constexpr float TF_ONE = 128.f;                  // the constant 1 of zt in scaled units
constexpr float TF_RSCALE = 8192.f;              // responsibilities in [0, 1] -> [0, 2^13]

constexpr uint32_t TF_OFF_A = TF_BSTAGES * TF_PAIR;
constexpr uint32_t TF_OFF_ZS = TF_OFF_A + 2 * TF_PAIR;
constexpr uint32_t TF_OFF_BARS = TF_OFF_ZS + 2 * TF_PAIR;

struct TfBars {
    uint64_t zs_full[2], zs_empty[2], a_full[2], a_empty[2];
    uint64_t b_full[TF_BSTAGES], b_empty[TF_BSTAGES];
    uint64_t acc_ready, acc_drained;
    uint64_t peer_a_full[2], peer_b_full[TF_BSTAGES], peer_acc_drained;   // leader CTA only: events forwarded by the peer's relay
    uint32_t tmem_base;
};
constexpr uint32_t TF_SMEM = TF_OFF_BARS + sizeof(TfBars);

// folded row r, thread t -> feature (i, j) of zt (the constant has index D); i < 0: unused slot
__host__ __device__ inline void tf_slot_pair(int D, int r, int t, int& i, int& j) {
    if (r < TF_H - 1) {
        const int ih = TF_H + r;
        if (t <= ih) { i = ih; j = t; } else { i = TF_DC - 2 - ih; j = TF_DC - 1 - t; }
    } else if (r == TF_H - 1) { i = TF_DC - 1; j = t; }
    else if (r == TF_H) {
        if (t < TF_H) { i = TF_H - 1; j = t; }
        else if (t == TF_H) { i = D; j = D; return; }
        else { i = -1; j = -1; return; }
    } else { i = D; j = t; if (t >= D) i = -1; return; }
    if (i >= D || j >= D) { i = -1; j = -1; }
}

__device__ __forceinline__ void red_add_f64(double* p, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// scale of the data inside this kernel: max |z| * sz in [64, 128)
__host__ __device__ __forceinline__ float tf_scale(float maxabs) { return pow2_scale_for(maxabs) * (1.f / 128.f); }

// 8 products a*b of split FP16 operands -> split FP16 result (Dekker product on the FP16 FMA pipe):
//   p = fl(ah*bh);  lo = fl(fl((ah*bh - p) + ah*bl) + al*bh)      (ah*bh - p is exact)
__device__ __forceinline__ void tf_prod8(const uint4& ah, const uint4& al, const uint4& bh, const uint4& bl, uint4& hi, uint4& lo) {
    const __half2* a_h = reinterpret_cast<const __half2*>(&ah);
    const __half2* a_l = reinterpret_cast<const __half2*>(&al);
    const __half2* b_h = reinterpret_cast<const __half2*>(&bh);
    const __half2* b_l = reinterpret_cast<const __half2*>(&bl);
    __half2* h = reinterpret_cast<__half2*>(&hi);
    __half2* l = reinterpret_cast<__half2*>(&lo);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __half2 p = __hmul2(a_h[e], b_h[e]);
        __half2 r = __hfma2(a_h[e], b_h[e], __hneg2(p));
        r = __hfma2(a_h[e], b_l[e], r);
        r = __hfma2(a_l[e], b_h[e], r);
        h[e] = p;
        l[e] = r;
    }
}

// ---- operand images ---------------------------------------------------------------------------
// data: one CTA per 64-point block; thread = (column t, 32-point half).  zimg block = [hi|lo][128 columns][64 points].
__global__ void __launch_bounds__(256)
tc_fstats_zimg_kernel(const float* __restrict__ Z, int64_t N, int D, int64_t ldz,
                      const unsigned int* __restrict__ maxbits, unsigned char* __restrict__ zimg,
                      const unsigned int* __restrict__ gate, unsigned int gate_value) {
    if (gate != nullptr && __ldg(gate) != gate_value) return;
    const float sz = tf_scale(__uint_as_float(__ldg(maxbits)));
    const int t = threadIdx.x & 127, ph = threadIdx.x >> 7;
    const int64_t nb = (int64_t)blockIdx.x * TF_KB + ph * 32;
    float zn[32];
    if (t < D && nb + 32 <= N) {
        const float* src = Z + nb * ldz + t;
#pragma unroll
        for (int p = 0; p < 32; ++p) zn[p] = __ldg(src + (int64_t)p * ldz);
    } else {
#pragma unroll
        for (int p = 0; p < 32; ++p) zn[p] = (nb + p < N && t < D) ? __ldg(Z + (nb + p) * ldz + t) : 0.f;
    }
    unsigned char* blk = zimg + (size_t)blockIdx.x * TF_PAIR;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = zn[8 * q + e] * sz;
        uint4 hi, lo;
        split8(x, hi, lo);
        const uint32_t o = sw128_chunk_off(t, ph * 4 + q);
        *reinterpret_cast<uint4*>(blk + o) = hi;
        *reinterpret_cast<uint4*>(blk + TF_TILE + o) = lo;
    }
}

// responsibilities: one CTA per (64-point block, component block); thread = (component row, 32-point half).
// rimg block (cb, kblock) = [hi|lo][128 components][64 points] at ((cb * nkb_cap + kblock) * TF_PAIR).
__global__ void __launch_bounds__(256)
tc_fstats_rimg_kernel(const float* __restrict__ R, const float* __restrict__ lse, int64_t N, int64_t ldr, int K, int rvec4, int64_t nkb_cap,
                      unsigned char* __restrict__ rimg, const unsigned int* __restrict__ gate, unsigned int gate_value) {
    if (gate != nullptr && __ldg(gate) != gate_value) return;
    const int ca = threadIdx.x >> 1, hh = threadIdx.x & 1;
    const int cb = blockIdx.y;
    const int64_t nb = (int64_t)blockIdx.x * TF_KB + hh * 32;
    const bool kok = cb * 128 + ca < K;
    const float* src = R + (int64_t)(cb * 128 + ca) * ldr + nb;
    float rn[32];
    if (rvec4 && kok && nb + 32 <= N) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src) + q);
            rn[4 * q] = v.x; rn[4 * q + 1] = v.y; rn[4 * q + 2] = v.z; rn[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int p = 0; p < 32; ++p) rn[p] = (kok && nb + p < N) ? __ldg(src + p) : 0.f;
    }
    if (lse != nullptr) {                  // R holds log-joints (fused log-normaliser of tc_estep2.cu): r = exp(a - lse_n)
#pragma unroll
        for (int p = 0; p < 32; ++p) rn[p] = (kok && nb + p < N) ? fast_exp(rn[p] - __ldg(lse + nb + p)) : 0.f;
    }
    unsigned char* blk = rimg + ((size_t)cb * nkb_cap + blockIdx.x) * TF_PAIR;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = rn[8 * q + e] * TF_RSCALE;
        uint4 hi, lo;
        split8(x, hi, lo);
        const uint32_t o = sw128_chunk_off(ca, hh * 4 + q);
        *reinterpret_cast<uint4*>(blk + o) = hi;
        *reinterpret_cast<uint4*>(blk + TF_TILE + o) = lo;
    }
}

// Diagnostics: clocks the MMA issuers spent waiting, by cause, summed over clusters (mimo_tc_fstats_stall_clocks):
// [0] A (responsibility) tile landed, [1] peer's A tile, [2] B stage formed, [3] peer's B stage, [4] accumulator drained,
// [5] total clocks of the issuer loops, [6] stages issued
__device__ unsigned long long g_tf_stall[8];

// ---- main kernel ------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TF_THREADS, 1)
tc_fstats_kernel(const unsigned char* __restrict__ zimg, const unsigned char* __restrict__ rimg, int64_t nkb_cap,
                 int64_t N, double* __restrict__ partial, unsigned int* __restrict__ pace,
                 int cbps, int fbs, int slabs, int64_t slab_points, int flush_kb, int pace_epochs,
                 const unsigned int* __restrict__ gate, unsigned int gate_value) {
    if (gate != nullptr && __ldg(gate) != gate_value) return;     // device-side choice: the pair-list statistics ran instead
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem;
    unsigned char* sA = smem + TF_OFF_A;
    unsigned char* sZ = smem + TF_OFF_ZS;
    TfBars* bars = reinterpret_cast<TfBars*>(smem + TF_OFF_BARS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->zs_full[b], 1); mbar_init(&bars->zs_empty[b], TF_PRODUCERS);
            mbar_init(&bars->a_full[b], 1);  mbar_init(&bars->a_empty[b], 1);
            mbar_init(&bars->peer_a_full[b], 1);
        }
        for (int b = 0; b < TF_BSTAGES; ++b) {
            mbar_init(&bars->b_full[b], TF_PRODUCERS); mbar_init(&bars->b_empty[b], 1); mbar_init(&bars->peer_b_full[b], 1);
        }
        mbar_init(&bars->acc_ready, 1);
        mbar_init(&bars->acc_drained, TF_PRODUCERS);
        mbar_init(&bars->peer_acc_drained, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (warp == 16) tmem_alloc2(&bars->tmem_base, 512);
    tc_fence_before();
    cluster_sync_all();                                   // barriers of both CTAs initialised, TMEM allocated
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const int n_units = cbps * fbs * slabs;

    // Static schedule: unit u = (slab, feature block, pair of component blocks) -> cluster u % n_clusters.  Every
    // cluster of a slab streams the same points at the same pace, so the images shared by the units of a slab
    // are read from HBM once and served from L2 afterwards.  All roles walk the same unit / block / row sequence.
    // Inside a unit CTA `rank` owns component block 2 cbp + rank (its 128 TMEM lanes) and forms the folded rows
    // row0 + 2 s + rank (its 128 of the 256 B rows of stage s).
#define TF_UNIT_DECODE                                                                   \
        const int slab = u / (cbps * fbs);                                               \
        const int rem = u - slab * (cbps * fbs);                                         \
        const int fb = rem / cbps, cb = 2 * (rem - fb * cbps) + rank;                    \
        const int row0 = fb * TF_FBROWS;                                                 \
        const int nst = min(TF_FBROWS, TF_ROWS - row0) / 2;                              \
        const int64_t p0 = (int64_t)slab * slab_points;                                  \
        const int64_t p1 = min(N, p0 + slab_points);                                     \
        const int nkb = (p1 > p0) ? (int)((p1 - p0 + TF_KB - 1) / TF_KB) : 0;           \
        const int64_t kblock0 = p0 / TF_KB;                                              \
        (void)cb; (void)kblock0; (void)row0;

    if (warp < 16) {
        // ================= producers / drain =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        const int t = tid & 127, pq = tid >> 7;          // own column, 16-point quarter of the block
        uint32_t zc = 0, bc = 0, dc = 0;
        for (int u = cluster_id; u < n_units; u += n_clusters) {
            TF_UNIT_DECODE
            for (int kb = 0; kb < nkb; ++kb, ++zc) {
                // ---- own column t and mirror column 127 - t of the point block (split FP16, 16 points) ----
                const uint32_t zb = zc & 1;
                const unsigned char* zh_s = sZ + zb * TF_PAIR;
                const unsigned char* zl_s = zh_s + TF_TILE;
                mbar_wait(&bars->zs_full[zb], (zc >> 1) & 1);
                uint4 zh[2], zl[2], mh[2], ml[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const uint32_t o = sw128_chunk_off(t, pq * 2 + q), om = sw128_chunk_off(TF_DC - 1 - t, pq * 2 + q);
                    zh[q] = *reinterpret_cast<const uint4*>(zh_s + o);
                    zl[q] = *reinterpret_cast<const uint4*>(zl_s + o);
                    mh[q] = *reinterpret_cast<const uint4*>(zh_s + om);
                    ml[q] = *reinterpret_cast<const uint4*>(zl_s + om);
                }
                // ---- B stages: this CTA's folded row of the stage = 128 slots x 64 points of products ----
#pragma unroll 1
                for (int st = 0; st < nst; ++st, ++bc) {
                    const uint32_t bst = bc % TF_BSTAGES;
                    mbar_wait(&bars->b_empty[bst], ((bc / TF_BSTAGES) & 1) ^ 1);
                    const int r = row0 + 2 * st + rank;
                    unsigned char* dst = sB + (size_t)bst * TF_PAIR + (uint32_t)t * 128u;
                    const int sw = t & 7;
                    // products of the broadcast column `irow` with the thread's own / mirror column
                    auto emit = [&](const uint4 (&oh)[2], const uint4 (&ol)[2], int irow) {
                        uint4 ah[2], al[2];
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const uint32_t o = sw128_chunk_off(irow, pq * 2 + q);
                            ah[q] = *reinterpret_cast<const uint4*>(zh_s + o);
                            al[q] = *reinterpret_cast<const uint4*>(zl_s + o);
                        }
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            uint4 hi, lo;
                            tf_prod8(ah[q], al[q], oh[q], ol[q], hi, lo);
                            const uint32_t o = (uint32_t)(((pq * 2 + q) ^ sw) << 4);
                            *reinterpret_cast<uint4*>(dst + o) = hi;
                            *reinterpret_cast<uint4*>(dst + TF_TILE + o) = lo;
                        }
                    };
                    if (r < TF_H) {
                        int irow;
                        bool mirror = false;
                        if (r < TF_H - 1) {
                            const int ih = TF_H + r;
                            if (t <= ih) irow = ih; else { irow = TF_DC - 2 - ih; mirror = true; }
                        } else irow = TF_DC - 1;
                        if (mirror) emit(mh, ml, irow); else emit(zh, zl, irow);
                    } else if (r == TF_H && t < TF_H) {
                        emit(zh, zl, TF_H - 1);
                    } else {
                        // constant row (1 * z_t, exact power-of-two scaling), the 1 * 1 slot, or an unused slot
                        const __half2 one2 = __float2half2_rn(TF_ONE);
                        const __half2 cc2 = __float2half2_rn((r == TF_H && t == TF_H) ? TF_ONE * TF_ONE : 0.f);
                        const bool crow = r > TF_H;
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            uint4 hi, lo;
                            __half2* h = reinterpret_cast<__half2*>(&hi);
                            __half2* l = reinterpret_cast<__half2*>(&lo);
                            const __half2* oh = reinterpret_cast<const __half2*>(&zh[q]);
                            const __half2* ol = reinterpret_cast<const __half2*>(&zl[q]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                h[e] = crow ? __hmul2(oh[e], one2) : cc2;
                                l[e] = crow ? __hmul2(ol[e], one2) : __float2half2_rn(0.f);
                            }
                            const uint32_t o = (uint32_t)(((pq * 2 + q) ^ sw) << 4);
                            *reinterpret_cast<uint4*>(dst + o) = hi;
                            *reinterpret_cast<uint4*>(dst + TF_TILE + o) = lo;
                        }
                    }
                    fence_proxy_async();
                    mbar_arrive(&bars->b_full[bst]);
                }
                mbar_arrive(&bars->zs_empty[zb]);                        // done reading this point block
                // ---- drain the FP32 accumulators (this CTA's 128 components x all columns) into the FP64 partials ----
                if ((kb + 1) % flush_kb == 0 || kb + 1 == nkb) {
                    mbar_wait(&bars->acc_ready, dc & 1);
                    tc_fence_after();
                    const int qd = warp & 3, cg = warp >> 2;             // TMEM lane quarter, 128-column group (= folded row)
                    if (cg < 2 * nst) {
                        double* pk = partial + ((size_t)cb * TF_ROWS * TF_DC + (size_t)(row0 + cg) * TF_DC) * 128 + qd * 32 + lane;
                        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + cg * 128;
#pragma unroll 1
                        for (int c0 = 0; c0 < 128; c0 += 32) {
                            float v[32];
                            tmem_ld32(taddr + c0, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (v[j] != 0.f) red_add_f64(pk + (size_t)(c0 + j) * 128, (double)v[j]);
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(&bars->acc_drained);
                    ++dc;
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == 16 && lane == 0 && rank == 0) {
            // ================= MMA issuer (leader CTA, one thread) =================
            const uint32_t idesc = make_idesc_f16(256, 256);
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
            uint32_t ac = 0, bc = 0, dc = 0;
            long long w_a = 0, w_pa = 0, w_b = 0, w_pb = 0, w_d = 0, n_st = 0;
            const long long t_begin = clock64();
            for (int u = cluster_id; u < n_units; u += n_clusters) {
                TF_UNIT_DECODE
                bool fresh = true;
                for (int kb = 0; kb < nkb; ++kb, ++ac) {
                    const uint32_t abuf = ac & 1, apar = (ac >> 1) & 1;
                    long long t0 = clock64();
                    mbar_wait(&bars->a_full[abuf], apar);
                    long long t1 = clock64();
                    mbar_wait_cluster(&bars->peer_a_full[abuf], apar);
                    w_a += t1 - t0; w_pa += clock64() - t1;
                    tc_fence_after();
                    const uint64_t ah = make_desc_sw128(a0 + abuf * TF_PAIR);
                    const uint64_t al = make_desc_sw128(a0 + abuf * TF_PAIR + TF_TILE);
                    for (int st = 0; st < nst; ++st, ++bc) {
                        const uint32_t bst = bc % TF_BSTAGES, bpar = (bc / TF_BSTAGES) & 1;
                        t0 = clock64();
                        mbar_wait(&bars->b_full[bst], bpar);
                        t1 = clock64();
                        mbar_wait_cluster(&bars->peer_b_full[bst], bpar);
                        w_b += t1 - t0; w_pb += clock64() - t1; ++n_st;
                        tc_fence_after();
                        const uint64_t bh = make_desc_sw128(b0 + bst * TF_PAIR);
                        const uint64_t bl = make_desc_sw128(b0 + bst * TF_PAIR + TF_TILE);
                        const uint32_t d = tmem_base + st * 256;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            umma2_f16(d, al + 2 * kk, bh + 2 * kk, idesc, (fresh && kk == 0) ? 0u : 1u);
                            umma2_f16(d, ah + 2 * kk, bl + 2 * kk, idesc, 1);
                            umma2_f16(d, ah + 2 * kk, bh + 2 * kk, idesc, 1);
                        }
                        umma2_commit(&bars->b_empty[bst]);
                    }
                    umma2_commit(&bars->a_empty[abuf]);
                    fresh = false;
                    if ((kb + 1) % flush_kb == 0 || kb + 1 == nkb) {
                        umma2_commit(&bars->acc_ready);
                        t0 = clock64();
                        mbar_wait(&bars->acc_drained, dc & 1);     // both CTAs' accumulators read out: may be overwritten
                        mbar_wait_cluster(&bars->peer_acc_drained, dc & 1);
                        w_d += clock64() - t0;
                        tc_fence_after();
                        ++dc;
                        fresh = true;
                    }
                }
            }
            atomicAdd(&g_tf_stall[0], (unsigned long long)w_a); atomicAdd(&g_tf_stall[1], (unsigned long long)w_pa);
            atomicAdd(&g_tf_stall[2], (unsigned long long)w_b); atomicAdd(&g_tf_stall[3], (unsigned long long)w_pb);
            atomicAdd(&g_tf_stall[4], (unsigned long long)w_d); atomicAdd(&g_tf_stall[5], (unsigned long long)(clock64() - t_begin));
            atomicAdd(&g_tf_stall[6], (unsigned long long)n_st);
        } else if (warp == 16 && lane == 0) {
            // ================= relay (peer CTA): forward local events to the leader's issuer =================
            uint32_t ac = 0, bc = 0, dc = 0;
            const uint32_t r_drained = map_to_rank(smem_u32(&bars->peer_acc_drained), 0);
            for (int u = cluster_id; u < n_units; u += n_clusters) {
                TF_UNIT_DECODE
                for (int kb = 0; kb < nkb; ++kb, ++ac) {
                    const uint32_t abuf = ac & 1;
                    mbar_wait(&bars->a_full[abuf], (ac >> 1) & 1);
                    mbar_arrive_remote(map_to_rank(smem_u32(&bars->peer_a_full[abuf]), 0));
                    for (int st = 0; st < nst; ++st, ++bc) {
                        const uint32_t bst = bc % TF_BSTAGES;
                        mbar_wait(&bars->b_full[bst], (bc / TF_BSTAGES) & 1);
                        mbar_arrive_remote(map_to_rank(smem_u32(&bars->peer_b_full[bst]), 0));
                    }
                    if ((kb + 1) % flush_kb == 0 || kb + 1 == nkb) {
                        mbar_wait(&bars->acc_drained, dc & 1);
                        mbar_arrive_remote(r_drained);
                        ++dc;
                    }
                }
            }
        } else if (warp == 17 && lane == 0) {
            // ================= operand loader (one thread per CTA, TMA engine bulk copies) =================
            // Pacing: the bulk copies are asynchronous, so nothing couples the speed of a cluster to its L2 hit
            // rate and clusters would drift apart until every image tile is fetched from HBM once per cluster.
            // Every `flush_kb` blocks (an epoch) the leader's loader announces itself on a global counter and
            // does not start epoch e before all clusters (all co-resident: one CTA per SM) have started epoch
            // e - 1; the peer's loader is tied to the leader by the shared a_empty / b_empty completions.
            uint32_t zc = 0, epoch = 0;
            for (int u = cluster_id; u < n_units; u += n_clusters) {
                TF_UNIT_DECODE
                for (int kb = 0; kb < nkb; ++kb, ++zc) {
                    if (rank == 0 && zc % (uint32_t)flush_kb == 0) {
                        const unsigned int target = epoch * (unsigned int)n_clusters;
                        while (*reinterpret_cast<volatile unsigned int*>(pace) < target) __nanosleep(100);
                        atomicAdd(pace, 1u);
                        ++epoch;
                    }
                    const uint32_t b = zc & 1, par = ((zc >> 1) & 1) ^ 1;
                    mbar_wait(&bars->zs_empty[b], par);
                    mbar_arrive_expect_tx(&bars->zs_full[b], TF_PAIR);
                    bulk_g2s(sZ + b * TF_PAIR, zimg + (size_t)(kblock0 + kb) * TF_PAIR, TF_PAIR, &bars->zs_full[b]);
                    mbar_wait(&bars->a_empty[b], par);
                    mbar_arrive_expect_tx(&bars->a_full[b], TF_PAIR);
                    bulk_g2s(sA + b * TF_PAIR, rimg + ((size_t)cb * nkb_cap + kblock0 + kb) * TF_PAIR, TF_PAIR, &bars->a_full[b]);
                }
            }
            if (rank == 0 && (int)epoch < pace_epochs) atomicAdd(pace, (unsigned int)pace_epochs - epoch);   // epochs this cluster never starts
        }
    }
#undef TF_UNIT_DECODE
    tc_fence_before();
    cluster_sync_all();                                   // no CTA leaves (or frees TMEM) while its partner may still signal it
    if (warp == 16) tmem_dealloc2(tmem_base, 512);
}

// stat[k][tri(i, j)] += scale(i, j) * partial[k / 128][slot][k % 128]
__global__ void tc_fstats_reduce_kernel(const double* __restrict__ partial, int K, int D, int F,
                                        const unsigned int* __restrict__ maxbits, double* __restrict__ stat) {
    const int k = blockIdx.x;
    const int cb = k >> 7, lanek = k & 127;
    const double sz = (double)tf_scale(__uint_as_float(*maxbits));
    const double one = (double)TF_ONE, rs = (double)TF_RSCALE;
    const double inv_dd = 1.0 / (rs * sz * sz), inv_d1 = 1.0 / (rs * sz * one), inv_11 = 1.0 / (rs * one * one);
    for (int slot = threadIdx.x; slot < TF_ROWS * TF_DC; slot += blockDim.x) {
        int i, j;
        tf_slot_pair(D, slot / TF_DC, slot % TF_DC, i, j);
        if (i < 0) continue;
        const double v = partial[((size_t)cb * TF_ROWS * TF_DC + slot) * 128 + lanek];
        const double sc = (i == D) ? ((j == D) ? inv_11 : inv_d1) : inv_dd;
        stat[(size_t)k * F + (size_t)i * (i + 1) / 2 + j] += v * sc;
    }
}

// ---- host side -------------------------------------------------------------------------------

static char* align1k(void* p) { return (char*)(((uintptr_t)p + 1023) / 1024 * 1024); }
static size_t up1k(size_t x) { return (x + 1023) / 1024 * 1024; }

bool tc_fstats_supported(int dtype, int D, int F) {
    return dtype == MIMO_F32 && D > TF_H && D <= TF_DC && F == (D + 1) * (D + 2) / 2;
}

struct TfLayout { int cbs; int64_t nkb_cap; size_t off_partial, off_zimg, off_rimg, bytes; };
static TfLayout tf_layout(int64_t chunk_points, int K) {
    TfLayout L;
    L.cbs = ((K + 127) / 128 + 1) / 2 * 2;            // component blocks, rounded up to whole CTA pairs (zero padded)
    L.nkb_cap = std::max<int64_t>(1, (chunk_points + TF_KB - 1) / TF_KB);
    size_t o = 1024;
    L.off_partial = o; o += up1k((size_t)L.cbs * TF_ROWS * TF_DC * 128 * sizeof(double));
    L.off_zimg = o;    o += (size_t)L.nkb_cap * TF_PAIR;
    L.off_rimg = o;    o += (size_t)L.cbs * L.nkb_cap * TF_PAIR;
    L.bytes = o;
    return L;
}

// [reserved (1 KB) | FP64 partials | data image of one chunk | responsibility image of one chunk]
size_t tc_fstats_workspace(int64_t chunk_points, int K) { return 1024 + tf_layout(chunk_points, K).bytes; }

int tc_fstats_begin(int64_t chunk_points, int K, void* ws, cudaStream_t st) {
    TfLayout L = tf_layout(chunk_points, K);
    MIMO_CUDA(cudaMemsetAsync(align1k(ws) + L.off_partial, 0, L.off_zimg - L.off_partial, st));
    return MIMO_OK;
}

static int g_flush_tiles_f = 64;          // 8192 points of FP32 accumulation between FP64 drains (measured: 16 -> 64 takes 9 % off the kernel,
                                          // the drain's red.global.add traffic; error of the largest entry stays below 1e-5)
void tc_fstats_set_flush_tiles(int t) { g_flush_tiles_f = t < 1 ? 64 : t; }

// one chunk of N <= plan_points points: operand images, then the GEMM; accumulates into the partial buffer
int tc_fstats_chunk(const float* Z, int64_t N, int D, int64_t ldz, const float* R, int64_t ldr, int K,
                    const unsigned int* maxbits, int64_t plan_points, void* ws, cudaStream_t st,
                    const unsigned int* gate, unsigned int gate_value, const float* lse) {
    if (N == 0) return MIMO_OK;
    TfLayout L = tf_layout(plan_points, K);
    char* base = align1k(ws);
    double* partial = (double*)(base + L.off_partial);
    unsigned char* zimg = (unsigned char*)(base + L.off_zimg);
    unsigned char* rimg = (unsigned char*)(base + L.off_rimg);
    const int64_t blocks = (N + TF_KB - 1) / TF_KB;
    MIMO_CHECK_ARG(blocks <= L.nkb_cap, "chunk larger than planned");
    const int rvec4 = (ldr % 4 == 0) && (((uintptr_t)R & 15) == 0);
    tc_fstats_zimg_kernel<<<(unsigned)blocks, 256, 0, st>>>(Z, N, D, ldz, maxbits, zimg, gate, gate_value);
    MIMO_LAUNCH_CHECK();
    tc_fstats_rimg_kernel<<<dim3((unsigned)blocks, (unsigned)L.cbs), 256, 0, st>>>(R, lse, N, ldr, K, rvec4, L.nkb_cap, rimg, gate, gate_value);
    MIMO_LAUNCH_CHECK();
    const int fbs = (TF_ROWS + TF_FBROWS - 1) / TF_FBROWS;
    const int cbps = L.cbs / 2, max_clusters = sm_count() / 2;
    // point slabs: the count (<= 16, >= 64 blocks each) that fills the clusters most evenly over whole rounds
    // of the static schedule, e.g. K = 1024: 68 units x 13 slabs = 884 = 11.95 rounds of 74 clusters
    int slabs = 1;
    {
        double best = 0.0;
        const int max_slabs = (int)std::min<int64_t>(16, std::max<int64_t>(1, blocks / 64));
        for (int sl = 1; sl <= max_slabs; ++sl) {
            const double units = (double)cbps * fbs * sl;
            const double eff = units / (std::ceil(units / max_clusters) * max_clusters);
            if (eff > best + 0.01) { best = eff; slabs = sl; }
        }
    }
    const int64_t slab_points = (blocks + slabs - 1) / slabs * TF_KB;
    MIMO_CUDA(cudaFuncSetAttribute(tc_fstats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TF_SMEM));
    const int n_units = cbps * fbs * slabs;
    const int clusters = std::min(n_units, max_clusters);
    const int grid = 2 * clusters;
    const int flush_kb = 2 * g_flush_tiles_f;
    // upper bound of the pacing epochs of one cluster: its units hold at most slab_points / 64 blocks each
    const int64_t units_per_cta = (n_units + clusters - 1) / clusters;
    const int pace_epochs = (int)((units_per_cta * (slab_points / TF_KB) + flush_kb - 1) / flush_kb) + 1;
    unsigned int* pace = (unsigned int*)base;
    MIMO_CUDA(cudaMemsetAsync(pace, 0, 4, st));
    tc_fstats_kernel<<<grid, TF_THREADS, TF_SMEM, st>>>(zimg, rimg, L.nkb_cap, N, partial, pace, cbps, fbs, slabs, slab_points,
                                                        flush_kb, pace_epochs, gate, gate_value);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

// read and reset the issuer stall counters (synchronises)
int tc_fstats_stall_clocks(unsigned long long* out_host8) {
    MIMO_CUDA(cudaDeviceSynchronize());
    MIMO_CUDA(cudaMemcpyFromSymbol(out_host8, g_tf_stall, 8 * sizeof(unsigned long long)));
    unsigned long long zero[8] = {};
    MIMO_CUDA(cudaMemcpyToSymbol(g_tf_stall, zero, sizeof(zero)));
    return MIMO_OK;
}

int tc_fstats_end(int64_t plan_points, int K, int D, int F, const unsigned int* maxbits, double* stat, void* ws, cudaStream_t st) {
    TfLayout L = tf_layout(plan_points, K);
    tc_fstats_reduce_kernel<<<K, 256, 0, st>>>((const double*)(align1k(ws) + L.off_partial), K, D, F, maxbits, stat);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

}  // namespace mimo
