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
// Work unit = (128-component block, 4 folded rows = 512 TMEM columns, slab of points), handed
// out by an atomic counter, slab-major so that concurrent CTAs stream the same points through L2.
// Per 64-point block: 256 producer threads write the A operand (R in the 3xFP16 split,
// tc_common.cuh) and, per 2 folded rows, one B stage of 256 x 64 products in the same split;
// one thread issues 4 x 3 tcgen05.mma (M=128, N=256, K=16) per stage.  Accumulators are FP32 in
// TMEM; every `flush` blocks they are drained with tcgen05.ld and added in FP64 (red.global) to
// the partial buffer [component block][slot][component lane], which tc_fstats_end folds into
// the packed (K, F) statistics once per sweep.
#include <algorithm>
#include "tc_common.cuh"
#include "internal.h"

namespace mimo {

using namespace tc;

constexpr int TF_THREADS = 384;                  // 2 producer / drain warpgroups + 1 warpgroup holding the MMA warp
constexpr int TF_KB = 64;                        // points per block (one 128-byte operand row)
constexpr int TF_DC = 128;                       // columns of the folded triangle
constexpr int TF_H = TF_DC / 2;
constexpr int TF_ROWS = TF_H + 2;                // folded rows
constexpr int TF_FBROWS = 4;                     // folded rows per unit (4 x 128 = 512 TMEM columns)
constexpr uint32_t TF_ATILE = 16384;             // [128 components][64 points] FP16
constexpr uint32_t TF_BTILE = 32768;             // [256 slots][64 points] FP16
constexpr uint32_t TF_BSTAGE = 2 * TF_BTILE;     // hi | lo
constexpr uint32_t TF_ZTILE = 16384;             // [128 columns][64 points] FP16, 16-byte chunks XOR-swizzled by row
constexpr float TF_ONE = 128.f;                  // the constant 1 of zt in scaled units
constexpr float TF_RSCALE = 8192.f;              // responsibilities in [0, 1] -> [0, 2^13]

constexpr uint32_t TF_OFF_A = 2 * TF_BSTAGE;
constexpr uint32_t TF_OFF_ZS = TF_OFF_A + 4 * TF_ATILE;
constexpr uint32_t TF_OFF_BARS = TF_OFF_ZS + 2 * TF_ZTILE;

struct TfBars {
    uint64_t a_full[2], a_empty[2], b_full[2], b_empty[2];
    uint64_t acc_ready, acc_drained;
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

// byte offset of the 16-byte chunk `c` (8 points) of column `row` in the swizzled point block
__device__ __forceinline__ uint32_t tf_z_off(int row, int c) { return (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4); }

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

__global__ void __launch_bounds__(TF_THREADS, 1)
tc_fstats_kernel(const float* __restrict__ Z, int64_t N, int D, int64_t ldz,
                 const float* __restrict__ R, int64_t ldr, int K, int rvec4,
                 const unsigned int* __restrict__ maxbits,
                 double* __restrict__ partial,
                 int cbs, int fbs, int slabs, int64_t slab_points, int flush_kb) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem;
    unsigned char* sA = smem + TF_OFF_A;
    unsigned char* sZh = smem + TF_OFF_ZS;
    unsigned char* sZl = sZh + TF_ZTILE;
    TfBars* bars = reinterpret_cast<TfBars*>(smem + TF_OFF_BARS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->a_full[b], 256); mbar_init(&bars->a_empty[b], 1);
            mbar_init(&bars->b_full[b], 256); mbar_init(&bars->b_empty[b], 1);
        }
        mbar_init(&bars->acc_ready, 1);
        mbar_init(&bars->acc_drained, 256);
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(&bars->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const int n_units = cbs * fbs * slabs;

    // pipeline counters (identical sequences in the producer and MMA roles)
    uint32_t ac = 0, bc = 0, dc = 0;

    // Static schedule: unit u = (slab, feature block, component block) -> CTA u % gridDim.  Every CTA of a slab
    // streams the same points at the same pace, so the point block and the responsibility rows shared by
    // the units of a component block are read from HBM once and served from L2 afterwards.
    // The two roles run their own copy of the unit loop so that each gets its own register budget.
    if (warp < 8) {
        // ================= producers / drain: 232 registers (four split 32-point columns live in registers) =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const float sz = tf_scale(__uint_as_float(__ldg(maxbits)));
        const int t = tid & 127, ph = tid >> 7;          // own column, 32-point half of the block
        const int ca = tid >> 1, hh = tid & 1;           // A operand: component row, 32-point half
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int slab = u / (cbs * fbs);
            const int rem = u - slab * (cbs * fbs);
            const int fb = rem / cbs, cb = rem - fb * cbs;
            const int k0 = cb * 128;
            const int row0 = fb * TF_FBROWS;
            const int nst = (min(TF_FBROWS, TF_ROWS - row0) + 1) / 2;        // B stages (2 folded rows each) per block
            const int64_t p0 = (int64_t)slab * slab_points;
            const int64_t p1 = min(N, p0 + slab_points);
            const int nkb = (p1 > p0) ? (int)((p1 - p0 + TF_KB - 1) / TF_KB) : 0;
            if (nkb == 0) continue;

            uint4 zh[4], zl[4], mh[4], ml[4];            // own column t and mirror column 127 - t: 32 points, hi / lo halves
            float zn[32], rn[32];                        // next block, in flight
            auto load_z = [&](int kb) {
                const int64_t nb = p0 + (int64_t)kb * TF_KB + ph * 32;
                if (t < D && nb + 32 <= p1) {
                    const float* src = Z + nb * ldz + t;
#pragma unroll
                    for (int p = 0; p < 32; ++p) zn[p] = __ldg(src + (int64_t)p * ldz);
                } else {
#pragma unroll
                    for (int p = 0; p < 32; ++p) zn[p] = (nb + p < p1 && t < D) ? __ldg(Z + (nb + p) * ldz + t) : 0.f;
                }
            };
            auto load_r = [&](int kb) {
                const int64_t nb = p0 + (int64_t)kb * TF_KB + hh * 32;
                const bool kok = k0 + ca < K;
                const float* src = R + (int64_t)(k0 + ca) * ldr + nb;
                if (rvec4 && kok && nb + 32 <= p1) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + q);
                        rn[4 * q] = v.x; rn[4 * q + 1] = v.y; rn[4 * q + 2] = v.z; rn[4 * q + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int p = 0; p < 32; ++p) rn[p] = (kok && nb + p < p1) ? __ldg(src + p) : 0.f;
                }
            };
            load_z(0);
            load_r(0);
            for (int kb = 0; kb < nkb; ++kb) {
                // ---- the point block: own column -> split FP16 in registers + shared copy; mirror column back ----
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float x[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) x[e] = zn[8 * q + e] * sz;
                    split8(x, zh[q], zl[q]);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");          // broadcast reads of the previous block are done
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t o = tf_z_off(t, ph * 4 + q);
                    *reinterpret_cast<uint4*>(sZh + o) = zh[q];
                    *reinterpret_cast<uint4*>(sZl + o) = zl[q];
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t o = tf_z_off(TF_DC - 1 - t, ph * 4 + q);
                    mh[q] = *reinterpret_cast<const uint4*>(sZh + o);
                    ml[q] = *reinterpret_cast<const uint4*>(sZl + o);
                }
                // ---- A operand: responsibilities of 128 components x 64 points, 3xFP16 split ----
                {
                    const uint32_t buf = ac & 1;
                    mbar_wait(&bars->a_empty[buf], ((ac >> 1) & 1) ^ 1);
                    unsigned char* ah = sA + (size_t)buf * 2 * TF_ATILE;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float x[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) x[e] = rn[8 * q + e] * TF_RSCALE;
                        uint4 hi, lo;
                        split8(x, hi, lo);
                        const uint32_t o = sw128_chunk_off(ca, hh * 4 + q);
                        *reinterpret_cast<uint4*>(ah + o) = hi;
                        *reinterpret_cast<uint4*>(ah + TF_ATILE + o) = lo;
                    }
                    fence_proxy_async();
                    mbar_arrive(&bars->a_full[buf]);
                    ++ac;
                }
                if (kb + 1 < nkb) { load_z(kb + 1); load_r(kb + 1); }   // in flight while this block's products are generated
                // ---- B stages: 2 folded rows = 256 slots x 64 points of products ----
                for (int s = 0; s < nst; ++s, ++bc) {
                    const uint32_t bst = bc & 1;
                    mbar_wait(&bars->b_empty[bst], ((bc >> 1) & 1) ^ 1);
                    unsigned char* stage = sB + (size_t)bst * TF_BSTAGE;
#pragma unroll 1
                    for (int rr = 0; rr < 2; ++rr) {
                        const int r = row0 + 2 * s + rr;
                        const int rowN = rr * 128 + t;
                        unsigned char* dst = stage + (uint32_t)rowN * 128u;
                        const int sw = rowN & 7;
                        // products of the broadcast column `irow` with the thread's own / mirror column
                        auto emit = [&](const uint4 (&oh)[4], const uint4 (&ol)[4], int irow) {
                            uint4 ah[4], al[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint32_t o = tf_z_off(irow, ph * 4 + q);
                                ah[q] = *reinterpret_cast<const uint4*>(sZh + o);
                                al[q] = *reinterpret_cast<const uint4*>(sZl + o);
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                uint4 hi, lo;
                                tf_prod8(ah[q], al[q], oh[q], ol[q], hi, lo);
                                const uint32_t o = (uint32_t)(((ph * 4 + q) ^ sw) << 4);
                                *reinterpret_cast<uint4*>(dst + o) = hi;
                                *reinterpret_cast<uint4*>(dst + TF_BTILE + o) = lo;
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
                            for (int q = 0; q < 4; ++q) {
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
                                const uint32_t o = (uint32_t)(((ph * 4 + q) ^ sw) << 4);
                                *reinterpret_cast<uint4*>(dst + o) = hi;
                                *reinterpret_cast<uint4*>(dst + TF_BTILE + o) = lo;
                            }
                        }
                    }
                    fence_proxy_async();
                    mbar_arrive(&bars->b_full[bst]);
                }
                // ---- drain the FP32 accumulators into the FP64 partials ----
                if ((kb + 1) % flush_kb == 0 || kb + 1 == nkb) {
                    mbar_wait(&bars->acc_ready, dc & 1);
                    tc_fence_after();
                    const int qd = warp & 3, half = warp >> 2;
                    const int ncol = nst * 128;                          // columns per warp half
                    double* pk = partial + ((size_t)cb * TF_ROWS * TF_DC + (size_t)row0 * TF_DC + (size_t)half * ncol) * 128
                               + qd * 32 + lane;
                    const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + half * ncol;
#pragma unroll 1
                    for (int c0 = 0; c0 < ncol; c0 += 32) {
                        float v[32];
                        tmem_ld32(taddr + c0, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (v[j] != 0.f) red_add_f64(pk + (size_t)(c0 + j) * 128, (double)v[j]);
                    }
                    tc_fence_before();
                    mbar_arrive(&bars->acc_drained);
                    ++dc;
                }
            }
        }
    } else {
        // ================= MMA warpgroup: one issuing thread, three idle warps =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 8 && lane == 0) {
            const uint32_t idesc = make_idesc_f16(128, 256);
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int slab = u / (cbs * fbs);
                const int rem = u - slab * (cbs * fbs);
                const int fb = rem / cbs;
                const int row0 = fb * TF_FBROWS;
                const int nst = (min(TF_FBROWS, TF_ROWS - row0) + 1) / 2;
                const int64_t p0 = (int64_t)slab * slab_points;
                const int64_t p1 = min(N, p0 + slab_points);
                const int nkb = (p1 > p0) ? (int)((p1 - p0 + TF_KB - 1) / TF_KB) : 0;
                bool fresh = true;
                for (int kb = 0; kb < nkb; ++kb) {
                    const uint32_t abuf = ac & 1;
                    mbar_wait(&bars->a_full[abuf], (ac >> 1) & 1);
                    tc_fence_after();
                    const uint64_t ah = make_desc_sw128(a0 + abuf * 2 * TF_ATILE);
                    const uint64_t al = make_desc_sw128(a0 + abuf * 2 * TF_ATILE + TF_ATILE);
                    for (int s = 0; s < nst; ++s, ++bc) {
                        const uint32_t bst = bc & 1;
                        mbar_wait(&bars->b_full[bst], (bc >> 1) & 1);
                        tc_fence_after();
                        const uint64_t bh = make_desc_sw128(b0 + bst * TF_BSTAGE);
                        const uint64_t bl = make_desc_sw128(b0 + bst * TF_BSTAGE + TF_BTILE);
                        const uint32_t d = tmem_base + s * 256;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            umma_f16(d, al + 2 * kk, bh + 2 * kk, idesc, (fresh && kk == 0) ? 0u : 1u);
                            umma_f16(d, ah + 2 * kk, bl + 2 * kk, idesc, 1);
                            umma_f16(d, ah + 2 * kk, bh + 2 * kk, idesc, 1);
                        }
                        umma_commit(&bars->b_empty[bst]);
                    }
                    umma_commit(&bars->a_empty[abuf]);
                    ++ac;
                    fresh = false;
                    if ((kb + 1) % flush_kb == 0 || kb + 1 == nkb) {
                        umma_commit(&bars->acc_ready);
                        mbar_wait(&bars->acc_drained, dc & 1);     // accumulators read out: may be overwritten
                        tc_fence_after();
                        ++dc;
                        fresh = true;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 512);
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

bool tc_fstats_supported(int dtype, int D, int F) {
    return dtype == MIMO_F32 && D > TF_H && D <= TF_DC && F == (D + 1) * (D + 2) / 2;
}

static size_t tf_partial_bytes(int K) { return (size_t)((K + 127) / 128) * TF_ROWS * TF_DC * 128 * sizeof(double); }

// [reserved (1 KB) | partial]
size_t tc_fstats_workspace(int K) { return 2048 + tf_partial_bytes(K); }

int tc_fstats_begin(int K, void* ws, cudaStream_t st) {
    MIMO_CUDA(cudaMemsetAsync(align1k(ws), 0, 1024 + tf_partial_bytes(K), st));
    return MIMO_OK;
}

static int g_flush_tiles_f = 16;
void tc_fstats_set_flush_tiles(int t) { g_flush_tiles_f = t < 1 ? 1 : t; }

// one chunk of points: accumulates into the partial buffer
int tc_fstats_chunk(const float* Z, int64_t N, int D, int64_t ldz, const float* R, int64_t ldr, int K,
                    const unsigned int* maxbits, void* ws, cudaStream_t st) {
    if (N == 0) return MIMO_OK;
    double* partial = (double*)(align1k(ws) + 1024);
    const int cbs = (K + 127) / 128, fbs = (TF_ROWS + TF_FBROWS - 1) / TF_FBROWS;
    // point slabs only when one (component block, feature block) grid does not fill the SMs
    int slabs = std::max(1, sm_count() / (cbs * fbs));
    const int64_t blocks = (N + TF_KB - 1) / TF_KB;
    slabs = (int)std::min<int64_t>(slabs, std::max<int64_t>(1, blocks / 8));
    const int64_t slab_points = (blocks + slabs - 1) / slabs * TF_KB;
    MIMO_CUDA(cudaFuncSetAttribute(tc_fstats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TF_SMEM));
    const int grid = std::min(cbs * fbs * slabs, sm_count());
    const int rvec4 = (ldr % 4 == 0) && (((uintptr_t)R & 15) == 0);
    tc_fstats_kernel<<<grid, TF_THREADS, TF_SMEM, st>>>(Z, N, D, ldz, R, ldr, K, rvec4, maxbits, partial,
                                                        cbs, fbs, slabs, slab_points, 2 * g_flush_tiles_f);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

int tc_fstats_end(int K, int D, int F, const unsigned int* maxbits, double* stat, void* ws, cudaStream_t st) {
    tc_fstats_reduce_kernel<<<K, 256, 0, st>>>((const double*)(align1k(ws) + 1024), K, D, F, maxbits, stat);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

}  // namespace mimo
