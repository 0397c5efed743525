// Tensor-core weighted sufficient statistics in feature form for SMALL dimensions (D <= 21), sm_100a (tcgen05 / TMEM).
//
//   stat[k][f(i,j)] += sum_n r[k][n] zt[n][i] zt[n][j]      zt = [z ; 1],  j <= i,  f = i (i + 1) / 2 + j
//
// replaces distributions/gaussian.py:491-505 and lingauss.py:306-325 (the einsums 'nd,kn,nl->kdl', 'kn,nd->kd', 'kn->k')
// for FP32 data on the shapes of cfg2 (D = 9, K = 128) and cfg4 (D = 16, K = 64) of BASELINE.json.
//
// One GEMM over the points, S (components x features) = R (components x points) . Phi (points x features), with all
// F = (D + 1)(D + 2) / 2 <= 253 features of the packed triangle in ONE accumulator (128 component lanes x round16(F)
// columns of tensor memory) that stays resident while the CTA's slab of points streams through.  Per 64-point block
// the 512 producer threads (a) stage the block's data transposed in shared memory, (b) write the responsibility tile
// [128 components][64 points] -- taken as given, or formed as exp(a - lse_n) from the log-joints when the E-step kernel
// supplied the log-normalisers (tc_estep2.cu, fused softmax) -- and (c) form the feature tile [F][64 points] as FP32
// products split into FP16 hi + lo (tc_common.cuh); one thread issues 4 K steps x 3 passes of tcgen05.mma (M = 128,
// N = round16(F), K = 16).  Every `flush` blocks the accumulator is drained (tcgen05.ld) and added in FP64 to the
// packed statistics with red.global.add.f64.  Two stages of operand tiles overlap the producers with the tensor pipe.
#include <algorithm>
#include "tc_common.cuh"
#include "internal.h"

namespace mimo {

using namespace tc;

constexpr int SS_NP = 512;                      // producer / drain threads: 16 warps (8 ran the block loop at 0.9 warp instructions per clock:
                                                // latency-bound with two warps per scheduler)
constexpr int SS_NPW = SS_NP / 32;
constexpr int SS_THREADS = SS_NP + 32;          // + MMA warp
constexpr int SS_ZQ = (64 * 21 + SS_NP - 1) / SS_NP;   // data values a thread stages per block
constexpr int SS_RQ = 128 * 8 / SS_NP;          // responsibility items (component row, 8-point chunk) per thread
constexpr int SS_KB = 64;                       // points per block (one 128-byte operand row)
constexpr int SS_DMAX = 21;                     // (D + 1)(D + 2) / 2 <= 253 features
constexpr int SS_FMAX = 256;
constexpr uint32_t SS_ATILE = 16384;            // [128 components][64 points] FP16
constexpr float SS_ONE = 128.f;                 // the constant 1 of zt in scaled units
constexpr float SS_RSCALE = 8192.f;             // responsibilities in [0, 1] -> [0, 2^13]
constexpr int SS_ZLD = SS_KB + 4;               // row stride of the transposed data tile (floats)

struct SsBars {
    uint64_t full[2], empty[2];
    uint64_t acc_ready, acc_drained;
    uint32_t tmem_base;
};

__device__ __forceinline__ void ss_red_add_f64(double* p, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// scale of the data inside this kernel: max |z| * sz in [64, 128)  (so every product stays below 2^14)
__host__ __device__ __forceinline__ float ss_scale(float maxabs) { return pow2_scale_for(maxabs) * (1.f / 128.f); }

// stage layout: [A hi | A lo | B hi (Fpad rows) | B lo (Fpad rows)]
__global__ void __launch_bounds__(SS_THREADS, 1)
tc_sstats_kernel(const float* __restrict__ Z, int64_t N, int D, int64_t ldz,
                 const float* __restrict__ R, int64_t ldr, const float* __restrict__ lse, int K, int F, int Fpad,
                 const unsigned int* __restrict__ maxbits, double* __restrict__ stat,
                 int mtiles, int slabs, int64_t slab_blocks, int flush,
                 const unsigned int* __restrict__ gate, unsigned int gate_value) {
    if (gate != nullptr && __ldg(gate) != gate_value) return;
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t btile = (uint32_t)Fpad * 128u;                       // bytes of one B tile (hi or lo)
    const uint32_t stage_bytes = 2 * SS_ATILE + 2 * btile;
    float* zsT = reinterpret_cast<float*>(smem + 2 * stage_bytes);     // [D + 1][SS_ZLD] scaled data of the block, row D = the constant
    float* lse_s = zsT + (SS_DMAX + 1) * SS_ZLD;                         // [64]
    double* fscale = reinterpret_cast<double*>(lse_s + SS_KB);          // [SS_FMAX] 1 / scale of each feature
    unsigned char* fij = reinterpret_cast<unsigned char*>(fscale + SS_FMAX);   // [SS_FMAX][2] (i, j) of each feature
    SsBars* bars = reinterpret_cast<SsBars*>(fij + 2 * SS_FMAX);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float sz = ss_scale(__uint_as_float(__ldg(maxbits)));
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) { mbar_init(&bars->full[b], SS_NP); mbar_init(&bars->empty[b], 1); }
        mbar_init(&bars->acc_ready, 1);
        mbar_init(&bars->acc_drained, SS_NP);
        fence_barrier_init();
    }
    // feature tables + zero rows F .. Fpad-1 of both stages' B tiles (never written again)
    for (int f = tid; f < SS_FMAX; f += SS_THREADS) {
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= f) ++i;
        const int j = f - i * (i + 1) / 2;
        fij[2 * f] = (unsigned char)i; fij[2 * f + 1] = (unsigned char)j;
        const double dsz = (double)sz, one = (double)SS_ONE, rs = (double)SS_RSCALE;
        fscale[f] = f < F ? 1.0 / (rs * (i == D ? one : dsz) * (j == D ? one : dsz)) : 0.0;
    }
    for (int idx = tid; idx < 2 * 2 * (Fpad - F) * 8; idx += SS_THREADS) {
        const int ch = idx & 7, r = F + ((idx >> 3) % (Fpad - F)), hl = (idx >> 3) / (Fpad - F) & 1, st = (idx >> 3) / (Fpad - F) >> 1;
        *reinterpret_cast<uint4*>(smem + st * stage_bytes + 2 * SS_ATILE + hl * btile + sw128_chunk_off(r, ch)) = make_uint4(0, 0, 0, 0);
    }
    if (warp == SS_NPW) tmem_alloc(&bars->tmem_base, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const int n_units = mtiles * slabs;
    const int64_t n_blocks = (N + SS_KB - 1) / SS_KB;

    if (warp < SS_NPW) {
        // ================= producers / drain =================
        uint32_t bc = 0, dc = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int mt = u / slabs, slab = u - mt * slabs;
            const int64_t b0 = (int64_t)slab * slab_blocks, b1 = min(n_blocks, b0 + slab_blocks);
            const int k0 = mt * 128;
            const int kt = min(128, K - k0);                              // component rows of this tile; the rest stay zero
            const int rvec4 = ((ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(R) & 15) == 0);
            for (int idx = tid; idx < 2 * 2 * (128 - kt) * 8; idx += SS_NP) {
                const int ch = idx & 7, r = kt + ((idx >> 3) % (128 - kt)), hl = ((idx >> 3) / (128 - kt)) & 1, stz = ((idx >> 3) / (128 - kt)) >> 1;
                *reinterpret_cast<uint4*>(smem + stz * stage_bytes + hl * SS_ATILE + sw128_chunk_off(r, ch)) = make_uint4(0, 0, 0, 0);
            }
            // the global loads of a block are issued one block ahead (into registers) so that their latency hides behind the
            // feature products of the block before
            float zr[SS_ZQ], lr = 0.f;                                    // data values per thread, one log-normaliser
            float4 rx[SS_RQ][2];                                          // responsibility items of 8 points
            auto fetch = [&](int64_t blk) {
                const int64_t n0 = blk * SS_KB;
                const bool live = blk < b1;
#pragma unroll
                for (int q = 0; q < SS_ZQ; ++q) {
                    const int idx = tid + SS_NP * q;
                    const int p = idx / D, i = idx - p * D;
                    zr[q] = (live && idx < SS_KB * D && n0 + p < N) ? __ldg(Z + (n0 + p) * ldz + i) : 0.f;
                }
                lr = (live && lse != nullptr && tid < SS_KB && n0 + tid < N) ? __ldg(lse + n0 + tid) : 0.f;
#pragma unroll
                for (int q = 0; q < SS_RQ; ++q) {
                    const int item = tid + SS_NP * q;
                    const int ca = item >> 3, c = item & 7;
                    rx[q][0] = rx[q][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (live && item < kt * 8) {
                        const float* src = R + (int64_t)(k0 + ca) * ldr + n0;
                        if (n0 + SS_KB <= N && rvec4) {
                            rx[q][0] = __ldg(reinterpret_cast<const float4*>(src) + c);
                            rx[q][1] = __ldg(reinterpret_cast<const float4*>(src) + 8 + c);
                        } else {
                            float t[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) { const int p = (e < 4 ? 4 * c + e : 28 + 4 * c + e); t[e] = (n0 + p < N) ? __ldg(src + p) : 0.f; }
                            rx[q][0] = make_float4(t[0], t[1], t[2], t[3]);
                            rx[q][1] = make_float4(t[4], t[5], t[6], t[7]);
                        }
                    }
                }
            };
            fetch(b0);
            for (int64_t blk = b0; blk < b1; ++blk, ++bc) {
                const uint32_t st = bc & 1;
                const int64_t n0 = blk * SS_KB;
                mbar_wait(&bars->empty[st], ((bc >> 1) & 1) ^ 1);         // the MMAs that read this stage (and zsT two blocks ago) are done
                asm volatile("bar.sync 1, %0;" ::"n"(SS_NP) : "memory");   // everyone finished reading zsT / lse_s of the previous block
                // (a) the block's data, transposed and scaled; the constant row
#pragma unroll
                for (int q = 0; q < SS_ZQ; ++q) {
                    const int idx = tid + SS_NP * q;
                    if (idx < SS_KB * D) { const int p = idx / D, i = idx - p * D; zsT[i * SS_ZLD + p] = zr[q] * sz; }
                }
                if (tid < SS_KB) {
                    zsT[D * SS_ZLD + tid] = (n0 + tid < N) ? SS_ONE : 0.f;
                    lse_s[tid] = lr;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(SS_NP) : "memory");
                unsigned char* sA = smem + st * stage_bytes;
                unsigned char* sBt = sA + 2 * SS_ATILE;
                // Operand slot s = 8 c + e of the block holds point 4 c + e (e < 4) or 32 + 4 c + e - 4: both tiles use the same
                // order (the contraction does not care), and the eight lanes of a row then read 128 contiguous bytes.
                // (b) responsibilities: item = (component row, 8-slot chunk), spread over all threads whatever K is
#pragma unroll
                for (int q = 0; q < SS_RQ; ++q) {
                    const int item = tid + SS_NP * q;
                    if (item >= kt * 8) break;
                    const int ca = item >> 3, c = item & 7;
                    float x[8] = {rx[q][0].x, rx[q][0].y, rx[q][0].z, rx[q][0].w, rx[q][1].x, rx[q][1].y, rx[q][1].z, rx[q][1].w};
                    if (lse != nullptr) {
                        const float4 l0 = *reinterpret_cast<const float4*>(lse_s + 4 * c), l1 = *reinterpret_cast<const float4*>(lse_s + 32 + 4 * c);
                        const float ll[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
                        for (int e = 0; e < 8; ++e) { const int p = (e < 4 ? 4 * c + e : 28 + 4 * c + e); x[e] = (n0 + p < N) ? fast_exp(x[e] - ll[e]) : 0.f; }
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) x[e] *= SS_RSCALE;
                    uint4 hi, lo;
                    split8(x, hi, lo);
                    const uint32_t o = sw128_chunk_off(ca, c);
                    *reinterpret_cast<uint4*>(sA + o) = hi;
                    *reinterpret_cast<uint4*>(sA + SS_ATILE + o) = lo;
                }
                fetch(blk + 1);                                           // in flight while the features are formed
                // (c) features: item = (feature f, 8-slot chunk c)   (loading a thread's items together before multiplying
                // them was measured slower: 2.8 vs 2.0 ms at cfg4)
                for (int item = tid; item < F * 8; item += SS_NP) {
                    const int f = item >> 3, c = item & 7;
                    const int i = fij[2 * f], j = fij[2 * f + 1];
                    const float4 a0 = *reinterpret_cast<const float4*>(zsT + i * SS_ZLD + 4 * c), a1 = *reinterpret_cast<const float4*>(zsT + i * SS_ZLD + 32 + 4 * c);
                    const float4 c0 = *reinterpret_cast<const float4*>(zsT + j * SS_ZLD + 4 * c), c1 = *reinterpret_cast<const float4*>(zsT + j * SS_ZLD + 32 + 4 * c);
                    const float x[8] = {a0.x * c0.x, a0.y * c0.y, a0.z * c0.z, a0.w * c0.w, a1.x * c1.x, a1.y * c1.y, a1.z * c1.z, a1.w * c1.w};
                    uint4 hi, lo;
                    split8(x, hi, lo);
                    const uint32_t o = sw128_chunk_off(f, c);
                    *reinterpret_cast<uint4*>(sBt + o) = hi;
                    *reinterpret_cast<uint4*>(sBt + btile + o) = lo;
                }
                fence_proxy_async();
                mbar_arrive(&bars->full[st]);
                // ---- drain the FP32 accumulator into the FP64 statistics ----
                const int64_t done = blk - b0 + 1;
                if (done % flush == 0 || blk + 1 == b1) {
                    mbar_wait(&bars->acc_ready, dc & 1);
                    tc_fence_after();
                    const int qd = warp & 3, cg = warp >> 2;
                    const int k = k0 + qd * 32 + lane;
                    const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16);
                    for (int c16 = cg; c16 < Fpad / 16; c16 += SS_NPW / 4) {
                        float v[16];
                        tmem_ld16(taddr + 16 * c16, v);
                        tmem_ld_wait();
                        if (k < K) {
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                const int f = 16 * c16 + e;
                                if (f < F && v[e] != 0.f) ss_red_add_f64(stat + (size_t)k * F + f, (double)v[e] * fscale[f]);
                            }
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(&bars->acc_drained);
                    ++dc;
                }
            }
        }
    } else if (lane == 0) {
        // ================= MMA issuer (one thread) =================
        const uint32_t idesc = make_idesc_f16(128, Fpad);
        uint32_t bc = 0, dc = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int mt = u / slabs, slab = u - mt * slabs;
            const int64_t b0 = (int64_t)slab * slab_blocks, b1 = min(n_blocks, b0 + slab_blocks);
            (void)mt;
            bool fresh = true;
            for (int64_t blk = b0; blk < b1; ++blk, ++bc) {
                const uint32_t st = bc & 1;
                mbar_wait(&bars->full[st], (bc >> 1) & 1);
                tc_fence_after();
                const uint32_t a0 = smem_u32(smem + st * stage_bytes);
                const uint64_t ah = make_desc_sw128(a0), al = make_desc_sw128(a0 + SS_ATILE);
                const uint64_t bh = make_desc_sw128(a0 + 2 * SS_ATILE), bl = make_desc_sw128(a0 + 2 * SS_ATILE + btile);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    umma_f16(tmem_base, al + 2 * kk, bh + 2 * kk, idesc, (fresh && kk == 0) ? 0u : 1u);
                    umma_f16(tmem_base, ah + 2 * kk, bl + 2 * kk, idesc, 1);
                    umma_f16(tmem_base, ah + 2 * kk, bh + 2 * kk, idesc, 1);
                }
                umma_commit(&bars->empty[st]);
                fresh = false;
                const int64_t done = blk - b0 + 1;
                if (done % flush == 0 || blk + 1 == b1) {
                    umma_commit(&bars->acc_ready);
                    mbar_wait(&bars->acc_drained, dc & 1);               // read out: may be overwritten
                    tc_fence_after();
                    ++dc;
                    fresh = true;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == SS_NPW) tmem_dealloc(tmem_base, 256);
}

// ---- host side ---------------------------------------------------------------------------------------------------

static bool g_ss_enabled = true;
int tc_sstats_enable(int on) { int old = g_ss_enabled; g_ss_enabled = on != 0; return old; }

bool tc_sstats_supported(int dtype, int D, int F) {
    return g_ss_enabled && dtype == MIMO_F32 && D >= 1 && D <= SS_DMAX && F == (D + 1) * (D + 2) / 2;
}

static int g_ss_flush = 64;                    // 64-point blocks accumulated in FP32 between FP64 drains

// one chunk of points; accumulates straight into stat (K, F).  R: responsibilities, or log-joints when lse != nullptr
int tc_sstats_chunk(const float* Z, int64_t N, int D, int64_t ldz, const float* R, int64_t ldr, const float* lse, int K, int F,
                    const unsigned int* maxbits, double* stat, cudaStream_t st, const unsigned int* gate, unsigned int gate_value) {
    if (N == 0) return MIMO_OK;
    const int Fpad = (F + 15) / 16 * 16;
    const size_t smem = 2 * (2 * (size_t)SS_ATILE + 2 * (size_t)Fpad * 128) + (size_t)(SS_DMAX + 1) * SS_ZLD * 4 + SS_KB * 4
                      + SS_FMAX * 8 + 2 * SS_FMAX + sizeof(SsBars) + 16;
    MIMO_CUDA(cudaFuncSetAttribute(tc_sstats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int mtiles = (K + 127) / 128;
    const int64_t blocks = (N + SS_KB - 1) / SS_KB;
    const int sms = sm_count();
    int slabs = (int)std::min<int64_t>(blocks, std::max(1, sms / mtiles));
    const int64_t slab_blocks = (blocks + slabs - 1) / slabs;
    slabs = (int)((blocks + slab_blocks - 1) / slab_blocks);
    const int grid = std::min(mtiles * slabs, sms);
    tc_sstats_kernel<<<grid, SS_THREADS, smem, st>>>(Z, N, D, ldz, R, ldr, lse, K, F, Fpad, maxbits, stat, mtiles, slabs, slab_blocks,
                                                     g_ss_flush, gate, gate_value);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

}  // namespace mimo
