// Tensor-core weighted sufficient statistics for the full-covariance family, sm_100a.
//
//   S_k = sum_n r[k][n] z_n z_n^T ,   s_k = sum_n r[k][n] z_n ,   n_k = sum_n r[k][n]
//
// replaces distributions/gaussian.py:491-505 and lingauss.py:306-325 (the einsum
// 'nd,kn,nl->kdl' and its companions) for FP32 data with D <= 128 and soft responsibilities.
//
// S_k is a GEMM per component with the points as the contraction dimension:
//   A_k = (r_k * Z)^T  (features x points),   B = Z^T  (features x points),   D = A_k B^T.
// Work unit = (group of 4 components, slab of points): the CTA keeps the four 128 x 128 FP32
// accumulators in TMEM (4 x 128 = 512 columns) while the point tiles of its slab stream
// through.  Per 128-point tile the 256 producer threads hold the tile in registers
// (thread = one feature x 64 points, so global reads are coalesced along the features and
// the transposed, K-major 128B-swizzled shared-memory rows are written with conflict-free
// 16-byte stores), write B once, and for each of the 4 components write r*z in the 3xFP16
// split (tc_common.cuh) into a double-buffered A tile; one thread issues 2 x 4 x 3
// tcgen05.mma (M=128, N=round16(D), K=16) per component.  s_k and n_k fall out of the same
// FP32 products on the CUDA cores (FP64 accumulation across tiles).
// Every `flush_tiles` tiles the accumulators are drained (tcgen05.ld) into an FP64 partial
// buffer owned by the unit's (slab, component) -- plain read-modify-write, no atomics -- which
// bounds the FP32 accumulation length; tc_stats_reduce_kernel folds the slabs into the packed
// statistics once per sweep.
#include "tc_common.cuh"
#include "internal.h"

namespace mimo {

using namespace tc;

constexpr int TS_THREADS = 288;              // 8 producer / drain warps + MMA warp
constexpr int TS_G = 4;                      // components per unit (TMEM: 4 x 128 columns)
constexpr uint32_t TS_TILE_BYTES = 16384;    // 128 rows x 64 FP16

struct TsBarriers {
    uint64_t a_full[2], a_empty[2];
    uint64_t tile_done;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(TS_THREADS, 1)
tc_stats_kernel(const float* __restrict__ Z, int64_t N, int D, int64_t ldz,
                const float* __restrict__ R, const float* __restrict__ lse, int64_t ldr, int K,
                const unsigned int* __restrict__ maxbits,
                double* __restrict__ partial, double* __restrict__ stat, int F,
                int groups, int slabs, int64_t slab_points, int flush_tiles,
                const unsigned int* __restrict__ gate, unsigned int gate_value) {
    if (gate != nullptr && __ldg(gate) != gate_value) return;     // device-side choice: the pair-list statistics ran instead
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // Bz: [hi|lo][point block 2] ; A: [buf 2][hi|lo][point block 2] ; tiles of 16 KB
    unsigned char* sBz = smem_raw;
    unsigned char* sA = sBz + 4 * TS_TILE_BYTES;
    float* sR = reinterpret_cast<float*>(sA + 8 * TS_TILE_BYTES);        // [TS_G][128]
    TsBarriers* bars = reinterpret_cast<TsBarriers*>(sR + TS_G * 128);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) { mbar_init(&bars->a_full[b], 256); mbar_init(&bars->a_empty[b], 1); }
        mbar_init(&bars->tile_done, 1);
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(&bars->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const int n_units = groups * slabs;
    const int ND = (D + 15) & ~15;                           // MMA N

    if (warp < 8) {
        // ================= producers / drain =================
        const float sz = pow2_scale_for(__uint_as_float(__ldg(maxbits)));
        const double inv_sz = 1.0 / (double)sz;
        const int i = tid & 127, h = tid >> 7;               // feature, point half
        uint32_t ac = 0, tcount = 0;                          // A buffers used, tiles issued (this CTA)
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int slab = u / groups, grp = u - slab * groups;
            const int k0 = grp * TS_G;
            const int64_t p0 = (int64_t)slab * slab_points;
            const int64_t p1 = min(N, p0 + slab_points);
            if (p0 >= p1) continue;
            const int n_tiles = (int)((p1 - p0 + 127) / 128);
            double acc_rx[TS_G], acc_r[TS_G];
#pragma unroll
            for (int g = 0; g < TS_G; ++g) { acc_rx[g] = 0.0; acc_r[g] = 0.0; }
            float zn[64];
            {
                const int64_t nb = p0 + h * 64;
#pragma unroll
                for (int p = 0; p < 64; ++p) zn[p] = (nb + p < p1 && i < D) ? __ldg(Z + (nb + p) * ldz + i) : 0.f;
            }
            for (int tt = 0; tt < n_tiles; ++tt, ++tcount) {
                float z[64];
#pragma unroll
                for (int p = 0; p < 64; ++p) z[p] = zn[p] * sz;
                const int64_t t0 = p0 + (int64_t)tt * 128;
                // responsibilities of this tile for the unit's 4 components (issued early, stored after the wait)
                float rv[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    int idx = tid + e * 256;                 // [g][point]
                    int g = idx >> 7, p = idx & 127;
                    rv[e] = (k0 + g < K && t0 + p < p1) ? __ldg(R + (int64_t)(k0 + g) * ldr + t0 + p) : 0.f;
                    if (lse != nullptr && k0 + g < K && t0 + p < p1) rv[e] = fast_exp(rv[e] - __ldg(lse + t0 + p));   // R holds log-joints
                }
                if (tt + 1 < n_tiles) {                      // prefetch the next tile's column of Z
                    const int64_t nb = t0 + 128 + h * 64;
#pragma unroll
                    for (int p = 0; p < 64; ++p) zn[p] = (nb + p < p1 && i < D) ? __ldg(Z + (nb + p) * ldz + i) : 0.f;
                }
                // every MMA of the previous tile must have finished reading Bz / sR
                if (tcount > 0) mbar_wait(&bars->tile_done, (tcount - 1) & 1);
                {   // B = Z^T: row = feature i, 64 points of block h
                    unsigned char* rowh = sBz + (size_t)(0 * 2 + h) * TS_TILE_BYTES;
                    unsigned char* rowl = sBz + (size_t)(1 * 2 + h) * TS_TILE_BYTES;
#pragma unroll
                    for (int c8 = 0; c8 < 8; ++c8) {
                        float x[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) x[e] = z[c8 * 8 + e];
                        uint4 hi, lo;
                        split8(x, hi, lo);
                        const uint32_t o = sw128_chunk_off(i, c8);
                        *reinterpret_cast<uint4*>(rowh + o) = hi;
                        *reinterpret_cast<uint4*>(rowl + o) = lo;
                    }
                }
                sR[tid] = rv[0];
                sR[tid + 256] = rv[1];
                asm volatile("bar.sync 1, 256;" ::: "memory");               // sR visible to the producers
#pragma unroll
                for (int g = 0; g < TS_G; ++g, ++ac) {
                    const uint32_t buf = ac & 1;
                    mbar_wait(&bars->a_empty[buf], ((ac >> 1) & 1) ^ 1);
                    unsigned char* rowh = sA + (size_t)((buf * 2 + 0) * 2 + h) * TS_TILE_BYTES;
                    unsigned char* rowl = sA + (size_t)((buf * 2 + 1) * 2 + h) * TS_TILE_BYTES;
                    const float* rg = sR + g * 128 + h * 64;
                    float sum = 0.f, rsum = 0.f;
#pragma unroll
                    for (int c8 = 0; c8 < 8; ++c8) {
                        const float4 ra = *reinterpret_cast<const float4*>(rg + c8 * 8);
                        const float4 rb = *reinterpret_cast<const float4*>(rg + c8 * 8 + 4);
                        const float rr[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
                        float x[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) { x[e] = rr[e] * z[c8 * 8 + e]; sum += x[e]; rsum += rr[e]; }
                        uint4 hi, lo;
                        split8(x, hi, lo);
                        const uint32_t o = sw128_chunk_off(i, c8);
                        *reinterpret_cast<uint4*>(rowh + o) = hi;
                        *reinterpret_cast<uint4*>(rowl + o) = lo;
                    }
                    acc_rx[g] += (double)sum;
                    acc_r[g] += (double)rsum;
                    fence_proxy_async();
                    mbar_arrive(&bars->a_full[buf]);
                }
                // ---- drain the accumulators into the unit's FP64 partials ----
                if ((tt + 1) % flush_tiles == 0 || tt + 1 == n_tiles) {
                    mbar_wait(&bars->tile_done, tcount & 1);
                    tc_fence_after();
                    const int qd = warp & 3, ch = warp >> 2;                 // lane quarter, column half
                    const int row = qd * 32 + lane;                          // accumulator row = feature i'
#pragma unroll 1
                    for (int g = 0; g < TS_G; ++g) {
                        if (k0 + g >= K) break;
                        double* pk = partial + ((size_t)slab * groups * TS_G + (k0 + g)) * 16384;
#pragma unroll 1
                        for (int b = 0; b < 2; ++b) {
                            const int col0 = ch * 64 + b * 32;
                            if (col0 >= D) break;
                            float v[32];
                            tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + g * 128 + col0, v);
                            tmem_ld_wait();
                            if (row < D) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    const int col = col0 + j;
                                    if (col <= row) pk[(size_t)col * 128 + row] += (double)v[j];   // [j][i]: lanes contiguous
                                }
                            }
                        }
                    }
                    tc_fence_before();
                    asm volatile("bar.sync 1, 256;" ::: "memory");           // all drained before the next tile's A is signalled
                }
            }
            // ---- s_k and n_k of this unit ----
#pragma unroll
            for (int g = 0; g < TS_G; ++g) {
                if (k0 + g < K) {
                    if (i < D) atomicAdd(&stat[(size_t)(k0 + g) * F + (size_t)D * (D + 1) / 2 + i], acc_rx[g] * inv_sz);
                    if (i == 0) atomicAdd(&stat[(size_t)(k0 + g) * F + F - 1], acc_r[g]);
                }
            }
        }
    } else {
        // ================= MMA issuer (one thread) =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_f16(128, ND);
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sBz);
            uint32_t ac = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int slab = u / groups;
                const int64_t p0 = (int64_t)slab * slab_points;
                const int64_t p1 = min(N, p0 + slab_points);
                if (p0 >= p1) continue;
                const int n_tiles = (int)((p1 - p0 + 127) / 128);
                bool fresh = true;
                for (int tt = 0; tt < n_tiles; ++tt) {
                    for (int g = 0; g < TS_G; ++g, ++ac) {
                        const uint32_t buf = ac & 1;
                        mbar_wait(&bars->a_full[buf], (ac >> 1) & 1);
                        tc_fence_after();
                        const uint32_t d = tmem_base + g * 128;
#pragma unroll
                        for (int pb = 0; pb < 2; ++pb) {
                            const uint64_t ah = make_desc_sw128(a0 + ((buf * 2 + 0) * 2 + pb) * TS_TILE_BYTES);
                            const uint64_t al = make_desc_sw128(a0 + ((buf * 2 + 1) * 2 + pb) * TS_TILE_BYTES);
                            const uint64_t bh = make_desc_sw128(b0 + (0 * 2 + pb) * TS_TILE_BYTES);
                            const uint64_t bl = make_desc_sw128(b0 + (1 * 2 + pb) * TS_TILE_BYTES);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                umma_f16(d, al + 2 * kk, bh + 2 * kk, idesc, (fresh && pb == 0 && kk == 0) ? 0u : 1u);
                                umma_f16(d, ah + 2 * kk, bl + 2 * kk, idesc, 1);
                                umma_f16(d, ah + 2 * kk, bh + 2 * kk, idesc, 1);
                            }
                        }
                        umma_commit(&bars->a_empty[buf]);
                    }
                    umma_commit(&bars->tile_done);
                    fresh = ((tt + 1) % flush_tiles == 0);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// stat[k][tri(i, j)] += (1 / sz^2) * sum_slab partial[slab][k][j][i]      (j <= i < D)
__global__ void tc_stats_reduce_kernel(const double* __restrict__ partial, int slabs, int kstride, int K, int D, int F,
                                       const unsigned int* __restrict__ maxbits, double* __restrict__ stat) {
    const int k = blockIdx.x;
    const float sz = tc::pow2_scale_for(__uint_as_float(*maxbits));
    const double inv = 1.0 / ((double)sz * (double)sz);
    for (int e = threadIdx.x; e < D * 128; e += blockDim.x) {
        int j = e >> 7, i = e & 127;
        if (i >= D || j > i) continue;
        double s = 0.0;
        for (int sl = 0; sl < slabs; ++sl) s += partial[((size_t)sl * kstride + k) * 16384 + (size_t)j * 128 + i];
        stat[(size_t)k * F + (size_t)i * (i + 1) / 2 + j] += s * inv;
    }
}

// ---- host side -------------------------------------------------------------------------------

static size_t a256(size_t x) { return (x + 255) / 256 * 256; }
static char* align1k(void* p) { return (char*)(((uintptr_t)p + 1023) / 1024 * 1024); }

bool tc_stats_supported(int dtype, int D, int F) {
    return dtype == MIMO_F32 && D >= 1 && D <= 128 && F == (D + 1) * (D + 2) / 2;
}

struct TsPlan { int groups, slabs; int64_t slab_points; };

// slabs: the smallest count that fills the SMs evenly (units = groups * slabs, one unit per CTA at a time)
static TsPlan ts_plan(int64_t chunk_points, int K) {
    TsPlan P;
    P.groups = (K + TS_G - 1) / TS_G;
    const int sms = sm_count();
    int64_t tiles = std::max<int64_t>(1, (chunk_points + 127) / 128);
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 64 && s <= tiles; ++s) {
        double units = (double)P.groups * s;
        double eff = units / (std::ceil(units / sms) * sms);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
        if (eff >= 0.95) { best = s; break; }
    }
    P.slabs = best;
    int64_t per = (tiles + P.slabs - 1) / P.slabs;
    P.slab_points = per * 128;
    return P;
}

// partial buffer: sized for the largest chunk the sweep will pass (slabs <= 64)
size_t tc_stats_workspace(int64_t chunk_points, int K) {
    TsPlan P = ts_plan(chunk_points, K);
    return (size_t)P.slabs * P.groups * TS_G * 16384 * sizeof(double) + 2048;
}

int tc_stats_begin(int64_t chunk_points, int K, void* ws, cudaStream_t st) {
    TsPlan P = ts_plan(chunk_points, K);
    MIMO_CUDA(cudaMemsetAsync(align1k(ws), 0, (size_t)P.slabs * P.groups * TS_G * 16384 * sizeof(double), st));
    return MIMO_OK;
}

static int g_flush_tiles = 16;
void tc_set_flush_tiles(int t) { g_flush_tiles = t < 1 ? 16 : t; }

// one chunk of points: accumulates into the partial buffer (plan of `plan_points`, the sweep's chunk size)
int tc_stats_chunk(const float* Z, int64_t N, int D, int64_t ldz, const float* R, int64_t ldr, int K, int F,
                   const unsigned int* maxbits, double* stat, int64_t plan_points, void* ws, cudaStream_t st,
                   const unsigned int* gate, unsigned int gate_value, const float* lse) {
    if (N == 0) return MIMO_OK;
    TsPlan P = ts_plan(plan_points, K);
    size_t smem = 12 * (size_t)TS_TILE_BYTES + TS_G * 128 * sizeof(float) + sizeof(TsBarriers) + 1024;
    MIMO_CUDA(cudaFuncSetAttribute(tc_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = std::min(P.groups * P.slabs, sm_count());
    tc_stats_kernel<<<grid, TS_THREADS, smem, st>>>(Z, N, D, ldz, R, lse, ldr, K, maxbits, (double*)align1k(ws), stat, F,
                                                    P.groups, P.slabs, P.slab_points, g_flush_tiles, gate, gate_value);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

int tc_stats_end(int64_t plan_points, int K, int D, int F, const unsigned int* maxbits, double* stat, void* ws, cudaStream_t st) {
    TsPlan P = ts_plan(plan_points, K);
    tc_stats_reduce_kernel<<<K, 256, 0, st>>>((const double*)align1k(ws), P.slabs, P.groups * TS_G, K, D, F, maxbits, stat);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

}  // namespace mimo
