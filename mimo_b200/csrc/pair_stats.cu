// Sufficient statistics over a LIST of (component, point) pairs grouped by component -- the sparse form of
//
//   stat[k][tri(i, j)] += sum_{n in list(k)} r[k][n] * zt[n][i] * zt[n][j]        zt = [z ; 1],  j <= i
//
// (distributions/gaussian.py:491-505 with the dense one-hot / responsibility matrix of utils/data.py:160-169,
//  mixtures/gmm.py:236, 289-297; lingauss.py:306-325 on z = [x | y]).
//
// Three callers:
//   * Gibbs (hard labels): list(k) = the points labelled k (counting sort of stats.cu), r = 1;
//   * mean field behind the screened E-step (tc_screen.cu): list(k) = the candidate points of component k,
//     r = exp(a - lse_n) from the refined log-joint.  Every pair outside the list has a responsibility below e^-40,
//     so the list carries the whole statistic to FP32 resolution; when the list would be long (overlapping
//     components) the dense tensor-core kernels run instead;
//   * mean field on the CUDA-core path, 16 <= D < 24: list(k) = the points with r >= e^-40 after the dense softmax
//     (resp_list_* in tc_screen.cu), dense CUDA-core statistics above a break-even list length.
//
// One work item = (component, slab of <= PS_SLAB listed points).  The CTA gathers 32 rows at a time into a
// double-buffered shared-memory tile with cp.async (the gather of the next 32 rows runs under the arithmetic of the
// current ones) and accumulates the lower triangle of the (D+1) x (D+1) outer-product sum in 8 x 8 register tiles,
// one tile per thread (FP32 FMA pipe; 153 tiles at D = 128; the row weight is applied to the 8 row-side values as
// they are read); the slab total is added to the FP64 statistics with one atomic per element.  For small D several
// thread groups split the 32 rows.
#include <algorithm>
#include "common.cuh"
#include "internal.h"

namespace mimo {

constexpr int PS_THREADS = 160;
constexpr int PS_PT = 32;

// column c of a staged row: 4 floats of padding after every 32 so that the 8-float groups t and t + 4 of one row
// start in different banks (the 128-bit reads of 8 consecutive threads then never collide)
__device__ __forceinline__ int ps_off(int c) { return c + ((c >> 5) << 2); }

__device__ __forceinline__ void ps_cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void ps_cp_async4(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}

__global__ void __launch_bounds__(PS_THREADS, 3)
pair_stats_kernel(const float* __restrict__ Z, int D, int64_t ldz, int vec4,
                  const int32_t* __restrict__ perm, const int32_t* __restrict__ offsets, const int32_t* __restrict__ slabs, int K,
                  const float* __restrict__ R, int64_t ldr, const float* __restrict__ lse,
                  const unsigned int* __restrict__ gate, unsigned int gate_value,
                  double* __restrict__ stat, int F) {
    if (gate != nullptr && __ldg(gate) != gate_value) return;
    const int T = (D + 8) >> 3;                       // 8-wide column groups of zt (D + 1 columns)
    const int W8 = T << 3;
    const int RS = W8 + 4 * ((T + 3) >> 2);           // staged row stride (floats, multiple of 4)
    const int ntiles = T * (T + 1) / 2;
    const int G = max(1, PS_THREADS / ntiles);        // thread groups splitting the rows of a tile
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Bs0 = reinterpret_cast<float*>(smem_raw);  // 2 x [PT][RS]  zt, double buffered (rows gathered with cp.async)
    __shared__ float s_w[2][PS_PT];                   // weight of each staged row

    const int tid = threadIdx.x;
    const int g = tid / ntiles, e = tid - g * ntiles;
    const bool active = g < G;
    int ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= e) ++ti;        // tile e = ti (ti + 1) / 2 + tj of the lower triangle
    const int tj = e - ti * (ti + 1) / 2;
    const int offa = ps_off(ti << 3), offb = ps_off(tj << 3);
    const int n_items = slabs[K];

    // gather of the listed rows [p0, p0 + np) into buffer b: data columns asynchronously, the 1 / padding columns and
    // the row weights with plain stores
    auto prefetch = [&](int k, int b, int p0, int np) {
        float* Bs = Bs0 + (size_t)b * PS_PT * RS;
        if (tid < np) {
            float w = 1.f;                            // hard labels
            if (R) {
                const int n = __ldg(perm + p0 + tid);
                w = __ldg(R + (int64_t)k * ldr + n);                            // responsibility ...
                if (lse) w = __expf(w - __ldg(lse + n));                        // ... or log-joint and the point's log-normaliser
            }
            s_w[b][tid] = w;
        }
        if (vec4) {
            const int q = D >> 2;
            for (int idx = tid; idx < np * q; idx += PS_THREADS) {
                const int p = idx / q, c = (idx - p * q) << 2;
                ps_cp_async16(Bs + p * RS + ps_off(c), Z + (int64_t)__ldg(perm + p0 + p) * ldz + c);
            }
        } else {
            for (int idx = tid; idx < np * D; idx += PS_THREADS) {
                const int p = idx / D, c = idx - p * D;
                ps_cp_async4(Bs + p * RS + ps_off(c), Z + (int64_t)__ldg(perm + p0 + p) * ldz + c);
            }
        }
        const int tail = W8 - D;                      // the 1 column and the zero padding of the last group
        for (int idx = tid; idx < np * tail; idx += PS_THREADS) {
            const int p = idx / tail, c = D + (idx - p * tail);
            Bs[p * RS + ps_off(c)] = (c == D) ? 1.f : 0.f;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int lo = 0, hi = K;                           // item -> component: last k with slabs[k] <= item
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (slabs[mid] <= item) lo = mid; else hi = mid;
        }
        const int k = lo;
        const int beg = offsets[k] + (item - slabs[k]) * PS_SLAB;
        const int end = min(offsets[k + 1], beg + PS_SLAB);
        if (beg >= end) continue;
        float acc[8][8];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;

        __syncthreads();                              // previous item done with both buffers
        prefetch(k, 0, beg, min(PS_PT, end - beg));
        int buf = 0;
        for (int p0 = beg; p0 < end; p0 += PS_PT, buf ^= 1) {
            const int np = min(PS_PT, end - p0);
            if (p0 + PS_PT < end) {                   // the other buffer was released by the barrier that ended the previous tile
                prefetch(k, buf ^ 1, p0 + PS_PT, min(PS_PT, end - p0 - PS_PT));
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncthreads();                          // this tile's rows and weights visible to everyone
            if (active) {
                const float* Bs = Bs0 + (size_t)buf * PS_PT * RS;
#pragma unroll 2
                for (int p = g; p < np; p += G) {
                    const float w = s_w[buf][p];
                    const float4 a0 = *reinterpret_cast<const float4*>(Bs + p * RS + offa);
                    const float4 a1 = *reinterpret_cast<const float4*>(Bs + p * RS + offa + 4);
                    const float4 b0 = *reinterpret_cast<const float4*>(Bs + p * RS + offb);
                    const float4 b1 = *reinterpret_cast<const float4*>(Bs + p * RS + offb + 4);
                    const float a[8] = {w * a0.x, w * a0.y, w * a0.z, w * a0.w, w * a1.x, w * a1.y, w * a1.z, w * a1.w};
                    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int x = 0; x < 8; ++x)
#pragma unroll
                        for (int y = 0; y < 8; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
                }
            }
            __syncthreads();                          // tile consumed: its buffer may be refilled
        }
        if (active) {
            double* out = stat + (int64_t)k * F;
#pragma unroll
            for (int x = 0; x < 8; ++x) {
                const int i = (ti << 3) + x;
                if (i > D) continue;
#pragma unroll
                for (int y = 0; y < 8; ++y) {
                    const int j = (tj << 3) + y;
                    if (j <= i) atomicAdd(out + (int64_t)i * (i + 1) / 2 + j, (double)acc[x][y]);
                }
            }
        }
    }
}

// the statistics must be the packed lower triangle of zt zt^T (the layout of quad_features) in FP32 data
bool pair_stats_supported(int dtype, int D, int F) {
    return dtype == MIMO_F32 && D >= 8 && D <= 128 && F == (D + 1) * (D + 2) / 2;
}

int pair_stats(const float* Z, int D, int64_t ldz, const int32_t* perm, const int32_t* offsets, const int32_t* slabs, int K,
               const float* R, int64_t ldr, const float* lse, const unsigned int* gate, unsigned int gate_value,
               double* stat, int F, cudaStream_t st) {
    MIMO_CHECK_ARG(Z && perm && offsets && slabs && stat, "null pointer");
    MIMO_CHECK_ARG(pair_stats_supported(MIMO_F32, D, F), "pair statistics: unsupported shape");
    const int T = (D + 8) >> 3;
    const int RS = 8 * T + 4 * ((T + 3) >> 2);
    const size_t smem = (size_t)2 * PS_PT * RS * sizeof(float);
    const int vec4 = (D % 4 == 0) && (ldz % 4 == 0) && (((uintptr_t)Z & 15) == 0);
    MIMO_CUDA(cudaFuncSetAttribute(pair_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pair_stats_kernel<<<sm_count() * 6, PS_THREADS, smem, st>>>(Z, D, ldz, vec4, perm, offsets, slabs, K, R, ldr, lse,
                                                                gate, gate_value, stat, F);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

}  // namespace mimo
