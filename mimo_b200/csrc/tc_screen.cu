// Screened E-step: one cheap FP16 tensor-core pass over ALL (point, component) pairs, an exact pass only over
// the pairs that can matter.
//
//   a[k][n] = cst[k] - 0.5 * q[k][n],   q = || W_k [z_n ; 1] ||^2
//   (distributions/gaussian.py:510-523, bayesian.py:287-301 followed by the logsumexp of mixtures/gmm.py:72-75, 256-259)
//
// The responsibilities and the log-normaliser only depend on the components within a few tens of nats of a
// point's best component.  The screening pass computes a rigorous UPPER bound of every a[k][n] at a fraction of the
// dense cost:
//   * rows: with Q (32 x Rp) a matrix with orthonormal rows, || Q y ||^2 <= || y ||^2, so the 32-row operand
//     W'_k = Q W_k gives q' <= q.  Q is a sub-sampled Hadamard transform with fixed random column signs (a
//     Johnson-Lindenstrauss projection: E q' = q 32 / Rp whatever the geometry of W_k); a quarter of the MMA work
//     and of the accumulator traffic at Rp = 128;
//   * precision: operands rounded to FP16 (tc_estep2.cu, PASSES = 1), a third of the passes.  With y' = W' z and y~ its
//     FP16-operand value,  || y~ - y' ||_2 <= B = 1.05 * 2^-10 * max_k ||W'_k||_F * max_n ||z_n||_2  + (FP32 rounding
//     of the projection) + 1e-3   (two roundings of 2^-11 per product; Cauchy-Schwarz over the row, then the rows),
//     hence  U_kn = cst_k - max(0, sqrt(q~) - B)^2 / 2  >=  a[k][n].
// The pass also returns each point's best component under the screening values (the guess).  The guesses are
// recomputed exactly first (list A, one pair per point): L_n = a[guess_n][n] is a lower bound of the point's best
// log-joint.  Then (n, k) is a CANDIDATE iff U_kn >= L_n - 40 (list B): every other pair lies more than 40 nats below
// the point's maximum, all of them together change exp-sums by < K e^-40 relative -- far below FP32 resolution --
// so their screening values (upper bounds, themselves below the cut) are kept.  Both lists are grouped by component
// (counting sort) and recomputed in FP32 on the CUDA cores, overwriting the scratch; the sufficient statistics are then
// summed over the two lists only (pair_stats.cu).  When more than 4 % of the pairs are candidates (overlapping
// components, early sweeps) refinement would cost more than it saves: a device-side flag then makes the dense 3-pass
// tensor-core kernel and the dense statistics kernels run instead -- no host round trip either way -- and the rest of
// the sweep moves one screening TIER up (device-side state): tier 0 the projection, tier 1 all operand rows in one
// FP16 pass (q' = q: a tighter bound for moderately separated components at 4x the accumulator traffic), tier 2 no
// screening.
#include <algorithm>
#include "tc_common.cuh"
#include "internal.h"

namespace mimo {

using tc::screen_bound;

constexpr float SCREEN_T0 = 40.f;
constexpr int SCREEN_ROWS = 32;              // rows of the screening operands
constexpr int RF_THREADS = 512;
constexpr int RF_CPT = 8;                    // candidates per thread (x 4 rows: 32 accumulators)

// max_n ||z_n||_2: one warp per row (grid-stride)
__global__ void screen_rownorm_kernel(const float* __restrict__ Z, int64_t N, int D, int64_t ldz, unsigned int* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float m = 0.f;
    for (int64_t n = warp0; n < N; n += nwarps) {
        float s = 0.f;
        for (int j = lane; j < D; j += 32) { const float v = __ldg(Z + n * ldz + j); s = fmaf(v, v, s); }
        s = warp_sum(s);
        m = fmaxf(m, s);
    }
    if (lane == 0) atomicMax(flags + 3, __float_as_uint(sqrtf(m) * 1.0000005f));
}

// max_k ||W_k[:, :cols]||_F -> flags[slot]   (slot 2: screening operands, data columns only -- the offset column is
// applied in FP32 in the epilogue; slot 4: full operands, all columns -- scale of the projection's FP32 rounding)
__global__ void screen_wnorm_kernel(const float* __restrict__ W, int K, int Rp, int Dpp, int cols, unsigned int* __restrict__ flags, int slot) {
    __shared__ float red[32];
    const int k = blockIdx.x;
    float s = 0.f;
    for (int idx = threadIdx.x; idx < Rp * cols; idx += blockDim.x) {
        const int i = idx / cols, j = idx - i * cols;
        const float w = W[((size_t)k * Rp + i) * Dpp + j];
        s = fmaf(w, w, s);
    }
    s = block_sum<float>(s, red);
    if (threadIdx.x == 0) atomicMax(flags + slot, __float_as_uint(sqrtf(s) * 1.0000005f));
}

// W'_k = Q W_k  (all Dpp columns: data, offset, padding).  Q[s][r] = sign(r) H[sel(s)][r] / sqrt(Rp) with H the
// Sylvester-Hadamard matrix, H[a][r] = (-1)^popc(a & r): distinct rows of H are orthogonal, so Q has orthonormal rows.
// grid = K, block = 256, dynamic shared memory = Rp * Dpp floats.
__global__ void __launch_bounds__(256)
screen_project_kernel(const float* __restrict__ W, int K, int Rp, int Dpp, float* __restrict__ Wp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Ws = reinterpret_cast<float*>(smem_raw);
    const int k = blockIdx.x;
    for (int idx = threadIdx.x; idx < Rp * Dpp; idx += blockDim.x) {
        const int r = idx / Dpp;
        const float sgn = ((((unsigned int)r * 2654435761u) >> 15) & 1u) ? -1.f : 1.f;
        Ws[idx] = sgn * W[(size_t)k * Rp * Dpp + idx];
    }
    __syncthreads();
    const float inv = rsqrtf((float)Rp);
    const int step = Rp / SCREEN_ROWS;
    for (int idx = threadIdx.x; idx < SCREEN_ROWS * Dpp; idx += blockDim.x) {
        const int s = idx / Dpp, j = idx - s * Dpp;
        const int a = s * step + step / 2;                       // the selected Hadamard row
        float acc = 0.f;
        for (int r = 0; r < Rp; ++r) {
            const float w = Ws[r * Dpp + j];
            acc += (__popc(a & r) & 1) ? -w : w;
        }
        Wp[((size_t)k * SCREEN_ROWS + s) * Dpp + j] = acc * inv;
    }
}

// list A: the better of the two accumulator halves' guesses; histogram over components
__global__ void __launch_bounds__(256)
screen_guess_kernel(const float* __restrict__ best_val, const int* __restrict__ best_k, int64_t ldl, int64_t n, int K,
                    int* __restrict__ guess_k, int* __restrict__ histA, const unsigned int* __restrict__ level) {
    if (*level >= 2u) return;                             // earlier chunks of this sweep exhausted the screening tiers: straight to the dense pass
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v0 = best_val[i], v1 = best_val[ldl + i];
    int k = (v1 > v0) ? best_k[ldl + i] : best_k[i];
    k = min(max(k, 0), K - 1);
    guess_k[i] = k;
    atomicAdd(histA + k, 1);
}

__global__ void __launch_bounds__(256)
screen_scatterA_kernel(const int* __restrict__ guess_k, int64_t n, int* __restrict__ cursor, int* __restrict__ perm,
                       const unsigned int* __restrict__ level) {
    if (*level >= 2u) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[atomicAdd(cursor + guess_k[i], 1)] = (int)i;
}

// list B: every pair whose upper bound reaches the point's threshold L_n - 40 (its guess excluded: already exact)
// counters: [0] list B entries found, [1] dense flag (set by screen_scan_kernel), [2] always 0
// grid = (point tiles of EMIT_PTS, component groups of EMIT_KG): a CTA streams EMIT_KG rows of 4 KB each (the rows of
// the scratch lie megabytes apart, so a CTA that walked all K rows would touch K pages of the TLB per kilobyte read);
// thread = 4 consecutive points (one 128-bit streaming load per row).  Candidates are rare: one ballot per row decides
// whether any lane of the warp has one at all.
constexpr int EMIT_PTS = 1024;
constexpr int EMIT_KG = 16;

__global__ void __launch_bounds__(256)
screen_emit_kernel(const float* __restrict__ a, int K, int64_t n, int64_t ldo, int vec4, const float* __restrict__ cst,
                   const unsigned int* __restrict__ flags_proj, const unsigned int* __restrict__ flags_full,
                   const float* __restrict__ lower, const int* __restrict__ guess_k,
                   int2* __restrict__ list, unsigned int cap, unsigned int* __restrict__ counters, int* __restrict__ hist,
                   const unsigned int* __restrict__ level) {
    if (*level >= 2u) return;
    const float B = screen_bound(*level == 0u ? flags_proj : flags_full);      // the bound of the tier that produced the values
    const int64_t i0 = (int64_t)blockIdx.x * EMIT_PTS + threadIdx.x * 4;
    const int k_lo = blockIdx.y * EMIT_KG, k_hi = min(K, k_lo + EMIT_KG);
    const int lane = threadIdx.x & 31;
    float t[4];
    int gk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const bool valid = i0 + e < n;
        const float lw = valid ? lower[i0 + e] : 0.f;
        t[e] = valid ? lw - SCREEN_T0 - 0.01f * (1.f + 1e-4f * fabsf(lw)) : INFINITY;     // FP32 slack of the exact value
        gk[e] = valid ? guess_k[i0 + e] : -1;
    }
    for (int k = k_lo; k < k_hi; ++k) {
        float av[4] = {0.f, 0.f, 0.f, 0.f};
        const float* row = a + (int64_t)k * ldo + i0;
        if (vec4 && i0 + 3 < n) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(row));
            av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) if (i0 + e < n) av[e] = __ldcs(row + e);
        }
        const float c = __ldg(cst + k);
        unsigned int mine = 0u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float s = fmaxf(0.f, sqrtf(fmaxf(0.f, 2.f * (c - av[e]))) - B);
            if ((c - 0.5f * s * s) >= t[e] && k != gk[e]) mine |= 1u << e;
        }
        if (__ballot_sync(0xffffffffu, mine != 0u) == 0u) continue;          // the common case: nothing in this warp
        // warp-aggregated append: exclusive prefix of the per-lane counts
        const int cnt = __popc(mine);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned int base = 0;
        if (lane == 0) { base = atomicAdd(counters, (unsigned int)total); atomicAdd(hist + k, total); }
        base = __shfl_sync(0xffffffffu, base, 0) + (unsigned int)(incl - cnt);
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (mine & (1u << e)) { if (base < cap) list[base] = make_int2(k, (int)(i0 + e)); ++base; }
    }
}

// single block: exclusive scan of hist -> offsets[K+1], cursor := offsets, slabs (work items of pair_stats.cu);
// flag (optional): dense fallback when the list is too long; level (optional): the screening tier of the sweep
// (0 projected rows, 1 all rows in one FP16 pass, 2 none).  A chunk whose tier yields too many candidates takes the
// dense pass and moves the REST of the sweep one tier up: chunks of one data set look alike, and a dense chunk costs
// ~10x a projected screening pass, so a failed tier is not worth retrying.
__global__ void screen_scan_kernel(const int* __restrict__ hist, int K, int* __restrict__ offsets, int* __restrict__ cursor,
                                   int* __restrict__ slabs, unsigned int* __restrict__ counters, unsigned int max_cands,
                                   unsigned int* __restrict__ level, unsigned int n_points) {
    if (threadIdx.x == 0) {
        // sweep totals behind the level word (tc_screen_totals): [2,3] candidate pairs of the refined chunks,
        // [4,5] their points, [6] chunks that took the dense pass, [7] chunks
        unsigned long long* tot = level ? reinterpret_cast<unsigned long long*>(level + 2) : nullptr;
        if (level && *level >= 2u) { if (counters) { counters[1] = 1u; level[6] += 1u; level[7] += 1u; } return; }
        int run = 0, items = 0;
        for (int k = 0; k < K; ++k) {
            offsets[k] = run; cursor[k] = run; slabs[k] = items;
            run += hist[k];
            items += (hist[k] + PS_SLAB - 1) / PS_SLAB;
        }
        offsets[K] = run;
        slabs[K] = items;
        if (counters) {
            counters[1] = (counters[0] > max_cands) ? 1u : 0u;
            if (level) {
                if (counters[1]) { *level += 1u; level[6] += 1u; }
                else { tot[0] += (unsigned long long)counters[0] + n_points; tot[1] += n_points; }
                level[7] += 1u;
            }
        }
    }
}

__global__ void screen_scatter_kernel(const int2* __restrict__ list, const unsigned int* __restrict__ counters,
                                      int* __restrict__ cursor, int* __restrict__ perm) {
    if (counters[1] != 0u) return;                       // dense second pass instead
    const unsigned int total = counters[0];
    for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int2 c = list[e];
        perm[atomicAdd(cursor + c.x, 1)] = c.y;
    }
}

// Exact (FP32 FMA) recomputation of a list of candidates grouped by component.
// Work item = (component, slab of <= PS_SLAB listed candidates) on a persistent grid (one CTA per SM): the component's
// operand block W_k stays in shared memory for the whole slab, candidate rows are gathered TILE_C at a time into a
// double-buffered tile with cp.async, so the gather of tile i + 1 runs under the arithmetic of tile i.
// Thread (rg, cg): rows rg + RG r (r < 4), candidates RF_CPT cg + c (c < RF_CPT) of the tile.
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int RP>
__global__ void __launch_bounds__(RF_THREADS, 1)
screen_refine_kernel(const float* __restrict__ Z, int D, int64_t ldz, int vec4,
                     const float* __restrict__ W, int K, int Dpp, const float* __restrict__ cst,
                     const int* __restrict__ perm, const int* __restrict__ offsets, const int* __restrict__ slabs,
                     const unsigned int* __restrict__ gate, unsigned int gate_min,
                     float* __restrict__ a, int64_t ldo, float* __restrict__ exact) {
    if (gate != nullptr && *gate >= gate_min) return;          // dense pass instead
    constexpr int RG = RP / 4, CG = RF_THREADS / RG, TILE_C = RF_CPT * CG;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Ws = reinterpret_cast<float*>(smem_raw);            // [RP][Dpp]
    float* Zs0 = Ws + (size_t)RP * Dpp;                        // 2 x [TILE_C][Dpp]   (column D = 1, beyond = 0)
    const int tid = threadIdx.x, rg = tid % RG, cg = tid / RG;
    const int n_items = slabs[K];
    int k_loaded = -1;

    // gather of one tile (asynchronous for the data columns; rows beyond the tile are zero-filled)
    auto prefetch = [&](float* Zs, int c0, int nc) {
        if (vec4) {
            const int q4 = D >> 2;
            for (int idx = tid; idx < TILE_C * q4; idx += RF_THREADS) {
                const int c = idx / q4, j = (idx - c * q4) << 2;
                const bool ok = c < nc;
                const float* src = ok ? Z + (int64_t)__ldg(perm + c0 + c) * ldz + j : Z;
                cp_async16(Zs + (size_t)c * Dpp + j, src, ok ? 16 : 0);
            }
        } else {
            for (int idx = tid; idx < TILE_C * D; idx += RF_THREADS) {
                const int c = idx / D, j = idx - c * D;
                const bool ok = c < nc;
                const float* src = ok ? Z + (int64_t)__ldg(perm + c0 + c) * ldz + j : Z;
                cp_async4(Zs + (size_t)c * Dpp + j, src, ok ? 4 : 0);
            }
        }
        const int tail = Dpp - D;
        for (int idx = tid; idx < TILE_C * tail; idx += RF_THREADS) {
            const int c = idx / tail, j = D + (idx - c * tail);
            Zs[(size_t)c * Dpp + j] = (c < nc && j == D) ? 1.f : 0.f;
        }
        cp_async_commit();
    };

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int lo = 0, hi = K;                                    // item -> component: last k with slabs[k] <= item
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (slabs[mid] <= item) lo = mid; else hi = mid;
        }
        const int k = lo;
        const int beg = offsets[k] + (item - slabs[k]) * PS_SLAB;
        const int end = min(offsets[k + 1], beg + PS_SLAB);
        if (beg >= end) continue;
        __syncthreads();                                       // previous item done with Ws and both tiles
        if (k != k_loaded) {
            for (int idx = tid; idx < RP * Dpp / 4; idx += RF_THREADS)
                reinterpret_cast<float4*>(Ws)[idx] = __ldg(reinterpret_cast<const float4*>(W + (size_t)k * RP * Dpp) + idx);
            k_loaded = k;
        }
        const float ck = __ldg(cst + k);
        const int ntiles = (end - beg + TILE_C - 1) / TILE_C;
        prefetch(Zs0, beg, min(TILE_C, end - beg));
        for (int t = 0; t < ntiles; ++t) {
            float* Zs = Zs0 + (size_t)(t & 1) * TILE_C * Dpp;
            const int c0 = beg + t * TILE_C, nc = min(TILE_C, end - c0);
            if (t + 1 < ntiles) {                              // the other buffer was released by the barrier that ended tile t - 1
                prefetch(Zs0 + (size_t)((t + 1) & 1) * TILE_C * Dpp, c0 + TILE_C, min(TILE_C, end - c0 - TILE_C));
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();                                   // tile t (and Ws) visible to everyone
            float acc[4][RF_CPT];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < RF_CPT; ++c) acc[r][c] = 0.f;
#pragma unroll 2
            for (int j = 0; j < Dpp; j += 4) {
                float4 w[4], z[RF_CPT];
#pragma unroll
                for (int r = 0; r < 4; ++r) w[r] = *reinterpret_cast<const float4*>(Ws + (size_t)(rg + RG * r) * Dpp + j);
#pragma unroll
                for (int c = 0; c < RF_CPT; ++c) z[c] = *reinterpret_cast<const float4*>(Zs + (size_t)(RF_CPT * cg + c) * Dpp + j);
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < RF_CPT; ++c) {
                        acc[r][c] = fmaf(w[r].x, z[c].x, acc[r][c]);
                        acc[r][c] = fmaf(w[r].y, z[c].y, acc[r][c]);
                        acc[r][c] = fmaf(w[r].z, z[c].z, acc[r][c]);
                        acc[r][c] = fmaf(w[r].w, z[c].w, acc[r][c]);
                    }
            }
#pragma unroll
            for (int c = 0; c < RF_CPT; ++c) {
                float q = 0.f;
#pragma unroll
                for (int r = 0; r < 4; ++r) q = fmaf(acc[r][c], acc[r][c], q);
#pragma unroll
                for (int o = RG / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);    // over the RG lanes of this candidate group
                const int ci = RF_CPT * cg + c;
                if (rg == 0 && ci < nc) {
                    const int n = perm[c0 + ci];
                    const float val = ck - 0.5f * q;
                    a[(int64_t)k * ldo + n] = val;
                    if (exact) exact[n] = val;                 // list A: the point's guess, exactly (a lower bound of its best log-joint)
                }
            }
            __syncthreads();                                   // tile t consumed: its buffer may be refilled
        }
    }
}

// ---- log-normaliser over the lists only ---------------------------------------------------------
// lse_n = log sum_k exp(a[k][n]) restricted to the point's guess and its list-B candidates: every other pair lies
// more than 40 nats below the maximum (mixtures/gmm.py:72-75, 256-259 up to K e^-40 relative).  All kernels return
// at once when the dense flag is set (the dense softmax kernel runs instead).
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else          atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void __launch_bounds__(256)
screen_lse_init_kernel(const float* __restrict__ lower, int64_t n, const unsigned int* __restrict__ counters,
                       float* __restrict__ m, float* __restrict__ s) {
    if (counters[1] != 0u) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { m[i] = lower[i]; s[i] = 0.f; }
}

// phase 0: m_n = max(m_n, a) over list B;  phase 1: s_n += exp(a - m_n) over list B
__global__ void __launch_bounds__(256)
screen_lse_list_kernel(const int2* __restrict__ list, const unsigned int* __restrict__ counters, const float* __restrict__ a, int64_t ldo,
                       float* __restrict__ m, float* __restrict__ s, int phase) {
    if (counters[1] != 0u) return;
    const unsigned int total = counters[0];
    for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int2 c = list[e];
        const float v = a[(int64_t)c.x * ldo + c.y];
        if (phase == 0) atomic_max_float(m + c.y, v);
        else            atomicAdd(s + c.y, __expf(v - m[c.y]));
    }
}

__global__ void __launch_bounds__(256)
screen_lse_final_kernel(const float* __restrict__ lower, const float* __restrict__ m, const float* __restrict__ s, int64_t n,
                        const unsigned int* __restrict__ counters, float* __restrict__ lse, double* __restrict__ lse_sum) {
    if (counters[1] != 0u) return;
    __shared__ double red[32];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double mine = 0.0;
    if (i < n) {
        const float mm = m[i];
        const double tot = (double)s[i] + (double)expf(lower[i] - mm);
        mine = (double)mm + log(tot);
        lse[i] = (float)mine;
    }
    if (lse_sum) {
        const double tot = block_sum<double>(mine, red);
        if (threadIdx.x == 0) atomicAdd(lse_sum, tot);
    }
}

// ---- host side -------------------------------------------------------------------------------

static size_t a256(size_t x) { return (x + 255) / 256 * 256; }
static char* align256(void* p) { return (char*)(((uintptr_t)p + 255) / 256 * 256); }

bool tc_screen_supported(int D, int Rp) { return D >= 24 && D <= 128 && (Rp == 32 || Rp == 64 || Rp == 128); }

struct ScreenLists { size_t hist, offsets, cursor, slabs, perm; };
struct ScreenLayout {
    unsigned int cap; int64_t ldl;
    size_t off_counters; ScreenLists A, B;
    size_t off_best_val, off_best_k, off_lower, off_guess, off_m, off_s, off_lse, off_level, off_list, bytes;
};
static ScreenLayout screen_layout(int64_t chunk_points, int K) {
    ScreenLayout L;
    const double pairs = (double)chunk_points * K;
    L.cap = (unsigned int)std::min<double>(2.0e9, 0.05 * pairs + 1024.0);
    L.ldl = (int64_t)(a256((size_t)chunk_points * 4) / 4);
    const size_t kk = a256((size_t)(K + 1) * 4), pp = (size_t)L.ldl * 4;
    size_t o = 0;
    L.off_counters = o; o += 256;
    L.A.hist = o; o += kk;  L.B.hist = o; o += kk;                 // zeroed together with the counters every chunk
    L.A.offsets = o; o += kk;  L.A.cursor = o; o += kk;  L.A.slabs = o; o += kk;
    L.B.offsets = o; o += kk;  L.B.cursor = o; o += kk;  L.B.slabs = o; o += kk;
    L.off_best_val = o; o += 2 * pp;                               // one row per accumulator half
    L.off_best_k = o;   o += 2 * pp;
    L.off_lower = o;    o += pp;
    L.off_guess = o;    o += pp;
    L.off_m = o;        o += pp;
    L.off_s = o;        o += pp;
    L.off_lse = o;      o += pp;
    L.off_level = o;    o += 256;
    L.A.perm = o;       o += pp;
    L.off_list = o;     o += a256((size_t)L.cap * 8);
    L.B.perm = o;       o += a256((size_t)L.cap * 4);
    L.bytes = o;
    return L;
}
size_t tc_screen_workspace(int64_t chunk_points, int K) { return screen_layout(chunk_points, K).bytes + 256; }

// [projected operands W' (K, 32, Dpp) | their operand image]; nothing when the operands already have <= 32 rows
size_t tc_screen_operand_workspace(int K, int Rp, int Dpp, int D) {
    if (Rp <= SCREEN_ROWS) return 0;
    return 256 + a256((size_t)K * SCREEN_ROWS * Dpp * 4) + tc_operand_workspace(K, SCREEN_ROWS, D);
}
static float* screen_wproj(void* sops_ws) { return (float*)align256(sops_ws); }
static void* screen_ops_ws(void* ops_ws, void* sops_ws, int K, int Rp, int Dpp) {
    if (Rp <= SCREEN_ROWS) return ops_ws;
    return align256(sops_ws) + a256((size_t)K * SCREEN_ROWS * Dpp * 4);
}
int tc_screen_rows(int Rp) { return Rp <= SCREEN_ROWS ? Rp : SCREEN_ROWS; }

// once per sweep, after tc_data_scale + tc_prepare_operands on ops_ws: the norms of the error bound, the projected
// operands and their image
int tc_screen_prepare(const float* Z, int64_t N, int D, int64_t ldz, const float* W, const float* cst, int K, int Rp, int Dpp,
                      void* ops_ws, void* sops_ws, cudaStream_t st) {
    unsigned int* flags = tc_flags(ops_ws);
    if (N > 0) {
        const int grid = (int)std::min<int64_t>((N + 7) / 8, (int64_t)sm_count() * 16);
        screen_rownorm_kernel<<<grid, 256, 0, st>>>(Z, N, D, ldz, flags);
        MIMO_LAUNCH_CHECK();
    }
    if (Rp <= SCREEN_ROWS) {
        screen_wnorm_kernel<<<K, 256, 0, st>>>(W, K, Rp, Dpp, D, flags, 2);
        MIMO_LAUNCH_CHECK();
        return MIMO_OK;
    }
    screen_wnorm_kernel<<<K, 256, 0, st>>>(W, K, Rp, Dpp, D, flags, 2);        // tier 1 (all rows): data columns
    screen_wnorm_kernel<<<K, 256, 0, st>>>(W, K, Rp, Dpp, Dpp, flags, 4);      // FP32 rounding of the projection
    MIMO_LAUNCH_CHECK();
    float* Wp = screen_wproj(sops_ws);
    const size_t smem = (size_t)Rp * Dpp * sizeof(float);
    MIMO_CUDA(cudaFuncSetAttribute(screen_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    screen_project_kernel<<<K, 256, smem, st>>>(W, K, Rp, Dpp, Wp);
    MIMO_LAUNCH_CHECK();
    void* sws = screen_ops_ws(ops_ws, sops_ws, K, Rp, Dpp);
    unsigned int* sflags = tc_flags(sws);
    MIMO_CUDA(cudaMemcpyAsync(sflags, flags, 256, cudaMemcpyDeviceToDevice, st));      // data scale, ||z||, ||W||
    MIMO_CUDA(cudaMemsetAsync(sflags + 2, 0, 4, st));
    screen_wnorm_kernel<<<K, 256, 0, st>>>(Wp, K, SCREEN_ROWS, Dpp, D, sflags, 2);
    MIMO_LAUNCH_CHECK();
    return tc_prepare_operands(Wp, cst, K, SCREEN_ROWS, Dpp, D, sws, st);
}

// the screening pass over one chunk: upper bounds of every pair into `out`, the per-half guesses into the workspace
int tc_screen_pass(const float* Z, int64_t n, int D, int64_t ldz, int K, int Rp, int Dpp, float* out, int64_t ldo,
                   void* ops_ws, void* sops_ws, int64_t plan_points, void* ws, cudaStream_t st) {
    ScreenLayout L = screen_layout(plan_points, K);
    char* base = align256(ws);
    const unsigned int* level = (const unsigned int*)(base + L.off_level);
    if (Rp > SCREEN_ROWS) {                              // tier 0: projected rows
        int rc = tc_estep_pass(Z, n, D, ldz, K, SCREEN_ROWS, out, ldo, screen_ops_ws(ops_ws, sops_ws, K, Rp, Dpp), 1, level, 0u,
                               (float*)(base + L.off_best_val), (int*)(base + L.off_best_k), L.ldl, st);
        if (rc) return rc;
    }
    // tier 1: all rows, one FP16 pass
    return tc_estep_pass(Z, n, D, ldz, K, Rp, out, ldo, ops_ws, 1, level, 1u,
                         (float*)(base + L.off_best_val), (int*)(base + L.off_best_k), L.ldl, st);
}

__global__ void screen_level_kernel(unsigned int* level, unsigned int v) {
    *level = v;
    for (int i = 1; i < 8; ++i) level[i] = 0u;           // sweep totals (screen_scan_kernel)
}

// once per sweep: the screening tier the first chunk starts on (0 projected rows -- needs Rp > 32 --, 1 all rows)
int tc_screen_begin(int64_t plan_points, int K, int Rp, int start_level, void* ws, cudaStream_t st) {
    unsigned int* level = (unsigned int*)(align256(ws) + screen_layout(plan_points, K).off_level);
    const unsigned int v = (Rp <= SCREEN_ROWS || start_level >= 1) ? 1u : 0u;
    screen_level_kernel<<<1, 1, 0, st>>>(level, v);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

const unsigned int* tc_screen_gate(void* ws, int64_t plan_points, int K) {
    return (const unsigned int*)(align256(ws) + screen_layout(plan_points, K).off_counters) + 1;
}

// the lists of the chunk, grouped by component: which = 0 the guesses (one per point), 1 the other candidates
// (valid after tc_screen_refine when the dense flag is clear)
void tc_screen_lists(void* ws, int64_t plan_points, int K, int which, const int32_t** perm, const int32_t** offsets, const int32_t** slabs) {
    ScreenLayout L = screen_layout(plan_points, K);
    char* base = align256(ws);
    const ScreenLists& S = which ? L.B : L.A;
    *perm = (const int32_t*)(base + S.perm);
    *offsets = (const int32_t*)(base + S.offsets);
    *slabs = (const int32_t*)(base + S.slabs);
}

static const unsigned int* g_last_counters = nullptr;
static const unsigned int* g_last_level = nullptr;
static int64_t g_last_points = 0;

void tc_screen_forget() { g_last_counters = nullptr; g_last_level = nullptr; g_last_points = 0; }

// screening tier the most recent sweep ended on (0 projection, 1 all rows, 2 none; -1 unknown); synchronises
int tc_screen_level() {
    if (!g_last_level) return -1;
    unsigned int v = 0;
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(&v, g_last_level, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int)v;
}   // the workspace the counters live in is going away

// totals of the most recent screened sweep: {candidate pairs (guesses + list B) of the refined chunks, points of the
// refined chunks, chunks that took the dense pass, chunks, tier the sweep ended on}; synchronises
int tc_screen_totals(unsigned long long* out_host5) {
    for (int i = 0; i < 5; ++i) out_host5[i] = 0ull;
    if (!g_last_level) return MIMO_OK;
    unsigned int h[8];
    MIMO_CUDA(cudaDeviceSynchronize());
    MIMO_CUDA(cudaMemcpy(h, g_last_level, sizeof(h), cudaMemcpyDeviceToHost));
    out_host5[0] = ((unsigned long long)h[3] << 32) | h[2];
    out_host5[1] = ((unsigned long long)h[5] << 32) | h[4];
    out_host5[2] = h[6]; out_host5[3] = h[7]; out_host5[4] = h[0];
    return MIMO_OK;
}

// {candidates (guesses + list B), dense flag} of the most recent screened chunk (synchronises the device; tests / bench reporting)
int tc_screen_last(unsigned int* out_host2) {
    out_host2[0] = out_host2[1] = 0u;
    if (!g_last_counters) return MIMO_OK;
    MIMO_CUDA(cudaDeviceSynchronize());
    MIMO_CUDA(cudaMemcpy(out_host2, g_last_counters, 8, cudaMemcpyDeviceToHost));
    out_host2[0] += (unsigned int)g_last_points;
    return MIMO_OK;
}

template <int RP>
static int launch_refine(const float* Z, int D, int64_t ldz, const float* W, int K, int Dpp, const float* cst,
                         const int* perm, const int* offsets, const int* slabs, const unsigned int* gate, unsigned int gate_min,
                         float* a, int64_t ldo, float* exact, cudaStream_t st) {
    constexpr int TILE_C = RF_CPT * (RF_THREADS / (RP / 4));
    const size_t smem = (size_t)(RP + 2 * TILE_C) * Dpp * sizeof(float);
    if (smem > 227 * 1024) { set_error("screened E-step: refinement tile needs %zu B of shared memory", smem); return MIMO_EUNSUPPORTED; }
    const int vec4 = (D % 4 == 0) && (Dpp % 4 == 0) && (ldz % 4 == 0) && (((uintptr_t)Z & 15) == 0);
    MIMO_CUDA(cudaFuncSetAttribute(screen_refine_kernel<RP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    screen_refine_kernel<RP><<<sm_count(), RF_THREADS, smem, st>>>(Z, D, ldz, vec4, W, K, Dpp, cst, perm, offsets, slabs, gate, gate_min,
                                                                    a, ldo, exact);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}
static int refine(const float* Z, int D, int64_t ldz, const float* W, int K, int Rp, int Dpp, const float* cst,
                  const int* perm, const int* offsets, const int* slabs, const unsigned int* gate, unsigned int gate_min,
                  float* a, int64_t ldo, float* exact, cudaStream_t st) {
    if (Rp == 32) return launch_refine<32>(Z, D, ldz, W, K, Dpp, cst, perm, offsets, slabs, gate, gate_min, a, ldo, exact, st);
    if (Rp == 64) return launch_refine<64>(Z, D, ldz, W, K, Dpp, cst, perm, offsets, slabs, gate, gate_min, a, ldo, exact, st);
    if (Rp == 128) return launch_refine<128>(Z, D, ldz, W, K, Dpp, cst, perm, offsets, slabs, gate, gate_min, a, ldo, exact, st);
    set_error("screened E-step: unsupported Rp=%d", Rp);
    return MIMO_EUNSUPPORTED;
}

// after the screening pass of a chunk: exact values of the guesses (list A), then the candidates (list B) and the
// device flag the dense pass is gated on
int tc_screen_select(const float* Z, int D, int64_t ldz, const float* W, const float* cst, int K, int Rp, int Dpp,
                     float* a, int64_t n, int64_t ldo, void* ops_ws, void* sops_ws, int64_t plan_points, void* ws, cudaStream_t st) {
    ScreenLayout L = screen_layout(plan_points, K);
    char* base = align256(ws);
    unsigned int* counters = (unsigned int*)(base + L.off_counters);
    g_last_counters = counters;
    g_last_level = (const unsigned int*)(base + L.off_level);
    g_last_points = n;
    MIMO_CUDA(cudaMemsetAsync(base, 0, L.A.offsets, st));                   // counters + both histograms
    const int grid = cdiv(n, 256);
    int* guess_k = (int*)(base + L.off_guess);
    float* lower = (float*)(base + L.off_lower);
    unsigned int* level = (unsigned int*)(base + L.off_level);
    screen_guess_kernel<<<grid, 256, 0, st>>>((const float*)(base + L.off_best_val), (const int*)(base + L.off_best_k), L.ldl, n, K,
                                              guess_k, (int*)(base + L.A.hist), level);
    screen_scan_kernel<<<1, 32, 0, st>>>((const int*)(base + L.A.hist), K, (int*)(base + L.A.offsets), (int*)(base + L.A.cursor),
                                         (int*)(base + L.A.slabs), nullptr, 0u, nullptr, 0u);
    screen_scatterA_kernel<<<grid, 256, 0, st>>>(guess_k, n, (int*)(base + L.A.cursor), (int*)(base + L.A.perm), level);
    MIMO_LAUNCH_CHECK();
    int rc = refine(Z, D, ldz, W, K, Rp, Dpp, cst, (const int*)(base + L.A.perm), (const int*)(base + L.A.offsets), (const int*)(base + L.A.slabs), level, 2u, a, ldo, lower, st);
    if (rc) return rc;
    const int evec4 = (ldo % 4 == 0) && (((uintptr_t)a & 15) == 0);
    screen_emit_kernel<<<dim3(cdiv(n, EMIT_PTS), cdiv(K, EMIT_KG)), 256, 0, st>>>(
        a, K, n, ldo, evec4, cst, tc_flags(screen_ops_ws(ops_ws, sops_ws, K, Rp, Dpp)), tc_flags(ops_ws),
        lower, guess_k, (int2*)(base + L.off_list), L.cap, counters, (int*)(base + L.B.hist), level);
    const double maxc = std::min<double>((double)L.cap, 0.04 * (double)n * K);
    screen_scan_kernel<<<1, 32, 0, st>>>((const int*)(base + L.B.hist), K, (int*)(base + L.B.offsets), (int*)(base + L.B.cursor),
                                         (int*)(base + L.B.slabs), counters, (unsigned int)maxc, level, (unsigned int)n);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

// exact values of list B (returns immediately on the device when the dense flag is set)
int tc_screen_refine(const float* Z, int D, int64_t ldz, const float* W, int K, int Rp, int Dpp, const float* cst,
                     float* a, int64_t ldo, int64_t plan_points, void* ws, cudaStream_t st) {
    ScreenLayout L = screen_layout(plan_points, K);
    char* base = align256(ws);
    const unsigned int* counters = (const unsigned int*)(base + L.off_counters);
    screen_scatter_kernel<<<sm_count() * 4, 256, 0, st>>>((const int2*)(base + L.off_list), counters, (int*)(base + L.B.cursor), (int*)(base + L.B.perm));
    MIMO_LAUNCH_CHECK();
    return refine(Z, D, ldz, W, K, Rp, Dpp, cst, (const int*)(base + L.B.perm), (const int*)(base + L.B.offsets), (const int*)(base + L.B.slabs), counters + 1, 1u, a, ldo, nullptr, st);
}

// log-normalisers of the chunk from the lists (after tc_screen_refine): per-point lse into the workspace, their sum
// added to *lse_sum.  No-op on the device when the chunk fell back to the dense pass.
int tc_screen_lse(const float* a, int K, int64_t n, int64_t ldo, double* lse_sum, int64_t plan_points, void* ws, cudaStream_t st) {
    ScreenLayout L = screen_layout(plan_points, K);
    char* base = align256(ws);
    const unsigned int* counters = (const unsigned int*)(base + L.off_counters);
    const float* lower = (const float*)(base + L.off_lower);
    float* m = (float*)(base + L.off_m);
    float* s = (float*)(base + L.off_s);
    const int grid = cdiv(n, 256);
    screen_lse_init_kernel<<<grid, 256, 0, st>>>(lower, n, counters, m, s);
    screen_lse_list_kernel<<<sm_count() * 4, 256, 0, st>>>((const int2*)(base + L.off_list), counters, a, ldo, m, s, 0);
    screen_lse_list_kernel<<<sm_count() * 4, 256, 0, st>>>((const int2*)(base + L.off_list), counters, a, ldo, m, s, 1);
    screen_lse_final_kernel<<<grid, 256, 0, st>>>(lower, m, s, n, counters, (float*)(base + L.off_lse), lse_sum);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}
const float* tc_screen_lse_values(void* ws, int64_t plan_points, int K) {
    return (const float*)(align256(ws) + screen_layout(plan_points, K).off_lse);
}


// ---- responsibility lists for the CUDA-core path (small D) ---------------------------------------------
// After the dense softmax of a chunk the responsibilities of most pairs are (next to) zero once components are
// separated.  The pairs with r >= e^-40 are listed, grouped by component, and the statistics summed over the list
// (pair_stats.cu); the dropped mass is < N e^-40 per component.  Dense CUDA-core statistics (stats.cu) when more
// than rl_max_frac(D) of the pairs qualify -- the same device-side flag mechanism as above.
constexpr float RL_TAU = 4.2e-18f;           // e^-40
// Break-even: a listed pair costs 64 FMAs per 8 x 8 tile of the triangle (plus a row gather), a dense pair F FMAs; measured
// on cfg2 (D = 9) and cfg4 (D = 16) the list wins below ~0.1 F / (64 tiles) of the pairs.
static double rl_max_frac(int D) {
    const int T = (D + 8) >> 3, ntiles = T * (T + 1) / 2, F = (D + 1) * (D + 2) / 2;
    return std::min(0.15, 0.1 * (double)F / (64.0 * ntiles));
}

__global__ void __launch_bounds__(256)
resp_emit_kernel(const float* __restrict__ r, int K, int64_t n, int64_t ldr, int2* __restrict__ list, unsigned int cap,
                 unsigned int* __restrict__ counters, int* __restrict__ hist) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const int lane = threadIdx.x & 31;
    for (int k0 = 0; k0 < K; k0 += 8) {
        float rv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) rv[u] = (valid && k0 + u < K) ? r[(int64_t)(k0 + u) * ldr + i] : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = k0 + u;
            if (k >= K) break;
            const bool cand = rv[u] >= RL_TAU;
            const unsigned int m = __ballot_sync(0xffffffffu, cand);
            if (m) {
                unsigned int base = 0;
                const int leader = __ffs(m) - 1;
                if (lane == leader) { base = atomicAdd(counters, (unsigned int)__popc(m)); atomicAdd(hist + k, __popc(m)); }
                base = __shfl_sync(0xffffffffu, base, leader);
                if (cand) {
                    const unsigned int slot = base + __popc(m & ((1u << lane) - 1u));
                    if (slot < cap) list[slot] = make_int2(k, (int)i);
                }
            }
        }
    }
}

struct RespListLayout { unsigned int cap; size_t off_counters, hist, offsets, cursor, slabs, list, perm, bytes; };
static RespListLayout resp_list_layout(int64_t chunk_points, int K) {
    RespListLayout L;
    L.cap = (unsigned int)std::min<double>(2.0e9, 0.16 * (double)chunk_points * K + 1024.0);
    const size_t kk = a256((size_t)(K + 1) * 4);
    size_t o = 0;
    L.off_counters = o; o += 256;
    L.hist = o; o += kk;
    L.offsets = o; o += kk;  L.cursor = o; o += kk;  L.slabs = o; o += kk;
    L.list = o; o += a256((size_t)L.cap * 8);
    L.perm = o; o += a256((size_t)L.cap * 4);
    L.bytes = o;
    return L;
}
size_t resp_list_workspace(int64_t chunk_points, int K) { return resp_list_layout(chunk_points, K).bytes + 256; }

// lists of the chunk's responsibilities R (K, n); sets the dense flag *resp_list_gate() when the list is too long
int resp_list_build(const float* R, int K, int64_t n, int64_t ldr, int D, int64_t plan_points, void* ws, cudaStream_t st) {
    RespListLayout L = resp_list_layout(plan_points, K);
    char* base = align256(ws);
    unsigned int* counters = (unsigned int*)(base + L.off_counters);
    MIMO_CUDA(cudaMemsetAsync(base, 0, L.offsets, st));                     // counters + histogram
    resp_emit_kernel<<<cdiv(n, 256), 256, 0, st>>>(R, K, n, ldr, (int2*)(base + L.list), L.cap, counters, (int*)(base + L.hist));
    const double maxc = std::min<double>((double)L.cap, rl_max_frac(D) * (double)n * K);
    screen_scan_kernel<<<1, 32, 0, st>>>((const int*)(base + L.hist), K, (int*)(base + L.offsets), (int*)(base + L.cursor),
                                         (int*)(base + L.slabs), counters, (unsigned int)maxc, nullptr, 0u);
    screen_scatter_kernel<<<sm_count() * 4, 256, 0, st>>>((const int2*)(base + L.list), counters, (int*)(base + L.cursor), (int*)(base + L.perm));
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}
const unsigned int* resp_list_gate(void* ws, int64_t plan_points, int K) {
    return (const unsigned int*)(align256(ws) + resp_list_layout(plan_points, K).off_counters) + 1;
}
void resp_list_get(void* ws, int64_t plan_points, int K, const int32_t** perm, const int32_t** offsets, const int32_t** slabs) {
    RespListLayout L = resp_list_layout(plan_points, K);
    char* base = align256(ws);
    *perm = (const int32_t*)(base + L.perm);
    *offsets = (const int32_t*)(base + L.offsets);
    *slabs = (const int32_t*)(base + L.slabs);
}

}  // namespace mimo
