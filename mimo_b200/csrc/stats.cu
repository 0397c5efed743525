// Weighted sufficient statistics, CUDA-core path, sm_100a.
//
//   stat[k][f] += sum_n r[k][n] * zt[n][fi[f]] * zt[n][fj[f]]        zt = [z ; 1]
//
// soft (mean-field): a register-tiled GEMM  R (K x points) . Phi (points x F)  whose
//   B operand -- the feature tile Phi -- is generated in shared memory from the point
//   tile, so neither the one-hot / responsibility-weighted copies of the data nor the
//   per-point outer products of the reference ever exist.  FP32 partial sums are
//   folded into FP64 every 1024 points; the CTA's slab total is added to the global
//   FP64 buffer with one atomic per element.
// hard (Gibbs): counting sort of the point indices by label, then a segmented dense
//   reduction per component with FP64 accumulation in registers.
//
// Reference call sites replaced: distributions/gaussian.py:491-505, 819-832;
// lingauss.py:306-325; categorical.py:35-46; utils/data.py:160-169.
#include "common.cuh"
#include "internal.h"
#include <algorithm>

namespace mimo {

constexpr int SS_THREADS = 256;
constexpr int SS_BK = 64;     // components per CTA tile   (16 thread rows x 4)
constexpr int SS_BF = 128;    // features per CTA tile     (16 thread cols x 8)
constexpr int SS_BP = 32;     // points per inner block
constexpr int SS_FOLD = 32;   // inner blocks between FP32 -> FP64 folds

template <typename T>
__global__ void __launch_bounds__(SS_THREADS)
stats_soft_kernel(const T* __restrict__ Z, int64_t N, int D, int64_t ldz,
                  const T* __restrict__ resp, int64_t ldr, int K,
                  const int32_t* __restrict__ fi, const int32_t* __restrict__ fj, int F,
                  double* __restrict__ stat, int64_t slab, const unsigned int* __restrict__ gate, unsigned int gate_value) {
    if (gate != nullptr && *gate != gate_value) return;          // the pair-list statistics ran instead
    constexpr int AS = SS_BK + 4, BS = SS_BF + 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ZTS = D + 2;                                   // zt row stride
    T* Zt = reinterpret_cast<T*>(smem_raw);                  // [BP][ZTS]
    T* As = Zt + SS_BP * ZTS + ((4 - (SS_BP * ZTS) % 4) % 4);  // [BP][AS]  r, transposed (keep 16B alignment)
    T* Bs = As + SS_BP * AS;                                 // [BP][BS]  generated features
    __shared__ int sfi[SS_BF], sfj[SS_BF];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int f0 = blockIdx.x * SS_BF, k0 = blockIdx.y * SS_BK;
    const int64_t p_begin = (int64_t)blockIdx.z * slab;
    const int64_t p_end = min(N, p_begin + slab);
    if (tid < SS_BF) {
        bool ok = f0 + tid < F;
        sfi[tid] = ok ? fi[f0 + tid] : D;
        sfj[tid] = ok ? fj[f0 + tid] : D;
    }

    float  accf[4][8];
    double accd[4][8];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 8; ++n) { accf[m][n] = 0.f; accd[m][n] = 0.0; }

    int fold = 0;
    for (int64_t p0 = p_begin; p0 < p_end; p0 += SS_BP) {
        __syncthreads();
        for (int idx = tid; idx < SS_BP * D; idx += SS_THREADS) {
            int p = idx / D, j = idx - p * D;
            Zt[p * ZTS + j] = (p0 + p < p_end) ? Z[(p0 + p) * ldz + j] : T(0);
        }
        if (tid < SS_BP) Zt[tid * ZTS + D] = T(1);
        for (int idx = tid; idx < SS_BP * SS_BK; idx += SS_THREADS) {
            int kk = idx / SS_BP, p = idx - kk * SS_BP;
            bool ok = (k0 + kk < K) && (p0 + p < p_end);
            As[p * AS + kk] = ok ? resp[(int64_t)(k0 + kk) * ldr + p0 + p] : T(0);
        }
        __syncthreads();
        for (int idx = tid; idx < SS_BP * SS_BF; idx += SS_THREADS) {
            int p = idx / SS_BF, f = idx - p * SS_BF;
            Bs[p * BS + f] = Zt[p * ZTS + sfi[f]] * Zt[p * ZTS + sfj[f]];
        }
        __syncthreads();
#pragma unroll 4
        for (int p = 0; p < SS_BP; ++p) {
            T a[4], b[8];
            lds_vec<T, 4>(a, As + p * AS + ty * 4);
            lds_vec<T, 8>(b, Bs + p * BS + tx * 8);
            if constexpr (sizeof(T) == 4) {
#pragma unroll
                for (int m = 0; m < 4; ++m)
#pragma unroll
                    for (int n = 0; n < 8; ++n) accf[m][n] = fmaf(a[m], b[n], accf[m][n]);
            } else {
#pragma unroll
                for (int m = 0; m < 4; ++m)
#pragma unroll
                    for (int n = 0; n < 8; ++n) accd[m][n] = fma((double)a[m], (double)b[n], accd[m][n]);
            }
        }
        if constexpr (sizeof(T) == 4) {
            if (++fold == SS_FOLD) {
                fold = 0;
#pragma unroll
                for (int m = 0; m < 4; ++m)
#pragma unroll
                    for (int n = 0; n < 8; ++n) { accd[m][n] += (double)accf[m][n]; accf[m][n] = 0.f; }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        int k = k0 + ty * 4 + m;
        if (k >= K) continue;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            int f = f0 + tx * 8 + n;
            if (f < F) atomicAdd(&stat[(int64_t)k * F + f], accd[m][n] + (double)accf[m][n]);
        }
    }
}

int stats_soft(int dtype, const void* Z, int64_t N, int D, int64_t ldz, const void* resp, int64_t ldr, int K,
               const int32_t* fi, const int32_t* fj, int F, double* stat, cudaStream_t st,
               const unsigned int* gate, unsigned int gate_value) {
    MIMO_CHECK_ARG(dtype == MIMO_F32 || dtype == MIMO_F64, "dtype");
    MIMO_CHECK_ARG(Z && resp && fi && fj && stat, "null pointer");
    MIMO_CHECK_ARG(N >= 0 && D >= 1 && K >= 1 && F >= 1 && ldz >= D && ldr >= N, "shape");
    if (N == 0) return MIMO_OK;
    size_t es = dtype == MIMO_F32 ? 4 : 8;
    size_t smem = ((size_t)SS_BP * (D + 2) + 4 + SS_BP * (SS_BK + 4) + SS_BP * (SS_BF + 4)) * es;
    if (smem > 200 * 1024) { set_error("soft stats: D=%d too large for shared memory", D); return MIMO_EUNSUPPORTED; }
    int ft = cdiv(F, SS_BF), kt = cdiv(K, SS_BK);
    int64_t want = (int64_t)4 * sm_count() / ((int64_t)ft * kt);
    int64_t slabs = std::max<int64_t>(1, std::min<int64_t>(want, cdiv(N, SS_BP * 8)));
    slabs = std::min<int64_t>(slabs, 65535);
    int64_t slab = ((N + slabs - 1) / slabs + SS_BP - 1) / SS_BP * SS_BP;
    slabs = cdiv(N, slab);
    dim3 grid(ft, kt, (unsigned)slabs);
    if (dtype == MIMO_F32) {
        auto kern = stats_soft_kernel<float>;
        MIMO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, SS_THREADS, smem, st>>>((const float*)Z, N, D, ldz, (const float*)resp, ldr, K, fi, fj, F, stat, slab, gate, gate_value);
    } else {
        auto kern = stats_soft_kernel<double>;
        MIMO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, SS_THREADS, smem, st>>>((const double*)Z, N, D, ldz, (const double*)resp, ldr, K, fi, fj, F, stat, slab, gate, gate_value);
    }
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

// ---- hard statistics: counting sort + segmented reduction ------------------------

__global__ void label_hist_kernel(const int32_t* __restrict__ labels, int64_t N, int K,
                                  int32_t* __restrict__ counts, int32_t* __restrict__ bad) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) {
        int z = labels[i];
        if (z < 0 || z >= K) atomicExch(bad, 1);
        else atomicAdd(&counts[z], 1);
    }
}

// Counting sort with block-local histograms (K <= LB_KMAX): one global atomic per (block, component) instead of one per
// point -- with a few hundred components the per-point atomics of the kernels above serialise on as many addresses
// (cfg3 of BASELINE.json, N = 100M, K = 256: 33 ms for histogram + scatter against 5 ms for the statistics themselves).
constexpr int LB_KMAX = 4096;
constexpr int LB_TILE = 4096;                  // points per block of the scatter: 256 threads x 16
__global__ void __launch_bounds__(256)
label_hist_block_kernel(const int32_t* __restrict__ labels, int64_t N, int K,
                        int32_t* __restrict__ counts, int32_t* __restrict__ bad) {
    extern __shared__ int32_t lb_h[];
    for (int k = threadIdx.x; k < K; k += 256) lb_h[k] = 0;
    __syncthreads();
    bool oob = false;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < N; i += (int64_t)gridDim.x * 256) {
        const int z = __ldg(labels + i);
        if (z < 0 || z >= K) oob = true;
        else atomicAdd(&lb_h[z], 1);
    }
    if (oob) atomicExch(bad, 1);
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += 256) if (lb_h[k]) atomicAdd(&counts[k], lb_h[k]);
}
// one block per LB_TILE points: local ranks from a shared-memory histogram, one range reservation per component
__global__ void __launch_bounds__(256)
label_scatter_block_kernel(const int32_t* __restrict__ labels, int64_t N, int K,
                           int32_t* __restrict__ cursor, int32_t* __restrict__ perm) {
    extern __shared__ int32_t lb_h[];          // [K] counts, then [K] bases
    int32_t* base = lb_h + K;
    for (int k = threadIdx.x; k < K; k += 256) lb_h[k] = 0;
    __syncthreads();
    const int64_t i0 = (int64_t)blockIdx.x * LB_TILE + threadIdx.x;
    int z[LB_TILE / 256], rk[LB_TILE / 256];
#pragma unroll
    for (int j = 0; j < LB_TILE / 256; ++j) {
        const int64_t i = i0 + j * 256;
        z[j] = i < N ? __ldg(labels + i) : -1;
        if (z[j] >= K) z[j] = -1;
        rk[j] = z[j] >= 0 ? atomicAdd(&lb_h[z[j]], 1) : 0;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += 256) base[k] = lb_h[k] ? atomicAdd(&cursor[k], lb_h[k]) : 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < LB_TILE / 256; ++j)
        if (z[j] >= 0) perm[base[z[j]] + rk[j]] = (int32_t)(i0 + j * 256);
}

// single block: exclusive scan of counts -> offsets[K+1]; cursor := offsets;
// slabs[k] = first work item (SH_SEG-point slab) of component k, slabs[K] = number of items
__global__ void label_scan_kernel(const int32_t* __restrict__ counts, int K, int32_t* __restrict__ offsets,
                                  int32_t* __restrict__ cursor, int32_t* __restrict__ slabs, int seg) {
    if (threadIdx.x == 0) {
        int run = 0, items = 0;
        for (int k = 0; k < K; ++k) {
            offsets[k] = run; cursor[k] = run; slabs[k] = items;
            run += counts[k];
            items += (counts[k] + seg - 1) / seg;
        }
        offsets[K] = run;
        slabs[K] = items;
    }
}

__global__ void label_scatter_kernel(const int32_t* __restrict__ labels, int64_t N, int K,
                                     int32_t* __restrict__ cursor, int32_t* __restrict__ perm) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) {
        int z = labels[i];
        if (z >= 0 && z < K) perm[atomicAdd(&cursor[z], 1)] = (int32_t)i;
    }
}

constexpr int SH_THREADS = 256;
constexpr int SH_SEG = 2048;    // points of one component per CTA
constexpr int SH_PT = 32;       // gathered rows per shared-memory tile
constexpr int SH_MAXFT = 34;    // features per thread when F > 256  (F <= 8704)

template <typename T>
__global__ void __launch_bounds__(SH_THREADS)
stats_hard_kernel(const T* __restrict__ Z, int D, int64_t ldz, const int32_t* __restrict__ perm,
                  const int32_t* __restrict__ offsets, const int32_t* __restrict__ slabs, int K,
                  const int32_t* __restrict__ fi, const int32_t* __restrict__ fj, int F,
                  double* __restrict__ stat, const int32_t* __restrict__ kind, int fast_kinds) {
    if (kind != nullptr && ((fast_kinds >> *kind) & 1)) return;      // a fast kernel for this feature layout ran instead
    // work item -> (component, slab): binary search in the per-component slab prefix
    const int item = blockIdx.x;
    if (item >= slabs[K]) return;
    int lo = 0, hi = K;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (slabs[mid] <= item) lo = mid; else hi = mid;
    }
    const int k = lo;
    const int beg = offsets[k] + (item - slabs[k]) * SH_SEG;
    const int end = min(offsets[k + 1], beg + SH_SEG);
    if (beg >= end) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ZTS = D + 2;
    T* Zt = reinterpret_cast<T*>(smem_raw);                  // [PT][ZTS]
    const int tid = threadIdx.x;
    const int Fc = min(F, SH_THREADS);
    const int G = (F >= SH_THREADS) ? 1 : SH_THREADS / F;    // point groups working in parallel
    const int g = tid / Fc, fl = tid - g * Fc;
    const bool active = g < G;
    const int nft = (F + SH_THREADS - 1) / SH_THREADS;       // features per thread
    int myi[SH_MAXFT], myj[SH_MAXFT];
    double acc[SH_MAXFT];
#pragma unroll
    for (int t = 0; t < SH_MAXFT; ++t) {
        int f = fl + t * SH_THREADS;
        bool ok = active && t < nft && f < F;
        myi[t] = ok ? fi[f] : D; myj[t] = ok ? fj[f] : D; acc[t] = 0.0;
    }
    for (int p0 = beg; p0 < end; p0 += SH_PT) {
        int np = min(SH_PT, end - p0);
        __syncthreads();
        for (int idx = tid; idx < np * D; idx += SH_THREADS) {
            int p = idx / D, j = idx - p * D;
            Zt[p * ZTS + j] = Z[(int64_t)perm[p0 + p] * ldz + j];
        }
        if (tid < np) Zt[tid * ZTS + D] = T(1);
        __syncthreads();
        if (active) {
            for (int p = g; p < np; p += G) {
                const T* row = Zt + p * ZTS;
                if (nft == 1) {
                    acc[0] += (double)row[myi[0]] * (double)row[myj[0]];
                } else {
#pragma unroll
                    for (int t = 0; t < SH_MAXFT; ++t)
                        if (t < nft) acc[t] += (double)row[myi[t]] * (double)row[myj[t]];
                }
            }
        }
    }
    if (active) {
#pragma unroll
        for (int t = 0; t < SH_MAXFT; ++t) {
            int f = fl + t * SH_THREADS;
            if (t < nft && f < F) atomicAdd(&stat[(int64_t)k * F + f], acc[t]);
        }
    }
}

// Which canonical layout is the caller's feature table?  1: packed lower triangle of zt zt^T (quad_features), 2: the
// diagonal family [z_j | z_j^2 | 1], 0: anything else.  The fast kernels below are gated on the answer ON THE DEVICE, so an
// arbitrary table still gets the generic kernel.
__global__ void feature_kind_kernel(const int32_t* __restrict__ fi, const int32_t* __restrict__ fj, int F, int D, int32_t* __restrict__ kind) {
    __shared__ int tri_ok, diag_ok;
    if (threadIdx.x == 0) { tri_ok = (F == (D + 1) * (D + 2) / 2); diag_ok = (F == 2 * D + 1); }
    __syncthreads();
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        const int i = fi[f], j = fj[f];
        if (tri_ok) {                                  // f = i (i + 1) / 2 + j, j <= i
            if (i < 0 || i > D || j < 0 || j > i || i * (i + 1) / 2 + j != f) tri_ok = 0;
        }
        if (diag_ok) {
            const bool ok = (f < D) ? (i == f && j == D) : (f < 2 * D) ? (i == f - D && j == f - D) : (i == D && j == D);
            if (!ok) diag_ok = 0;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *kind = tri_ok ? 1 : (diag_ok ? 2 : 0);
}

// Hard statistics of the diagonal family: sum x, sum x^2 and the count per component over the label-sorted lists.
// Work item = (component, slab of <= PS_SLAB points); thread = (row group, column): coalesced row reads, FP32 within the
// slab, one FP64 atomic per feature and slab.  HBM-bound (one read of Z per sweep).
constexpr int DH_THREADS = 256;
__global__ void __launch_bounds__(DH_THREADS)
diag_hard_stats_kernel(const float* __restrict__ Z, int D, int64_t ldz, const int32_t* __restrict__ perm,
                       const int32_t* __restrict__ offsets, const int32_t* __restrict__ slabs, int K,
                       const int32_t* __restrict__ kind, double* __restrict__ stat, int F) {
    if (*kind != 2) return;
    __shared__ float s1[DH_THREADS], s2[DH_THREADS];
    const int tid = threadIdx.x;
    const int cols = min(D, DH_THREADS);                 // columns handled per pass
    const int G = DH_THREADS / cols;                     // row groups
    const int g = tid / cols, c = tid - g * cols;
    const int n_items = slabs[K];
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int lo = 0, hi = K;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (slabs[mid] <= item) lo = mid; else hi = mid; }
        const int k = lo;
        const int beg = offsets[k] + (item - slabs[k]) * PS_SLAB;
        const int end = min(offsets[k + 1], beg + PS_SLAB);
        if (beg >= end) continue;
        for (int c0 = 0; c0 < D; c0 += cols) {
            const int j = c0 + c;
            float a1 = 0.f, a2 = 0.f;
            if (g < G && j < D) {
                int p = beg + g;
                for (; p + 3 * G < end; p += 4 * G) {                      // four gathers in flight
                    const float x0 = __ldg(Z + (int64_t)perm[p] * ldz + j), x1 = __ldg(Z + (int64_t)perm[p + G] * ldz + j);
                    const float x2 = __ldg(Z + (int64_t)perm[p + 2 * G] * ldz + j), x3 = __ldg(Z + (int64_t)perm[p + 3 * G] * ldz + j);
                    a1 += (x0 + x1) + (x2 + x3);
                    a2 = fmaf(x0, x0, fmaf(x1, x1, fmaf(x2, x2, fmaf(x3, x3, a2))));
                }
                for (; p < end; p += G) { const float x = __ldg(Z + (int64_t)perm[p] * ldz + j); a1 += x; a2 = fmaf(x, x, a2); }
            }
            __syncthreads();
            s1[tid] = a1; s2[tid] = a2;
            __syncthreads();
            if (g == 0 && j < D) {
                double t1 = 0.0, t2 = 0.0;
                for (int gg = 0; gg < G; ++gg) { t1 += (double)s1[gg * cols + c]; t2 += (double)s2[gg * cols + c]; }
                atomicAdd(stat + (int64_t)k * F + j, t1);
                atomicAdd(stat + (int64_t)k * F + D + j, t2);
            }
        }
        if (tid == 0) atomicAdd(stat + (int64_t)k * F + 2 * D, (double)(end - beg));
    }
}

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

size_t stats_hard_workspace(int64_t N, int K) {
    return align256((size_t)(K + 1) * 4) * 5 + 256 + align256((size_t)N * 4);
}

int stats_hard(int dtype, const void* Z, int64_t N, int D, int64_t ldz, const int32_t* labels, int K,
               const int32_t* fi, const int32_t* fj, int F, double* stat,
               void* workspace, size_t workspace_bytes, bool check, cudaStream_t st) {
    MIMO_CHECK_ARG(dtype == MIMO_F32 || dtype == MIMO_F64, "dtype");
    MIMO_CHECK_ARG(Z && labels && fi && fj && stat && workspace, "null pointer");
    MIMO_CHECK_ARG(N >= 0 && N < (int64_t)2147483647 && D >= 1 && K >= 1 && F >= 1 && ldz >= D, "shape");
    MIMO_CHECK_ARG(workspace_bytes >= stats_hard_workspace(N, K), "workspace too small");
    if (F > SH_MAXFT * SH_THREADS) { set_error("hard stats: F=%d too large", F); return MIMO_EUNSUPPORTED; }
    if (N == 0) return MIMO_OK;
    char* ws = (char*)workspace;
    size_t seg = align256((size_t)(K + 1) * 4);
    int32_t* counts = (int32_t*)ws;
    int32_t* offsets = (int32_t*)(ws + seg);
    int32_t* cursor = (int32_t*)(ws + 2 * seg);
    int32_t* slabs = (int32_t*)(ws + 3 * seg);
    int32_t* slabs_fast = (int32_t*)(ws + 4 * seg);                      // second slab prefix (PS_SLAB points per item)
    int32_t* bad = (int32_t*)(ws + 5 * seg);                             // [0] bad label seen, [1] feature layout
    int32_t* perm = (int32_t*)(ws + 5 * seg + 256);
    MIMO_CUDA(cudaMemsetAsync(ws, 0, 5 * seg + 256, st));
    int grid = cdiv(N, 256);
    const bool block_sort = K <= LB_KMAX;
    if (block_sort) label_hist_block_kernel<<<std::min(cdiv(N, 4096), sm_count() * 16), 256, (size_t)K * 4, st>>>(labels, N, K, counts, bad);
    else label_hist_kernel<<<grid, 256, 0, st>>>(labels, N, K, counts, bad);
    // FP32 data with a canonical feature table (checked on the device): the register-tiled pair-list kernel
    // (pair_stats.cu) for the packed triangle, the streaming kernel above for the diagonal family; otherwise -- and in
    // FP64 -- the generic per-feature kernel.  Slab size of the lists: PS_SLAB for the fast kernels, SH_SEG for the generic.
    const bool pair = pair_stats_supported(dtype, D, F);
    const bool diag = dtype == MIMO_F32 && D >= 2 && F == 2 * D + 1;
    int32_t* kind = bad + 1;
    const int fast_kinds = (pair ? 2 : 0) | (diag ? 4 : 0);                  // bit k: layout k has a fast kernel here
    if (fast_kinds) feature_kind_kernel<<<1, 256, 0, st>>>(fi, fj, F, D, kind);
    label_scan_kernel<<<1, 32, 0, st>>>(counts, K, offsets, cursor, slabs, SH_SEG);
    if (fast_kinds) label_scan_kernel<<<1, 32, 0, st>>>(counts, K, offsets, cursor, slabs_fast, PS_SLAB);
    if (block_sort) label_scatter_block_kernel<<<cdiv(N, LB_TILE), 256, (size_t)K * 8, st>>>(labels, N, K, cursor, perm);
    else label_scatter_kernel<<<grid, 256, 0, st>>>(labels, N, K, cursor, perm);
    MIMO_LAUNCH_CHECK();
    if (check) {
        int32_t hbad = 0;
        MIMO_CUDA(cudaMemcpyAsync(&hbad, bad, 4, cudaMemcpyDeviceToHost, st));
        MIMO_CUDA(cudaStreamSynchronize(st));
        if (hbad) { set_error("labels outside [0, K)"); return MIMO_EINVAL; }
    }
    if (pair) {
        int rc = pair_stats((const float*)Z, D, ldz, perm, offsets, slabs_fast, K, nullptr, 0, nullptr, (const unsigned int*)kind, 1u, stat, F, st);
        if (rc) return rc;
    }
    if (diag) {
        diag_hard_stats_kernel<<<sm_count() * 8, DH_THREADS, 0, st>>>((const float*)Z, D, ldz, perm, offsets, slabs_fast, K, kind, stat, F);
        MIMO_LAUNCH_CHECK();
    }
    size_t es = dtype == MIMO_F32 ? 4 : 8;
    size_t smem = (size_t)SH_PT * (D + 2) * es;
    // every component contributes at most ceil(count/SEG) <= count/SEG + 1 slabs
    dim3 g2((unsigned)(cdiv(N, SH_SEG) + K));
    const int32_t* gk = fast_kinds ? kind : nullptr;
    if (smem > 200 * 1024) { set_error("hard stats: D=%d too large for shared memory", D); return MIMO_EUNSUPPORTED; }
    if (smem > 48 * 1024) {
        if (dtype == MIMO_F32) MIMO_CUDA(cudaFuncSetAttribute(stats_hard_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else MIMO_CUDA(cudaFuncSetAttribute(stats_hard_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (dtype == MIMO_F32)
        stats_hard_kernel<float><<<g2, SH_THREADS, smem, st>>>((const float*)Z, D, ldz, perm, offsets, slabs, K, fi, fj, F, stat, gk, fast_kinds);
    else
        stats_hard_kernel<double><<<g2, SH_THREADS, smem, st>>>((const double*)Z, D, ldz, perm, offsets, slabs, K, fi, fj, F, stat, gk, fast_kinds);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

}  // namespace mimo
