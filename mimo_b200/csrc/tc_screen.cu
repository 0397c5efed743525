// Screened E-step: one FP16 tensor-core pass over ALL (point, component) pairs, an exact pass only over the
// pairs that can matter.
//
//   a[k][n] = cst[k] - 0.5 * q[k][n],   q = || W_k [z_n ; 1] ||^2
//   (distributions/gaussian.py:510-523, bayesian.py:287-301 followed by the logsumexp of mixtures/gmm.py:72-75, 256-259)
//
// The responsibilities and the log-normaliser only depend on the components within a few tens of nats of a
// point's best component.  The single-pass kernel (tc_estep2.cu, PASSES = 1: operands rounded to FP16, a
// third of the tensor-pipe work) returns q~ with a RIGOROUS error bound: with y = W z and y~ its FP16-operand
// value,  || y~ - y ||_2 <= B' = 1.05 * 2^-10 * max_k ||W_k||_F * max_n ||z_n||_2 + 1e-3   (each product carries two
// roundings of 2^-11; Cauchy-Schwarz over the row, then over the rows), hence | sqrt(q~) - sqrt(q) | <= B'.
// Per point:   lower bound of the best log-joint   L_n = max_k [ cst_k - (sqrt(q~) + B')^2 / 2 ]
//              upper bound of component k           U_kn = cst_k - max(0, sqrt(q~) - B')^2 / 2
// and (n, k) is a CANDIDATE iff U_kn >= L_n - 40.  Every non-candidate has a true log-joint more than 40 nats below
// the true maximum: all of them together change exp-sums by < K e^-40 relative, far below FP32 resolution, so their
// single-pass values are kept.  Candidates are grouped by component (counting sort) and recomputed in FP32 on the
// CUDA cores (one component's operand block in shared memory, 4 x 4 register tiles), overwriting the scratch.
// When more than 4 % of the pairs are candidates (overlapping components, early sweeps) the refinement would cost
// more than it saves: a device-side flag then makes the dense 3-pass tensor-core kernel run instead and the
// refinement kernels return immediately -- no host round trip either way.
#include <algorithm>
#include "tc_common.cuh"
#include "internal.h"

namespace mimo {

using tc::screen_bound;

constexpr float SCREEN_T0 = 40.f;
constexpr int RF_THREADS = 256;
constexpr int RF_SPLIT = 8;                  // blocks per component

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

// max_k ||W_k[:, :D]||_F (the offset column is applied exactly in the epilogue and carries no rounding)
__global__ void screen_wnorm_kernel(const float* __restrict__ W, int K, int Rp, int Dpp, int D, unsigned int* __restrict__ flags) {
    __shared__ float red[32];
    const int k = blockIdx.x;
    float s = 0.f;
    for (int idx = threadIdx.x; idx < Rp * D; idx += blockDim.x) {
        const int i = idx / D, j = idx - i * D;
        const float w = W[((size_t)k * Rp + i) * Dpp + j];
        s = fmaf(w, w, s);
    }
    s = block_sum<float>(s, red);
    if (threadIdx.x == 0) atomicMax(flags + 2, __float_as_uint(sqrtf(s) * 1.0000005f));
}

// counters: [0] candidates found, [1] dense flag (set by screen_scan_kernel)
__global__ void __launch_bounds__(256)
screen_emit_kernel(const float* __restrict__ a, int K, int64_t n, int64_t ldo, const float* __restrict__ cst,
                   const unsigned int* __restrict__ flags, const float* __restrict__ lower, int64_t ldl,
                   int2* __restrict__ list, unsigned int cap, unsigned int* __restrict__ counters, int* __restrict__ hist) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const float B = screen_bound(flags);
    // L_n (lower bound of the point's best log-joint) was formed by the single-pass E-step epilogue, one value per
    // half of the accumulator columns
    const float t = valid ? fmaxf(lower[i], lower[ldl + i]) - SCREEN_T0 : INFINITY;
    const int lane = threadIdx.x & 31;
    for (int k0 = 0; k0 < K; k0 += 4) {
      float av[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) av[u] = (valid && k0 + u < K) ? a[(int64_t)(k0 + u) * ldo + i] : 0.f;   // 4 loads in flight
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + u;
        if (k >= K) break;
        bool cand = false;
        if (valid) {
            const float c = __ldg(cst + k);
            const float s = fmaxf(0.f, sqrtf(fmaxf(0.f, 2.f * (c - av[u]))) - B);
            cand = (c - 0.5f * s * s) >= t;
        }
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

// single block: exclusive scan of hist -> offsets[K+1], cursor := offsets; dense flag when too many candidates
__global__ void screen_scan_kernel(const int* __restrict__ hist, int K, int* __restrict__ offsets, int* __restrict__ cursor,
                                   int* __restrict__ slabs, unsigned int* __restrict__ counters, unsigned int max_cands) {
    if (threadIdx.x == 0) {
        int run = 0, items = 0;                           // slabs: work items (PS_SLAB listed points) of pair_stats.cu
        for (int k = 0; k < K; ++k) {
            offsets[k] = run; cursor[k] = run; slabs[k] = items;
            run += hist[k];
            items += (hist[k] + PS_SLAB - 1) / PS_SLAB;
        }
        offsets[K] = run;
        slabs[K] = items;
        counters[1] = (counters[0] > max_cands) ? 1u : 0u;
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

// Exact (FP32 FMA) recomputation of the candidates of one component.  grid = (K, RF_SPLIT).
// Thread (rg, cg): rows rg + RG r (r < 4), candidates 4 cg + c (c < 4) of a tile of TILE_C candidates.
template <int RP>
__global__ void __launch_bounds__(RF_THREADS)
screen_refine_kernel(const float* __restrict__ Z, int D, int64_t ldz,
                     const float* __restrict__ W, int Dpp, const float* __restrict__ cst,
                     const int* __restrict__ perm, const int* __restrict__ offsets, const unsigned int* __restrict__ counters,
                     float* __restrict__ a, int64_t ldo) {
    if (counters[1] != 0u) return;
    constexpr int RG = RP / 4, CG = RF_THREADS / RG, TILE_C = 4 * CG;
    const int k = blockIdx.x;
    const int beg = offsets[k], cnt = offsets[k + 1] - beg;
    const int tiles = (cnt + TILE_C - 1) / TILE_C;
    if ((int)blockIdx.y >= tiles) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Ws = reinterpret_cast<float*>(smem_raw);            // [RP][Dpp]
    float* Zs = Ws + (size_t)RP * Dpp;                         // [TILE_C][Dpp]   (column D = 1, beyond = 0)
    const int tid = threadIdx.x, rg = tid % RG, cg = tid / RG;
    for (int idx = tid; idx < RP * Dpp / 4; idx += RF_THREADS)
        reinterpret_cast<float4*>(Ws)[idx] = __ldg(reinterpret_cast<const float4*>(W + (size_t)k * RP * Dpp) + idx);
    const float ck = __ldg(cst + k);
    for (int t = blockIdx.y; t < tiles; t += gridDim.y) {
        const int c0 = beg + t * TILE_C, nc = min(TILE_C, beg + cnt - c0);
        __syncthreads();                                       // Ws ready / previous tile consumed
        for (int idx = tid; idx < TILE_C * Dpp; idx += RF_THREADS) {
            const int c = idx / Dpp, j = idx - c * Dpp;
            float v = 0.f;
            if (c < nc) v = (j < D) ? __ldg(Z + (int64_t)perm[c0 + c] * ldz + j) : (j == D ? 1.f : 0.f);
            Zs[idx] = v;
        }
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
        for (int j = 0; j < Dpp; j += 4) {
            float4 w[4], z[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) w[r] = *reinterpret_cast<const float4*>(Ws + (size_t)(rg + RG * r) * Dpp + j);
#pragma unroll
            for (int c = 0; c < 4; ++c) z[c] = *reinterpret_cast<const float4*>(Zs + (size_t)(4 * cg + c) * Dpp + j);
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    acc[r][c] = fmaf(w[r].x, z[c].x, acc[r][c]);
                    acc[r][c] = fmaf(w[r].y, z[c].y, acc[r][c]);
                    acc[r][c] = fmaf(w[r].z, z[c].z, acc[r][c]);
                    acc[r][c] = fmaf(w[r].w, z[c].w, acc[r][c]);
                }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float q = 0.f;
#pragma unroll
            for (int r = 0; r < 4; ++r) q = fmaf(acc[r][c], acc[r][c], q);
#pragma unroll
            for (int o = RG / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);    // over the RG lanes of this candidate group
            const int ci = 4 * cg + c;
            if (rg == 0 && ci < nc) a[(int64_t)k * ldo + perm[c0 + ci]] = ck - 0.5f * q;
        }
    }
}

// ---- host side -------------------------------------------------------------------------------

static size_t a256(size_t x) { return (x + 255) / 256 * 256; }

bool tc_screen_supported(int D, int Rp) { return D >= 24 && D <= 128 && (Rp == 32 || Rp == 64 || Rp == 128); }

struct ScreenLayout { unsigned int cap; size_t off_counters, off_hist, off_offsets, off_cursor, off_slabs, off_thr, off_list, off_perm, bytes; };
static ScreenLayout screen_layout(int64_t chunk_points, int K) {
    ScreenLayout L;
    const double pairs = (double)chunk_points * K;
    L.cap = (unsigned int)std::min<double>(2.0e9, 0.05 * pairs + 1024.0);
    size_t o = 0;
    L.off_counters = o; o += 256;
    L.off_hist = o;     o += a256((size_t)(K + 1) * 4);
    L.off_offsets = o;  o += a256((size_t)(K + 1) * 4);
    L.off_cursor = o;   o += a256((size_t)(K + 1) * 4);
    L.off_slabs = o;    o += a256((size_t)(K + 1) * 4);
    L.off_thr = o;      o += 2 * a256((size_t)chunk_points * 4);      // lower bounds, one row per accumulator half
    L.off_list = o;     o += a256((size_t)L.cap * 8);
    L.off_perm = o;     o += a256((size_t)L.cap * 4);
    L.bytes = o;
    return L;
}
static char* align256(void* p) { return (char*)(((uintptr_t)p + 255) / 256 * 256); }
size_t tc_screen_workspace(int64_t chunk_points, int K) { return screen_layout(chunk_points, K).bytes + 256; }

// once per sweep, after tc_data_scale (which zeroes the flags): the two norms of the error bound
int tc_screen_prepare(const float* Z, int64_t N, int D, int64_t ldz, const float* W, int K, int Rp, int Dpp,
                      unsigned int* flags, cudaStream_t st) {
    if (N > 0) {
        const int grid = (int)std::min<int64_t>((N + 7) / 8, (int64_t)sm_count() * 16);
        screen_rownorm_kernel<<<grid, 256, 0, st>>>(Z, N, D, ldz, flags);
        MIMO_LAUNCH_CHECK();
    }
    screen_wnorm_kernel<<<K, 256, 0, st>>>(W, K, Rp, Dpp, D, flags);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

// after the single-pass E-step of a chunk: find the candidates; sets the device flag the dense pass is gated on
// the two rows (leading dimension *ldl) the single-pass E-step writes its per-point lower bounds into
float* tc_screen_lower(void* ws, int64_t plan_points, int K, int64_t* ldl) {
    *ldl = (int64_t)(a256((size_t)plan_points * 4) / 4);
    return (float*)(align256(ws) + screen_layout(plan_points, K).off_thr);
}

const unsigned int* tc_screen_gate(void* ws, int64_t plan_points, int K) {
    return (const unsigned int*)(align256(ws) + screen_layout(plan_points, K).off_counters) + 1;
}

// the candidate lists of the chunk, grouped by component (valid after tc_screen_refine when the dense flag is clear)
void tc_screen_lists(void* ws, int64_t plan_points, int K, const int32_t** perm, const int32_t** offsets, const int32_t** slabs) {
    ScreenLayout L = screen_layout(plan_points, K);
    char* base = align256(ws);
    *perm = (const int32_t*)(base + L.off_perm);
    *offsets = (const int32_t*)(base + L.off_offsets);
    *slabs = (const int32_t*)(base + L.off_slabs);
}

static const unsigned int* g_last_counters = nullptr;

// {candidates, dense flag} of the most recent screened chunk (synchronises the device; tests / bench reporting)
int tc_screen_last(unsigned int* out_host2) {
    out_host2[0] = out_host2[1] = 0u;
    if (!g_last_counters) return MIMO_OK;
    MIMO_CUDA(cudaDeviceSynchronize());
    MIMO_CUDA(cudaMemcpy(out_host2, g_last_counters, 8, cudaMemcpyDeviceToHost));
    return MIMO_OK;
}

int tc_screen_select(const float* a, int K, int64_t n, int64_t ldo, const float* cst, const unsigned int* flags,
                     int64_t plan_points, void* ws, cudaStream_t st) {
    ScreenLayout L = screen_layout(plan_points, K);
    char* base = align256(ws);
    unsigned int* counters = (unsigned int*)(base + L.off_counters);
    g_last_counters = counters;
    int* hist = (int*)(base + L.off_hist);
    MIMO_CUDA(cudaMemsetAsync(base, 0, L.off_offsets, st));                 // counters + hist
    const int grid = cdiv(n, 256);
    screen_emit_kernel<<<grid, 256, 0, st>>>(a, K, n, ldo, cst, flags, (const float*)(base + L.off_thr),
                                             (int64_t)(a256((size_t)plan_points * 4) / 4),
                                             (int2*)(base + L.off_list), L.cap, counters, hist);
    const double maxc = std::min<double>((double)L.cap, 0.04 * (double)n * K);
    screen_scan_kernel<<<1, 32, 0, st>>>(hist, K, (int*)(base + L.off_offsets), (int*)(base + L.off_cursor), (int*)(base + L.off_slabs), counters,
                                         (unsigned int)maxc);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

// exact values of the candidates (returns immediately on the device when the dense flag is set)
int tc_screen_refine(const float* Z, int D, int64_t ldz, const float* W, int K, int Rp, int Dpp, const float* cst,
                     float* a, int64_t ldo, int64_t plan_points, void* ws, cudaStream_t st) {
    ScreenLayout L = screen_layout(plan_points, K);
    char* base = align256(ws);
    const unsigned int* counters = (const unsigned int*)(base + L.off_counters);
    int* perm = (int*)(base + L.off_perm);
    screen_scatter_kernel<<<sm_count() * 4, 256, 0, st>>>((const int2*)(base + L.off_list), counters, (int*)(base + L.off_cursor), perm);
    MIMO_LAUNCH_CHECK();
    const int* offsets = (const int*)(base + L.off_offsets);
    dim3 grid(K, RF_SPLIT);
#define RF_CASE(rp)                                                                                                   \
    if (Rp == rp) {                                                                                                   \
        constexpr int TILE_C = 4 * (RF_THREADS / (rp / 4));                                                           \
        const size_t smem = (size_t)(rp + TILE_C) * Dpp * sizeof(float);                                              \
        MIMO_CUDA(cudaFuncSetAttribute(screen_refine_kernel<rp>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        screen_refine_kernel<rp><<<grid, RF_THREADS, smem, st>>>(Z, D, ldz, W, Dpp, cst, perm, offsets, counters, a, ldo); \
        MIMO_LAUNCH_CHECK();                                                                                          \
        return MIMO_OK;                                                                                               \
    }
    RF_CASE(32) RF_CASE(64) RF_CASE(128)
#undef RF_CASE
    set_error("screened E-step: unsupported Rp=%d", Rp);
    return MIMO_EUNSUPPORTED;
}

}  // namespace mimo
