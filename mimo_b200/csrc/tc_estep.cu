// Tensor-core E-step (tcgen05 / TMEM / bulk-copy pipeline), sm_100a.
//
//   a[k][n] = cst[k] - 0.5 * || W_k [z_n ; 1] ||^2        (include/mimo_b200.h, "packed operand form")
//
// replaces the same reference call sites as quad_loglik_kernel (distributions/gaussian.py:510-523,
// lingauss.py:330-347, bayesian.py:287-301, 933-947) for FP32 data with D <= 128.
//
// One persistent CTA per SM walks "point pairs" (2 x 128 points).  Per pair the 256 epilogue
// threads convert the FP32 rows of Z to the 3xFP16 split (tc_common.cuh) and store them as the
// K-major, 128B-swizzled A operand; then for every 128-row chunk of the flattened operand
// matrix (K*Rp rows = whole components) a producer thread bulk-copies the pre-split, pre-swizzled
// B image (built once per sweep by tc_prep_operands_kernel), one thread issues
// 2 tiles x KB x 4 x 3 tcgen05.mma (M=128, N=128, K=16) into a double-buffered TMEM
// accumulator, and the epilogue threads (one point each: TMEM lane = point) add the offset
// column, square, sum over each component's Rp columns and write cst - q/2 to the (K, chunk)
// log-joint scratch with coalesced stores.  The offset W[k][i][D] is applied in the epilogue,
// not in the GEMM, so D = 128 needs no extra K step.
//
// K steps (16 columns) beyond D carry no data and are skipped.  (A variant that also skipped the
// zero lower-triangular blocks of Cholesky-factor operands with narrower MMAs was measured
// slower on B200: with both operands in shared memory a 128 x N x 16 MMA is bound by the
// 4 KB A-operand read, not by N.)
#include <string.h>
#include "tc_common.cuh"
#include "internal.h"

namespace mimo {

using namespace tc;

constexpr int TE_THREADS = 320;          // 8 epilogue / converter warps + MMA warp + producer warp
constexpr int TE_STAGES = 3;             // B pipeline depth, one stage = (chunk, 64-wide K block): hi + lo = 32 KB
constexpr uint32_t TE_TILE_BYTES = 16384;   // 128 rows x 64 FP16
constexpr uint32_t TE_STAGE_BYTES = 2 * TE_TILE_BYTES;

struct TeBarriers {
    uint64_t full[TE_STAGES], empty[TE_STAGES];
    uint64_t tmem_full[2], tmem_empty[2];
    uint64_t a_full;
    uint32_t tmem_base;
};

__global__ void tc_absmax_kernel(const float* __restrict__ Z, int64_t N, int D, int64_t ldz, int flat4, unsigned int* __restrict__ maxbits) {
    float m = 0.f;
    const int64_t total = N * (int64_t)D;
    if (flat4) {                     // contiguous rows, 16-byte aligned: a flat stream of float4
        const float4* z4 = reinterpret_cast<const float4*>(Z);
        const int64_t t4 = total >> 2;
        for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < t4; idx += (int64_t)gridDim.x * blockDim.x) {
            const float4 v = __ldg(z4 + idx);
            m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
        }
    } else {
        for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
            int64_t n = idx / D;
            int j = (int)(idx - n * D);
            m = fmaxf(m, fabsf(Z[n * ldz + j]));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(maxbits, __float_as_uint(m));   // non-negative floats order like their bits
}

// Operand image of one sweep.  grid = chunks of 128 flattened rows (whole components), block = 256.
//   Bimg   [chunk][kb][hi|lo][128 rows][64] FP16, swizzled exactly as the kernel's shared-memory stage
//   rowoff [chunk*128 + r] = W[k][i][D] * sw_k * sz        (offset column in accumulator units)
//   invS2  [k]             = 1 / (sw_k * sz)^2
__global__ void __launch_bounds__(256)
tc_prep_operands_kernel(const float* __restrict__ W, int K, int Rp, int Dpp, int D, int KB,
                        const unsigned int* __restrict__ maxbits,
                        __half* __restrict__ Bimg, float* __restrict__ rowoff, float* __restrict__ invS2) {
    __shared__ unsigned int cmax[16];
    __shared__ float csw[16];
    const int c = blockIdx.x, tid = threadIdx.x;
    const int cpc = 128 / Rp;                      // components per chunk
    const float sz = pow2_scale_for(__uint_as_float(*maxbits));
    if (tid < 16) cmax[tid] = 0u;
    __syncthreads();
    for (int idx = tid; idx < 128 * D; idx += 256) {
        int r = idx / D, j = idx - r * D;
        int64_t flat = (int64_t)c * 128 + r;
        int k = (int)(flat / Rp);
        if (k < K) atomicMax(&cmax[r / Rp], __float_as_uint(fabsf(W[flat * Dpp + j])));
    }
    __syncthreads();
    if (tid < cpc) {
        float sw = pow2_scale_for(__uint_as_float(cmax[tid]));
        csw[tid] = sw;
        int k = c * cpc + tid;
        if (k < K) { float s = sw * sz; invS2[k] = 1.f / (s * s); }
    }
    __syncthreads();
    if (tid < 128) {
        int64_t flat = (int64_t)c * 128 + tid;
        int k = (int)(flat / Rp);
        rowoff[flat] = (k < K) ? W[flat * Dpp + D] * csw[tid / Rp] * sz : 0.f;
    }
    char* img = reinterpret_cast<char*>(Bimg) + (size_t)c * KB * TE_STAGE_BYTES;
    for (int idx = tid; idx < 128 * KB * 8; idx += 256) {
        int r = idx / (KB * 8), ch = idx - r * (KB * 8);      // ch = 16-byte chunk along K (8 elements)
        int64_t flat = (int64_t)c * 128 + r;
        int k = (int)(flat / Rp);
        float sw = csw[r / Rp];
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int j = ch * 8 + e;
            x[e] = (k < K && j < D) ? W[flat * Dpp + j] * sw : 0.f;
        }
        uint4 hi, lo;
        split8(x, hi, lo);
        int kb = ch >> 3, cc = ch & 7;
        char* base = img + (size_t)kb * TE_STAGE_BYTES + sw128_chunk_off(r, cc);
        *reinterpret_cast<uint4*>(base) = hi;
        *reinterpret_cast<uint4*>(base + TE_TILE_BYTES) = lo;
    }
}

template <int RP>
__device__ __forceinline__ void te_consume(const float (&v)[32], const float* __restrict__ off, int col0, float& q,
                                           const float* __restrict__ cst, const float* __restrict__ invS2,
                                           int kbase, int K, bool pvalid, float* __restrict__ outp, int64_t ldo) {
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
        const float4 o = __ldg(reinterpret_cast<const float4*>(off + col0) + j4);
        const float oo[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int col = col0 + j4 * 4 + e;
            const float t = v[j4 * 4 + e] + oo[e];
            q = fmaf(t, t, q);
            if ((col + 1) % RP == 0) {
                const int k = kbase + col / RP;
                if (pvalid && k < K) outp[(int64_t)k * ldo] = __ldg(cst + k) - 0.5f * __ldg(invS2 + k) * q;
                q = 0.f;
            }
        }
    }
}

template <int KB, int RP>
__global__ void __launch_bounds__(TE_THREADS, 1)
tc_estep_kernel(const float* __restrict__ Z, int64_t N, int D, int64_t ldz, int vec4,
                const __half* __restrict__ Bimg, const float* __restrict__ rowoff,
                const float* __restrict__ cst, const float* __restrict__ invS2,
                const unsigned int* __restrict__ maxbits,
                int K, int n_chunks, float* __restrict__ out, int64_t ldo) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // A: [tile 2][hi|lo][kb KB] tiles of 16 KB;  B: [stage][hi|lo] tiles of 16 KB
    unsigned char* sA = smem_raw;
    unsigned char* sB = sA + (size_t)4 * KB * TE_TILE_BYTES;
    TeBarriers* bars = reinterpret_cast<TeBarriers*>(sB + (size_t)TE_STAGES * TE_STAGE_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_pairs = (N + 255) / 256;

    if (tid == 0) {
        for (int s = 0; s < TE_STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bars->tmem_full[b], 1); mbar_init(&bars->tmem_empty[b], 256); }
        mbar_init(&bars->a_full, 256);
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(&bars->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp < 8) {
        // ================= converter + epilogue warps =================
        const float sz = pow2_scale_for(__uint_as_float(__ldg(maxbits)));
        const int t = warp >> 2, qd = warp & 3;
        const int prow = qd * 32 + lane;                         // point row inside the tile = TMEM lane
        uint32_t gc = 0;
        for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
            const int64_t n0 = pair * 256;
            // ---- A operand: rows of Z -> 3xFP16 split, K-major swizzled.  Every MMA of the previous pair
            //      has completed (all threads waited on its last tmem_full), so A may be overwritten. ----
            const int f = lane * 4;                              // this lane's 4 features
            for (int r0 = warp; r0 < 256; r0 += 32) {            // 4 rows in flight per warp
                float x[4][4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = r0 + 8 * u;
                    const int64_t n = n0 + r;
#pragma unroll
                    for (int e = 0; e < 4; ++e) x[u][e] = 0.f;
                    if (n < N && f < D) {
                        const float* src = Z + n * ldz + f;
                        if (vec4 && f + 3 < D) {
                            float4 v = __ldg(reinterpret_cast<const float4*>(src));
                            x[u][0] = v.x; x[u][1] = v.y; x[u][2] = v.z; x[u][3] = v.w;
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) if (f + e < D) x[u][e] = __ldg(src + e);
                        }
                    }
                }
                if (f < KB * 64) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = r0 + 8 * u;
                        const int tt = r >> 7, rr = r & 127;
                        float xs[4] = {x[u][0] * sz, x[u][1] * sz, x[u][2] * sz, x[u][3] * sz};
                        uint2 hi, lo;
                        split4(xs, hi, lo);
                        const int kb = f >> 6, ch = (f & 63) >> 3;
                        unsigned char* base = sA + (size_t)((tt * 2 + 0) * KB + kb) * TE_TILE_BYTES
                                            + sw128_chunk_off(rr, ch) + (f & 7) * 2;
                        *reinterpret_cast<uint2*>(base) = hi;
                        *reinterpret_cast<uint2*>(base + (size_t)KB * TE_TILE_BYTES) = lo;
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(&bars->a_full);

            // ---- epilogue over the chunks of flattened operand rows ----
            const int64_t n = n0 + t * 128 + prow;
            const bool pvalid = n < N;
            float* outp = out + n;
            for (int c = 0; c < n_chunks; ++c, ++gc) {
                const uint32_t buf = gc & 1;
                mbar_wait(&bars->tmem_full[buf], (gc >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + buf * 256 + t * 128;
                const float* off = rowoff + (size_t)c * 128;
                const int kbase = c * (128 / RP);
                float va[32], vb[32];
                float q = 0.f;
                tmem_ld32(taddr, va);
                tmem_ld_wait();
                tmem_ld32(taddr + 32, vb);
                te_consume<RP>(va, off, 0, q, cst, invS2, kbase, K, pvalid, outp, ldo);
                tmem_ld_wait();
                tmem_ld32(taddr + 64, va);
                te_consume<RP>(vb, off, 32, q, cst, invS2, kbase, K, pvalid, outp, ldo);
                tmem_ld_wait();
                tmem_ld32(taddr + 96, vb);
                te_consume<RP>(va, off, 64, q, cst, invS2, kbase, K, pvalid, outp, ldo);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&bars->tmem_empty[buf]);             // accumulator drained: MMA may reuse it
                te_consume<RP>(vb, off, 96, q, cst, invS2, kbase, K, pvalid, outp, ldo);
            }
        }
    } else if (warp == 8) {
        // ================= MMA issuer (one thread) =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_f16(128, 128);
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
            const int S = (D + 15) >> 4;                                  // 16-wide K steps that hold data
            uint32_t stage = 0, phase = 0, gc = 0, it = 0;
            for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++it) {
                mbar_wait(&bars->a_full, it & 1);
                tc_fence_after();
                for (int c = 0; c < n_chunks; ++c, ++gc) {
                    const uint32_t buf = gc & 1;
                    mbar_wait(&bars->tmem_empty[buf], ((gc >> 1) & 1) ^ 1);
                    tc_fence_after();
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&bars->full[stage], phase);
                        tc_fence_after();
                        const uint64_t bh = make_desc_sw128(b0 + stage * TE_STAGE_BYTES);
                        const uint64_t bl = make_desc_sw128(b0 + stage * TE_STAGE_BYTES + TE_TILE_BYTES);
#pragma unroll
                        for (int tt = 0; tt < 2; ++tt) {
                            const uint64_t ah = make_desc_sw128(a0 + ((tt * 2 + 0) * KB + kb) * TE_TILE_BYTES);
                            const uint64_t al = make_desc_sw128(a0 + ((tt * 2 + 1) * KB + kb) * TE_TILE_BYTES);
                            const uint32_t d = tmem_base + buf * 256 + tt * 128;
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {           // 16-element K steps inside the 64-wide block: +32 B
                                if (kb * 4 + kk >= S) continue;
                                umma_f16(d, al + 2 * kk, bh + 2 * kk, idesc, (kb | kk) != 0);
                                umma_f16(d, ah + 2 * kk, bl + 2 * kk, idesc, 1);
                                umma_f16(d, ah + 2 * kk, bh + 2 * kk, idesc, 1);
                            }
                        }
                        umma_commit(&bars->empty[stage]);                // stage free once these MMAs have read it
                        if (++stage == TE_STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&bars->tmem_full[buf]);
                }
            }
        }
    } else {
        // ================= B producer (one thread, TMA engine bulk copies) =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                const unsigned char* src = reinterpret_cast<const unsigned char*>(Bimg);
                for (int s = 0; s < n_chunks * KB; ++s) {
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&bars->full[stage], TE_STAGE_BYTES);
                    bulk_g2s(sB + (size_t)stage * TE_STAGE_BYTES, src + (size_t)s * TE_STAGE_BYTES, TE_STAGE_BYTES, &bars->full[stage]);
                    if (++stage == TE_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// ---- host side -------------------------------------------------------------------------------

static size_t a256(size_t x) { return (x + 255) / 256 * 256; }

bool tc_estep_supported(int dtype, int D, int Rp) {
    return dtype == MIMO_F32 && D >= 1 && D <= 128 && Rp >= 8 && Rp <= 128 && (Rp & (Rp - 1)) == 0;
}

struct TcOperandLayout {
    int KB, n_chunks;                  // n_chunks: 128-row chunks, rounded up to an even count (CTA-pair kernel), zero padded
    size_t off_maxbits, off_invS2, off_rowoff, off_offs2, off_img, off_t3, off_t4, bytes;   // off_t3 / off_t4: image + offsets of tc_estep3.cu / tc_estep4.cu (always reserved for their shapes)
};
static TcOperandLayout tc_layout(int K, int Rp, int D) {
    TcOperandLayout L;
    L.KB = D <= 64 ? 1 : 2;
    L.n_chunks = (int)(((int64_t)K * Rp + 127) / 128);
    L.n_chunks = (L.n_chunks + 1) / 2 * 2;
    size_t o = 0;
    L.off_maxbits = o; o += 256;
    L.off_invS2 = o;   o += a256((size_t)K * 4);
    L.off_rowoff = o;  o += a256((size_t)L.n_chunks * 128 * 4);
    L.off_offs2 = o;   o += a256(tc2_offsets_bytes(K, Rp));
    o = (o + 1023) / 1024 * 1024;
    L.off_img = o;     o += (size_t)L.n_chunks * L.KB * TE_STAGE_BYTES;
    L.off_t3 = o;      if (D > 64 && Rp == 128) o += tc3_workspace(K);
    o = (o + 1023) / 1024 * 1024;
    L.off_t4 = o;      if (D > 64 && Rp == 128) o += tc4_workspace(K);
    L.bytes = o;
    return L;
}

size_t tc_operand_workspace(int K, int Rp, int D) { return tc_layout(K, Rp, D).bytes + 1024; }

static char* align1k(void* p) { return (char*)(((uintptr_t)p + 1023) / 1024 * 1024); }

// A caller that keeps the data resident across sweeps may pass max |Z| of that data (mimo_sweep_absmax_hint): the next
// SWEEP of this thread then skips the pass over Z below (25.6 GB = 4.5 ms per sweep at N = 100M, d = 64).  One-shot, and
// scoped to that sweep: sweep() takes the value on entry whatever kernels it ends up running (a sweep that stays on the
// CUDA cores used to leave it behind, and the next tensor-core call -- on other data -- scaled its operands with it) and
// hands it to tc_data_scale explicitly; the stand-alone entry points never see it.
static thread_local float g_absmax_hint = 0.f;
void tc_set_absmax_hint(float v) { g_absmax_hint = (v > 0.f && v < 3.0e38f) ? v : 0.f; }
float tc_take_absmax_hint() { const float v = g_absmax_hint; g_absmax_hint = 0.f; return v; }
__global__ void tc_store_bits_kernel(unsigned int* dst, unsigned int bits) { *dst = bits; }

// max |Z| over the resident data -> ws (the common power-of-two data scale of a sweep)
int tc_data_scale(const float* Z, int64_t N, int D, int64_t ldz, void* ws, cudaStream_t st, float absmax_hint) {
    char* base = align1k(ws);
    MIMO_CUDA(cudaMemsetAsync(base, 0, 256, st));
    if (absmax_hint > 0.f) {
        unsigned int bits;
        memcpy(&bits, &absmax_hint, 4);
        tc_store_bits_kernel<<<1, 1, 0, st>>>((unsigned int*)base, bits);
        MIMO_LAUNCH_CHECK();
        return MIMO_OK;
    }
    if (N > 0) {
        int64_t total = N * (int64_t)D;
        int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
        const int flat4 = (ldz == D) && (total % 4 == 0) && (((uintptr_t)Z & 15) == 0);
        tc_absmax_kernel<<<grid, 256, 0, st>>>(Z, N, D, ldz, flat4, (unsigned int*)base);
        MIMO_LAUNCH_CHECK();
    }
    return MIMO_OK;
}

const unsigned int* tc_maxbits(void* ws) { return (const unsigned int*)align1k(ws); }

// operand image for the current W and cst (after tc_data_scale on the same ws)
int tc_prepare_operands(const float* W, const float* cst, int K, int Rp, int Dpp, int D, void* ws, cudaStream_t st) {
    TcOperandLayout L = tc_layout(K, Rp, D);
    char* base = align1k(ws);
    tc_prep_operands_kernel<<<L.n_chunks, 256, 0, st>>>(W, K, Rp, Dpp, D, L.KB, (const unsigned int*)(base + L.off_maxbits),
                                                        (__half*)(base + L.off_img), (float*)(base + L.off_rowoff),
                                                        (float*)(base + L.off_invS2));
    MIMO_LAUNCH_CHECK();
    if (tc3_supported(D, Rp)) {
        int rc = tc3_prepare(W, cst, K, Dpp, D, (unsigned int*)(base + L.off_maxbits), base + L.off_t3, st);
        if (rc) return rc;
    }
    if (tc4_supported(D, Rp)) {
        int rc = tc4_prepare(W, cst, K, Dpp, D, (unsigned int*)(base + L.off_maxbits), base + L.off_t4, st);
        if (rc) return rc;
    }
    // per-chunk offsets / constants blocks of the CTA-pair kernel
    return tc2_prepare_offsets((const float*)(base + L.off_rowoff), (const float*)(base + L.off_invS2), cst, K, Rp,
                               (float*)(base + L.off_offs2), st);
}

unsigned int* tc_flags(void* ws) { return (unsigned int*)align1k(ws); }

// CTA-pair kernel with an explicit number of passes and an optional device-side gate (tc_screen.cu)
int tc_estep_pass(const float* Z, int64_t N, int D, int64_t ldz, int K, int Rp, float* out, int64_t ldo, void* ws,
                  int passes, const unsigned int* gate, unsigned int gate_value, float* lower, int* guess, int64_t ldl, cudaStream_t st,
                  float* lse_vals, double* lse_sum) {
    if (N == 0) return MIMO_OK;
    TcOperandLayout L = tc_layout(K, Rp, D);
    char* base = align1k(ws);
    if (passes == 3 && tc3_supported(D, Rp))
        return tc_estep3(Z, N, D, ldz, K, base + L.off_t3, (const unsigned int*)(base + L.off_maxbits), out, ldo, gate, gate_value, lse_vals, lse_sum, st);
    if (passes == 3 && tc4_supported(D, Rp))
        return tc_estep4(Z, N, D, ldz, K, base + L.off_t4, (const unsigned int*)(base + L.off_maxbits), out, ldo, gate, gate_value, lse_vals, lse_sum, st);
    return tc_estep2(Z, N, D, ldz, K, Rp, L.KB, (const void*)(base + L.off_img), (const float*)(base + L.off_offs2),
                     (const unsigned int*)(base + L.off_maxbits), out, ldo, passes, gate, gate_value, lower, guess, ldl, st, lse_vals, lse_sum);
}

template <int KB, int RP>
static int launch_estep(const float* Z, int64_t N, int D, int64_t ldz, const TcOperandLayout& L, char* base,
                        const float* cst, int K, float* out, int64_t ldo, cudaStream_t st) {
    size_t smem = (size_t)4 * KB * TE_TILE_BYTES + (size_t)TE_STAGES * TE_STAGE_BYTES + sizeof(TeBarriers) + 1024;
    auto kern = tc_estep_kernel<KB, RP>;
    MIMO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t pairs = (N + 255) / 256;
    int grid = (int)std::min<int64_t>(pairs, sm_count());
    int vec4 = (ldz % 4 == 0) && (((uintptr_t)Z & 15) == 0);
    kern<<<grid, TE_THREADS, smem, st>>>(Z, N, D, ldz, vec4, (const __half*)(base + L.off_img), (const float*)(base + L.off_rowoff),
                                         cst, (const float*)(base + L.off_invS2), (const unsigned int*)(base + L.off_maxbits),
                                         K, L.n_chunks, out, ldo);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

// E-step over N points with a prepared operand image
int tc_estep(const float* Z, int64_t N, int D, int64_t ldz, const float* cst, int K, int Rp,
             float* out, int64_t ldo, void* ws, cudaStream_t st, float* lse_vals, double* lse_sum) {
    if (N == 0) return MIMO_OK;
    TcOperandLayout L = tc_layout(K, Rp, D);
    char* base = align1k(ws);
    if (tc_mode() != 2 && tc3_supported(D, Rp))     // points operand in tensor memory, triangular skip
        return tc_estep3(Z, N, D, ldz, K, base + L.off_t3, (const unsigned int*)(base + L.off_maxbits), out, ldo, nullptr, 0u, lse_vals, lse_sum, st);
    if (tc_mode() != 2 && tc4_supported(D, Rp))     // four components per generation, zero block skipped at full MMA width
        return tc_estep4(Z, N, D, ldz, K, base + L.off_t4, (const unsigned int*)(base + L.off_maxbits), out, ldo, nullptr, 0u, lse_vals, lse_sum, st);
    // CTA-pair kernel (cta_group::2), dense 3-pass; the single-CTA kernel (mode 2) reads the plain image only
    if (tc_mode() != 2 || lse_vals)
        return tc_estep2(Z, N, D, ldz, K, Rp, L.KB, (const void*)(base + L.off_img), (const float*)(base + L.off_offs2),
                         (const unsigned int*)(base + L.off_maxbits), out, ldo, 3, nullptr, 0u, nullptr, nullptr, 0, st, lse_vals, lse_sum);
#define TE_CASE(kb, rp) if (L.KB == kb && Rp == rp) return launch_estep<kb, rp>(Z, N, D, ldz, L, base, cst, K, out, ldo, st);
    TE_CASE(1, 8) TE_CASE(1, 16) TE_CASE(1, 32) TE_CASE(1, 64) TE_CASE(1, 128)
    TE_CASE(2, 8) TE_CASE(2, 16) TE_CASE(2, 32) TE_CASE(2, 64) TE_CASE(2, 128)
#undef TE_CASE
    set_error("tensor-core E-step: unsupported shape D=%d Rp=%d", D, Rp);
    return MIMO_EUNSUPPORTED;
}

// stand-alone entry (mimo_loglik_quad_tc): scale + operand image + E-step
int loglik_quad_tc(const void* Z, int64_t N, int D, int64_t ldz, const void* W, const void* cst,
                   int K, int Rp, int Dpp, void* out, int64_t ldo, void* ws, size_t ws_bytes, cudaStream_t st) {
    MIMO_CHECK_ARG(Z && W && cst && out && ws, "null pointer");
    MIMO_CHECK_ARG(N >= 0 && K >= 1 && ldz >= D && ldo >= N, "shape");
    MIMO_CHECK_ARG(Dpp >= D + 1, "Dpp");
    if (!tc_estep_supported(MIMO_F32, D, Rp)) { set_error("tensor-core E-step: unsupported shape D=%d Rp=%d", D, Rp); return MIMO_EUNSUPPORTED; }
    MIMO_CHECK_ARG(ws_bytes >= tc_operand_workspace(K, Rp, D), "workspace too small");
    int rc = tc_data_scale((const float*)Z, N, D, ldz, ws, st);
    if (rc) return rc;
    rc = tc_prepare_operands((const float*)W, (const float*)cst, K, Rp, Dpp, D, ws, st);
    if (rc) return rc;
    return tc_estep((const float*)Z, N, D, ldz, (const float*)cst, K, Rp, (float*)out, ldo, ws, st);
}

}  // namespace mimo
