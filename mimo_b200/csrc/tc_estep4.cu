// Dense tensor-core E-step for 64 < D <= 128, Rp = 128 on CTA pairs, FOUR components per accumulator generation with the
// zero block of the Cholesky factors skipped at full MMA width, sm_100a.
//
//   a[k][n] = cst[k] - 0.5 * || W_k [z_n ; 1] ||^2        (include/mimo_b200.h, "packed operand form")
//
// replaces the same reference call sites as tc_estep2.cu (distributions/gaussian.py:510-523, bayesian.py:287-301,
// 933-947) on the dense path of the cfg5 shape (N = 50M, d = 128, K = 1024).
//
// tc_estep2.cu runs its 24 MMAs per component pair at 99 % tensor-pipe activity (profiles/r02_ncu_dense_cfg5.md): the
// kernel is MMA-bound, and a quarter of that work multiplies zeros -- the rows of W_k are Cholesky factors (U_k,
// sqrt(nu_k) C_k^T), row r is zero left of column r, so rows 64..127 contribute nothing in the first 64-wide K block.
// Narrower MMAs for that block gain nothing (a 256 x N x 16 MMA costs the same for N = 128 and N = 256, with the points
// operand in shared memory or in tensor memory: profiles/r02_estep_tmem_operand.md).  Here every MMA keeps N = 256:
// a generation holds FOUR components in the 512 TMEM columns, split by row half,
//     region F = columns [  0, 256): rows  0..63  of components c0 c1 c2 c3   (needs K blocks 1 and 0)
//     region S = columns [256, 512): rows 64..127 of components c0 c1 c2 c3   (needs K block 1 only)
// so a generation costs 3 steps of 12 MMAs (F x K block 1, F x K block 0, S x K block 1) instead of 4, and 3 operand
// tiles per CTA instead of 4.  The two regions are the double buffer: the epilogue drains F while the tensor pipe fills
// S and drains S while the next generation's F is filled; a thread carries its partial squared norm of one component
// from the F phase to the S phase in registers.  Operands that are not triangular (stacked ILR blocks; checked on the
// device, flags[9]) get the fourth step.  Same CTA-pair protocol, operand split and fused log-normaliser as tc_estep2.cu.
//
// MEASURED (B200, cfg5 shape, profiles/r02_estep_quad_generations.md): correct (parity 5e-7, 12 shapes), 25 % fewer MMAs
// and operand bytes -- and SLOWER: 58.7 ms (8 epilogue warps) / 60.7 ms (16) per 1 M-point chunk against 55.4 ms for
// tc_estep2.cu.  Region F has to be read out (128 lanes x 256 columns x 4 B = 2048 clk of the 64 B/clk tensor-memory
// read port, plus two barrier hops) inside the 1536 clk the S step takes, so the issuer stalls every generation, and the
// read-back of a component (1024 clk) is now as long as its MMAs (1152 clk): the two no longer hide each other.  The
// dense E-step is within ~10 % of what the port and the pipe allow together; the kernel stays OFF by default
// (mimo_tc_set_quad_generations(1) selects it; tests/test_gpu_tc.py runs both).
#include <algorithm>
#include <string.h>
#include "tc_common.cuh"
#include "internal.h"

namespace mimo {

using namespace tc;

constexpr int T4_NW = 16;                     // converter / epilogue warps: warp w reads TMEM lanes 32 (w % 4) .. and component w / 4 of a generation
constexpr int T4_THREADS = 32 * (T4_NW + 2);  // + MMA (relay) warp + producer warp
constexpr int T4_STAGES = 4;                  // B ring: one stage = this CTA's 128 rows x 64 K of one step, hi | lo
constexpr uint32_t T4_TILE = 16384;           // 128 rows x 64 FP16
constexpr uint32_t T4_STAGE = 2 * T4_TILE;
constexpr int T4_OFFBLK = 520;                // floats per generation: 512 row offsets (TMEM column order) | 4 cst | 4 1/scale^2
constexpr uint32_t T4_OFFBYTES = T4_OFFBLK * 4;
constexpr int T4_OFFRING = 4;

struct T4Bars {
    uint64_t full[T4_STAGES], empty[T4_STAGES], peer_full[T4_STAGES];
    uint64_t reg_full[2], reg_empty[2];       // region F / S of the accumulator
    uint64_t a_full, peer_a_full;
    uint64_t off_full[T4_OFFRING], off_empty[T4_OFFRING];
    uint32_t tmem_base;
    float2 comb[3][128];
};

// ---- operand image: grid = generations of 4 components, block = 256 ---------------------------------------------
// img [gen][rank][step F1 | F0 | S1 | S0][hi|lo][128 rows][64]: CTA `rank` supplies components 2 rank, 2 rank + 1 of the
// generation (64 rows each per step);  offs [gen][T4_OFFBLK]
__global__ void __launch_bounds__(256)
tc4_prep_kernel(const float* __restrict__ W, const float* __restrict__ cst, int K, int Dpp, int D,
                unsigned int* __restrict__ flags, unsigned char* __restrict__ img, float* __restrict__ offs) {
    __shared__ unsigned int cmax[4];
    __shared__ float csw[4];
    const int g = blockIdx.x, tid = threadIdx.x;
    const float sz = pow2_scale_for(__uint_as_float(flags[0]));
    if (tid < 4) cmax[tid] = 0u;
    __syncthreads();
    for (int idx = tid; idx < 4 * 128 * D; idx += 256) {
        const int ci = idx / (128 * D), rem = idx - ci * 128 * D;
        const int r = rem / D, j = rem - r * D;
        const int k = 4 * g + ci;
        if (k < K) atomicMax(&cmax[ci], __float_as_uint(fabsf(W[((size_t)k * 128 + r) * Dpp + j])));
    }
    __syncthreads();
    if (tid < 4) csw[tid] = pow2_scale_for(__uint_as_float(cmax[tid]));
    __syncthreads();
    float* ob = offs + (size_t)g * T4_OFFBLK;
    for (int col = tid; col < 512; col += 256) {
        const int region = col >> 8, ci = (col >> 6) & 3, r = (region << 6) | (col & 63);
        const int k = 4 * g + ci;
        ob[col] = k < K ? W[((size_t)k * 128 + r) * Dpp + D] * csw[ci] * sz : 0.f;
    }
    if (tid < 4) {
        const int k = 4 * g + tid;
        const float s = csw[tid] * sz;
        ob[512 + tid] = k < K ? cst[k] : 0.f;
        ob[516 + tid] = k < K ? 1.f / (s * s) : 0.f;
    }
    bool below = false;
    // (component ci, row r, 16-byte chunk ch of 8 K elements): 4 x 128 x 16 items
    for (int idx = tid; idx < 4 * 128 * 16; idx += 256) {
        const int ch = idx & 15, r = (idx >> 4) & 127, ci = idx >> 11;
        const int k = 4 * g + ci;
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int j = 8 * ch + e;
            x[e] = (k < K && j < D) ? W[((size_t)k * 128 + r) * Dpp + j] * csw[ci] : 0.f;
        }
        const int kb = ch >> 3, region = r >> 6;
        if (region == 1 && kb == 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) below |= (x[e] != 0.f);
        }
        uint4 hi, lo;
        split8(x, hi, lo);
        const int step = region == 0 ? (kb == 1 ? 0 : 1) : (kb == 1 ? 2 : 3);
        const int rank = ci >> 1, lrow = (ci & 1) * 64 + (r & 63);
        unsigned char* base = img + (((size_t)g * 2 + rank) * 4 + step) * T4_STAGE + sw128_chunk_off(lrow, ch & 7);
        *reinterpret_cast<uint4*>(base) = hi;
        *reinterpret_cast<uint4*>(base + T4_TILE) = lo;
    }
    if (below) flags[9] = 1u;
}

// 32 accumulator columns of one component -> partial squared norms (4 independent chains)
__device__ __forceinline__ void t4_sum32(const float (&v)[32], const float* __restrict__ off_s, float (&q)[4]) {
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
        const float4 o = *reinterpret_cast<const float4*>(off_s + j4 * 4);             // broadcast read
        const float t0 = v[j4 * 4] + o.x, t1 = v[j4 * 4 + 1] + o.y, t2 = v[j4 * 4 + 2] + o.z, t3 = v[j4 * 4 + 3] + o.w;
        q[0] = fmaf(t0, t0, q[0]); q[1] = fmaf(t1, t1, q[1]); q[2] = fmaf(t2, t2, q[2]); q[3] = fmaf(t3, t3, q[3]);
    }
}

// ---- main kernel ---------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T4_THREADS, 1)
tc_estep4_kernel(const float* __restrict__ Z, int64_t N, int D, int64_t ldz, int vec4,
                 const unsigned char* __restrict__ img, const float* __restrict__ offs,
                 const unsigned int* __restrict__ flags, int K, int n_gen, float* __restrict__ out, int64_t ldo,
                 const unsigned int* __restrict__ gate, unsigned int gate_value,
                 float* __restrict__ lse_vals, double* __restrict__ lse_sum) {
    if (gate != nullptr && __ldg(gate) != gate_value) return;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // A: [hi kb0 | hi kb1 | lo kb0 | lo kb1] tiles of 16 KB;  B: [stage][hi|lo];  offsets ring
    unsigned char* sA = smem_raw;
    unsigned char* sB = sA + 4 * T4_TILE;
    float* sOff = reinterpret_cast<float*>(sB + (size_t)T4_STAGES * T4_STAGE);
    T4Bars* bars = reinterpret_cast<T4Bars*>(reinterpret_cast<unsigned char*>(sOff) + T4_OFFRING * T4_OFFBYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int64_t n_passes = (N + 255) / 256;
    const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const bool dense0 = __ldg(flags + 9) != 0u;            // some operand has data left of its diagonal block: 4 steps
    const int n_steps = dense0 ? 4 : 3;

    if (tid == 0) {
        for (int s = 0; s < T4_STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); mbar_init(&bars->peer_full[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bars->reg_full[b], 1); mbar_init(&bars->reg_empty[b], 2 * T4_NW); }   // leader's: one arrival per epilogue warp of BOTH CTAs
        for (int b = 0; b < T4_OFFRING; ++b) { mbar_init(&bars->off_full[b], 1); mbar_init(&bars->off_empty[b], 32 * T4_NW); }
        mbar_init(&bars->a_full, 32 * T4_NW);
        mbar_init(&bars->peer_a_full, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (warp == T4_NW) tmem_alloc2(&bars->tmem_base, 512);
    tc_fence_before();
    cluster_sync_all();                                   // barriers of both CTAs initialised, TMEM allocated
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp < T4_NW) {
        // ================= converter + epilogue warps =================
        const float sz = pow2_scale_for(__uint_as_float(__ldg(flags)));
        const int ci = warp >> 2, qd = warp & 3;                     // ci: this warp group's component of a generation
        const int prow = qd * 32 + lane;                             // point row inside the tile = TMEM lane
        const uint32_t lane_base = tmem_base + ((uint32_t)(qd * 32) << 16);
        uint32_t gc = 0;
        for (int64_t pass = cluster_id; pass < n_passes; pass += n_clusters) {
            const int64_t n0 = pass * 256 + rank * 128;
            // ---- A operand: 128 rows of Z -> 3xFP16 split, K-major swizzled.  Every MMA of the previous pass has
            //      completed (all threads waited on its last region), so A may be overwritten. ----
            const int f = lane * 4;                                  // this lane's 4 features
            for (int r0 = warp; r0 < 128; r0 += 4 * T4_NW) {         // 4 rows in flight per warp
                float x[4][4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = r0 + T4_NW * u;
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
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int rr = r0 + T4_NW * u;
                    float xs[4] = {x[u][0] * sz, x[u][1] * sz, x[u][2] * sz, x[u][3] * sz};
                    uint2 hi, lo;
                    split4(xs, hi, lo);
                    const int kb = f >> 6, ch = (f & 63) >> 3;
                    unsigned char* base = sA + (size_t)kb * T4_TILE + sw128_chunk_off(rr, ch) + (f & 7) * 2;
                    *reinterpret_cast<uint2*>(base) = hi;
                    *reinterpret_cast<uint2*>(base + 2 * T4_TILE) = lo;
                }
            }
            fence_proxy_async();
            mbar_arrive(&bars->a_full);
            if (pass + n_clusters < n_passes) {                      // the rows of this cluster's next pass -> L2
                const int64_t nn0 = (pass + n_clusters) * 256 + rank * 128;
                const int64_t lines = ((int64_t)128 * ldz * 4 + 127) / 128;
                for (int64_t l = tid; l < lines; l += 32 * T4_NW) {
                    const char* pf = reinterpret_cast<const char*>(Z + nn0 * ldz) + l * 128;
                    if (pf < reinterpret_cast<const char*>(Z + N * ldz)) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
                }
            }

            // ---- epilogue over the generations ----
            const int64_t n = n0 + prow;
            const bool pvalid = n < N;
            float* outp = out + n;
            float lm = -INFINITY, ls = 0.f;                          // running (max, sum exp) over this group's components
            for (int g = 0; g < n_gen; ++g, ++gc) {
                const uint32_t par = gc & 1, ob = gc % T4_OFFRING;
                mbar_wait(&bars->off_full[ob], (gc / T4_OFFRING) & 1);
                const float* blk = sOff + ob * T4_OFFBLK;
                float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int region = 0; region < 2; ++region) {
                    mbar_wait(&bars->reg_full[region], par);
                    tc_fence_after();
                    const uint32_t taddr = lane_base + region * 256 + ci * 64;
                    float v0[32], v1[32];
                    tmem_ld32(taddr, v0);
                    tmem_ld32(taddr + 32, v1);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {                                 // region drained: straight to the leader's barrier
                        if (rank == 0) mbar_arrive(&bars->reg_empty[region]);
                        else mbar_arrive_remote_nofence(map_to_rank(smem_u32(&bars->reg_empty[region]), 0));
                    }
                    const float* o = blk + region * 256 + ci * 64;
                    t4_sum32(v0, o, q);
                    t4_sum32(v1, o + 32, q);
                }
                const int k = 4 * g + ci;
                if (k < K) {
                    const float qq = (q[0] + q[1]) + (q[2] + q[3]);
                    const float val = blk[512 + ci] - 0.5f * (blk[516 + ci] * qq);
                    if (pvalid) outp[(int64_t)k * ldo] = val;
                    const float mn = fmaxf(lm, val);                 // online log-sum-exp (fused log-normaliser)
                    ls = fmaf(ls, fast_exp(lm - mn), fast_exp(val - mn));
                    lm = mn;
                }
                mbar_arrive(&bars->off_empty[ob]);
            }
            if (lse_vals != nullptr) {
                // the four groups of a point meet: the next write of comb is a whole pass away (behind the a_full arrival of
                // every epilogue thread), so one barrier is enough
                if (ci > 0) bars->comb[ci - 1][prow] = make_float2(lm, ls);
                asm volatile("bar.sync 1, %0;" ::"n"(32 * T4_NW) : "memory");
                if (ci == 0) {
                    float M = lm;
#pragma unroll
                    for (int c2 = 0; c2 < 3; ++c2) M = fmaxf(M, bars->comb[c2][prow].x);
                    float S = ls * fast_exp(lm - M);
#pragma unroll
                    for (int c2 = 0; c2 < 3; ++c2) { const float2 o = bars->comb[c2][prow]; S = fmaf(o.y, fast_exp(o.x - M), S); }
                    const float lse = M + __logf(S);
                    double part = 0.0;
                    if (pvalid) { lse_vals[n] = lse; part = (double)lse; }
                    if (lse_sum != nullptr) {
#pragma unroll
                        for (int o2 = 16; o2 > 0; o2 >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o2);
                        if (lane == 0 && part != 0.0) atomicAdd(lse_sum, part);
                    }
                }
            }
        }
    } else if (warp == T4_NW) {
        if (lane == 0 && rank == 0) {
            // ================= MMA issuer (leader CTA, one thread) =================
            const uint32_t idesc = make_idesc_f16(256, 256);
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
            const int S = (D + 15) >> 4;                                  // 16-wide K steps that hold data (5..8)
            uint32_t stage = 0, phase = 0, gc = 0, it = 0;
            for (int64_t pass = cluster_id; pass < n_passes; pass += n_clusters, ++it) {
                mbar_wait(&bars->a_full, it & 1);
                mbar_wait_cluster(&bars->peer_a_full, it & 1);
                tc_fence_after();
                for (int g = 0; g < n_gen; ++g, ++gc) {
                    const uint32_t par = (gc & 1) ^ 1;
                    for (int step = 0; step < n_steps; ++step) {
                        // steps: F x K block 1, F x K block 0, S x K block 1, (S x K block 0)
                        const int region = step >> 1, kb = (step & 1) ^ 1;
                        if ((step & 1) == 0) {
                            mbar_wait_cluster(&bars->reg_empty[region], par);     // drained by the epilogue warps of both CTAs
                            tc_fence_after();
                        }
                        mbar_wait(&bars->full[stage], phase);
                        mbar_wait_cluster(&bars->peer_full[stage], phase);
                        tc_fence_after();
                        const uint32_t d = tmem_base + region * 256;
                        const uint32_t bs = b0 + stage * T4_STAGE;
                        const uint64_t bh = make_desc_sw128(bs), bl = make_desc_sw128(bs + T4_TILE);
                        const uint64_t ah = make_desc_sw128(a0 + kb * T4_TILE), al = make_desc_sw128(a0 + (2 + kb) * T4_TILE);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {                   // 16-element K steps inside the 64-wide block: +32 B
                            if (kb * 4 + kk >= S) continue;
                            const uint32_t acc = ((step & 1) == 0 && kk == 0) ? 0u : 1u;   // K block 1 comes first and always holds data
                            umma2_f16(d, al + 2 * kk, bh + 2 * kk, idesc, acc);
                            umma2_f16(d, ah + 2 * kk, bl + 2 * kk, idesc, 1);
                            umma2_f16(d, ah + 2 * kk, bh + 2 * kk, idesc, 1);
                        }
                        umma2_commit(&bars->empty[stage]);                // both CTAs' stage free once these MMAs have read it
                        if (++stage == T4_STAGES) { stage = 0; phase ^= 1; }
                        if (step == 1 || step == n_steps - 1) umma2_commit(&bars->reg_full[region]);
                    }
                }
            }
        } else if (lane == 0) {
            // ================= relay (peer CTA): forward local events to the leader's issuer =================
            const uint32_t r_a = map_to_rank(smem_u32(&bars->peer_a_full), 0);
            uint32_t stage = 0, phase = 0, it = 0;
            for (int64_t pass = cluster_id; pass < n_passes; pass += n_clusters, ++it) {
                mbar_wait(&bars->a_full, it & 1);
                mbar_arrive_remote(r_a);
                for (int s = 0; s < n_gen * n_steps; ++s) {
                    mbar_wait(&bars->full[stage], phase);
                    mbar_arrive_remote(map_to_rank(smem_u32(&bars->peer_full[stage]), 0));
                    if (++stage == T4_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ================= producer (one thread per CTA): this CTA's rows of every step + the generation's offsets =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, gc = 0;
            for (int64_t pass = cluster_id; pass < n_passes; pass += n_clusters) {
                for (int g = 0; g < n_gen; ++g, ++gc) {
                    const uint32_t ob = gc % T4_OFFRING;
                    mbar_wait(&bars->off_empty[ob], ((gc / T4_OFFRING) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bars->off_full[ob], T4_OFFBYTES);
                    bulk_g2s(sOff + ob * T4_OFFBLK, offs + (size_t)g * T4_OFFBLK, T4_OFFBYTES, &bars->off_full[ob]);
                    for (int step = 0; step < n_steps; ++step) {
                        mbar_wait(&bars->empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&bars->full[stage], T4_STAGE);
                        bulk_g2s(sB + (size_t)stage * T4_STAGE, img + (((size_t)g * 2 + rank) * 4 + step) * T4_STAGE, T4_STAGE, &bars->full[stage]);
                        if (++stage == T4_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();                                   // no CTA leaves (or frees TMEM) while its partner may still signal it
    if (warp == T4_NW) tmem_dealloc2(tmem_base, 512);
}

// ---- host side -----------------------------------------------------------------------------------------------------

static bool g_t4_enabled = false;          // measured slower than tc_estep2.cu (header)
int tc4_enable(int on) { int old = g_t4_enabled; g_t4_enabled = on != 0; return old; }

bool tc4_supported(int D, int Rp) { return g_t4_enabled && D > 64 && D <= 128 && Rp == 128; }

static size_t up1k4(size_t x) { return (x + 1023) / 1024 * 1024; }
static int t4_gens(int K) { return (K + 3) / 4; }
// [image gens x 2 x 4 x T4_STAGE | offsets gens x T4_OFFBLK floats], 1 KB aligned by the caller
size_t tc4_workspace(int K) { return (size_t)t4_gens(K) * 8 * T4_STAGE + up1k4((size_t)t4_gens(K) * T4_OFFBYTES); }

// flags: head of the operand workspace ([0] max |z| bits set by tc_data_scale; [9] written here)
int tc4_prepare(const float* W, const float* cst, int K, int Dpp, int D, unsigned int* flags, void* ws4, cudaStream_t st) {
    unsigned char* img = (unsigned char*)ws4;
    float* offs = (float*)(img + (size_t)t4_gens(K) * 8 * T4_STAGE);
    MIMO_CUDA(cudaMemsetAsync(flags + 9, 0, 4, st));
    tc4_prep_kernel<<<t4_gens(K), 256, 0, st>>>(W, cst, K, Dpp, D, flags, img, offs);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

int tc_estep4(const float* Z, int64_t N, int D, int64_t ldz, int K, const void* ws4, const unsigned int* flags,
              float* out, int64_t ldo, const unsigned int* gate, unsigned int gate_value,
              float* lse_vals, double* lse_sum, cudaStream_t st) {
    if (N == 0) return MIMO_OK;
    const unsigned char* img = (const unsigned char*)ws4;
    const float* offs = (const float*)(img + (size_t)t4_gens(K) * 8 * T4_STAGE);
    const size_t smem = 4 * (size_t)T4_TILE + (size_t)T4_STAGES * T4_STAGE + T4_OFFRING * T4_OFFBYTES + sizeof(T4Bars);
    MIMO_CUDA(cudaFuncSetAttribute(tc_estep4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t passes = (N + 255) / 256;
    const int clusters = (int)std::min<int64_t>(passes, sm_count() / 2);
    const int vec4 = (ldz % 4 == 0) && (((uintptr_t)Z & 15) == 0);
    tc_estep4_kernel<<<2 * clusters, T4_THREADS, smem, st>>>(Z, N, D, ldz, vec4, img, offs, flags, K, t4_gens(K), out, ldo, gate, gate_value,
                                                              lse_vals, lse_sum);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

}  // namespace mimo
