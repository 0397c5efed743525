// Tensor-core E-step on CTA PAIRS (tcgen05 cta_group::2), sm_100a.
//
//   a[k][n] = cst[k] - 0.5 * || W_k [z_n ; 1] ||^2        (include/mimo_b200.h, "packed operand form")
//
// Same math, operand image and 3xFP16 split as tc_estep.cu (reference call sites
// distributions/gaussian.py:510-523, lingauss.py:330-347, bayesian.py:287-301, 933-947); what
// changes is the shape of one MMA.  With both operands in shared memory a 128 x 128 x 16 MMA
// needs 128 B/clk of shared-memory reads -- the whole bandwidth of an SM -- and tc_estep.cu
// stalls at ~75 % of the tensor pipe.  Here two CTAs of a cluster (one TPC) issue ONE
// 256 x 256 x 16 MMA: each CTA supplies its own 128 points (A) and its own 128 operand rows
// (half of B), so per SM the reads drop to 64 B/clk, every B stage is copied from L2 by one of
// the two CTAs only, and the issuing thread has half as many instructions per flop.
//
// Per cluster pass = 2 x 128 points.  Per CTA: 8 converter / epilogue warps (thread = point,
// TMEM lane; warps 0-3 own accumulator columns 0..127 = the even 128-row chunk, warps 4-7
// columns 128..255 = the odd one), warp 8 = MMA issuer (leader CTA) or relay (peer CTA), warp 9
// = bulk-copy producer of this CTA's half of B and of the per-chunk offsets.  The leader's
// issuer must see BOTH CTAs' "A written", "B landed" and "accumulator drained" events: the
// epilogue warps of both CTAs arrive on the LEADER's accumulator barrier directly (one remote
// mbarrier.arrive per warp), the peer's relay thread forwards its "A written" / "B landed"
// events; completions travel the other way with tcgen05.commit ... multicast.
// The accumulator round trip MMA -> drain -> MMA is what bounds the single-pass (screening)
// variant -- 8 MMAs per 256 columns instead of 24 -- so an epilogue warp pulls its whole
// 128-column slice into registers with four tcgen05.ld in flight and releases the accumulator
// before doing any arithmetic on it.
//
// Narrow components (Rp <= 64, dense passes): 16 converter / epilogue warps instead of 8, each thread takes 64 of the
// chunk's 256 columns.  On the small-dimension shapes (cfg2 / cfg4 of BASELINE.json: one 16-wide K step feeds 256
// columns) the kernel is bound by the accumulator read-back and the per-column arithmetic behind it, and with two warps
// per scheduler that arithmetic ran at 1.7 warp instructions per clock (profiles/r02_small_d_kernels.md).  The rows of
// the NEXT pass are prefetched into L2 while this pass's chunks are drained.
// (A block-triangular variant for Cholesky-factor operands -- narrower MMAs for the K block where the lower row half is
// zero -- was measured slower twice, with the points operand in shared memory and in tensor memory: tc_estep3.cu.)
//
// Fused log-normaliser (PASSES = 3, lse_vals != nullptr): thread = point, so the running (max, sum) of the point's
// log-joints over all components is thread-local; the two column halves of a point meet once per pass through
// shared memory.  The softmax kernel and its two passes over the (K, chunk) scratch disappear from the dense
// sweep: the statistics pre-pass forms r = exp(a - lse_n) itself.
#include <algorithm>
#include "tc_common.cuh"
#include "internal.h"

namespace mimo {

using namespace tc;

// converter / epilogue warps of an instantiation (then one MMA / relay warp and one producer warp)
__host__ __device__ constexpr int t2_warps(int RP, int PASSES) { return (PASSES == 3 && RP <= 64) ? 16 : 8; }
constexpr int T2_STAGES = 4;                 // B ring: one stage = this CTA's 128 rows x 64 K, hi + lo = 32 KB
constexpr int T2_STAGES1 = 8;                // single-pass kernel: hi tiles only (16 KB), twice as many stages in the same space
constexpr uint32_t T2_TILE = 16384;          // 128 rows x 64 FP16
constexpr uint32_t T2_STAGE = 2 * T2_TILE;
constexpr int T2_OFFBLK = 320;               // floats per 256-row chunk: 256 row offsets | 32 cst | 32 1/scale^2
constexpr uint32_t T2_OFFBYTES = T2_OFFBLK * 4;
constexpr int T2_OFFRING = 8;                // offsets blocks in flight: deep enough that the producer never waits for the epilogue

struct T2Bars {
    uint64_t full[T2_STAGES1], empty[T2_STAGES1], peer_full[T2_STAGES1];
    uint64_t tmem_full[2], tmem_empty[2], peer_tmem_empty[2];
    uint64_t a_full, peer_a_full;
    uint64_t off_full[T2_OFFRING], off_empty[T2_OFFRING];
    uint32_t tmem_base;
    float2 comb[3][128];                     // fused log-normaliser: (max, sum) of the other column groups, per point
};

// offsets block of every 256-row chunk: [256 row offsets | cst of its components | 1/scale^2 of its components]
__global__ void tc2_offsets_kernel(const float* __restrict__ rowoff, const float* __restrict__ invS2, const float* __restrict__ cst,
                                   int K, int Rp, int n_chunks, float* __restrict__ offs2) {
    const int c2 = blockIdx.x, t = threadIdx.x;                 // 320 threads
    float v = 0.f;
    if (t < 256) {
        const int chunk = 2 * c2 + (t >> 7);
        v = chunk < n_chunks ? rowoff[(size_t)chunk * 128 + (t & 127)] : 0.f;
    } else {
        const int j = (t - 256) & 31;
        const int k = c2 * (256 / Rp) + j;
        if (j < 256 / Rp && k < K) v = (t < 288) ? cst[k] : invS2[k];
    }
    offs2[(size_t)c2 * T2_OFFBLK + t] = v;
}

// 32 accumulator columns -> partial squared norms (4 independent chains), component boundaries every RP columns
// SCREEN: also track the point's best component so far (value, index) -- the guess of tc_screen.cu
template <int RP, bool SCREEN>
__device__ __forceinline__ void t2_consume(const float (&v)[32], const float* __restrict__ off_s, int col0, float (&q)[4],
                                           const float* __restrict__ scal_s, int jbase, int kbase, int K, bool pvalid,
                                           float* __restrict__ outp, int64_t ldo, float& best, int& bestk,
                                           float& lm, float& ls) {
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
        const float4 o = *reinterpret_cast<const float4*>(off_s + col0 + j4 * 4);       // broadcast read
        const float oo[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int col = col0 + j4 * 4 + e;
            const float t = v[j4 * 4 + e] + oo[e];
            q[e] = fmaf(t, t, q[e]);
            if ((col + 1) % RP == 0) {
                const int j = jbase + col / RP;                    // component inside the 256-row chunk
                const int k = kbase + j;
                const float qq = (q[0] + q[1]) + (q[2] + q[3]);
                const float qt = scal_s[32 + j] * qq;              // || W_k [z ; 1] ||^2 in true units
                const float val = scal_s[j] - 0.5f * qt;
                if (pvalid && k < K) outp[(int64_t)k * ldo] = val;
                if (SCREEN && k < K && val > best) { best = val; bestk = k; }
                if (!SCREEN && k < K) {                            // online log-sum-exp (fused log-normaliser), branch-free
                    const float mn = fmaxf(lm, val);
                    ls = fmaf(ls, fast_exp(lm - mn), fast_exp(val - mn));
                    lm = mn;
                }
                q[0] = q[1] = q[2] = q[3] = 0.f;
            }
        }
    }
}

// PASSES = 3: Ah*Bh + Ah*Bl + Al*Bh (FP32-class);  PASSES = 1: Ah*Bh only (the screening pass of tc_screen.cu).
// gate (optional): the whole grid returns at once unless *gate == gate_value (device-side choice between the
// dense second pass and the per-candidate refinement without a host round trip).
template <int KB, int RP, int PASSES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(32 * (t2_warps(RP, PASSES) + 2), 1)
tc_estep2_kernel(const float* __restrict__ Z, int64_t N, int D, int64_t ldz, int vec4,
                 const __half* __restrict__ Bimg, const float* __restrict__ offs2,
                 const unsigned int* __restrict__ maxbits,
                 int K, int n_chunks2, float* __restrict__ out, int64_t ldo,
                 const unsigned int* __restrict__ gate, unsigned int gate_value,
                 float* __restrict__ lower, int* __restrict__ guess, int64_t ldl,
                 float* __restrict__ lse_vals, double* __restrict__ lse_sum) {
    if (gate != nullptr && __ldg(gate) != gate_value) return;
    constexpr int NW = t2_warps(RP, PASSES);                            // converter / epilogue warps
    constexpr int NG = NW / 4, CW = 256 / NG;                           // column groups of a chunk, columns per group
    constexpr uint32_t STAGE_TX = PASSES == 3 ? T2_STAGE : T2_TILE;   // bytes copied per stage (hi | lo, or hi only)
    constexpr int NST = PASSES == 3 ? T2_STAGES : T2_STAGES1;         // ring depth; a ring slot is STAGE_TX bytes
    static_assert(NST * STAGE_TX == T2_STAGES * T2_STAGE, "ring size");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // A: [hi|lo][kb KB] tiles of 16 KB;  B: [stage][hi|lo] tiles of 16 KB;  offsets ring: T2_OFFRING x T2_OFFBYTES
    unsigned char* sA = smem_raw;
    unsigned char* sB = sA + (size_t)2 * KB * T2_TILE;
    float* sOff = reinterpret_cast<float*>(sB + (size_t)T2_STAGES * T2_STAGE);
    T2Bars* bars = reinterpret_cast<T2Bars*>(reinterpret_cast<unsigned char*>(sOff) + T2_OFFRING * T2_OFFBYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int64_t n_passes = (N + 255) / 256;
    const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); mbar_init(&bars->peer_full[s], 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->tmem_full[b], 1); mbar_init(&bars->tmem_empty[b], 2 * NW); mbar_init(&bars->peer_tmem_empty[b], 1);   // tmem_empty (leader's): one arrival per epilogue warp of BOTH CTAs
        }
        for (int b = 0; b < T2_OFFRING; ++b) { mbar_init(&bars->off_full[b], 1); mbar_init(&bars->off_empty[b], 32 * NW); }
        mbar_init(&bars->a_full, 32 * NW);
        mbar_init(&bars->peer_a_full, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (warp == NW) tmem_alloc2(&bars->tmem_base, 512);
    tc_fence_before();
    cluster_sync_all();                                   // barriers of both CTAs initialised, TMEM allocated
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp < NW) {
        // ================= converter + epilogue warps =================
        const float sz = pow2_scale_for(__uint_as_float(__ldg(maxbits)));
        constexpr bool SCREEN = PASSES == 1;
        const int half = warp >> 2, qd = warp & 3;                   // half: column group of this warp (0 .. NG-1)
        const int prow = qd * 32 + lane;                             // point row inside the tile = TMEM lane
        uint32_t gc = 0;
        for (int64_t pass = cluster_id; pass < n_passes; pass += n_clusters) {
            const int64_t n0 = pass * 256 + rank * 128;
            // ---- A operand: 128 rows of Z -> 3xFP16 split, K-major swizzled.  Every MMA of the previous pass
            //      has completed (all threads waited on its last tmem_full), so A may be overwritten. ----
            const int f = lane * 4;                                  // this lane's 4 features
            for (int r0 = warp; r0 < 128; r0 += 4 * NW) {            // 4 rows in flight per warp
                float x[4][4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = r0 + NW * u;
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
                        const int rr = r0 + NW * u;
                        float xs[4] = {x[u][0] * sz, x[u][1] * sz, x[u][2] * sz, x[u][3] * sz};
                        uint2 hi, lo;
                        split4(xs, hi, lo);
                        const int kb = f >> 6, ch = (f & 63) >> 3;
                        unsigned char* base = sA + (size_t)kb * T2_TILE + sw128_chunk_off(rr, ch) + (f & 7) * 2;
                        *reinterpret_cast<uint2*>(base) = hi;
                        if (PASSES == 3) *reinterpret_cast<uint2*>(base + (size_t)KB * T2_TILE) = lo;
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(&bars->a_full);
            // the rows of this cluster's next pass -> L2 (one 128-byte line per lane; the conversion above then hits L2)
            if (pass + n_clusters < n_passes) {
                const int64_t nn0 = (pass + n_clusters) * 256 + rank * 128;
                const int64_t lines = ((int64_t)128 * ldz * 4 + 127) / 128;
                for (int64_t l = tid; l < lines; l += 32 * NW) {
                    const char* pf = reinterpret_cast<const char*>(Z + nn0 * ldz) + l * 128;
                    if (pf < reinterpret_cast<const char*>(Z + N * ldz)) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
                }
            }

            // ---- epilogue over the 256-row chunks: this warp group owns columns [CW half, +CW) ----
            const int64_t n = n0 + prow;
            const bool pvalid = n < N;
            float* outp = out + n;
            float Lb = -INFINITY;
            int Lk = 0;
            float lm = -INFINITY, ls = 0.f;                          // running (max, sum exp) over this half's components
            for (int c = 0; c < n_chunks2; ++c, ++gc) {
                const uint32_t buf = gc & 1, par = (gc >> 1) & 1;
                const uint32_t ob = gc % T2_OFFRING;
                mbar_wait(&bars->off_full[ob], (gc / T2_OFFRING) & 1);
                mbar_wait(&bars->tmem_full[buf], par);
                tc_fence_after();
                const float* blk = sOff + ob * T2_OFFBLK;
                const float* scal_s = blk + 256;
                // the whole column slice into registers with its loads in flight, then the accumulator is free again: the
                // drain costs one TMEM latency on the MMA -> epilogue -> MMA chain, not one per load plus the math
                float v[CW / 32][32];
                float q[4] = {0.f, 0.f, 0.f, 0.f};
                const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + buf * 256 + half * CW;
#pragma unroll
                for (int g = 0; g < CW / 32; ++g) tmem_ld32(taddr + 32 * g, v[g]);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {                                 // straight to the leader's barrier (no relay hop)
                    if (rank == 0) mbar_arrive(&bars->tmem_empty[buf]);
                    else mbar_arrive_remote_nofence(map_to_rank(smem_u32(&bars->tmem_empty[buf]), 0));
                }
                const float* off_s = blk + half * CW;
                const int kbase = c * (256 / RP), jbase = half * (CW / RP);
#pragma unroll
                for (int g = 0; g < CW / 32; ++g)
                    t2_consume<RP, SCREEN>(v[g], off_s, 32 * g, q, scal_s, jbase, kbase, K, pvalid, outp, ldo, Lb, Lk, lm, ls);
                mbar_arrive(&bars->off_empty[ob]);
            }
            if (SCREEN && pvalid) { lower[(int64_t)half * ldl + n] = Lb; guess[(int64_t)half * ldl + n] = Lk; }
            if (!SCREEN && lse_vals != nullptr) {
                // the column groups of a point meet: the next write of comb is a whole pass away (behind the a_full
                // arrival of every epilogue thread), so one barrier is enough
                if (half > 0) bars->comb[half - 1][prow] = make_float2(lm, ls);
                asm volatile("bar.sync 1, %0;" ::"n"(32 * NW) : "memory");
                if (half == 0) {
                    float M = lm;
#pragma unroll
                    for (int g = 0; g < NG - 1; ++g) M = fmaxf(M, bars->comb[g][prow].x);
                    float S = ls * fast_exp(lm - M);
#pragma unroll
                    for (int g = 0; g < NG - 1; ++g) { const float2 o = bars->comb[g][prow]; S = fmaf(o.y, fast_exp(o.x - M), S); }
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
    } else if (warp == NW) {
        if (lane == 0 && rank == 0) {
            // ================= MMA issuer (leader CTA, one thread) =================
            const uint32_t idesc = make_idesc_f16(256, 256);
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
            const int S = (D + 15) >> 4;                                  // 16-wide K steps that hold data
            uint32_t stage = 0, phase = 0, gc = 0, it = 0;
            for (int64_t pass = cluster_id; pass < n_passes; pass += n_clusters, ++it) {
                mbar_wait(&bars->a_full, it & 1);
                mbar_wait_cluster(&bars->peer_a_full, it & 1);
                tc_fence_after();
                for (int c = 0; c < n_chunks2; ++c, ++gc) {
                    const uint32_t buf = gc & 1, par = ((gc >> 1) & 1) ^ 1;
                    mbar_wait_cluster(&bars->tmem_empty[buf], par);      // drained by the epilogue warps of both CTAs
                    tc_fence_after();
                    const uint32_t d = tmem_base + buf * 256;
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&bars->full[stage], phase);
                        mbar_wait_cluster(&bars->peer_full[stage], phase);
                        tc_fence_after();
                        const uint32_t bs = b0 + stage * STAGE_TX;
                        const uint64_t bh = make_desc_sw128(bs);
                        const uint64_t bl = make_desc_sw128(bs + T2_TILE);
                        const uint64_t ah = make_desc_sw128(a0 + kb * T2_TILE);
                        const uint64_t al = make_desc_sw128(a0 + (KB + kb) * T2_TILE);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {               // 16-element K steps inside the 64-wide block: +32 B
                            if (kb * 4 + kk >= S) continue;
                            const uint32_t acc = (uint32_t)((kb | kk) != 0);
                            if (PASSES == 3) {
                                umma2_f16(d, al + 2 * kk, bh + 2 * kk, idesc, acc);
                                umma2_f16(d, ah + 2 * kk, bl + 2 * kk, idesc, 1);
                                umma2_f16(d, ah + 2 * kk, bh + 2 * kk, idesc, 1);
                            } else {
                                umma2_f16(d, ah + 2 * kk, bh + 2 * kk, idesc, acc);
                            }
                        }
                        umma2_commit(&bars->empty[stage]);               // both CTAs' stage free once these MMAs have read it
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
                    umma2_commit(&bars->tmem_full[buf]);
                }
            }
        } else if (lane == 0) {
            // ================= relay (peer CTA): forward local events to the leader's issuer =================
            const uint32_t r_a = map_to_rank(smem_u32(&bars->peer_a_full), 0);
            uint32_t stage = 0, phase = 0, gc = 0, it = 0;
            for (int64_t pass = cluster_id; pass < n_passes; pass += n_clusters, ++it) {
                mbar_wait(&bars->a_full, it & 1);
                mbar_arrive_remote(r_a);
                for (int c = 0; c < n_chunks2; ++c, ++gc) {
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&bars->full[stage], phase);
                        mbar_arrive_remote(map_to_rank(smem_u32(&bars->peer_full[stage]), 0));
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else {
        // ================= producer (one thread per CTA): this CTA's half of B + the chunk offsets =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, gc = 0;
            for (int64_t pass = cluster_id; pass < n_passes; pass += n_clusters) {
                const unsigned char* src = reinterpret_cast<const unsigned char*>(Bimg);
                for (int c = 0; c < n_chunks2; ++c, ++gc) {
                    const uint32_t ob = gc % T2_OFFRING;
                    mbar_wait(&bars->off_empty[ob], ((gc / T2_OFFRING) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bars->off_full[ob], T2_OFFBYTES);
                    bulk_g2s(sOff + ob * T2_OFFBLK, offs2 + (size_t)c * T2_OFFBLK, T2_OFFBYTES, &bars->off_full[ob]);
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&bars->empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&bars->full[stage], STAGE_TX);
                        bulk_g2s(sB + (size_t)stage * STAGE_TX, src + ((size_t)(2 * c + rank) * KB + kb) * T2_STAGE, STAGE_TX, &bars->full[stage]);
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();                                   // no CTA leaves (or frees TMEM) while its partner may still signal it
    if (warp == NW) tmem_dealloc2(tmem_base, 512);
}

// ---- host side -------------------------------------------------------------------------------

size_t tc2_offsets_bytes(int K, int Rp) {
    const int64_t n_chunks = ((int64_t)K * Rp + 127) / 128;
    return (size_t)((n_chunks + 1) / 2) * T2_OFFBYTES;
}

int tc2_prepare_offsets(const float* rowoff, const float* invS2, const float* cst, int K, int Rp, float* offs2, cudaStream_t st) {
    const int n_chunks = (int)(((int64_t)K * Rp + 127) / 128);
    tc2_offsets_kernel<<<(n_chunks + 1) / 2, T2_OFFBLK, 0, st>>>(rowoff, invS2, cst, K, Rp, n_chunks, offs2);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

template <int KB, int RP, int PASSES>
static int launch_estep2(const float* Z, int64_t N, int D, int64_t ldz, const __half* Bimg, const float* offs2,
                         const unsigned int* maxbits, int K, int n_chunks2, float* out, int64_t ldo,
                         const unsigned int* gate, unsigned int gate_value, float* lower, int* guess, int64_t ldl,
                         float* lse_vals, double* lse_sum, cudaStream_t st) {
    const size_t smem = (size_t)2 * KB * T2_TILE + (size_t)T2_STAGES * T2_STAGE + T2_OFFRING * T2_OFFBYTES + sizeof(T2Bars);
    auto kern = tc_estep2_kernel<KB, RP, PASSES>;
    MIMO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t passes = (N + 255) / 256;
    const int clusters = (int)std::min<int64_t>(passes, sm_count() / 2);
    const int vec4 = (ldz % 4 == 0) && (((uintptr_t)Z & 15) == 0);
    kern<<<2 * clusters, 32 * (t2_warps(RP, PASSES) + 2), smem, st>>>(Z, N, D, ldz, vec4, Bimg, offs2, maxbits, K, n_chunks2, out, ldo, gate, gate_value, lower, guess, ldl, lse_vals, lse_sum);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

// E-step over N points with a prepared operand image (even number of 128-row chunks, zero padded) and offsets blocks.
// passes = 3 (FP32-class) or 1 (screening pass); gate: see the kernel.
int tc_estep2(const float* Z, int64_t N, int D, int64_t ldz, int K, int Rp, int KB, const void* Bimg_, const float* offs2,
              const unsigned int* maxbits, float* out, int64_t ldo, int passes, const unsigned int* gate, unsigned int gate_value,
              float* lower, int* guess, int64_t ldl, cudaStream_t st, float* lse_vals, double* lse_sum) {
    const __half* Bimg = (const __half*)Bimg_;
    if (N == 0) return MIMO_OK;
    const int n_chunks = (int)(((int64_t)K * Rp + 127) / 128);
    const int n_chunks2 = (n_chunks + 1) / 2;
#define T2_CASE(kb, rp) if (KB == kb && Rp == rp) { \
        if (passes == 1) return launch_estep2<kb, rp, 1>(Z, N, D, ldz, Bimg, offs2, maxbits, K, n_chunks2, out, ldo, gate, gate_value, lower, guess, ldl, nullptr, nullptr, st); \
        return launch_estep2<kb, rp, 3>(Z, N, D, ldz, Bimg, offs2, maxbits, K, n_chunks2, out, ldo, gate, gate_value, lower, guess, ldl, lse_vals, lse_sum, st); }
    T2_CASE(1, 8) T2_CASE(1, 16) T2_CASE(1, 32) T2_CASE(1, 64) T2_CASE(1, 128)
    T2_CASE(2, 8) T2_CASE(2, 16) T2_CASE(2, 32) T2_CASE(2, 64) T2_CASE(2, 128)
#undef T2_CASE
    set_error("tensor-core E-step (CTA pairs): unsupported shape D=%d Rp=%d", D, Rp);
    return MIMO_EUNSUPPORTED;
}

}  // namespace mimo
