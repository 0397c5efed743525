// Tensor-core E-step for 64 < D <= 128, Rp = 128 on CTA pairs with the POINTS OPERAND IN TENSOR MEMORY, sm_100a.
//
//   a[k][n] = cst[k] - 0.5 * || W_k [z_n ; 1] ||^2        (include/mimo_b200.h, "packed operand form")
//
// replaces the same reference call sites as tc_estep2.cu (distributions/gaussian.py:510-523, bayesian.py:287-301,
// 933-947) on the dense path of the cfg5 shape.
//
// Why a second kernel.  tc_estep2.cu reads both operands of every tcgen05.mma from shared memory; the 4 KB read of the
// 128 x 16 points tile then takes as long as a 256-column MMA, so narrower MMAs (which is what skipping the zero half
// of a Cholesky factor needs) gain nothing there (measured twice: profiles/r01 and profiles/r02_tri_ss_mode.md).
// Here the split points tile (hi | lo, 128 points x 128 K elements, two FP16 per 32-bit cell = 128 columns) is written
// ONCE per pass into tensor memory by the epilogue threads (thread = point = TMEM lane, tcgen05.st) and every MMA takes
// its A operand from there: the MMA time is then proportional to N and the shared-memory traffic is the operand rows only.
//
// Triangular skip.  The rows of W_k are Cholesky factors (U_k, sqrt(nu_k) C_k^T): row r is zero left of column r.  For
// the K step j (columns 16 j .. 16 j + 15) only rows below 16 (j + 1) contribute, a PREFIX of the rows, so the MMA of
// that step has N_j = G (j / (G / 16) + 1) columns (G = 16 or 32 rows of granularity): 56 % (62 %) of the full work.
// One accumulator buffer = one component = 128 columns in row order; three buffers + the 128 columns of A fill TMEM.
// A cta_group::2 MMA takes the first N/2 operand rows from the leader CTA and the second N/2 from its peer, so the rows
// a CTA supplies differ from step to step; the operand image is built for exactly that (tc3_prep_kernel): position
// (row i, K step j) of CTA `rank`'s tile holds row rank * N_j / 2 + i of W_k.  Operands that are not triangular (stacked
// ILR blocks; checked on the device, flags[8]) get the image and the MMAs of the full 128 rows.
//
// MEASURED (B200, cfg5 shape, profiles/r02_estep_tmem_operand.md): correct (parity 6e-7) but SLOWER than tc_estep2.cu --
// 75.3 ms per 1 M-point chunk against 54.6 ms, the same for G = 16 and G = 32.  Every MMA re-reads its 128 x 16 slice of
// A (4 KB per CTA) through the tensor-memory read port, the same 64 B/clk port the epilogue's tcgen05.ld uses: 24 MMAs x
// 4 KB + 64 KB of accumulator per component = 2560 clk of port time, which is what the kernel takes (2507 clk).  With
// the operand in shared memory that read goes over the 128 B/clk shared-memory path instead.  The kernel therefore
// stays OFF by default (mimo_tc_set_triangular(16) turns it on for A/B runs and for its parity test).
//
// Epilogue.  The two epilogue warp groups take every other component: a thread reads its point's 128 accumulator columns
// (four tcgen05.ld in flight), releases the buffer, adds the offset column, squares and sums, stores cst - q/2 and keeps a
// running (max, sum exp) -- the log-normaliser of the point falls out at the end of the pass (fused softmax).  What
// bounds the kernel is this read-back (128 columns x 4 B per pair at 64 B/clk/SM), not the tensor pipe.
#include <algorithm>
#include "tc_common.cuh"
#include "internal.h"

namespace mimo {

using namespace tc;

constexpr int T3_THREADS = 320;               // 8 converter / epilogue warps + MMA (relay) warp + producer warp
constexpr int T3_NBUF = 3;                    // accumulator buffers of 128 columns
constexpr uint32_t T3_ACC0 = 128;             // first accumulator column; A: hi in columns [0, 64), lo in [64, 128)
constexpr int T3_STAGES = 5;                  // B ring: one stage = one component's rows for this CTA
constexpr uint32_t T3_SUB = 8192;             // a 64-row x 64-K tile (SW128)
constexpr uint32_t T3_STAGE = 4 * T3_SUB;     // [hi kb1 | lo kb1 | hi kb0 | lo kb0]; triangular: the kb0 tiles hold 32 rows (4 KB each)
constexpr uint32_t T3_STAGE_TRI = 2 * T3_SUB + T3_SUB;
constexpr int T3_OFFBLK = 136;                // floats per component: 128 row offsets | cst | 1/scale^2 | pad
constexpr uint32_t T3_OFFBYTES = T3_OFFBLK * 4;
constexpr int T3_OFFRING = 8;

struct T3Bars {
    uint64_t full[T3_STAGES], empty[T3_STAGES], peer_full[T3_STAGES];
    uint64_t tmem_full[T3_NBUF], tmem_empty[T3_NBUF];
    uint64_t a_full, peer_a_full;
    uint64_t off_full[T3_OFFRING], off_empty[T3_OFFRING];
    uint32_t tmem_base;
    float2 comb[128];                         // (max, sum) of the odd warp group, per point
};

// operand rows per CTA of K step j
template <int G>
__host__ __device__ __forceinline__ int t3_rows(int j, bool dense) { return dense ? 64 : (G / 2) * (j / (G / 16) + 1); }

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem, both CTAs] (+)= A[tmem, both CTAs] * B^T: M = 256, N columns (N/2 rows of B per CTA), K = 16
__device__ __forceinline__ void umma2_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---- operand image ------------------------------------------------------------------------------------------------
// flags[8] != 0: some W_k has data left of its G-row staircase (not a Cholesky factor)
template <int G>
__global__ void tc3_scan_kernel(const float* __restrict__ W, int K, int Dpp, int D, unsigned int* __restrict__ flags) {
    const int k = blockIdx.x;
    bool below = false;
    for (int idx = threadIdx.x; idx < 128 * D; idx += blockDim.x) {
        const int r = idx / D, j = idx - r * D;
        if (j < (r / G) * G && W[((size_t)k * 128 + r) * Dpp + j] != 0.f) below = true;
    }
    if (below) flags[8] = 1u;
}

// grid = K components, block = 256.  img [k][rank][T3_STAGE], offs [k][T3_OFFBLK]
template <int G>
__global__ void __launch_bounds__(256)
tc3_prep_kernel(const float* __restrict__ W, const float* __restrict__ cst, int K, int Dpp, int D,
                const unsigned int* __restrict__ flags, unsigned char* __restrict__ img, float* __restrict__ offs) {
    __shared__ unsigned int cmax;
    const int k = blockIdx.x, tid = threadIdx.x;
    const float* Wk = W + (size_t)k * 128 * Dpp;
    const bool dense = __ldg(flags + 8) != 0u;
    const float sz = pow2_scale_for(__uint_as_float(__ldg(flags)));
    if (tid == 0) cmax = 0u;
    __syncthreads();
    unsigned int m = 0u;
    for (int idx = tid; idx < 128 * D; idx += 256) {
        const int r = idx / D, j = idx - r * D;
        m = max(m, __float_as_uint(fabsf(Wk[(size_t)r * Dpp + j])));
    }
    atomicMax(&cmax, m);
    __syncthreads();
    const float sw = pow2_scale_for(__uint_as_float(cmax));
    float* ob = offs + (size_t)k * T3_OFFBLK;
    if (tid < 128) ob[tid] = Wk[(size_t)tid * Dpp + D] * sw * sz;
    else if (tid == 128) ob[128] = cst[k];
    else if (tid == 129) { const float s = sw * sz; ob[129] = 1.f / (s * s); }
    else if (tid < T3_OFFBLK) ob[tid] = 0.f;
    // tiles: (rank, kb, row i < 64, 16-byte chunk ch < 8); K step of a chunk = 4 kb + ch / 2
    const uint32_t off_kb0 = 2 * T3_SUB, lo_kb0 = dense ? T3_SUB : T3_SUB / 2;
    for (int idx = tid; idx < 2 * 2 * 64 * 8; idx += 256) {
        const int ch = idx & 7, i = (idx >> 3) & 63, kb = (idx >> 9) & 1, rank = idx >> 10;
        if (!dense && kb == 0 && i >= 32) continue;                 // the triangular kb0 tiles hold 32 rows
        const int j = 4 * kb + (ch >> 1);
        const int nj = t3_rows<G>(j, dense);
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int col = 64 * kb + 8 * ch + e;
            x[e] = (i < nj && col < D) ? Wk[(size_t)(rank * nj + i) * Dpp + col] * sw : 0.f;
        }
        uint4 hi, lo;
        split8(x, hi, lo);
        unsigned char* base = img + ((size_t)k * 2 + rank) * T3_STAGE + (kb ? 0u : off_kb0) + sw128_chunk_off(i, ch);
        *reinterpret_cast<uint4*>(base) = hi;
        *reinterpret_cast<uint4*>(base + (kb ? T3_SUB : lo_kb0)) = lo;
    }
}

// ---- main kernel ---------------------------------------------------------------------------------------------------
// gate (optional): the whole grid returns at once unless *gate == gate_value (tc_screen.cu's device-side choice).
template <int G>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T3_THREADS, 1)
tc_estep3_kernel(const float* __restrict__ Z, int64_t N, int D, int64_t ldz, int vec4,
                 const unsigned char* __restrict__ img, const float* __restrict__ offs,
                 const unsigned int* __restrict__ flags, int K, float* __restrict__ out, int64_t ldo,
                 const unsigned int* __restrict__ gate, unsigned int gate_value,
                 float* __restrict__ lse_vals, double* __restrict__ lse_sum) {
    if (gate != nullptr && __ldg(gate) != gate_value) return;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sB = smem_raw;
    float* sOff = reinterpret_cast<float*>(sB + (size_t)T3_STAGES * T3_STAGE);
    T3Bars* bars = reinterpret_cast<T3Bars*>(reinterpret_cast<unsigned char*>(sOff) + T3_OFFRING * T3_OFFBYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int64_t n_passes = (N + 255) / 256;
    const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const bool dense = __ldg(flags + 8) != 0u;

    if (tid == 0) {
        for (int s = 0; s < T3_STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); mbar_init(&bars->peer_full[s], 1); }
        for (int b = 0; b < T3_NBUF; ++b) { mbar_init(&bars->tmem_full[b], 1); mbar_init(&bars->tmem_empty[b], 8); }   // leader's: one arrival per warp of the owning group, both CTAs
        for (int b = 0; b < T3_OFFRING; ++b) { mbar_init(&bars->off_full[b], 1); mbar_init(&bars->off_empty[b], 128); }
        mbar_init(&bars->a_full, 256);
        mbar_init(&bars->peer_a_full, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (warp == 8) tmem_alloc2(&bars->tmem_base, 512);
    tc_fence_before();
    cluster_sync_all();                                   // barriers of both CTAs initialised, TMEM allocated
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp < 8) {
        // ================= converter + epilogue warps =================
        const float sz = pow2_scale_for(__uint_as_float(__ldg(flags)));
        const int grp = warp >> 2, qd = warp & 3;
        const int prow = qd * 32 + lane;                             // point row inside the tile = TMEM lane
        const uint32_t lane_base = tmem_base + ((uint32_t)(qd * 32) << 16);
        uint32_t gc0 = 0;                                            // components issued before this pass
        for (int64_t pass = cluster_id; pass < n_passes; pass += n_clusters, gc0 += (uint32_t)K) {
            const int64_t n = pass * 256 + rank * 128 + prow;
            const bool pvalid = n < N;
            // ---- A operand: this point's K elements [64 grp, 64 grp + 64) -> split FP16 pairs -> tensor memory.
            //      Every MMA of the previous pass has completed (bar.sync at its end), so A may be overwritten. ----
            {
                const float* src = Z + n * ldz + 64 * grp;
#pragma unroll
                for (int q = 0; q < 4; ++q) {                        // 16 elements = 8 columns of hi and of lo
                    float x[16];
                    const int e0 = 64 * grp + 16 * q;
                    if (pvalid && vec4 && e0 + 16 <= D) {
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const float4 t = __ldg(reinterpret_cast<const float4*>(src + 16 * q) + v);
                            x[4 * v] = t.x; x[4 * v + 1] = t.y; x[4 * v + 2] = t.z; x[4 * v + 3] = t.w;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e) x[e] = (pvalid && e0 + e < D) ? __ldg(src + 16 * q + e) : 0.f;
                    }
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        __half h0, l0, h1, l1;
                        split_f16(x[2 * e] * sz, h0, l0);
                        split_f16(x[2 * e + 1] * sz, h1, l1);
                        const __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);   // even K element in the low half
                        hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
                        lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
                    }
                    tmem_st8(lane_base + 32 * grp + 8 * q, hi);
                    tmem_st8(lane_base + 64 + 32 * grp + 8 * q, lo);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&bars->a_full);
            }

            // ---- epilogue: this warp group takes every other component ----
            float* outp = out + n;
            float lm = -INFINITY, ls = 0.f;                          // running (max, sum exp) over this group's components
            for (int c = 0; c < K; ++c) {
                const uint32_t gc = gc0 + (uint32_t)c;
                if ((int)(gc & 1u) != grp) continue;
                const uint32_t buf = gc % T3_NBUF, ob = gc % T3_OFFRING;
                mbar_wait(&bars->off_full[ob], (gc / T3_OFFRING) & 1);
                mbar_wait(&bars->tmem_full[buf], (gc / T3_NBUF) & 1);
                tc_fence_after();
                const uint32_t taddr = lane_base + T3_ACC0 + buf * 128;
                float v0[32], v1[32], v2[32], v3[32];
                tmem_ld32(taddr, v0);
                tmem_ld32(taddr + 32, v1);
                tmem_ld32(taddr + 64, v2);
                tmem_ld32(taddr + 96, v3);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {                                     // accumulator drained: the leader's issuer may reuse it
                    if (rank == 0) mbar_arrive(&bars->tmem_empty[buf]);
                    else mbar_arrive_remote_nofence(map_to_rank(smem_u32(&bars->tmem_empty[buf]), 0));
                }
                const float* o = sOff + ob * T3_OFFBLK;
                float q[4] = {0.f, 0.f, 0.f, 0.f};
                auto sum32 = [&](const float (&v)[32], const float* __restrict__ off_s) {
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 f = *reinterpret_cast<const float4*>(off_s + j4 * 4);             // broadcast read
                        const float t0 = v[j4 * 4] + f.x, t1 = v[j4 * 4 + 1] + f.y, t2 = v[j4 * 4 + 2] + f.z, t3 = v[j4 * 4 + 3] + f.w;
                        q[0] = fmaf(t0, t0, q[0]); q[1] = fmaf(t1, t1, q[1]); q[2] = fmaf(t2, t2, q[2]); q[3] = fmaf(t3, t3, q[3]);
                    }
                };
                sum32(v0, o); sum32(v1, o + 32); sum32(v2, o + 64); sum32(v3, o + 96);
                const float qq = (q[0] + q[1]) + (q[2] + q[3]);
                const float val = o[128] - 0.5f * (o[129] * qq);
                mbar_arrive(&bars->off_empty[ob]);
                if (pvalid) outp[(int64_t)c * ldo] = val;
                const float mn = fmaxf(lm, val);                                                       // online log-sum-exp
                ls = fmaf(ls, fast_exp(lm - mn), fast_exp(val - mn));
                lm = mn;
            }
            // ---- end of the pass: the two groups of a point meet.  The barrier also orders the next pass's writes of A
            //      behind the last MMAs of this one (whoever read the last accumulator waited for them). ----
            if (grp == 1) bars->comb[prow] = make_float2(lm, ls);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (grp == 0 && lse_vals != nullptr) {
                const float2 o = bars->comb[prow];
                const float M = fmaxf(lm, o.x);
                const float S = ls * fast_exp(lm - M) + o.y * fast_exp(o.x - M);
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
    } else if (warp == 8) {
        if (lane == 0 && rank == 0) {
            // ================= MMA issuer (leader CTA, one thread) =================
            const uint32_t b0 = smem_u32(sB);
            const int S = (D + 15) >> 4;                                  // 16-wide K steps that hold data
            const uint32_t o_h1 = 0, o_l1 = T3_SUB, o_h0 = 2 * T3_SUB, o_l0 = 2 * T3_SUB + (dense ? T3_SUB : T3_SUB / 2);
            uint32_t stage = 0, phase = 0, gc = 0, it = 0;
            for (int64_t pass = cluster_id; pass < n_passes; pass += n_clusters, ++it) {
                mbar_wait(&bars->a_full, it & 1);
                mbar_wait_cluster(&bars->peer_a_full, it & 1);
                tc_fence_after();
                for (int c = 0; c < K; ++c, ++gc) {
                    const uint32_t buf = gc % T3_NBUF;
                    mbar_wait_cluster(&bars->tmem_empty[buf], ((gc / T3_NBUF) & 1) ^ 1);   // drained by the owning warp group of both CTAs
                    mbar_wait(&bars->full[stage], phase);
                    mbar_wait_cluster(&bars->peer_full[stage], phase);
                    tc_fence_after();
                    const uint32_t d = tmem_base + T3_ACC0 + buf * 128;
                    const uint32_t bs = b0 + stage * T3_STAGE;
                    const uint64_t bh1 = make_desc_sw128(bs + o_h1), bl1 = make_desc_sw128(bs + o_l1);
                    const uint64_t bh0 = make_desc_sw128(bs + o_h0), bl0 = make_desc_sw128(bs + o_l0);
                    // K step 7 first: its 128 columns initialise the whole accumulator (all-zero operand rows when D <= 112)
                    {
                        const uint32_t idesc = make_idesc_f16(256, 128);
                        if (S == 8) {
                            umma2_ts_f16(d, tmem_base + 64 + 56, bh1 + 6, idesc, 0);
                            umma2_ts_f16(d, tmem_base + 56, bl1 + 6, idesc, 1);
                            umma2_ts_f16(d, tmem_base + 56, bh1 + 6, idesc, 1);
                        } else {
                            umma2_ts_f16(d, tmem_base + 56, bh1 + 6, idesc, 0);
                        }
                    }
#pragma unroll
                    for (int j = 6; j >= 0; --j) {
                        if (j >= S) continue;
                        const uint32_t idesc = make_idesc_f16(256, 2 * t3_rows<G>(j, dense));
                        const uint64_t bh = (j >= 4 ? bh1 : bh0) + 2 * (j & 3), bl = (j >= 4 ? bl1 : bl0) + 2 * (j & 3);
                        umma2_ts_f16(d, tmem_base + 64 + 8 * j, bh, idesc, 1);
                        umma2_ts_f16(d, tmem_base + 8 * j, bl, idesc, 1);
                        umma2_ts_f16(d, tmem_base + 8 * j, bh, idesc, 1);
                    }
                    umma2_commit(&bars->empty[stage]);                   // both CTAs' stage free once these MMAs have read it
                    if (++stage == T3_STAGES) { stage = 0; phase ^= 1; }
                    umma2_commit(&bars->tmem_full[buf]);
                }
            }
        } else if (lane == 0) {
            // ================= relay (peer CTA): forward local events to the leader's issuer =================
            const uint32_t r_a = map_to_rank(smem_u32(&bars->peer_a_full), 0);
            uint32_t stage = 0, phase = 0, it = 0;
            for (int64_t pass = cluster_id; pass < n_passes; pass += n_clusters, ++it) {
                mbar_wait(&bars->a_full, it & 1);
                mbar_arrive_remote(r_a);
                for (int c = 0; c < K; ++c) {
                    mbar_wait(&bars->full[stage], phase);
                    mbar_arrive_remote(map_to_rank(smem_u32(&bars->peer_full[stage]), 0));
                    if (++stage == T3_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ================= producer (one thread per CTA): this CTA's operand rows + the component's offsets =================
        if (lane == 0) {
            const uint32_t tx = dense ? T3_STAGE : T3_STAGE_TRI;
            uint32_t stage = 0, phase = 0, gc = 0;
            for (int64_t pass = cluster_id; pass < n_passes; pass += n_clusters) {
                for (int c = 0; c < K; ++c, ++gc) {
                    const uint32_t ob = gc % T3_OFFRING;
                    mbar_wait(&bars->off_empty[ob], ((gc / T3_OFFRING) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bars->off_full[ob], T3_OFFBYTES);
                    bulk_g2s(sOff + ob * T3_OFFBLK, offs + (size_t)c * T3_OFFBLK, T3_OFFBYTES, &bars->off_full[ob]);
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&bars->full[stage], tx);
                    bulk_g2s(sB + (size_t)stage * T3_STAGE, img + ((size_t)c * 2 + rank) * T3_STAGE, tx, &bars->full[stage]);
                    if (++stage == T3_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();                                   // no CTA leaves (or frees TMEM) while its partner may still signal it
    if (warp == 8) tmem_dealloc2(tmem_base, 512);
}

// ---- host side -----------------------------------------------------------------------------------------------------

// 0 = kernel off (the default: measured SLOWER than tc_estep2.cu, see the header); 16 or 32 = rows of granularity of the skip
static int g_t3_gran = 0;
int tc3_set_granularity(int g) { int old = g_t3_gran; g_t3_gran = (g == 16 || g == 32) ? g : 0; return old; }

bool tc3_supported(int D, int Rp) { return g_t3_gran != 0 && D > 64 && D <= 128 && Rp == 128; }

static size_t up1k3(size_t x) { return (x + 1023) / 1024 * 1024; }
// [image K x 2 x T3_STAGE | offsets K x T3_OFFBLK floats], 1 KB aligned by the caller
size_t tc3_workspace(int K) { return (size_t)K * 2 * T3_STAGE + up1k3((size_t)K * T3_OFFBYTES); }

// flags: head of the operand workspace ([0] max |z| bits set by tc_data_scale; [8] written here)
int tc3_prepare(const float* W, const float* cst, int K, int Dpp, int D, unsigned int* flags, void* ws3, cudaStream_t st) {
    unsigned char* img = (unsigned char*)ws3;
    float* offs = (float*)(img + (size_t)K * 2 * T3_STAGE);
    MIMO_CUDA(cudaMemsetAsync(flags + 8, 0, 4, st));
    if (g_t3_gran == 32) {
        tc3_scan_kernel<32><<<K, 256, 0, st>>>(W, K, Dpp, D, flags);
        tc3_prep_kernel<32><<<K, 256, 0, st>>>(W, cst, K, Dpp, D, flags, img, offs);
    } else {
        tc3_scan_kernel<16><<<K, 256, 0, st>>>(W, K, Dpp, D, flags);
        tc3_prep_kernel<16><<<K, 256, 0, st>>>(W, cst, K, Dpp, D, flags, img, offs);
    }
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

int tc_estep3(const float* Z, int64_t N, int D, int64_t ldz, int K, const void* ws3, const unsigned int* flags,
              float* out, int64_t ldo, const unsigned int* gate, unsigned int gate_value,
              float* lse_vals, double* lse_sum, cudaStream_t st) {
    if (N == 0) return MIMO_OK;
    const unsigned char* img = (const unsigned char*)ws3;
    const float* offs = (const float*)(img + (size_t)K * 2 * T3_STAGE);
    const size_t smem = (size_t)T3_STAGES * T3_STAGE + T3_OFFRING * T3_OFFBYTES + sizeof(T3Bars);
    const int64_t passes = (N + 255) / 256;
    const int clusters = (int)std::min<int64_t>(passes, sm_count() / 2);
    const int vec4 = (ldz % 4 == 0) && (((uintptr_t)Z & 15) == 0);
    if (g_t3_gran == 32) {
        auto kern = tc_estep3_kernel<32>;
        MIMO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<2 * clusters, T3_THREADS, smem, st>>>(Z, N, D, ldz, vec4, img, offs, flags, K, out, ldo, gate, gate_value, lse_vals, lse_sum);
    } else {
        auto kern = tc_estep3_kernel<16>;
        MIMO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<2 * clusters, T3_THREADS, smem, st>>>(Z, N, D, ldz, vec4, img, offs, flags, K, out, ldo, gate, gate_value, lse_vals, lse_sum);
    }
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

}  // namespace mimo
