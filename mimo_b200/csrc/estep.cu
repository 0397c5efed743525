// Per-point E-step kernels, CUDA-core path (FP32 / FP64), sm_100a.
//
//   quad_loglik_kernel : a[k][n] = cst[k] - 0.5 * || W_k [z_n ; 1] ||^2
//       register-tiled GEMM  Y = Zt (points x Dpp) . Wcat^T (Dpp x K*Rp)  with the
//       square-and-sum over each component's Rp rows fused into the epilogue, so
//       the (points x K*Rp) product never leaves registers.
//   diag_loglik_kernel : a[k][n] = cst[k] - 0.5 * sum_j (S_kj z_nj - T_kj)^2
//   softmax_kernel     : log-normaliser, responsibilities, inverse-CDF label draw.
//
// Reference call sites replaced: distributions/gaussian.py:510-523, 837-850;
// lingauss.py:330-347; bayesian.py:287-301, 446-460, 933-947; mixtures/gmm.py:72-75,
// 256-259; utils/stats.py:8-21.
#include "common.cuh"
#include <type_traits>

namespace mimo {

constexpr int QUAD_THREADS = 256;
constexpr int QUAD_BN = 128;   // flat (component,row) columns per chunk: 16 thread columns x 8
constexpr int QUAD_PAD = 4;

// grid.x = point tiles of BM = 16*TM points; every CTA walks all K*Rp rows in chunks of 128.
// Thread (tx, ty) owns rows {4tx..4tx+3} U {64+4tx..64+4tx+3} of the chunk and, for TM = 8,
// points {4ty..4ty+3} U {64+4ty..64+4ty+3}: consecutive lanes read consecutive 16-byte
// vectors of shared memory (conflict-free 128-bit loads).
template <typename T, int TM>
__global__ void __launch_bounds__(QUAD_THREADS)
quad_loglik_kernel(const T* __restrict__ Z, int64_t N, int D, int64_t ldz,
                   const T* __restrict__ W, const T* __restrict__ cst, int K, int Rp, int Dpp,
                   T* __restrict__ out, int64_t ldo) {
    constexpr int BM = 16 * TM;
    constexpr int ZS = BM + QUAD_PAD;        // smem row strides (keep 16-byte alignment)
    constexpr int WS = QUAD_BN + QUAD_PAD;
    constexpr int VEC = 16 / sizeof(T);      // elements per 128-bit global load
    using V = typename VecOf<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Zs = reinterpret_cast<T*>(smem_raw);          // [Dpp][ZS]   transposed point tile, row D == 1
    T* Ws = Zs + (size_t)Dpp * ZS;                   // [Dpp][WS]   transposed operand chunk

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t n0 = (int64_t)blockIdx.x * BM;
    const int64_t total_rows = (int64_t)K * Rp;

    // ---- stage the point tile, transposed: lanes walk points, so the shared-memory stores are
    //      conflict-free and each lane streams its own row of Z (sector reuse through L1) ----
    for (int idx = tid; idx < BM * D; idx += QUAD_THREADS) {
        int p = idx % BM, j = idx / BM;
        int64_t n = n0 + p;
        Zs[j * ZS + p] = (n < N) ? Z[n * ldz + j] : T(0);
    }
    for (int idx = tid; idx < BM * (Dpp - D); idx += QUAD_THREADS) {
        int j = D + idx / BM, p = idx % BM;
        Zs[j * ZS + p] = (j == D) ? T(1) : T(0);
    }

    const int nvec = Dpp / VEC;                      // 128-bit vectors per operand row
    for (int64_t row0 = 0; row0 < total_rows; row0 += QUAD_BN) {
        __syncthreads();                             // Zs ready / previous chunk consumed
        // operand chunk: lane -> row (conflict-free transposed stores), 4 independent 128-bit loads in flight
        for (int v0 = tid; v0 < QUAD_BN * nvec; v0 += 4 * QUAD_THREADS) {
            V tmp[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int v = v0 + u * QUAD_THREADS;
                int r = v & (QUAD_BN - 1), jv = v >> 7;
                int64_t row = row0 + r;
                if (v < QUAD_BN * nvec && row < total_rows)
                    tmp[u] = __ldg(reinterpret_cast<const V*>(W + row * Dpp) + jv);
                else
                    tmp[u] = V{};
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int v = v0 + u * QUAD_THREADS;
                if (v < QUAD_BN * nvec) {
                    int r = v & (QUAD_BN - 1), jv = v >> 7;
                    const T* e = reinterpret_cast<const T*>(&tmp[u]);
#pragma unroll
                    for (int c = 0; c < VEC; ++c) Ws[(jv * VEC + c) * WS + r] = e[c];
                }
            }
        }
        __syncthreads();

        T acc[TM][8];
#pragma unroll
        for (int m = 0; m < TM; ++m)
#pragma unroll
            for (int n = 0; n < 8; ++n) acc[m][n] = T(0);

        const T* zp = Zs + ty * 4;
        const T* wp = Ws + tx * 4;
#pragma unroll 4
        for (int j = 0; j < Dpp; ++j) {
            T a[TM], b[8];
            {
                T lo[4];
                lds_vec<T, 4>(lo, zp + j * ZS);
#pragma unroll
                for (int m = 0; m < 4; ++m) a[m] = lo[m];
                if constexpr (TM == 8) {
                    T hi[4];
                    lds_vec<T, 4>(hi, zp + j * ZS + 64);
#pragma unroll
                    for (int m = 0; m < 4; ++m) a[4 + m] = hi[m];
                }
                T blo[4], bhi[4];
                lds_vec<T, 4>(blo, wp + j * WS);
                lds_vec<T, 4>(bhi, wp + j * WS + 64);
#pragma unroll
                for (int n = 0; n < 4; ++n) { b[n] = blo[n]; b[4 + n] = bhi[n]; }
            }
#pragma unroll
            for (int m = 0; m < TM; ++m)
#pragma unroll
                for (int n = 0; n < 8; ++n) acc[m][n] = fma(a[m], b[n], acc[m][n]);
        }

        // ---- epilogue: sum of squares over each component's rows ----
        // half h of the thread's rows = chunk rows 64h + 4tx .. +3, all inside one component
        T q[2][TM];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int m = 0; m < TM; ++m) {
                T s = T(0);
#pragma unroll
                for (int n = 0; n < 4; ++n) s = fma(acc[m][4 * h + n], acc[m][4 * h + n], s);
                q[h][m] = s;
            }
        const int G4 = (Rp >= 64) ? 16 : (Rp >> 2);   // lanes (consecutive tx) sharing a component half
        for (int o = 1; o < G4; o <<= 1) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int m = 0; m < TM; ++m) q[h][m] += __shfl_xor_sync(0xffffffffu, q[h][m], o);
        }
        if ((tx & (G4 - 1)) == 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (Rp == 128 && h == 1) break;       // both halves belong to the same component
                int64_t row = row0 + 64 * h + 4 * tx;
                if (row >= total_rows) continue;
                int k = (int)(row / Rp);
                T ck = cst[k];
#pragma unroll
                for (int m = 0; m < TM; ++m) {
                    int p = (m < 4) ? (ty * 4 + m) : (64 + ty * 4 + (m - 4));
                    T qq = (Rp == 128) ? (q[0][m] + q[1][m]) : q[h][m];
                    if (n0 + p < N) out[(int64_t)k * ldo + n0 + p] = ck - T(0.5) * qq;
                }
            }
        }
    }
}

constexpr int DIAG_THREADS = 256;
constexpr int DIAG_BM = 128;     // points per CTA (32 lanes x 4 points)
constexpr int DIAG_KC = 32;      // components staged per chunk (8 warps x 4)

template <typename T>
__global__ void __launch_bounds__(DIAG_THREADS)
diag_loglik_kernel(const T* __restrict__ Z, int64_t N, int D, int64_t ldz,
                   const T* __restrict__ S, const T* __restrict__ Tm, const T* __restrict__ cst, int K,
                   T* __restrict__ out, int64_t ldo, const unsigned int* __restrict__ gate, unsigned int gate_value) {
    if (gate != nullptr && *gate != gate_value) return;      // the tensor-core kernel of tc_diag.cu ran instead
    constexpr int ZS = DIAG_BM + QUAD_PAD;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Zs = reinterpret_cast<T*>(smem_raw);              // [D][ZS]
    T* STs = Zs + (size_t)D * ZS;                        // [KC][D][2]  interleaved (s, t)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n0 = (int64_t)blockIdx.x * DIAG_BM;
    for (int idx = tid; idx < DIAG_BM * D; idx += DIAG_THREADS) {
        int p = idx / D, j = idx - p * D;
        int64_t n = n0 + p;
        Zs[j * ZS + p] = (n < N) ? Z[n * ldz + j] : T(0);
    }
    for (int k0 = 0; k0 < K; k0 += DIAG_KC) {
        __syncthreads();
        for (int idx = tid; idx < DIAG_KC * D; idx += DIAG_THREADS) {
            int kk = idx / D, j = idx - kk * D;
            bool ok = (k0 + kk) < K;
            STs[2 * idx]     = ok ? S[(int64_t)(k0 + kk) * D + j] : T(0);
            STs[2 * idx + 1] = ok ? Tm[(int64_t)(k0 + kk) * D + j] : T(0);
        }
        __syncthreads();
        T q[4][4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int m = 0; m < 4; ++m) q[c][m] = T(0);
        const T* zp = Zs + lane * 4;
        const T* st = STs + (size_t)(warp * 4) * D * 2;
        for (int j = 0; j < D; ++j) {
            T z[4];
            lds_vec<T, 4>(z, zp + j * ZS);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                T s = st[(c * D + j) * 2], t = st[(c * D + j) * 2 + 1];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    T y = fma(s, z[m], -t);
                    q[c][m] = fma(y, y, q[c][m]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int k = k0 + warp * 4 + c;
            if (k < K) {
                T ck = cst[k];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    int64_t n = n0 + lane * 4 + m;
                    if (n < N) out[(int64_t)k * ldo + n] = ck - T(0.5) * q[c][m];
                }
            }
        }
    }
}

// One thread per point over a (K, ldo) log-joint tile.  Follows the reference's
// order of operations: lse = max + log(sum exp(a - max)) (scipy logsumexp), p = exp(a - lse),
// cdf = sequential FP64 cumsum over k, label = #{k : u * cdf[K-1] > cdf[k]}.
template <typename T>
__global__ void __launch_bounds__(256)
softmax_kernel(T* __restrict__ a, int K, int64_t n, int64_t ldo, int flags,
               T* __restrict__ lse_out, const double* __restrict__ uniforms, uint64_t seed, uint64_t point_offset,
               int32_t* __restrict__ labels, double* __restrict__ lse_sum,
               const unsigned int* __restrict__ gate, unsigned int gate_value) {
    if (gate != nullptr && *gate != gate_value) return;      // the list-based log-normaliser of tc_screen.cu ran instead
    __shared__ double red[32];
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double my_lse = 0.0;
    if (i < n) {
        T* col = a + i;
        T mx = col[0];
        for (int k = 1; k < K; ++k) mx = max(mx, col[(int64_t)k * ldo]);
        double s = 0.0;
        for (int k = 0; k < K; ++k) s += (double)exp_t<T>(col[(int64_t)k * ldo] - mx);
        T lse = mx + (T)log(s);
        my_lse = (double)mx + log(s);
        if (flags & MIMO_WRITE_LSE) lse_out[i] = lse;
        if (flags & MIMO_DRAW_LABELS) {
            double tot = 0.0;
            for (int k = 0; k < K; ++k) tot += (double)exp_t<T>(col[(int64_t)k * ldo] - lse);
            double u = uniforms ? uniforms[i] : philox_uniform(seed, point_offset + (uint64_t)i);
            double thr = u * tot, cdf = 0.0;
            int z = 0;
            for (int k = 0; k < K; ++k) {
                T p = exp_t<T>(col[(int64_t)k * ldo] - lse);
                cdf += (double)p;
                z += (thr > cdf) ? 1 : 0;
                if (flags & MIMO_WRITE_RESP) col[(int64_t)k * ldo] = p;
            }
            labels[i] = z;
        } else if (flags & MIMO_WRITE_RESP) {
            for (int k = 0; k < K; ++k) col[(int64_t)k * ldo] = exp_t<T>(col[(int64_t)k * ldo] - lse);
        }
    }
    if (flags & MIMO_ACC_LSE) {
        double tot = block_sum<double>(my_lse, red);
        if (threadIdx.x == 0) atomicAdd(lse_sum, tot);
    }
}

// ---- host launchers ------------------------------------------------------------

template <typename T, int TM>
static int launch_quad(const void* Z, int64_t N, int D, int64_t ldz, const void* W, const void* cst,
                       int K, int Rp, int Dpp, void* out, int64_t ldo, cudaStream_t st) {
    constexpr int BM = 16 * TM;
    size_t smem = (size_t)Dpp * ((BM + QUAD_PAD) + (QUAD_BN + QUAD_PAD)) * sizeof(T);
    if (smem > 227 * 1024) { set_error("quad E-step: D=%d needs %zu B of shared memory", D, smem); return MIMO_EUNSUPPORTED; }
    auto kern = quad_loglik_kernel<T, TM>;
    MIMO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<cdiv(N, BM), QUAD_THREADS, smem, st>>>((const T*)Z, N, D, ldz, (const T*)W, (const T*)cst,
                                                   K, Rp, Dpp, (T*)out, ldo);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

int loglik_quad(int dtype, const void* Z, int64_t N, int D, int64_t ldz, const void* W, const void* cst,
                int K, int Rp, int Dpp, void* out, int64_t ldo, cudaStream_t st) {
    MIMO_CHECK_ARG(dtype == MIMO_F32 || dtype == MIMO_F64, "dtype");
    MIMO_CHECK_ARG(Z && W && cst && out, "null pointer");
    MIMO_CHECK_ARG(N >= 0 && D >= 1 && K >= 1 && ldz >= D && ldo >= N, "shape");
    MIMO_CHECK_ARG(Rp >= 8 && Rp <= 128 && (Rp & (Rp - 1)) == 0, "Rp must be a power of two in [8,128]");
    MIMO_CHECK_ARG(Dpp >= D + 1 && Dpp % 4 == 0, "Dpp must be >= D+1 and a multiple of 4");
    if (N == 0) return MIMO_OK;
    if (dtype == MIMO_F32) return launch_quad<float, 8>(Z, N, D, ldz, W, cst, K, Rp, Dpp, out, ldo, st);
    return launch_quad<double, 4>(Z, N, D, ldz, W, cst, K, Rp, Dpp, out, ldo, st);
}

template <typename T>
static int launch_diag(const void* Z, int64_t N, int D, int64_t ldz, const void* S, const void* Tm, const void* cst,
                       int K, void* out, int64_t ldo, cudaStream_t st, const unsigned int* gate, unsigned int gate_value) {
    size_t smem = ((size_t)D * (DIAG_BM + QUAD_PAD) + (size_t)DIAG_KC * D * 2) * sizeof(T);
    if (smem > 227 * 1024) { set_error("diag E-step: D=%d needs %zu B of shared memory", D, smem); return MIMO_EUNSUPPORTED; }
    auto kern = diag_loglik_kernel<T>;
    MIMO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<cdiv(N, DIAG_BM), DIAG_THREADS, smem, st>>>((const T*)Z, N, D, ldz, (const T*)S, (const T*)Tm,
                                                        (const T*)cst, K, (T*)out, ldo, gate, gate_value);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

int loglik_diag(int dtype, const void* Z, int64_t N, int D, int64_t ldz, const void* S, const void* Tm,
                const void* cst, int K, void* out, int64_t ldo, cudaStream_t st, const unsigned int* gate, unsigned int gate_value) {
    MIMO_CHECK_ARG(dtype == MIMO_F32 || dtype == MIMO_F64, "dtype");
    MIMO_CHECK_ARG(Z && S && Tm && cst && out, "null pointer");
    MIMO_CHECK_ARG(N >= 0 && D >= 1 && K >= 1 && ldz >= D && ldo >= N, "shape");
    if (N == 0) return MIMO_OK;
    if (dtype == MIMO_F32) return launch_diag<float>(Z, N, D, ldz, S, Tm, cst, K, out, ldo, st, gate, gate_value);
    return launch_diag<double>(Z, N, D, ldz, S, Tm, cst, K, out, ldo, st, gate, gate_value);
}

int softmax(int dtype, void* a, int K, int64_t n, int64_t ldo, int flags, void* lse, const void* uniforms,
            uint64_t seed, uint64_t point_offset, int32_t* labels, double* lse_sum, cudaStream_t st,
            const unsigned int* gate, unsigned int gate_value) {
    MIMO_CHECK_ARG(dtype == MIMO_F32 || dtype == MIMO_F64, "dtype");
    MIMO_CHECK_ARG(a && K >= 1 && n >= 0 && ldo >= n, "shape");
    MIMO_CHECK_ARG(!(flags & MIMO_WRITE_LSE) || lse, "lse output missing");
    MIMO_CHECK_ARG(!(flags & MIMO_DRAW_LABELS) || labels, "labels output missing");
    MIMO_CHECK_ARG(!(flags & MIMO_ACC_LSE) || lse_sum, "lse_sum output missing");
    if (n == 0) return MIMO_OK;
    int grid = cdiv(n, 256);
    if (dtype == MIMO_F32)
        softmax_kernel<float><<<grid, 256, 0, st>>>((float*)a, K, n, ldo, flags, (float*)lse, (const double*)uniforms,
                                                    seed, point_offset, labels, lse_sum, gate, gate_value);
    else
        softmax_kernel<double><<<grid, 256, 0, st>>>((double*)a, K, n, ldo, flags, (double*)lse, (const double*)uniforms,
                                                     seed, point_offset, labels, lse_sum, gate, gate_value);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

}  // namespace mimo
