// Shared device/host helpers for the mimo_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/mimo_b200.h"

namespace mimo {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char* fmt, ...);

#define MIMO_CHECK_ARG(cond, msg)                                        \
    do { if (!(cond)) { ::mimo::set_error("invalid argument: %s (%s:%d)", msg, __FILE__, __LINE__); \
                        return MIMO_EINVAL; } } while (0)
#define MIMO_CUDA(expr)                                                  \
    do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) {            \
             ::mimo::set_error("CUDA error %s at %s:%d", cudaGetErrorString(e__), __FILE__, __LINE__); \
             return MIMO_ECUDA; } } while (0)
#define MIMO_LAUNCH_CHECK() MIMO_CUDA(cudaGetLastError())

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
int sm_count();

// ---- Philox4x32-10 (counter-based; keyed by seed, counter = global point index) ----
__host__ __device__ inline uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
__host__ __device__ inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi32(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = mulhi32(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
// 53-bit uniform in [0,1) from (seed, index): independent of how points are sharded.
__host__ __device__ inline double philox_uniform(uint64_t seed, uint64_t index) {
    uint32_t c[4] = {(uint32_t)index, (uint32_t)(index >> 32), 0x6d696d6fu, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint64_t bits = (((uint64_t)c[0] << 32) | c[1]) >> 11;
    return (double)bits * (1.0 / 9007199254740992.0);
}

#ifdef __CUDACC__
// ---- special functions (FP64) -------------------------------------------------
// digamma: recurrence up to x >= 10, then the asymptotic series.  Checked against
// scipy.special.digamma to < 1e-13 relative on (0, 1e6] in tests/test_gpu_special.py.
__device__ inline double digamma_d(double x) {
    double r = 0.0;
    while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
    double f = 1.0 / (x * x);
    double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0
               + f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
    return r + log(x) - 0.5 / x + t;
}
// log multivariate gamma  (scipy.special.multigammaln)
__device__ inline double multigammaln_d(double a, int d) {
    double s = 0.25 * d * (d - 1) * 1.1447298858494001741434;  // log(pi)
    for (int i = 0; i < d; ++i) s += lgamma(a - 0.5 * i);
    return s;
}

// ---- reductions -------------------------------------------------------------
template <typename T>
__device__ inline T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum; result valid in thread 0.  `red` is >= 32 elements of shared memory.
template <typename T>
__device__ inline T block_sum(T v, T* red) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (lane < (int)((blockDim.x + 31) >> 5)) ? red[lane] : T(0);
        v = warp_sum(v);
    }
    return v;
}

template <typename T> struct VecOf;
template <> struct VecOf<float>  { using type = float4;  static constexpr int n = 4; };
template <> struct VecOf<double> { using type = double2; static constexpr int n = 2; };
// N contiguous, 16-byte aligned elements -> registers with 128-bit loads
template <typename T, int N>
__device__ __forceinline__ void lds_vec(T (&dst)[N], const T* src) {
    using V = typename VecOf<T>::type;
    constexpr int PER = VecOf<T>::n;
#pragma unroll
    for (int i = 0; i < N / PER; ++i)
        *reinterpret_cast<V*>(&dst[i * PER]) = reinterpret_cast<const V*>(src)[i];
}

template <typename T> __device__ inline T exp_t(T x);
template <> __device__ inline float exp_t<float>(float x) { return expf(x); }
template <> __device__ inline double exp_t<double>(double x) { return exp(x); }
template <typename T> __device__ inline T log_t(T x);
template <> __device__ inline float log_t<float>(float x) { return logf(x); }
template <> __device__ inline double log_t<double>(double x) { return log(x); }
#endif

}  // namespace mimo
