// CTA-cooperative dense linear algebra on small FP64 matrices (n <= ~160), used by the
// batched posterior kernels.  Every routine is called by ALL threads of the CTA and
// contains __syncthreads(); matrices are row-major with leading dimension n and may
// live in shared or global memory (generic pointers).
#pragma once
#include "common.cuh"

namespace mimo {

__device__ __forceinline__ int tri_idx(int a, int b) {     // packed lower-triangular index of (a,b)
    return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a;
}

// A = L L^T in place (lower triangle overwritten by L; strict upper triangle untouched).
// Returns false (uniformly) if a pivot is not positive.  logdet_half = sum log L_ii.
__device__ inline bool cta_chol_lower(double* A, int n, double* logdet_half) {
    const int tid = threadIdx.x, nt = blockDim.x;
    double ld = 0.0;
    for (int j = 0; j < n; ++j) {
        __syncthreads();
        double djj = A[j * n + j];
        if (!(djj > 0.0) || !isfinite(djj)) return false;
        double l = sqrt(djj);
        ld += log(l);
        __syncthreads();
        if (tid == 0) A[j * n + j] = l;
        for (int i = j + 1 + tid; i < n; i += nt) A[i * n + j] /= l;
        __syncthreads();
        const int m = n - j - 1;
        for (int idx = tid; idx < m * m; idx += nt) {
            int i = j + 1 + idx / m, k = j + 1 + idx % m;
            if (k <= i) A[i * n + k] -= A[i * n + j] * A[k * n + j];
        }
    }
    __syncthreads();
    if (logdet_half) *logdet_half = ld;
    return true;
}

// X = L^{-1} (lower triangular, strict upper part zeroed).  X must not alias L.
__device__ inline void cta_tri_inv_lower(const double* L, double* X, int n) {
    const int tid = threadIdx.x, nt = blockDim.x;
    __syncthreads();
    for (int j = tid; j < n; j += nt) {
        for (int i = 0; i < j; ++i) X[i * n + j] = 0.0;
        X[j * n + j] = 1.0 / L[j * n + j];
        for (int i = j + 1; i < n; ++i) {
            double s = 0.0;
            for (int m = j; m < i; ++m) s = fma(L[i * n + m], X[m * n + j], s);
            X[i * n + j] = -s / L[i * n + i];
        }
    }
    __syncthreads();
}

// P = X^T X for lower-triangular X (full symmetric result).  P must not alias X.
__device__ inline void cta_gram_lower(const double* X, double* P, int n) {
    const int tid = threadIdx.x, nt = blockDim.x;
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += nt) {
        int i = idx / n, j = idx - i * n;
        if (j > i) continue;
        double s = 0.0;
        for (int m = i; m < n; ++m) s = fma(X[m * n + i], X[m * n + j], s);
        P[i * n + j] = s;
        P[j * n + i] = s;
    }
    __syncthreads();
}

// Tm = C * A for lower-triangular C and A (lower-triangular result, strict upper zeroed).
__device__ inline void cta_trmm_lower(const double* C, const double* A, double* Tm, int n) {
    const int tid = threadIdx.x, nt = blockDim.x;
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += nt) {
        int i = idx / n, j = idx - i * n;
        double s = 0.0;
        if (j <= i)
            for (int m = j; m <= i; ++m) s = fma(C[i * n + m], A[m * n + j], s);
        Tm[idx] = s;
    }
    __syncthreads();
}

// A <- C * A in place for lower-triangular C and A: column j of the product only needs
// column j of A, and walking i downwards never reads an entry already overwritten.
__device__ inline void cta_trmm_lower_inplace(const double* C, double* A, int n) {
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        for (int i = n - 1; i >= j; --i) {
            double s = 0.0;
            for (int m = j; m <= i; ++m) s = fma(C[i * n + m], A[m * n + j], s);
            A[i * n + j] = s;
        }
    }
    __syncthreads();
}

// Solve T^T v = z for lower-triangular T (back substitution), nrhs right-hand sides stored
// as columns of V (ld = ldv): V[:, r] <- T^-T V[:, r].  One warp per right-hand side.
__device__ inline void cta_solve_lower_T(const double* Tm, int n, double* V, int ldv, int nrhs) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    for (int r = warp; r < nrhs; r += nw) {
        for (int i = n - 1; i >= 0; --i) {
            double s = 0.0;
            for (int m = i + 1 + lane; m < n; m += 32) s = fma(Tm[m * n + i], V[m * ldv + r], s);
            s = warp_sum(s);
            if (lane == 0) V[i * ldv + r] = (V[i * ldv + r] - s) / Tm[i * n + i];
            __syncwarp();
        }
    }
    __syncthreads();
}

}  // namespace mimo
