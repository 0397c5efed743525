// Prediction path of the mixtures of linear-Gaussian experts, sm_100a.
//
// replaces mixtures/ilr.py:325-430 (meanfield_prediction: predictive moments of every expert, their mixture / mode,
// negative log predictive density) together with distributions/bayesian.py:876-912, 949-985 (posterior predictive of
// the (tied) Matrix-Normal-Wishart: mu_kn = M_k x~_n, c_kn = 1 + x~_n^T K_k^-1 x~_n, Lambda_kn = Psi_k df_k / c_kn) and
// utils/stats.py:53-79 (the Student-t form of the basis weights).  The (K, N, o, o) arrays of the reference are never
// built: one thread walks the K experts of its point and keeps the mixture sums in registers (FP64).
#include <algorithm>
#include "common.cuh"
#include "internal.h"

namespace mimo {

constexpr int PR_MAXO = 4;          // output dimensions

// a[k][n] <- add[k] + log1p(2 (c0[k] - a[k][n]) / df[k]): the reference's Student-t "log-density" (stats.py:53-79 keeps
// the -(df + d)/2 factor inside the constant) from the Gaussian-form log-joint a = c0 - delta / 2 of the E-step kernel
template <typename T>
__global__ void studentt_from_quad_kernel(T* __restrict__ a, int K, int64_t N, int64_t lda, const double* __restrict__ c0,
                                          const double* __restrict__ add, const double* __restrict__ df) {
    const int k = blockIdx.y;
    const double ck = c0[k], ak = add[k], inv = 1.0 / df[k];
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        const double delta = fmax(0.0, 2.0 * (ck - (double)a[(int64_t)k * lda + n]));
        a[(int64_t)k * lda + n] = (T)(ak + log1p(delta * inv));
    }
}

// thread = point.  W (K, ldw): predictive weights (columns sum to one).  Per expert k: M (o, c), Kinv (c, c), Sigma =
// Psi^-1 (o, o), Psi (o, o), logdet Psi, df; tied: Sigma / Psi / logdet / df are shared (stride 0).
// mode 0: mixture mean and covariance (ilr.py:375-383); mode 1: the moments of the expert with the largest weight.
// studentt: covariances scaled by df / (df - 2) (ilr.py:356-361).  Y (optional): nlpd_n = -logsumexp_k(log N(y_n;
// mu_kn, Lambda_kn) + log(w_kn + eps)) (ilr.py:407-411; the reference always takes the Gaussian form here).
template <typename T>
__global__ void __launch_bounds__(128)
predict_lingauss_kernel(const T* __restrict__ X, int64_t N, int64_t ldx, int din, int affine,
                        const T* __restrict__ W, int64_t ldw, int K,
                        const double* __restrict__ M, const double* __restrict__ Kinv, const double* __restrict__ Sig,
                        const double* __restrict__ Psi, const double* __restrict__ logdet, const double* __restrict__ df,
                        int o, int c, int tied, int mode, int studentt,
                        const T* __restrict__ Y, int64_t ldy, double eps,
                        T* __restrict__ mu_out, T* __restrict__ cov_out, T* __restrict__ nlpd_out) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const T* x = X + n * ldx;
    double mu[PR_MAXO], cov[PR_MAXO * PR_MAXO], y[PR_MAXO];
#pragma unroll
    for (int i = 0; i < PR_MAXO; ++i) { mu[i] = 0.0; y[i] = (Y != nullptr && i < o) ? (double)Y[n * ldy + i] : 0.0; }
#pragma unroll
    for (int i = 0; i < PR_MAXO * PR_MAXO; ++i) cov[i] = 0.0;
    double wbest = -1.0, lmax = -INFINITY, lsum = 0.0;
    const int ps = tied ? 0 : 1;
    for (int k = 0; k < K; ++k) {
        const double w = (double)W[(int64_t)k * ldw + n];
        const double* Mk = M + (size_t)k * o * c;
        const double* Ki = Kinv + (size_t)k * c * c;
        const double* Sk = Sig + (size_t)k * ps * o * o;
        const double* Pk = Psi + (size_t)k * ps * o * o;
        const double dfk = df[k * ps];
        // c_kn = 1 + x~^T K^-1 x~,  mu_kn = M x~   (x~ = [x ; 1] for affine experts)
        double q = 0.0, mk[PR_MAXO];
#pragma unroll
        for (int i = 0; i < PR_MAXO; ++i) mk[i] = 0.0;
        for (int i = 0; i < c; ++i) {
            const double xi = (i < din) ? (double)x[i] : 1.0;
            double t = 0.0;
            for (int j = 0; j < c; ++j) t += Ki[i * c + j] * ((j < din) ? (double)x[j] : 1.0);
            q += xi * t;
#pragma unroll
            for (int r = 0; r < PR_MAXO; ++r) if (r < o) mk[r] += Mk[r * c + i] * xi;
        }
        (void)affine;
        const double ckn = 1.0 + q;
        const double cs = ckn / dfk * (studentt ? dfk / (dfk - 2.0) : 1.0);     // covar_kn = Sigma_k * cs
        if (mode == 0) {
#pragma unroll
            for (int r = 0; r < PR_MAXO; ++r) {
                if (r < o) {
                    mu[r] += w * mk[r];
#pragma unroll
                    for (int s = 0; s < PR_MAXO; ++s) if (s < o) cov[r * PR_MAXO + s] += w * (Sk[r * o + s] * cs + mk[r] * mk[s]);
                }
            }
        } else if (w > wbest) {
            wbest = w;
#pragma unroll
            for (int r = 0; r < PR_MAXO; ++r) {
                if (r < o) {
                    mu[r] = mk[r];
#pragma unroll
                    for (int s = 0; s < PR_MAXO; ++s) if (s < o) cov[r * PR_MAXO + s] = Sk[r * o + s] * cs;
                }
            }
        }
        if (Y != nullptr) {
            double md = 0.0;
#pragma unroll
            for (int r = 0; r < PR_MAXO; ++r) {
                if (r < o) {
#pragma unroll
                    for (int s = 0; s < PR_MAXO; ++s) if (s < o) md += (y[r] - mk[r]) * Pk[r * o + s] * (y[s] - mk[s]);
                }
            }
            const double sc = dfk / ckn;
            const double lp = -0.5 * sc * md + 0.5 * (logdet[k * ps] + o * log(sc)) - 0.5 * o * 1.8378770664093453 + log(w + eps);
            if (lp > lmax) { lsum = lsum * exp(lmax - lp) + 1.0; lmax = lp; }
            else lsum += exp(lp - lmax);
        }
    }
    for (int r = 0; r < o; ++r) {
        mu_out[n * o + r] = (T)mu[r];
        for (int s = 0; s < o; ++s)
            cov_out[(n * o + r) * o + s] = (T)(cov[r * PR_MAXO + s] - (mode == 0 ? mu[r] * mu[s] : 0.0));
    }
    if (Y != nullptr && nlpd_out != nullptr) nlpd_out[n] = (T)(-(lmax + log(lsum)));
}

int studentt_from_quad(int dtype, void* a, int K, int64_t N, int64_t lda, const double* c0, const double* add, const double* df,
                       cudaStream_t st) {
    MIMO_CHECK_ARG(dtype == MIMO_F32 || dtype == MIMO_F64, "dtype");
    MIMO_CHECK_ARG(a && c0 && add && df && K >= 1 && N >= 0 && lda >= N, "shape");
    if (N == 0) return MIMO_OK;
    dim3 grid((unsigned)std::min<int64_t>((N + 255) / 256, 4096), (unsigned)K);
    if (dtype == MIMO_F32) studentt_from_quad_kernel<float><<<grid, 256, 0, st>>>((float*)a, K, N, lda, c0, add, df);
    else studentt_from_quad_kernel<double><<<grid, 256, 0, st>>>((double*)a, K, N, lda, c0, add, df);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

int predict_lingauss(int dtype, const void* X, int64_t N, int64_t ldx, int din, int affine, const void* W, int64_t ldw, int K,
                     const double* M, const double* Kinv, const double* Sig, const double* Psi, const double* logdet, const double* df,
                     int o, int tied, int mode, int studentt, const void* Y, int64_t ldy, double eps,
                     void* mu_out, void* cov_out, void* nlpd_out, cudaStream_t st) {
    MIMO_CHECK_ARG(dtype == MIMO_F32 || dtype == MIMO_F64, "dtype");
    MIMO_CHECK_ARG(X && W && M && Kinv && Sig && Psi && logdet && df && mu_out && cov_out, "null pointer");
    MIMO_CHECK_ARG(N >= 0 && din >= 1 && ldx >= din && ldw >= N && K >= 1 && (mode == 0 || mode == 1), "shape");
    if (o < 1 || o > PR_MAXO) { set_error("prediction: output dimension %d not in [1, %d]", o, PR_MAXO); return MIMO_EUNSUPPORTED; }
    if (N == 0) return MIMO_OK;
    const int c = din + (affine ? 1 : 0);
    const unsigned grid = (unsigned)((N + 127) / 128);
    if (dtype == MIMO_F32)
        predict_lingauss_kernel<float><<<grid, 128, 0, st>>>((const float*)X, N, ldx, din, affine, (const float*)W, ldw, K, M, Kinv, Sig, Psi,
                                                             logdet, df, o, c, tied, mode, studentt, (const float*)Y, ldy, eps,
                                                             (float*)mu_out, (float*)cov_out, (float*)nlpd_out);
    else
        predict_lingauss_kernel<double><<<grid, 128, 0, st>>>((const double*)X, N, ldx, din, affine, (const double*)W, ldw, K, M, Kinv, Sig, Psi,
                                                              logdet, df, o, c, tied, mode, studentt, (const double*)Y, ldy, eps,
                                                              (double*)mu_out, (double*)cov_out, (double*)nlpd_out);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

}  // namespace mimo
