// Batched per-component posterior kernels (FP64), sm_100a: one CTA per component.
//
// Conjugate updates are done in "inverse-scale" form: with P = Psi'^-1 = L L^T the
// whitening operand the E-step needs is L^-1 itself, so neither the prior nor the
// posterior scale matrix is ever inverted explicitly on the sweep path:
//   Normal-Wishart   kappa' = kappa0 + n,  m' = (kappa0 m0 + Sx)/kappa',  nu' = nu0 + n,
//                    P = Psi0^-1 + kappa0 m0 m0^T + Sxx - kappa' m' m'^T
//   Matrix-N-Wishart K' = K0 + Sxx,  M' = (M0 K0 + Syx) K'^-1,  nu' = nu0 + n,
//                    P = Psi0^-1 + M0 K0 M0^T + Syy - M' K' M'^T
//   Normal-Gamma     kappa' = kappa0 + n, m' likewise, alpha' = alpha0 + n/2,
//                    beta' = beta0 + (Sxx + kappa0 m0^2 - kappa' m'^2)/2
// (restated from distributions/composite.py:50-72, 313-337, 577-599 and verified against
// the reference through the oracle).  Each kernel also emits the packed E-step operands
// (W | S,T and cst), the API-visible standard parameters and the lower-bound term
// entropy - cross-entropy (bayesian.py:240-243).
#include "linalg.cuh"

namespace mimo {

constexpr int PT_THREADS = 256;
constexpr double LOG_2PI = 1.8378770664093454835606594728112;
constexpr double LOG_2 = 0.6931471805599453094172321214582;

__device__ __forceinline__ void store_op(void* base, int dtype, size_t idx, double v) {
    if (dtype == MIMO_F32) reinterpret_cast<float*>(base)[idx] = (float)v;
    else reinterpret_cast<double*>(base)[idx] = v;
}
__device__ __forceinline__ void add_op(void* base, int dtype, size_t idx, double v) {
    if (dtype == MIMO_F32) reinterpret_cast<float*>(base)[idx] += (float)v;
    else reinterpret_cast<double*>(base)[idx] += v;
}
__device__ inline void flag_fail(int32_t* info, int code, int k) {
    if (threadIdx.x == 0 && atomicCAS(&info[0], 0, code) == 0) info[1] = k;
}
// E[log det Lambda] under Wishart(Psi, nu):  sum_i digamma((nu - i)/2) + d log 2 + log det Psi
__device__ inline double wishart_elogdet(double nu, int d, double logdet_psi) {
    double s = 0.0;
    for (int i = 0; i < d; ++i) s += digamma_d(0.5 * (nu - i));
    return s + d * LOG_2 + logdet_psi;
}
// Wishart log-partition (wishart.py:129-132) with log det Psi supplied
__device__ inline double wishart_logz(double nu, int d, double logdet_psi) {
    return 0.5 * nu * d * LOG_2 + multigammaln_d(0.5 * nu, d) + 0.5 * nu * logdet_psi;
}

// Two n*n FP64 work matrices per CTA: shared memory when they fit, else global scratch.
struct WorkMats { double* b1; double* b2; };
__device__ inline WorkMats work_mats(int n, double* gscratch, bool use_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WorkMats w;
    if (use_smem) { w.b1 = reinterpret_cast<double*>(smem_raw); w.b2 = w.b1 + n * n; }
    else { w.b1 = gscratch + (size_t)blockIdx.x * 2 * n * n; w.b2 = w.b1 + n * n; }
    return w;
}

// ---------------------------------------------------------------------------------
// Normal-Wishart
// ---------------------------------------------------------------------------------
struct NWArgs {
    int K, d, tied, mode;
    const double *m0, *kappa0, *psi0, *nu0;
    const double* stat; int F; const int32_t* stat_idx;
    const double* variates;
    double *post_m, *post_kappa, *post_psi, *post_nu, *lik_mu, *lik_lmbda, *vlb;
    int op_dtype; void* W; void* cst; int Rp, Dpp, row_off; const int32_t* col_map;
    double *Pk, *Pbar, *scal, *mpost, *nubar, *gscratch;
    int use_smem; int32_t* info;
};

__global__ void __launch_bounds__(PT_THREADS) nw_phase_a(NWArgs a) {
    const int k = blockIdx.x, d = a.d, tid = threadIdx.x;
    WorkMats w = work_mats(d, a.gscratch, a.use_smem);
    const double* psi0 = a.psi0 + (size_t)k * d * d;
    for (int idx = tid; idx < d * d; idx += PT_THREADS) w.b1[idx] = psi0[idx];
    double ld0;
    if (!cta_chol_lower(w.b1, d, &ld0)) { flag_fail(a.info, MIMO_ENOTPD, k); return; }
    cta_tri_inv_lower(w.b1, w.b2, d);
    cta_gram_lower(w.b2, w.b1, d);                      // b1 = Psi0^-1
    const double* st = a.stat + (size_t)k * a.F;
    const int pc = a.stat_idx[d];
    const double n = st[tri_idx(pc, pc)];
    const double kap0 = a.kappa0[k], kap = kap0 + n;
    const double* m0 = a.m0 + (size_t)k * d;
    double* mp = a.mpost + (size_t)k * d;
    for (int i = tid; i < d; i += PT_THREADS) mp[i] = (kap0 * m0[i] + st[tri_idx(pc, a.stat_idx[i])]) / kap;
    __syncthreads();
    double* P = a.Pk + (size_t)k * d * d;
    for (int idx = tid; idx < d * d; idx += PT_THREADS) {
        int i = idx / d, j = idx - i * d;
        double sxx = st[tri_idx(a.stat_idx[i], a.stat_idx[j])];
        P[idx] = w.b1[idx] + kap0 * m0[i] * m0[j] + sxx - kap * mp[i] * mp[j];
    }
    if (tid == 0) {
        double* s = a.scal + (size_t)k * 8;
        s[0] = n; s[1] = kap; s[2] = a.nu0[k] + n; s[3] = 2.0 * ld0;
    }
}

// mean over components of (K, len) rows -> out (len); one thread per element
__global__ void mean_over_k(const double* __restrict__ x, int K, int len, int stride, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) {
        double s = 0.0;
        for (int k = 0; k < K; ++k) s += x[(size_t)k * stride + i];
        out[i] = s / K;
    }
}

__global__ void __launch_bounds__(PT_THREADS) nw_phase_b(NWArgs a) {
    __shared__ double red[32];
    __shared__ double sh_vec[512];                      // mu / solve scratch (d <= 512)
    const int k = blockIdx.x, d = a.d, tid = threadIdx.x;
    if (a.info[0] != 0) return;
    WorkMats w = work_mats(d, a.gscratch, a.use_smem);
    const double* Pk = a.Pk + (size_t)k * d * d;
    const double* Pu = a.tied ? a.Pbar : Pk;            // P actually used
    const double* sc = a.scal + (size_t)k * 8;
    const double n = sc[0], kap = sc[1], logdet_psi0 = sc[3];
    const double nu = a.tied ? a.nubar[0] : sc[2];
    const double* mp = a.mpost + (size_t)k * d;
    const double* st = a.stat + (size_t)k * a.F;
    (void)n;
    for (int idx = tid; idx < d * d; idx += PT_THREADS) w.b1[idx] = Pu[idx];
    double ldP;
    if (!cta_chol_lower(w.b1, d, &ldP)) { flag_fail(a.info, MIMO_ENOTPD, k); return; }
    const double logdet_psi = -2.0 * ldP;
    cta_tri_inv_lower(w.b1, w.b2, d);                   // b2 = L^-1 ;  Psi' = b2^T b2
    const double elogdet = wishart_elogdet(nu, d, logdet_psi);

    if (a.post_m) for (int i = tid; i < d; i += PT_THREADS) a.post_m[(size_t)k * d + i] = mp[i];
    if (tid == 0) {
        if (a.post_kappa) a.post_kappa[k] = kap;
        if (a.post_nu) a.post_nu[k] = nu;
    }
    const size_t wbase = ((size_t)k * a.Rp + a.row_off) * a.Dpp;
    const int cD = a.col_map ? a.col_map[d] : 0;

    if (a.mode == 0 || a.mode == 2) {
        // whitening rows  s * L^-1 [I | -m']   (mean-field: s^2 = nu';  MAP: s^2 = nu' - d)
        const double s2 = (a.mode == 0) ? nu : (nu - d);
        const double s = sqrt(s2);
        if (a.W) {
            for (int idx = tid; idx < d * d; idx += PT_THREADS) {
                int i = idx / d, j = idx - i * d;
                store_op(a.W, a.op_dtype, wbase + (size_t)i * a.Dpp + a.col_map[j], s * w.b2[idx]);
            }
            for (int i = tid; i < d; i += PT_THREADS) {
                double o = 0.0;
                for (int j = 0; j <= i; ++j) o = fma(w.b2[i * d + j], mp[j], o);
                store_op(a.W, a.op_dtype, wbase + (size_t)i * a.Dpp + cD, -s * o);
            }
            if (tid == 0) {
                double c = (a.mode == 0)
                    ? (-0.5 * d / kap + 0.5 * elogdet - 0.5 * d * LOG_2PI)
                    : (0.5 * (d * log(s2) + logdet_psi) - 0.5 * d * LOG_2PI);
                add_op(a.cst, a.op_dtype, k, c);
            }
        }
    }
    // Psi' (needed for the API, the lower bound, Gibbs and MAP parameters)
    const bool need_psi = a.post_psi || a.vlb || a.mode == 1 || (a.mode == 2 && a.lik_lmbda);
    if (need_psi) cta_gram_lower(w.b2, w.b1, d);        // b1 = Psi'
    if (a.post_psi) for (int idx = tid; idx < d * d; idx += PT_THREADS) a.post_psi[(size_t)k * d * d + idx] = w.b1[idx];
    if (a.mode == 2) {
        if (a.lik_mu) for (int i = tid; i < d; i += PT_THREADS) a.lik_mu[(size_t)k * d + i] = mp[i];
        if (a.lik_lmbda) for (int idx = tid; idx < d * d; idx += PT_THREADS)
            a.lik_lmbda[(size_t)k * d * d + idx] = (nu - d) * w.b1[idx];
    }

    if (a.vlb) {
        // entropy - cross-entropy = logZ(q) - logZ(p) - <eta(q) - eta(p), E_q[t]>   (composite.py:120-134)
        const double kap0 = a.kappa0[k], nu0 = a.nu0[k];
        const double* m0 = a.m0 + (size_t)k * d;
        double part = 0.0;
        for (int i = tid; i < d; i += PT_THREADS) {
            double e0 = 0.0;                            // (Psi' m')_i
            for (int j = 0; j < d; ++j) e0 = fma(w.b1[i * d + j], mp[j], e0);
            e0 *= nu;
            part += (kap * mp[i] - kap0 * m0[i]) * e0 + (kap - kap0) * (-0.5 * mp[i] * e0);
        }
        for (int idx = tid; idx < d * d; idx += PT_THREADS) {
            int i = idx / d, j = idx - i * d;
            double d2 = Pu[idx] - Pk[idx] + st[tri_idx(a.stat_idx[i], a.stat_idx[j])];
            part += d2 * (-0.5 * nu * w.b1[idx]);
        }
        double tot = block_sum<double>(part, red);
        if (tid == 0) {
            tot += (kap - kap0) * (-0.5 * d / kap) + (nu - nu0) * 0.5 * elogdet;
            double lzq = -0.5 * d * log(kap) + wishart_logz(nu, d, logdet_psi);
            double lzp = -0.5 * d * log(kap0) + wishart_logz(nu0, d, logdet_psi0);
            a.vlb[k] = lzq - lzp - tot;
        }
        __syncthreads();
    }

    if (a.mode == 1) {
        // Gibbs: Lambda = T T^T, T = chol(Psi') A (Bartlett, wishart.py:72-92);
        // mu = m' + T^-T z / sqrt(kappa')  (== gaussian.py:311-313 with U = sqrt(kappa') T^T)
        if (!cta_chol_lower(w.b1, d, nullptr)) { flag_fail(a.info, MIMO_ENOTPD, k); return; }
        const int nt = d * (d - 1) / 2;
        const double* var = a.variates + (size_t)k * (nt + 2 * d);
        // b2 <- A (lower): strict lower = normals in tril order, diagonal = sqrt(chi-square)
        for (int idx = tid; idx < d * d; idx += PT_THREADS) {
            int i = idx / d, j = idx - i * d;
            w.b2[idx] = (j < i) ? var[i * (i - 1) / 2 + j] : (j == i ? sqrt(var[nt + i]) : 0.0);
        }
        __syncthreads();
        cta_trmm_lower_inplace(w.b1, w.b2, d);           // b2 <- C * A = T (lower)
        double* mu = sh_vec;
        for (int i = tid; i < d; i += PT_THREADS) mu[i] = var[nt + d + i];
        cta_solve_lower_T(w.b2, d, mu, 1, 1);            // mu <- T^-T z
        const double rs = 1.0 / sqrt(kap);
        for (int i = tid; i < d; i += PT_THREADS) mu[i] = mp[i] + rs * mu[i];
        __syncthreads();
        if (a.lik_mu) for (int i = tid; i < d; i += PT_THREADS) a.lik_mu[(size_t)k * d + i] = mu[i];
        if (a.lik_lmbda) {
            for (int idx = tid; idx < d * d; idx += PT_THREADS) {
                int i = idx / d, j = idx - i * d;
                int mm = min(i, j);
                double s = 0.0;
                for (int m = 0; m <= mm; ++m) s = fma(w.b2[i * d + m], w.b2[j * d + m], s);
                a.lik_lmbda[(size_t)k * d * d + idx] = s;
            }
        }
        if (a.W) {
            // rows of U = T^T:  W[i][j] = T[j][i];  offset = -(T^T mu)_i;  cst += sum log T_ii - d/2 log 2pi
            for (int idx = tid; idx < d * d; idx += PT_THREADS) {
                int i = idx / d, j = idx - i * d;
                store_op(a.W, a.op_dtype, wbase + (size_t)i * a.Dpp + a.col_map[j], w.b2[j * d + i]);
            }
            for (int i = tid; i < d; i += PT_THREADS) {
                double o = 0.0;
                for (int j = i; j < d; ++j) o = fma(w.b2[j * d + i], mu[j], o);
                store_op(a.W, a.op_dtype, wbase + (size_t)i * a.Dpp + cD, -o);
            }
            if (tid == 0) {
                double c = -0.5 * d * LOG_2PI;
                for (int i = 0; i < d; ++i) c += log(w.b2[i * d + i]);
                add_op(a.cst, a.op_dtype, k, c);
            }
        }
    }
}

static size_t a256(size_t x) { return (x + 255) / 256 * 256; }

size_t nw_workspace(int K, int d) {
    size_t dd = (size_t)d * d;
    return a256(8 * ((size_t)K * dd)) + a256(8 * dd) + a256(8 * (size_t)K * 8) + a256(8 * (size_t)K * d)
         + 256 + a256(8 * (size_t)K * 2 * dd);
}

static bool mats_fit_smem(int n) { return (size_t)2 * n * n * 8 <= 160 * 1024; }

int nw_posterior(int K, int d, int tied, int mode,
                 const double* m0, const double* kappa0, const double* psi0, const double* nu0,
                 const double* stat, int F, const int32_t* stat_idx, int Dp, const double* variates,
                 double* post_m, double* post_kappa, double* post_psi, double* post_nu,
                 double* lik_mu, double* lik_lmbda, double* vlb,
                 int op_dtype, void* W, void* cst, int Rp, int Dpp, int row_off, const int32_t* col_map,
                 void* workspace, size_t workspace_bytes, int32_t* info, cudaStream_t st) {
    MIMO_CHECK_ARG(K >= 1 && d >= 1 && d <= 256, "shape (d <= 256)");
    MIMO_CHECK_ARG(m0 && kappa0 && psi0 && nu0 && stat && stat_idx && info && workspace, "null pointer");
    MIMO_CHECK_ARG(mode >= 0 && mode <= 3, "mode");
    MIMO_CHECK_ARG(mode != 1 || variates, "Gibbs mode needs variates");
    MIMO_CHECK_ARG(F >= Dp * (Dp + 1) / 2 && Dp >= d + 1, "statistics layout");
    MIMO_CHECK_ARG(!W || (cst && col_map && row_off + d <= Rp && Dpp >= Dp), "operand placement");
    MIMO_CHECK_ARG(workspace_bytes >= nw_workspace(K, d), "workspace too small");
    size_t dd = (size_t)d * d;
    char* ws = (char*)workspace;
    NWArgs a;
    a.K = K; a.d = d; a.tied = tied; a.mode = mode;
    a.m0 = m0; a.kappa0 = kappa0; a.psi0 = psi0; a.nu0 = nu0;
    a.stat = stat; a.F = F; a.stat_idx = stat_idx; a.variates = variates;
    a.post_m = post_m; a.post_kappa = post_kappa; a.post_psi = post_psi; a.post_nu = post_nu;
    a.lik_mu = lik_mu; a.lik_lmbda = lik_lmbda; a.vlb = vlb;
    a.op_dtype = op_dtype; a.W = (mode == 3) ? nullptr : W; a.cst = cst; a.Rp = Rp; a.Dpp = Dpp;
    a.row_off = row_off; a.col_map = col_map;
    a.Pk = (double*)ws; ws += a256(8 * K * dd);
    a.Pbar = (double*)ws; ws += a256(8 * dd);
    a.scal = (double*)ws; ws += a256(8 * (size_t)K * 8);
    a.mpost = (double*)ws; ws += a256(8 * (size_t)K * d);
    a.nubar = (double*)ws; ws += 256;
    a.gscratch = (double*)ws;
    a.use_smem = mats_fit_smem(d);
    a.info = info;
    size_t smem = a.use_smem ? 2 * dd * 8 : 0;
    if (smem > 48 * 1024) {
        MIMO_CUDA(cudaFuncSetAttribute(nw_phase_a, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MIMO_CUDA(cudaFuncSetAttribute(nw_phase_b, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    nw_phase_a<<<K, PT_THREADS, smem, st>>>(a);
    if (tied) {
        mean_over_k<<<cdiv(dd, 256), 256, 0, st>>>(a.Pk, K, (int)dd, (int)dd, a.Pbar);
        mean_over_k<<<1, 32, 0, st>>>(a.scal + 2, K, 1, 8, a.nubar);
    }
    nw_phase_b<<<K, PT_THREADS, smem, st>>>(a);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

// ---------------------------------------------------------------------------------
// Normal-Gamma (diagonal family): one CTA per component, one thread per dimension
// ---------------------------------------------------------------------------------
struct NGArgs {
    int K, d, tied, mode, bug_compat;
    const double *m0, *kappa0, *alpha0, *beta0;
    const double* stat; int F; const double* variates;
    double *post_m, *post_kappa, *post_alpha, *post_beta, *lik_mu, *lik_l, *vlb;
    int op_dtype; void *S, *T, *cst;
    double *ab;        // (K, 2, d) per-component alpha', beta'   (workspace)
    double *abbar;     // (2, d) tied means
};

// stat layout of the diag family: [sum r x_j (d) | sum r x_j^2 (d) | sum r]
__global__ void ng_phase_a(NGArgs a) {
    const int k = blockIdx.x, d = a.d;
    const double* st = a.stat + (size_t)k * a.F;
    const double n = st[2 * d];
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        size_t o = (size_t)k * d + j;
        double kap0 = a.kappa0[o], kap = kap0 + n;
        double m = (kap0 * a.m0[o] + st[j]) / kap;
        double al = a.alpha0[o] + 0.5 * n;
        double be = a.beta0[o] + 0.5 * (st[d + j] + kap0 * a.m0[o] * a.m0[o] - kap * m * m);
        a.ab[((size_t)k * 2) * d + j] = al;
        a.ab[((size_t)k * 2 + 1) * d + j] = be;
    }
}

__global__ void __launch_bounds__(PT_THREADS) ng_phase_b(NGArgs a) {
    __shared__ double red[32];
    const int k = blockIdx.x, d = a.d, tid = threadIdx.x;
    const double* st = a.stat + (size_t)k * a.F;
    const double n = st[2 * d];
    double c_part = 0.0, v_part = 0.0;
    for (int j = tid; j < d; j += PT_THREADS) {
        size_t o = (size_t)k * d + j;
        double kap0 = a.kappa0[o], kap = kap0 + n, m0 = a.m0[o];
        double m = (kap0 * m0 + st[j]) / kap;
        double al = a.tied ? a.abbar[j] : a.ab[((size_t)k * 2) * d + j];
        double be = a.tied ? a.abbar[d + j] : a.ab[((size_t)k * 2 + 1) * d + j];
        if (a.bug_compat) { al = a.alpha0[o]; be = a.beta0[o]; }   // composite.py:472-484
        if (a.post_m) a.post_m[o] = m;
        if (a.post_kappa) a.post_kappa[o] = kap;
        if (a.post_alpha) a.post_alpha[o] = al;
        if (a.post_beta) a.post_beta[o] = be;
        double lam = 0.0, mu = m, cj = 0.0;
        if (a.mode == 0) {            // expected operands  (composite.py:371-382)
            lam = al / be;
            cj = -0.5 / kap + 0.5 * (digamma_d(al) - log(be));
        } else if (a.mode == 1) {     // Gibbs  (composite.py:347-351, gaussian.py:644-646)
            lam = a.variates[(size_t)k * 2 * d + j];
            mu = m + a.variates[(size_t)k * 2 * d + d + j] / sqrt(kap * lam);
            cj = 0.5 * log(lam);
        } else if (a.mode == 2) {     // MAP  (composite.py:342-345)
            lam = (al - 0.5) / be;
            cj = 0.5 * log(lam);
        }
        if (a.mode != 3) {
            if (a.S) {
                double s = sqrt(lam);
                store_op(a.S, a.op_dtype, o, s);
                store_op(a.T, a.op_dtype, o, s * mu);
            }
            c_part += cj - 0.5 * LOG_2PI;
            if (a.mode != 0) {
                if (a.lik_mu) a.lik_mu[o] = mu;
                if (a.lik_l) a.lik_l[o] = lam;
            }
        }
        if (a.vlb) {
            double al0 = a.alpha0[o], be0 = a.beta0[o];
            double e0 = al / be * m, e1 = -0.5 * (1.0 / kap + m * e0);
            double e2 = 0.5 * (digamma_d(al) - log(be)), e3 = -0.5 * al / be;
            double dot = (kap * m - kap0 * m0) * e0 + (kap - kap0) * e1
                       + (2.0 * al - 2.0 * al0) * e2 + (2.0 * be + kap * m * m - 2.0 * be0 - kap0 * m0 * m0) * e3;
            double lzq = -0.5 * log(kap) + lgamma(al) - al * log(be);
            double lzp = -0.5 * log(kap0) + lgamma(al0) - al0 * log(be0);
            v_part += lzq - lzp - dot;
        }
    }
    double c = block_sum<double>(c_part, red);
    if (tid == 0 && a.mode != 3 && a.cst) add_op(a.cst, a.op_dtype, k, c);
    if (a.vlb) {
        double v = block_sum<double>(v_part, red);
        if (tid == 0) a.vlb[k] = v;
    }
}

size_t ng_workspace(int K, int d) { return a256(8 * (size_t)K * 2 * d) + a256(8 * (size_t)2 * d); }

int ng_posterior(int K, int d, int tied, int mode, int bug_compat,
                 const double* m0, const double* kappa0, const double* alpha0, const double* beta0,
                 const double* stat, int F, const double* variates,
                 double* post_m, double* post_kappa, double* post_alpha, double* post_beta,
                 double* lik_mu, double* lik_l, double* vlb,
                 int op_dtype, void* S, void* T, void* cst, void* workspace, size_t workspace_bytes,
                 int32_t* info, cudaStream_t st) {
    (void)info;
    MIMO_CHECK_ARG(K >= 1 && d >= 1, "shape");
    MIMO_CHECK_ARG(m0 && kappa0 && alpha0 && beta0 && stat && workspace, "null pointer");
    MIMO_CHECK_ARG(F >= 2 * d + 1, "statistics layout");
    MIMO_CHECK_ARG(mode >= 0 && mode <= 3 && (mode != 1 || variates), "mode / variates");
    MIMO_CHECK_ARG(workspace_bytes >= ng_workspace(K, d), "workspace too small");
    NGArgs a;
    a.K = K; a.d = d; a.tied = tied; a.mode = mode; a.bug_compat = bug_compat;
    a.m0 = m0; a.kappa0 = kappa0; a.alpha0 = alpha0; a.beta0 = beta0;
    a.stat = stat; a.F = F; a.variates = variates;
    a.post_m = post_m; a.post_kappa = post_kappa; a.post_alpha = post_alpha; a.post_beta = post_beta;
    a.lik_mu = lik_mu; a.lik_l = lik_l; a.vlb = vlb;
    a.op_dtype = op_dtype; a.S = S; a.T = T; a.cst = cst;
    a.ab = (double*)workspace;
    a.abbar = (double*)((char*)workspace + a256(8 * (size_t)K * 2 * d));
    ng_phase_a<<<K, 128, 0, st>>>(a);
    if (tied) mean_over_k<<<cdiv(2 * d, 256), 256, 0, st>>>(a.ab, K, 2 * d, 2 * d, a.abbar);
    ng_phase_b<<<K, PT_THREADS, 0, st>>>(a);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

// ---------------------------------------------------------------------------------
// Matrix-Normal-Wishart (linear-Gaussian experts).  c = column_dim, o = row_dim.
// Work per component is small (c, o are input/output dims), so plain CTA loops.
// ---------------------------------------------------------------------------------
struct MNWArgs {
    int K, c, o, tied, mode;
    const double *M0, *K0, *psi0, *nu0;
    const double* stat; int F; const int32_t* stat_idx;
    const double* variates;
    double *post_M, *post_K, *post_psi, *post_nu, *lik_A, *lik_lmbda, *vlb;
    int op_dtype; void* W; void* cst; int Rp, Dpp, row_off; const int32_t* col_map;
    // workspace
    double *Pk, *Pbar, *scal, *Mp, *Kp, *Ginv, *nubar, *gscratch;
    int32_t* info;
};

// scratch per CTA (global): n = max(c, o);  b1, b2 : n*n each
__global__ void __launch_bounds__(PT_THREADS) mnw_phase_a(MNWArgs a) {
    const int k = blockIdx.x, c = a.c, o = a.o, tid = threadIdx.x;
    const int n = max(c, o);
    double* b1 = a.gscratch + (size_t)k * 2 * n * n;
    double* b2 = b1 + n * n;
    const double* st = a.stat + (size_t)k * a.F;
    const int32_t* xi = a.stat_idx;            // positions of xt[0..c)
    const int32_t* yi = a.stat_idx + c;        // positions of y[0..o)
    const double* M0 = a.M0 + (size_t)k * o * c;
    const double* K0 = a.K0 + (size_t)k * c * c;
    double* Kp = a.Kp + (size_t)k * c * c;
    double* Mp = a.Mp + (size_t)k * o * c;
    double* Gi = a.Ginv + (size_t)k * c * c;
    double* sc = a.scal + (size_t)k * 8;
    // count: the constant feature is the largest index in zt = Dp - 1, addressed as (pc, pc);
    // for the affine model pc = xi[c-1]; keep it general: passed as stat_idx[c + o].
    const int pc = a.stat_idx[c + o];
    const double cnt = st[tri_idx(pc, pc)];

    // log det K0 (for the prior log-partition)
    for (int idx = tid; idx < c * c; idx += PT_THREADS) b1[idx] = K0[idx];
    double ldK0;
    if (!cta_chol_lower(b1, c, &ldK0)) { flag_fail(a.info, MIMO_ENOTPD, k); return; }
    // K' = K0 + Sxx, G = chol(K'), Ginv
    for (int idx = tid; idx < c * c; idx += PT_THREADS) {
        int i = idx / c, j = idx - i * c;
        double v = K0[idx] + st[tri_idx(xi[i], xi[j])];
        Kp[idx] = v; b1[idx] = v;
    }
    double ldK;
    if (!cta_chol_lower(b1, c, &ldK)) { flag_fail(a.info, MIMO_ENOTPD, k); return; }
    cta_tri_inv_lower(b1, Gi, c);
    cta_gram_lower(Gi, b1, c);                 // b1 = K'^-1
    // Nn = M0 K0 + Syx  (o x c) -> b2 ;  M' = Nn K'^-1
    for (int idx = tid; idx < o * c; idx += PT_THREADS) {
        int i = idx / c, j = idx - i * c;
        double s = st[tri_idx(yi[i], xi[j])];
        for (int m = 0; m < c; ++m) s = fma(M0[i * c + m], K0[m * c + j], s);
        b2[idx] = s;
    }
    __syncthreads();
    for (int idx = tid; idx < o * c; idx += PT_THREADS) {
        int i = idx / c, j = idx - i * c;
        double s = 0.0;
        for (int m = 0; m < c; ++m) s = fma(b2[i * c + m], b1[m * c + j], s);
        Mp[idx] = s;
    }
    __syncthreads();
    // P_k = Psi0^-1 + M0 K0 M0^T + Syy - Nn M'^T     (o x o)
    double* P = a.Pk + (size_t)k * o * o;
    for (int idx = tid; idx < o * o; idx += PT_THREADS) {
        int i = idx / o, j = idx - i * o;
        double s = st[tri_idx(yi[i], yi[j])];
        for (int m = 0; m < c; ++m) {
            double mk = 0.0;                   // (M0 K0)[i][m]
            for (int q = 0; q < c; ++q) mk = fma(M0[i * c + q], K0[q * c + m], mk);
            s = fma(mk, M0[j * c + m], s);
            s = fma(-b2[i * c + m], Mp[j * c + m], s);
        }
        P[idx] = s;
    }
    __syncthreads();
    // + Psi0^-1
    const double* psi0 = a.psi0 + (size_t)k * o * o;
    for (int idx = tid; idx < o * o; idx += PT_THREADS) b1[idx] = psi0[idx];
    double ld0;
    if (!cta_chol_lower(b1, o, &ld0)) { flag_fail(a.info, MIMO_ENOTPD, k); return; }
    cta_tri_inv_lower(b1, b2, o);
    cta_gram_lower(b2, b1, o);
    for (int idx = tid; idx < o * o; idx += PT_THREADS) P[idx] += b1[idx];
    if (tid == 0) { sc[0] = cnt; sc[1] = 2.0 * ldK; sc[2] = a.nu0[k] + cnt; sc[3] = 2.0 * ld0; sc[4] = 2.0 * ldK0; }
}

__global__ void __launch_bounds__(PT_THREADS) mnw_phase_b(MNWArgs a) {
    __shared__ double red[32];
    const int k = blockIdx.x, c = a.c, o = a.o, tid = threadIdx.x;
    if (a.info[0] != 0) return;
    const int n = max(c, o);
    double* b1 = a.gscratch + (size_t)k * 2 * n * n;
    double* b2 = b1 + n * n;
    const double* st = a.stat + (size_t)k * a.F;
    const int32_t* xi = a.stat_idx;
    const int32_t* yi = a.stat_idx + c;
    const double* Pk = a.Pk + (size_t)k * o * o;
    const double* Pu = a.tied ? a.Pbar : Pk;
    const double* sc = a.scal + (size_t)k * 8;
    const double logdetK = sc[1], logdet_psi0 = sc[3], logdetK0 = sc[4];
    const double nu = a.tied ? a.nubar[0] : sc[2];
    const double* Mp = a.Mp + (size_t)k * o * c;
    const double* Kp = a.Kp + (size_t)k * c * c;
    const double* Gi = a.Ginv + (size_t)k * c * c;

    for (int idx = tid; idx < o * o; idx += PT_THREADS) b1[idx] = Pu[idx];
    double ldP;
    if (!cta_chol_lower(b1, o, &ldP)) { flag_fail(a.info, MIMO_ENOTPD, k); return; }
    const double logdet_psi = -2.0 * ldP;
    cta_tri_inv_lower(b1, b2, o);              // b2 = L^-1
    const double elogdet = wishart_elogdet(nu, o, logdet_psi);

    if (a.post_M) for (int idx = tid; idx < o * c; idx += PT_THREADS) a.post_M[(size_t)k * o * c + idx] = Mp[idx];
    if (a.post_K) for (int idx = tid; idx < c * c; idx += PT_THREADS) a.post_K[(size_t)k * c * c + idx] = Kp[idx];
    if (tid == 0 && a.post_nu) a.post_nu[k] = nu;

    const size_t wbase = ((size_t)k * a.Rp + a.row_off) * a.Dpp;
    if ((a.mode == 0 || a.mode == 2) && a.W) {
        const double s2 = (a.mode == 0) ? nu : (nu - o);
        const double s = sqrt(s2);
        // rows 0..o):  s L^-1 [ -M' | I ]
        for (int idx = tid; idx < o * c; idx += PT_THREADS) {
            int i = idx / c, j = idx - i * c;
            double v = 0.0;
            for (int m = 0; m <= i; ++m) v = fma(b2[i * o + m], Mp[m * c + j], v);
            store_op(a.W, a.op_dtype, wbase + (size_t)i * a.Dpp + a.col_map[j], -s * v);
        }
        for (int idx = tid; idx < o * o; idx += PT_THREADS) {
            int i = idx / o, m = idx - i * o;
            store_op(a.W, a.op_dtype, wbase + (size_t)i * a.Dpp + a.col_map[c + m], s * b2[idx]);
        }
        if (a.mode == 0) {
            // rows o..o+c):  sqrt(o) G^-1 on the xt columns   (the  -(o/2) xt^T K'^-1 xt  term)
            const double so = sqrt((double)o);
            for (int idx = tid; idx < c * c; idx += PT_THREADS) {
                int i = idx / c, j = idx - i * c;
                store_op(a.W, a.op_dtype, wbase + (size_t)(o + i) * a.Dpp + a.col_map[j], so * Gi[idx]);
            }
        }
        if (tid == 0) {
            double cc = (a.mode == 0) ? (0.5 * elogdet - 0.5 * o * LOG_2PI)
                                      : (0.5 * (o * log(s2) + logdet_psi) - 0.5 * o * LOG_2PI);
            add_op(a.cst, a.op_dtype, k, cc);
        }
    }
    const bool need_psi = a.post_psi || a.vlb || a.mode == 1 || (a.mode == 2 && a.lik_lmbda);
    if (need_psi) cta_gram_lower(b2, b1, o);   // b1 = Psi'
    if (a.post_psi) for (int idx = tid; idx < o * o; idx += PT_THREADS) a.post_psi[(size_t)k * o * o + idx] = b1[idx];
    if (a.mode == 2) {
        if (a.lik_A) for (int idx = tid; idx < o * c; idx += PT_THREADS) a.lik_A[(size_t)k * o * c + idx] = Mp[idx];
        if (a.lik_lmbda) for (int idx = tid; idx < o * o; idx += PT_THREADS)
            a.lik_lmbda[(size_t)k * o * o + idx] = (nu - o) * b1[idx];
    }

    if (a.vlb) {
        const double* M0 = a.M0 + (size_t)k * o * c;
        (void)M0;
        double part = 0.0;
        // <Syx, nu Psi' M'>
        for (int idx = tid; idx < o * c; idx += PT_THREADS) {
            int i = idx / c, j = idx - i * c;
            double e0 = 0.0;
            for (int m = 0; m < o; ++m) e0 = fma(b1[i * o + m], Mp[m * c + j], e0);
            part += st[tri_idx(yi[i], xi[j])] * nu * e0;
        }
        // <Sxx, -0.5 (o K'^-1 + nu M'^T Psi' M')>
        for (int idx = tid; idx < c * c; idx += PT_THREADS) {
            int i = idx / c, j = idx - i * c;
            double kinv = 0.0;
            for (int m = max(i, j); m < c; ++m) kinv = fma(Gi[m * c + i], Gi[m * c + j], kinv);
            double q = 0.0;
            for (int p = 0; p < o; ++p) {
                double t = 0.0;
                for (int m = 0; m < o; ++m) t = fma(b1[p * o + m], Mp[m * c + j], t);
                q = fma(Mp[p * c + i], t, q);
            }
            part += st[tri_idx(xi[i], xi[j])] * (-0.5) * (o * kinv + nu * q);
        }
        // <P_used - P_k + Syy, -0.5 nu Psi'>
        for (int idx = tid; idx < o * o; idx += PT_THREADS) {
            int i = idx / o, j = idx - i * o;
            part += (Pu[idx] - Pk[idx] + st[tri_idx(yi[i], yi[j])]) * (-0.5 * nu * b1[idx]);
        }
        double tot = block_sum<double>(part, red);
        if (tid == 0) {
            const double nu0 = a.nu0[k];
            tot += (nu - nu0) * 0.5 * elogdet;
            double lzq = -0.5 * o * logdetK + wishart_logz(nu, o, logdet_psi);
            double lzp = -0.5 * o * logdetK0 + wishart_logz(nu0, o, logdet_psi0);
            a.vlb[k] = lzq - lzp - tot;
        }
        __syncthreads();
    }

    if (a.mode == 1) {
        // Lambda = T T^T (Bartlett);  A = M' + T^-T Zv G^-1,  Zv = unvec_F(z)   (matrix.py:98-125
        // with chol_upper(kron(K', Lambda)) = kron(G^T, T^T))
        if (!cta_chol_lower(b1, o, nullptr)) { flag_fail(a.info, MIMO_ENOTPD, k); return; }
        const int nt = o * (o - 1) / 2;
        const double* var = a.variates + (size_t)k * (nt + o + o * c);
        for (int idx = tid; idx < o * o; idx += PT_THREADS) {
            int i = idx / o, j = idx - i * o;
            b2[idx] = (j < i) ? var[i * (i - 1) / 2 + j] : (j == i ? sqrt(var[nt + i]) : 0.0);
        }
        __syncthreads();
        double* Tm = a.gscratch + (size_t)a.K * 2 * n * n + (size_t)k * (o * o + o * c);   // T then V
        cta_trmm_lower(b1, b2, Tm, o);
        double* V = Tm + o * o;                // (o x c):  Zv G^-1
        const double* z = var + nt + o;
        for (int idx = tid; idx < o * c; idx += PT_THREADS) {
            int i = idx / c, j = idx - i * c;
            double s = 0.0;
            for (int m = j; m < c; ++m) s = fma(z[m * o + i], Gi[m * c + j], s);
            V[idx] = s;
        }
        cta_solve_lower_T(Tm, o, V, c, c);
        for (int idx = tid; idx < o * c; idx += PT_THREADS) V[idx] += Mp[idx];
        __syncthreads();
        if (a.lik_A) for (int idx = tid; idx < o * c; idx += PT_THREADS) a.lik_A[(size_t)k * o * c + idx] = V[idx];
        if (a.lik_lmbda) for (int idx = tid; idx < o * o; idx += PT_THREADS) {
            int i = idx / o, j = idx - i * o;
            double s = 0.0;
            for (int m = 0; m <= min(i, j); ++m) s = fma(Tm[i * o + m], Tm[j * o + m], s);
            a.lik_lmbda[(size_t)k * o * o + idx] = s;
        }
        if (a.W) {
            // rows of U = T^T:  W[i][xt_j] = -(U A)_ij ,  W[i][y_m] = U_im = T[m][i]
            for (int idx = tid; idx < o * c; idx += PT_THREADS) {
                int i = idx / c, j = idx - i * c;
                double v = 0.0;
                for (int m = i; m < o; ++m) v = fma(Tm[m * o + i], V[m * c + j], v);
                store_op(a.W, a.op_dtype, wbase + (size_t)i * a.Dpp + a.col_map[j], -v);
            }
            for (int idx = tid; idx < o * o; idx += PT_THREADS) {
                int i = idx / o, m = idx - i * o;
                store_op(a.W, a.op_dtype, wbase + (size_t)i * a.Dpp + a.col_map[c + m], Tm[m * o + i]);
            }
            if (tid == 0) {
                double cc = -0.5 * o * LOG_2PI;
                for (int i = 0; i < o; ++i) cc += log(Tm[i * o + i]);
                add_op(a.cst, a.op_dtype, k, cc);
            }
        }
    }
}

size_t mnw_workspace(int K, int c, int o) {
    size_t n = (size_t)std::max(c, o);
    return a256(8 * (size_t)K * o * o) + a256(8 * (size_t)o * o) + a256(8 * (size_t)K * 8)
         + a256(8 * (size_t)K * o * c) + 2 * a256(8 * (size_t)K * c * c) + 256
         + a256(8 * ((size_t)K * 2 * n * n + (size_t)K * (o * o + o * c)));
}

int mnw_posterior(int K, int c, int o, int tied, int mode,
                  const double* M0, const double* K0, const double* psi0, const double* nu0,
                  const double* stat, int F, const int32_t* stat_idx, int Dp, const double* variates,
                  double* post_M, double* post_K, double* post_psi, double* post_nu,
                  double* lik_A, double* lik_lmbda, double* vlb,
                  int op_dtype, void* W, void* cst, int Rp, int Dpp, int row_off, const int32_t* col_map,
                  void* workspace, size_t workspace_bytes, int32_t* info, cudaStream_t st) {
    MIMO_CHECK_ARG(K >= 1 && c >= 1 && o >= 1 && c <= 512 && o <= 512, "shape");
    MIMO_CHECK_ARG(M0 && K0 && psi0 && nu0 && stat && stat_idx && info && workspace, "null pointer");
    MIMO_CHECK_ARG(mode >= 0 && mode <= 3 && (mode != 1 || variates), "mode / variates");
    MIMO_CHECK_ARG(F >= Dp * (Dp + 1) / 2, "statistics layout");
    int rows = (mode == 0) ? o + c : o;
    MIMO_CHECK_ARG(!W || mode == 3 || (cst && col_map && row_off + rows <= Rp), "operand placement");
    MIMO_CHECK_ARG(workspace_bytes >= mnw_workspace(K, c, o), "workspace too small");
    char* ws = (char*)workspace;
    MNWArgs a;
    a.K = K; a.c = c; a.o = o; a.tied = tied; a.mode = mode;
    a.M0 = M0; a.K0 = K0; a.psi0 = psi0; a.nu0 = nu0;
    a.stat = stat; a.F = F; a.stat_idx = stat_idx; a.variates = variates;
    a.post_M = post_M; a.post_K = post_K; a.post_psi = post_psi; a.post_nu = post_nu;
    a.lik_A = lik_A; a.lik_lmbda = lik_lmbda; a.vlb = vlb;
    a.op_dtype = op_dtype; a.W = (mode == 3) ? nullptr : W; a.cst = cst; a.Rp = Rp; a.Dpp = Dpp;
    a.row_off = row_off; a.col_map = col_map;
    a.Pk = (double*)ws; ws += a256(8 * (size_t)K * o * o);
    a.Pbar = (double*)ws; ws += a256(8 * (size_t)o * o);
    a.scal = (double*)ws; ws += a256(8 * (size_t)K * 8);
    a.Mp = (double*)ws; ws += a256(8 * (size_t)K * o * c);
    a.Kp = (double*)ws; ws += a256(8 * (size_t)K * c * c);
    a.Ginv = (double*)ws; ws += a256(8 * (size_t)K * c * c);
    a.nubar = (double*)ws; ws += 256;
    a.gscratch = (double*)ws;
    a.info = info;
    mnw_phase_a<<<K, PT_THREADS, 0, st>>>(a);
    if (tied) {
        mean_over_k<<<cdiv(o * o, 256), 256, 0, st>>>(a.Pk, K, o * o, o * o, a.Pbar);
        mean_over_k<<<1, 32, 0, st>>>(a.scal + 2, K, 1, 8, a.nubar);
    }
    mnw_phase_b<<<K, PT_THREADS, 0, st>>>(a);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

// ---------------------------------------------------------------------------------
// Gating: Dirichlet / truncated stick-breaking.  One CTA; K is small (<= a few thousand).
// ---------------------------------------------------------------------------------
__global__ void gating_kernel(int K, int kind, int mode, const double* prior_a, const double* prior_b,
                              const double* stat, int F, int count_feature, const double* variates,
                              double* post_a, double* post_b, double* probs, double* vlb,
                              int op_dtype, void* cst, double* tmp, int32_t* info) {
    // tmp: (8, K) scratch.  The transcendental work (digamma / lgamma per component) is spread over
    // the threads; thread 0 only does the K-long prefix sums.
    __shared__ double red[32];
    double* pa = tmp; double* pb = tmp + K; double* lw = tmp + 2 * K; double* aux = tmp + 3 * K;
    double* e_a = tmp + 4 * K; double* e_b = tmp + 5 * K; double* lz = tmp + 6 * K;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int k = tid; k < K; k += nt) pa[k] = prior_a[k] + stat[(size_t)k * F + count_feature];
    if (kind == 1 && tid == 0) {              // tail counts (bayesian.py:143)
        double run = 0.0;
        for (int k = K - 1; k >= 0; --k) { pb[k] = prior_b[k] + run; run += stat[(size_t)k * F + count_feature]; }
    }
    __syncthreads();
    double s_part = 0.0, s0_part = 0.0;
    for (int k = tid; k < K; k += nt) { s_part += pa[k]; s0_part += prior_a[k]; }
    double sa = block_sum<double>(s_part, red);
    __shared__ double sh_sa, sh_sa0;
    if (tid == 0) sh_sa = sa;
    double sa0 = block_sum<double>(s0_part, red);
    if (tid == 0) sh_sa0 = sa0;
    __syncthreads();
    sa = sh_sa; sa0 = sh_sa0;
    const bool need_dig = (mode == 0) || (vlb != nullptr);
    double v_part = 0.0;
    if (kind == 0) {
        const double ds = need_dig ? digamma_d(sa) : 0.0;
        for (int k = tid; k < K; k += nt) {
            double dg = need_dig ? digamma_d(pa[k]) - ds : 0.0;
            if (mode == 0) lw[k] = dg;
            else if (mode == 2) {
                if (!(pa[k] > 1.0)) { if (atomicCAS(&info[0], 0, MIMO_EINVAL) == 0) info[1] = k; }
                lw[k] = log((pa[k] - 1.0) / (sa - K));
            } else if (mode == 4) lw[k] = log(pa[k] / sa);
            if (vlb) v_part += lgamma(pa[k]) - lgamma(prior_a[k]) - (pa[k] - prior_a[k]) * dg;
        }
        if (mode == 1) {
            double g_part = 0.0;
            for (int k = tid; k < K; k += nt) g_part += variates[k];
            double sg = block_sum<double>(g_part, red);
            __shared__ double sh_sg;
            if (tid == 0) sh_sg = sg;
            __syncthreads();
            for (int k = tid; k < K; k += nt) lw[k] = log(fmax(variates[k] / sh_sg, 2.220446049250313e-16));
        }
        if (vlb) {   // bayesian.py:93-96, dirichlet.py:78-97
            double v = block_sum<double>(v_part, red);
            if (tid == 0) vlb[0] = v - lgamma(sa) + lgamma(sa0);
        }
    } else {
        for (int k = tid; k < K; k += nt) {
            if (need_dig) {
                double dsum = digamma_d(pa[k] + pb[k]);
                e_a[k] = digamma_d(pa[k]) - dsum;       // E log v_k
                e_b[k] = digamma_d(pb[k]) - dsum;       // E log (1 - v_k)
            }
            if (vlb) {
                lz[k] = lgamma(pa[k]) + lgamma(pb[k]) - lgamma(pa[k] + pb[k])
                      - (lgamma(prior_a[k]) + lgamma(prior_b[k]) - lgamma(prior_a[k] + prior_b[k]));
                v_part += lz[k] - (pa[k] - prior_a[k]) * e_a[k] - (pb[k] - prior_b[k]) * e_b[k];
            }
            if (mode != 0) {
                double b;
                if (k == K - 1) b = 1.0;
                else if (mode == 1) b = variates[k];
                else if (mode == 4) b = pa[k] / (pa[k] + pb[k]);
                else {                    // mode of a Beta(g, d)   (dirichlet.py:152-170)
                    double g = pa[k], dd = pb[k];
                    if (g > 1.0 && dd > 1.0) b = (g - 1.0) / (g + dd - 2.0);
                    else if (g == 1.0 && dd == 1.0) b = 1.0;
                    else if (g < 1.0 && dd < 1.0) b = 1.0;
                    else if (g <= 1.0 && dd > 1.0) b = 0.0;
                    else if (g > 1.0 && dd <= 1.0) b = 1.0;
                    else { b = 1.0; if (atomicCAS(&info[0], 0, MIMO_EINVAL) == 0) info[1] = k; }
                }
                aux[k] = b;
            }
        }
        if (vlb) {   // bayesian.py:173-176, dirichlet.py:195-214
            double v = block_sum<double>(v_part, red);
            if (tid == 0) vlb[0] = v;
        }
        __syncthreads();
        if (tid == 0) {
            if (mode == 0) {
                double run = 0.0;         // prefix sum of E log(1 - v_j), gmm.py:250-252
                for (int k = 0; k < K; ++k) { lw[k] = e_a[k] + run; run += e_b[k]; }
            } else {
                double rest = 1.0;        // running prod (1 - v_j), dirichlet.py:181-184
                for (int k = 0; k < K; ++k) { double b = aux[k]; lw[k] = log(b * rest); rest *= (1.0 - b); }
            }
        }
    }
    __syncthreads();
    for (int k = tid; k < K; k += nt) {
        if (post_a) post_a[k] = pa[k];
        if (post_b && kind == 1) post_b[k] = pb[k];
        if (probs) probs[k] = exp(lw[k]);
        if (cst) store_op(cst, op_dtype, k, lw[k]);
    }
}

size_t gating_workspace(int K) { return a256(8 * (size_t)8 * K); }

int gating_posterior(int K, int kind, int mode, const double* prior_a, const double* prior_b,
                     const double* stat, int F, int count_feature, const double* variates,
                     double* post_a, double* post_b, double* probs, double* vlb,
                     int op_dtype, void* cst, void* workspace, size_t workspace_bytes,
                     int32_t* info, cudaStream_t st) {
    MIMO_CHECK_ARG(K >= 1 && (kind == 0 || kind == 1), "shape / kind");
    MIMO_CHECK_ARG(prior_a && stat && info && workspace && (kind == 0 || prior_b), "null pointer");
    MIMO_CHECK_ARG(mode == 0 || mode == 1 || mode == 2 || mode == 4, "mode");
    MIMO_CHECK_ARG(mode != 1 || variates, "Gibbs mode needs variates");
    MIMO_CHECK_ARG(workspace_bytes >= gating_workspace(K), "workspace too small");
    gating_kernel<<<1, 256, 0, st>>>(K, kind, mode, prior_a, prior_b, stat, F, count_feature, variates,
                                     post_a, post_b, probs, vlb, op_dtype, cst, (double*)workspace, info);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

// ---------------------------------------------------------------------------------
// Likelihood parameters -> operands, and EM M-steps
// ---------------------------------------------------------------------------------
// Gaussian: lmbda = L L^T  =>  U = L^T is the reference's upper factor (gaussian.py:298);
// rows of W are rows of U, offset -U mu, cst += sum log L_ii - d/2 log 2pi.
// Linear-Gaussian (c > 0): W[i][xt_j] = -(U A)_ij, W[i][y_m] = U_im.
__global__ void __launch_bounds__(PT_THREADS)
operands_kernel(int K, int d, int c, const double* mu_or_A, const double* lmbda,
                int op_dtype, void* W, void* cst, int Rp, int Dpp, int row_off, const int32_t* col_map,
                double* gscratch, int32_t* info) {
    const int k = blockIdx.x, tid = threadIdx.x;
    double* L = gscratch + (size_t)k * d * d;
    for (int idx = tid; idx < d * d; idx += PT_THREADS) L[idx] = lmbda[(size_t)k * d * d + idx];
    double ld;
    if (!cta_chol_lower(L, d, &ld)) { flag_fail(info, MIMO_ENOTPD, k); return; }
    const size_t wbase = ((size_t)k * Rp + row_off) * Dpp;
    if (c == 0) {
        const double* mu = mu_or_A + (size_t)k * d;
        for (int idx = tid; idx < d * d; idx += PT_THREADS) {
            int i = idx / d, j = idx - i * d;
            store_op(W, op_dtype, wbase + (size_t)i * Dpp + col_map[j], j >= i ? L[j * d + i] : 0.0);
        }
        for (int i = tid; i < d; i += PT_THREADS) {
            double o = 0.0;
            for (int j = i; j < d; ++j) o = fma(L[j * d + i], mu[j], o);
            store_op(W, op_dtype, wbase + (size_t)i * Dpp + col_map[d], -o);
        }
    } else {
        const double* A = mu_or_A + (size_t)k * d * c;     // (o = d) x c
        for (int idx = tid; idx < d * c; idx += PT_THREADS) {
            int i = idx / c, j = idx - i * c;
            double v = 0.0;
            for (int m = i; m < d; ++m) v = fma(L[m * d + i], A[m * c + j], v);
            store_op(W, op_dtype, wbase + (size_t)i * Dpp + col_map[j], -v);
        }
        for (int idx = tid; idx < d * d; idx += PT_THREADS) {
            int i = idx / d, m = idx - i * d;
            store_op(W, op_dtype, wbase + (size_t)i * Dpp + col_map[c + m], m >= i ? L[m * d + i] : 0.0);
        }
    }
    if (tid == 0) add_op(cst, op_dtype, k, ld - 0.5 * d * LOG_2PI);
}

int operands_gauss(int K, int d, int c, const double* mu_or_A, const double* lmbda,
                   int op_dtype, void* W, void* cst, int Rp, int Dpp, int row_off, const int32_t* col_map,
                   void* workspace, size_t workspace_bytes, int32_t* info, cudaStream_t st) {
    MIMO_CHECK_ARG(K >= 1 && d >= 1 && c >= 0, "shape");
    MIMO_CHECK_ARG(mu_or_A && lmbda && W && cst && col_map && info && workspace, "null pointer");
    MIMO_CHECK_ARG(row_off + d <= Rp, "operand placement");
    MIMO_CHECK_ARG(workspace_bytes >= 8 * (size_t)K * d * d, "workspace too small");
    operands_kernel<<<K, PT_THREADS, 0, st>>>(K, d, c, mu_or_A, lmbda, op_dtype, W, cst, Rp, Dpp, row_off,
                                              col_map, (double*)workspace, info);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

__global__ void operands_diag_kernel(int K, int d, const double* mu, const double* lam,
                                     int op_dtype, void* S, void* T, void* cst) {
    __shared__ double red[32];
    const int k = blockIdx.x;
    double part = 0.0;
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        size_t o = (size_t)k * d + j;
        double s = sqrt(lam[o]);
        store_op(S, op_dtype, o, s);
        store_op(T, op_dtype, o, s * mu[o]);
        part += 0.5 * log(lam[o]) - 0.5 * LOG_2PI;
    }
    double c = block_sum<double>(part, red);
    if (threadIdx.x == 0) add_op(cst, op_dtype, k, c);
}

int operands_gauss_diag(int K, int d, const double* mu, const double* lam, int op_dtype,
                        void* S, void* T, void* cst, cudaStream_t st) {
    MIMO_CHECK_ARG(K >= 1 && d >= 1 && mu && lam && S && T && cst, "arguments");
    operands_diag_kernel<<<K, 128, 0, st>>>(K, d, mu, lam, op_dtype, S, T, cst);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

// EM M-step, full covariance (gaussian.py:525-542; tied :550-572): mu = Sx/n,
// Sigma = Sxx/n - mu mu^T (symmetrised, +1e-16 I), lmbda = Sigma^-1; a failed Cholesky
// stands in for the reference's eigvalsh assertion.
__global__ void __launch_bounds__(PT_THREADS)
mstep_gauss_sigma(int K, int d, int tied, const double* stat, int F, const int32_t* stat_idx,
                  double* mu, double* sig /* (K,d,d) */) {
    const int k = blockIdx.x, tid = threadIdx.x;
    const double* st = stat + (size_t)k * F;
    const int pc = stat_idx[d];
    const double n = st[tri_idx(pc, pc)];
    for (int i = tid; i < d; i += PT_THREADS) mu[(size_t)k * d + i] = st[tri_idx(pc, stat_idx[i])] / n;
    __syncthreads();
    for (int idx = tid; idx < d * d; idx += PT_THREADS) {
        int i = idx / d, j = idx - i * d;
        double sxx = st[tri_idx(stat_idx[i], stat_idx[j])];
        double mm = mu[(size_t)k * d + i] * mu[(size_t)k * d + j];
        // tied: accumulate n_k-weighted pieces; the mean kernel below sums them
        sig[(size_t)k * d * d + idx] = tied ? (sxx - n * mm) : (sxx / n - mm);
    }
}
__global__ void __launch_bounds__(PT_THREADS)
mstep_invert(int K, int d, int tied, const double* sig_in, const double* ntot, double* lmbda,
             double* gscratch, int32_t* info) {
    const int k = blockIdx.x, tid = threadIdx.x;
    double* b1 = gscratch + (size_t)k * 2 * d * d;
    double* b2 = b1 + d * d;
    const double* s = tied ? sig_in : sig_in + (size_t)k * d * d;
    const double scale = tied ? (double)K / ntot[0] : 1.0;     // mean over k * K / sum n = sum / sum n
    for (int idx = tid; idx < d * d; idx += PT_THREADS) {
        int i = idx / d, j = idx - i * d;
        b1[idx] = 0.5 * scale * (s[i * d + j] + s[j * d + i]) + (i == j ? 1e-16 : 0.0);
    }
    if (!cta_chol_lower(b1, d, nullptr)) { flag_fail(info, MIMO_ENOTPD, k); return; }
    cta_tri_inv_lower(b1, b2, d);
    cta_gram_lower(b2, b1, d);
    for (int idx = tid; idx < d * d; idx += PT_THREADS) lmbda[(size_t)k * d * d + idx] = b1[idx];
}
__global__ void sum_counts(const double* stat, int K, int F, int f, double* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < K; ++k) s += stat[(size_t)k * F + f];
        out[0] = s;
    }
}

size_t mstep_workspace(int K, int d) { return a256(8 * (size_t)K * d * d) * 3 + a256(8 * (size_t)d * d) + 256; }

int mstep_gauss(int K, int d, int tied, const double* stat, int F, const int32_t* stat_idx, int Dp,
                double* mu, double* lmbda, void* workspace, size_t workspace_bytes, int32_t* info, cudaStream_t st) {
    MIMO_CHECK_ARG(K >= 1 && d >= 1 && stat && stat_idx && mu && lmbda && workspace && info, "arguments");
    MIMO_CHECK_ARG(F >= Dp * (Dp + 1) / 2 && workspace_bytes >= mstep_workspace(K, d), "layout / workspace");
    char* ws = (char*)workspace;
    double* sig = (double*)ws; ws += a256(8 * (size_t)K * d * d);
    double* scratch = (double*)ws; ws += 2 * a256(8 * (size_t)K * d * d);
    double* sbar = (double*)ws; ws += a256(8 * (size_t)d * d);
    double* ntot = (double*)ws;
    mstep_gauss_sigma<<<K, PT_THREADS, 0, st>>>(K, d, tied, stat, F, stat_idx, mu, sig);
    if (tied) {
        int pc = Dp - 1;
        mean_over_k<<<cdiv(d * d, 256), 256, 0, st>>>(sig, K, d * d, d * d, sbar);
        sum_counts<<<1, 32, 0, st>>>(stat, K, F, pc * (pc + 1) / 2 + pc, ntot);
    }
    mstep_invert<<<K, PT_THREADS, 0, st>>>(K, d, tied, tied ? sbar : sig, ntot, lmbda, scratch, info);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

__global__ void mstep_diag_kernel(int K, int d, int tied, const double* stat, int F, double* mu, double* lam,
                                  double* acc /* (d) tied accumulator, pre-zeroed */, const double* ntot, int phase) {
    const int k = blockIdx.x;
    const double* st = stat + (size_t)k * F;
    const double n = st[2 * d];
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        size_t o = (size_t)k * d + j;
        double m = st[j] / n;
        if (phase == 0) {
            mu[o] = m;
            if (!tied) lam[o] = 1.0 / (st[d + j] / n - m * m + 1e-16);
            else atomicAdd(&acc[j], st[d + j] - n * m * m);
        } else {
            lam[o] = 1.0 / (acc[j] / ntot[0] + 1e-16);
        }
    }
}

int mstep_gauss_diag(int K, int d, int tied, const double* stat, int F, double* mu, double* lam,
                     void* workspace, size_t workspace_bytes, cudaStream_t st) {
    MIMO_CHECK_ARG(K >= 1 && d >= 1 && stat && mu && lam && F >= 2 * d + 1, "arguments");
    MIMO_CHECK_ARG(!tied || (workspace && workspace_bytes >= 8 * (size_t)(d + 1)), "workspace too small");
    double* acc = (double*)workspace;
    double* ntot = acc ? acc + d : nullptr;
    if (tied) {
        MIMO_CUDA(cudaMemsetAsync(acc, 0, 8 * (size_t)(d + 1), st));
        sum_counts<<<1, 32, 0, st>>>(stat, K, F, 2 * d, ntot);
    }
    mstep_diag_kernel<<<K, 128, 0, st>>>(K, d, tied, stat, F, mu, lam, acc, ntot, 0);
    if (tied) mstep_diag_kernel<<<K, 128, 0, st>>>(K, d, tied, stat, F, mu, lam, acc, ntot, 1);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

// Linear-Gaussian M-step (lingauss.py:350-367; tied :384-400): A = Syx Sxx^-1,
// Sigma = (Syy - A Syx^T)/n.  Sxx is SPD, so the solve goes through its Cholesky factor.
__global__ void __launch_bounds__(PT_THREADS)
mstep_lingauss_kernel(int K, int c, int o, int tied, const double* stat, int F, const int32_t* stat_idx,
                      double* A, double* sig, double* gscratch, int32_t* info) {
    const int k = blockIdx.x, tid = threadIdx.x;
    const double* st = stat + (size_t)k * F;
    const int32_t* xi = stat_idx; const int32_t* yi = stat_idx + c;
    const int pc = stat_idx[c + o];
    const double n = st[tri_idx(pc, pc)];
    double* b1 = gscratch + (size_t)k * 2 * c * c;
    double* b2 = b1 + c * c;
    for (int idx = tid; idx < c * c; idx += PT_THREADS) b1[idx] = st[tri_idx(xi[idx / c], xi[idx % c])];
    if (!cta_chol_lower(b1, c, nullptr)) { flag_fail(info, MIMO_ENOTPD, k); return; }
    cta_tri_inv_lower(b1, b2, c);
    cta_gram_lower(b2, b1, c);                          // Sxx^-1
    double* Ak = A + (size_t)k * o * c;
    for (int idx = tid; idx < o * c; idx += PT_THREADS) {
        int i = idx / c, j = idx - i * c;
        double s = 0.0;
        for (int m = 0; m < c; ++m) s = fma(st[tri_idx(yi[i], xi[m])], b1[m * c + j], s);
        Ak[idx] = s;
    }
    __syncthreads();
    for (int idx = tid; idx < o * o; idx += PT_THREADS) {
        int i = idx / o, j = idx - i * o;
        double s = st[tri_idx(yi[i], yi[j])];
        for (int m = 0; m < c; ++m) s = fma(-Ak[i * c + m], st[tri_idx(yi[j], xi[m])], s);
        sig[(size_t)k * o * o + idx] = tied ? s : s / n;
    }
}

size_t mstep_lingauss_workspace(int K, int c, int o) {
    size_t n = (size_t)std::max(c, o);
    return a256(8 * (size_t)K * o * o) + a256(8 * (size_t)K * 2 * n * n) + a256(8 * (size_t)o * o) + 256;
}

int mstep_lingauss(int K, int c, int o, int tied, const double* stat, int F, const int32_t* stat_idx, int Dp,
                   double* A, double* lmbda, void* workspace, size_t workspace_bytes, int32_t* info, cudaStream_t st) {
    MIMO_CHECK_ARG(K >= 1 && c >= 1 && o >= 1 && stat && stat_idx && A && lmbda && workspace && info, "arguments");
    MIMO_CHECK_ARG(F >= Dp * (Dp + 1) / 2 && workspace_bytes >= mstep_lingauss_workspace(K, c, o), "layout / workspace");
    size_t n = (size_t)std::max(c, o);
    char* ws = (char*)workspace;
    double* sig = (double*)ws; ws += a256(8 * (size_t)K * o * o);
    double* scratch = (double*)ws; ws += a256(8 * (size_t)K * 2 * n * n);
    double* sbar = (double*)ws; ws += a256(8 * (size_t)o * o);
    double* ntot = (double*)ws;
    mstep_lingauss_kernel<<<K, PT_THREADS, 0, st>>>(K, c, o, tied, stat, F, stat_idx, A, sig, scratch, info);
    if (tied) {
        int pc = Dp - 1;
        mean_over_k<<<cdiv(o * o, 256), 256, 0, st>>>(sig, K, o * o, o * o, sbar);
        sum_counts<<<1, 32, 0, st>>>(stat, K, F, pc * (pc + 1) / 2 + pc, ntot);
    }
    mstep_invert<<<K, PT_THREADS, 0, st>>>(K, o, tied, tied ? sbar : sig, ntot, lmbda, scratch, info);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
}

}  // namespace mimo
