// gp_posterior.cu — batched MNIW posterior and Student-t predictive on a grid: the step AFTER the hot path
// (SURVEY.md 8f item 3).  The reference's figure scripts turn the traced statistics of K iterations into standard
// parameters with jax.vmap(prior_mniw_2naturalPara_inv) (SingleMassOscillator_Figures.py:58-89; function:
// src/BayesianInferrence.py:35-45) and then loop prior_mniw_Predictive over the iterations on the plot grid
// (SingleMassOscillator_Figures.py:131-140; function: src/BayesianInferrence.py:64-89), keeping only
// diag(col_scale) and the mean.  Here one CTA owns one statistics set k:
//   eta1_k = L L^T (in the workspace, L2-resident);  mean_k^T = L^-T L^-1 eta0_k;  Psi_k = eta2_k - mean_k eta0_k;
//   for every grid point g:  w = L^-1 phi_g,  col_scale_gg = 1 + |w|^2  (= phi_g V phi_g^T + 1 with V = eta1^-1, no inverse
//   is formed),  pred_mean_g = mean_k phi_g.
// The reference's full (M x M) col_cov and (G x G) col_scale are never needed by its callers and are not produced.
#include "common.cuh"
#include "sweep_args.cuh"

constexpr int GT = 256;       // threads per CTA = grid points in flight

struct GpPostArgs {
    int M, n, K, G;
    const double* eta0;       // (K, M, n)
    const double* eta1;       // (K, M, M)
    const double* eta2;       // (K, n, n)
    const double* eta3;       // (K)
    const double* basis;      // (G, M)
    double* mean;             // (K, n, M)
    double* row_scale;        // (K, n, n)
    double* df;               // (K)
    double* pred_mean;        // (K, G, n)
    double* pred_cs;          // (K, G)
    int* status;              // (K)
    double* wsL;              // (K, M, M)   Cholesky factor
    double* wsY;              // (K, M, n)   L^-1 eta0, then mean^T
    double* wsW;              // (K, M, GT)  forward-substitution vectors of the grid points in flight
};

__global__ void __launch_bounds__(GT) gp_posterior_kernel(const __grid_constant__ GpPostArgs a) {
    const int k = blockIdx.x, tid = threadIdx.x;
    const int M = a.M, n = a.n, G = a.G;
    const double* eta0 = a.eta0 + (size_t)k * M * n;
    const double* eta1 = a.eta1 + (size_t)k * M * M;
    double* L = a.wsL + (size_t)k * M * M;
    double* Y = a.wsY + (size_t)k * M * n;
    double* W = a.wsW + (size_t)k * M * GT;
    __shared__ int s_status;
    __shared__ double s_piv;
    if (tid == 0) s_status = 0;
    for (size_t e = tid; e < (size_t)M * M; e += GT) {
        const int i = (int)(e / M), j = (int)(e % M);
        if (j <= i) L[e] = 0.5 * (eta1[e] + eta1[(size_t)j * M + i]);        // symmetrised lower triangle
    }
    for (int e = tid; e < M * n; e += GT) Y[e] = eta0[e];
    __syncthreads();
    // ---- 1. right-looking column Cholesky (src/BayesianInferrence.py:12: jnp.linalg.cholesky)
    for (int j = 0; j < M; ++j) {
        if (tid == 0) {
            const double d = L[(size_t)j * M + j];
            if (!(d > 0.0) && s_status == 0) s_status = j + 1;
            s_piv = sqrt(d);
            L[(size_t)j * M + j] = s_piv;
        }
        __syncthreads();
        const double piv = s_piv;
        for (int i = j + 1 + tid; i < M; i += GT) L[(size_t)i * M + j] /= piv;
        __syncthreads();
        const int rem = M - 1 - j;                                            // trailing rows / columns j+1 .. M-1
        for (int e = tid; e < rem * rem; e += GT) {
            const int i = j + 1 + e / rem, c = j + 1 + e % rem;
            if (c <= i) L[(size_t)i * M + c] = fma(-L[(size_t)i * M + j], L[(size_t)c * M + j], L[(size_t)i * M + c]);
        }
        __syncthreads();
    }
    // ---- 2. mean^T = L^-T L^-1 eta0  (cho_solve, :13): thread c < n owns a column
    if (tid < n) {
        for (int i = 0; i < M; ++i) {
            double v = Y[(size_t)i * n + tid];
            for (int p = 0; p < i; ++p) v = fma(-L[(size_t)i * M + p], Y[(size_t)p * n + tid], v);
            Y[(size_t)i * n + tid] = v / L[(size_t)i * M + i];
        }
        for (int i = M - 1; i >= 0; --i) {
            double v = Y[(size_t)i * n + tid];
            for (int p = i + 1; p < M; ++p) v = fma(-L[(size_t)p * M + i], Y[(size_t)p * n + tid], v);
            Y[(size_t)i * n + tid] = v / L[(size_t)i * M + i];
        }
    }
    __syncthreads();
    for (int e = tid; e < M * n; e += GT) a.mean[(size_t)k * n * M + (size_t)(e % n) * M + e / n] = Y[e];   // (n, M)
    // ---- 3. row_scale = eta2 - mean eta0 (:42), df = eta3
    if (tid < n * n) {
        const int r = tid / n, c = tid % n;
        double s = 0.0;
        for (int m = 0; m < M; ++m) s = fma(Y[(size_t)m * n + r], eta0[(size_t)m * n + c], s);
        a.row_scale[(size_t)k * n * n + tid] = a.eta2[(size_t)k * n * n + tid] - s;
    }
    if (tid == 0) { a.df[k] = a.eta3[k]; a.status[k] = s_status; }
    // ---- 4. predictive on the grid (src/BayesianInferrence.py:64-89): one grid point per thread
    for (int g0 = 0; g0 < G; g0 += GT) {
        const int g = g0 + tid;
        if (g < G) {
            const double* phi = a.basis + (size_t)g * M;
            double q = 0.0;
            for (int i = 0; i < M; ++i) {
                double v = phi[i];
                const double* Li = L + (size_t)i * M;
                for (int p = 0; p < i; ++p) v = fma(-Li[p], W[(size_t)p * GT + tid], v);
                v /= Li[i];
                W[(size_t)i * GT + tid] = v;
                q = fma(v, v, q);
            }
            a.pred_cs[(size_t)k * G + g] = q + 1.0;                          // diag(basis V basis^T + I)
            for (int c = 0; c < n; ++c) {
                double s = 0.0;
                for (int m = 0; m < M; ++m) s = fma(phi[m], Y[(size_t)m * n + c], s);
                a.pred_mean[((size_t)k * G + g) * n + c] = s;                // basis mean^T
            }
        }
    }
}

static size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

extern "C" size_t pgas_mniw_posterior_batch_workspace_bytes(int32_t M, int32_t n, int32_t K) {
    if (M < 1 || n < 1 || K < 1) return 0;
    return al256(sizeof(double) * (size_t)K * M * M) + al256(sizeof(double) * (size_t)K * M * n) + al256(sizeof(double) * (size_t)K * M * GT) + 256;
}

extern "C" int pgas_mniw_posterior_batch_f64(int32_t M, int32_t n, int32_t K, const double* eta0, const double* eta1, const double* eta2,
                                             const double* eta3, const double* basis, int32_t G, double* mean_out,
                                             double* row_scale_out, double* df_out, double* pred_mean_out, double* pred_colscale_out,
                                             int32_t* status_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!eta0 || !eta1 || !eta2 || !eta3 || !mean_out || !row_scale_out || !df_out || !status_out || !workspace)
        PGAS_FAIL(-1, "pgas_mniw_posterior_batch_f64: null argument");
    if (G > 0 && (!basis || !pred_mean_out || !pred_colscale_out)) PGAS_FAIL(-1, "pgas_mniw_posterior_batch_f64: grid outputs missing");
    if (M < 1 || n < 1 || n > PGAS_MAX_NX || K < 1 || G < 0) PGAS_FAIL(-2, "bad sizes (M=%d n=%d K=%d G=%d)", M, n, K, G);
    if (workspace_bytes < pgas_mniw_posterior_batch_workspace_bytes(M, n, K)) PGAS_FAIL(-5, "workspace too small for the posterior batch");
    GpPostArgs a;
    a.M = M; a.n = n; a.K = K; a.G = G;
    a.eta0 = eta0; a.eta1 = eta1; a.eta2 = eta2; a.eta3 = eta3; a.basis = basis;
    a.mean = mean_out; a.row_scale = row_scale_out; a.df = df_out; a.pred_mean = pred_mean_out; a.pred_cs = pred_colscale_out;
    a.status = status_out;
    char* base = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    a.wsL = (double*)base;
    a.wsY = (double*)(base + al256(sizeof(double) * (size_t)K * M * M));
    a.wsW = (double*)((char*)a.wsY + al256(sizeof(double) * (size_t)K * M * n));
    gp_posterior_kernel<<<K, GT, 0, (cudaStream_t)stream>>>(a);
    PGAS_KERNEL_CHECK();
    return 0;
}
