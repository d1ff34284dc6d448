// mniw_draw.cu — posterior draw (A, Sigma) ~ MNIW(eta) (second half of PGAS.sample_params,
// reference src/PGAS.py:306-343 with prior_mniw_2naturalPara_inv, src/BayesianInferrence.py:35-45).
//
// The reference forms V = eta1^-1 explicitly (Cholesky + cho_solve against the identity, 2 M^3
// flops), factors it again (V_chol = chol(V), M^3/3) and multiplies Nrm V_chol.  Here ONE
// factorisation is done: the reverse ("UL") Cholesky eta1 = U U^T, U upper triangular, obtained
// as the ordinary lower Cholesky L' of the index-reversed matrix J eta1 J (U = J L' J).  Then
//   mean^T       = eta1^-1 eta0 = U^-T U^-1 eta0           (two triangular solves, n_x columns)
//   V_chol       = U^-T  exactly (the lower Cholesky factor of V is unique and U^-T is lower
//                  triangular with positive diagonal and U^-T U^-1 = V)
//   Nrm V_chol   = X with U X^T = Nrm^T                    (one triangular solve, n_x columns)
//   Nrm V_chol^T = X with U^T X^T = Nrm^T                  (PGAS_FLAG_VCHOL_TRANSPOSE)
// so the draw costs M^3/3 + O(n_x M^2) flops instead of 2.67 M^3 and needs no inverse.
// The inverse-Wishart part (p = n_x <= 4) follows src/PGAS.py:312-335 literally in one thread.
// One CTA per chain; the factor lives in the caller's workspace (L2-resident for M <= ~2000).
#include "common.cuh"
#include "sweep_args.cuh"
#include <algorithm>

constexpr int DT = 256;      // threads
constexpr int NB = 32;       // panel width
constexpr int TS = 64;       // trailing-update tile

struct DrawArgs {
    int M, nx, n_chains, flags;
    double eta3;
    const double* eta0;      // (n_chains, M, nx)
    const double* eta1;      // (n_chains, M, M)
    const double* eta2;      // (n_chains, nx, nx)
    long long eta_stride0, eta_stride1, eta_stride2;    // 0 when a single eta is shared by all chains
    int rng_mode;
    unsigned long long seed;
    unsigned chain_base, iteration;
    const double* chi2;      // (n_chains, nx)
    const double* G;         // (n_chains, nx, nx)
    const double* Nrm;       // (n_chains, nx, M)
    double* A;               // (n_chains, nx, M)
    double* S;               // (n_chains, nx, nx)
    int* status;             // (n_chains)
    double* wsB;             // (n_chains, M, M)      reversed eta1 -> L'
    double* wsF;             // (n_chains, M, 2 nx)   forward right-hand sides / solutions
    double* wsK;             // (n_chains, M, 2 nx)   backward right-hand sides / solutions
};

// chi-square(nu) = 2 Gamma(nu/2, 1) by Marsaglia-Tsang; attempts indexed by the Philox counter
__device__ double philox_chisquare(unsigned long long seed, unsigned chain, unsigned iter, unsigned idx, double nu) {
    double a = 0.5 * nu, boost = 1.0;
    unsigned attempt = 0;
    if (a < 1.0) {
        double u, ub;
        philox_uniform2(seed, PURPOSE_DRAW_CHI, chain, iter, 0x40000000u, idx, u, ub);
        boost = pow(u + (1.0 / 9007199254740992.0), 1.0 / a);
        a += 1.0;
    }
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (;; ++attempt) {
        double x, xb, u, ub;
        philox_normal2(seed, PURPOSE_DRAW_CHI, chain, iter, 2 * attempt, idx, x, xb);
        philox_uniform2(seed, PURPOSE_DRAW_CHI, chain, iter, 2 * attempt + 1, idx, u, ub);
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        if (log(u + (1.0 / 9007199254740992.0)) < 0.5 * x * x + d - d * v + d * log(v) || attempt > 1000) return 2.0 * d * v * boost;
    }
}

// BLOCKED = false: everything in one CTA per chain (small M).  BLOCKED = true: the factorisation and the forward solves were done by
// chol_prep_kernel / chol_panel_kernel / chol_update_kernel (multi-CTA, below) on the augmented matrix; this kernel does the
// backward solve, Psi, the inverse-Wishart part and A.
template <bool BLOCKED>
__global__ void __launch_bounds__(DT) mniw_draw_kernel(const __grid_constant__ DrawArgs a) {
    const int chain = blockIdx.x, tid = threadIdx.x;
    const int M = a.M, nx = a.nx, R2 = 2 * nx;
    const bool vt = (a.flags & PGAS_FLAG_VCHOL_TRANSPOSE) != 0;
    const double* eta0 = a.eta0 + (size_t)chain * a.eta_stride0;
    const double* eta1 = a.eta1 + (size_t)chain * a.eta_stride1;
    const double* eta2 = a.eta2 + (size_t)chain * a.eta_stride2;
    double* B = a.wsB + (size_t)chain * (BLOCKED ? (size_t)(M + R2) * M : (size_t)M * M);
    double* F = a.wsF + (size_t)chain * M * R2;
    double* K = a.wsK + (size_t)chain * M * R2;
    double* Aout = a.A + (size_t)chain * nx * M;

    __shared__ double Dg[NB][NB + 1];                 // diagonal block
    __shared__ __align__(16) double PiT[NB][TS];      // panel rows of tile i, transposed [p][row]
    __shared__ __align__(16) double PjT[NB][TS];
    __shared__ double xs[NB][2 * PGAS_MAX_NX];
    __shared__ double red[DT / 32][PGAS_MAX_NX * PGAS_MAX_NX];
    __shared__ double Sc[PGAS_MAX_NX][PGAS_MAX_NX];   // S_chol
    __shared__ int s_status;
    if (tid == 0) s_status = (BLOCKED && a.status) ? a.status[chain] : 0;

    if constexpr (!BLOCKED) {
    // ---- 0. B = J eta1 J (lower part), right-hand sides
    for (size_t e = tid; e < (size_t)M * M; e += DT) {
        const int i = (int)(e / M), j = (int)(e % M);
        if (j <= i) B[e] = eta1[(size_t)(M - 1 - i) * M + (M - 1 - j)];
    }
    for (int e = tid; e < M * nx; e += DT) {
        const int i = e / nx, k = e % nx;
        F[(size_t)i * R2 + k] = eta0[(size_t)(M - 1 - i) * nx + k];
        double z;
        if (a.rng_mode == 1) {
            z = a.Nrm[((size_t)chain * nx + k) * M + (M - 1 - i)];
        } else {
            // Nrm[k, m] = normal #(k*M + m): pairs share a Philox block
            const unsigned flat = (unsigned)(k * M + (M - 1 - i));
            double za, zb;
            philox_normal2(a.seed, PURPOSE_DRAW_N, a.chain_base + chain, a.iteration, 0u, flat >> 1, za, zb);
            z = (flat & 1) ? zb : za;
        }
        if (vt) K[(size_t)i * R2 + nx + k] = z; else F[(size_t)i * R2 + nx + k] = z;
    }
    __syncthreads();

    // ---- 1. blocked right-looking Cholesky B = L' L'^T (in place, lower)
    for (int kb = 0; kb < M; kb += NB) {
        const int nb = min(NB, M - kb);
        for (int e = tid; e < NB * NB; e += DT) {
            const int r = e / NB, c = e % NB;
            Dg[r][c] = (r < nb && c <= r) ? B[(size_t)(kb + r) * M + kb + c] : (r == c ? 1.0 : 0.0);
        }
        __syncthreads();
        for (int j = 0; j < nb; ++j) {
            if (tid == 0) {
                const double d = Dg[j][j];
                if (!(d > 0.0) && s_status == 0) s_status = kb + j + 1;
                Dg[j][j] = sqrt(d);
            }
            __syncthreads();
            if (tid > j && tid < nb) Dg[tid][j] /= Dg[j][j];
            __syncthreads();
            for (int e = tid; e < nb * nb; e += DT) {
                const int r = e / nb, c = e % nb;
                if (c > j && r >= c) Dg[r][c] -= Dg[r][j] * Dg[c][j];
            }
            __syncthreads();
        }
        for (int e = tid; e < nb * nb; e += DT) {
            const int r = e / nb, c = e % nb;
            if (c <= r) B[(size_t)(kb + r) * M + kb + c] = Dg[r][c];
        }
        // panel: rows below the diagonal block, X Dg^T = B[i, kb:kb+nb]
        for (int i = kb + nb + tid; i < M; i += DT) {
            double x[NB];
            double* row = B + (size_t)i * M + kb;
#pragma unroll
            for (int c = 0; c < NB; ++c) x[c] = (c < nb) ? row[c] : 0.0;
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                double v = x[c];
#pragma unroll
                for (int p = 0; p < c; ++p) v = fma(-x[p], Dg[c][p], v);
                x[c] = v / Dg[c][c];
            }
#pragma unroll
            for (int c = 0; c < NB; ++c) if (c < nb) row[c] = x[c];
        }
        __syncthreads();
        // trailing update: B[i,j] -= sum_p P[i,p] P[j,p] on the lower-triangular 64x64 tiles
        const int r0 = kb + nb;
        const int nt = (M - r0 + TS - 1) / TS;
        const int ty = tid / 16, tx = tid % 16;
        for (int bi = 0; bi < nt; ++bi)
            for (int bj = 0; bj <= bi; ++bj) {
                for (int e = tid; e < TS * NB; e += DT) {
                    const int rr = e / NB, p = e % NB;
                    const int gi = r0 + bi * TS + rr, gj = r0 + bj * TS + rr;
                    PiT[p][rr] = (gi < M && p < nb) ? B[(size_t)gi * M + kb + p] : 0.0;
                    PjT[p][rr] = (gj < M && p < nb) ? B[(size_t)gj * M + kb + p] : 0.0;
                }
                __syncthreads();
                double acc[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
#pragma unroll 4
                for (int p = 0; p < NB; ++p) {
                    const double2 a0 = *reinterpret_cast<const double2*>(&PiT[p][ty * 4]);
                    const double2 a1 = *reinterpret_cast<const double2*>(&PiT[p][ty * 4 + 2]);
                    const double2 b0 = *reinterpret_cast<const double2*>(&PjT[p][tx * 4]);
                    const double2 b1 = *reinterpret_cast<const double2*>(&PjT[p][tx * 4 + 2]);
                    const double av[4] = {a0.x, a0.y, a1.x, a1.y}, bv[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[r][c] = fma(av[r], bv[c], acc[r][c]);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int gi = r0 + bi * TS + ty * 4 + r, gj = r0 + bj * TS + tx * 4 + c;
                        if (gi < M && gj <= gi) B[(size_t)gi * M + gj] -= acc[r][c];
                    }
                __syncthreads();
            }
    }

    // ---- 2. forward solve L' F = rhs (all 2 nx columns; the unused ones are skipped by width)
    const int wf = vt ? nx : R2;               // active forward columns
    for (int kb = 0; kb < M; kb += NB) {
        const int nb = min(NB, M - kb);
        if (tid < wf) {
            for (int r = 0; r < nb; ++r) {
                double v = F[(size_t)(kb + r) * R2 + tid];
                for (int p = 0; p < r; ++p) v = fma(-B[(size_t)(kb + r) * M + kb + p], xs[p][tid], v);
                v /= B[(size_t)(kb + r) * M + kb + r];
                xs[r][tid] = v;
                F[(size_t)(kb + r) * R2 + tid] = v;
            }
        }
        __syncthreads();
        for (int i = kb + nb + tid; i < M; i += DT) {
            double acc[2 * PGAS_MAX_NX];
#pragma unroll
            for (int c = 0; c < 2 * PGAS_MAX_NX; ++c) acc[c] = 0.0;
            const double* row = B + (size_t)i * M + kb;
            for (int p = 0; p < nb; ++p) {
                const double l = row[p];
#pragma unroll
                for (int c = 0; c < 2 * PGAS_MAX_NX; ++c) if (c < wf) acc[c] = fma(l, xs[p][c], acc[c]);
            }
#pragma unroll
            for (int c = 0; c < 2 * PGAS_MAX_NX; ++c) if (c < wf) F[(size_t)i * R2 + c] -= acc[c];
        }
        __syncthreads();
    }
    } else {
        // the forward solves came out of the augmented factorisation: F[i][c] = Baug[M + c][i]
        for (int e = tid; e < M * R2; e += DT) {
            const int i = e % M, c = e / M;
            F[(size_t)i * R2 + c] = B[(size_t)(M + c) * M + i];
        }
        __syncthreads();
    }
    // ---- 3. backward solve L'^T K = [y | (Nrm part when transposed)]
    const int wk = vt ? R2 : nx;
    for (int e = tid; e < M * nx; e += DT) K[(size_t)(e / nx) * R2 + e % nx] = F[(size_t)(e / nx) * R2 + e % nx];
    __syncthreads();
    for (int kb = ((M - 1) / NB) * NB; kb >= 0; kb -= NB) {
        const int nb = min(NB, M - kb);
        for (int e = tid; e < NB * NB; e += DT) {
            const int r = e / NB, c = e % NB;
            Dg[r][c] = (r < nb && c <= r) ? B[(size_t)(kb + r) * M + kb + c] : 0.0;
        }
        __syncthreads();
        for (int e = tid; e < nb * wk; e += DT) xs[e / wk][e % wk] = K[(size_t)(kb + e / wk) * R2 + e % wk];
        __syncthreads();
        {   // one warp per right-hand side, block and right-hand sides in shared memory: lanes split the dot product over the rows below r
            const int lane = tid & 31, w = tid >> 5;
            if (w < wk) {
                for (int r = nb - 1; r >= 0; --r) {
                    const int p = r + 1 + lane;
                    double part = (p < nb) ? Dg[p][r] * xs[p][w] : 0.0;
                    part = warp_sum(part);
                    if (lane == 0) xs[r][w] = (xs[r][w] - part) / Dg[r][r];
                    __syncwarp();
                }
            }
        }
        __syncthreads();
        for (int e = tid; e < nb * wk; e += DT) K[(size_t)(kb + e / wk) * R2 + e % wk] = xs[e / wk][e % wk];
        for (int j = tid; j < kb; j += DT) {
            double acc[2 * PGAS_MAX_NX];
#pragma unroll
            for (int c = 0; c < 2 * PGAS_MAX_NX; ++c) acc[c] = 0.0;
            for (int p = 0; p < nb; ++p) {
                const double l = B[(size_t)(kb + p) * M + j];
#pragma unroll
                for (int c = 0; c < 2 * PGAS_MAX_NX; ++c) if (c < wk) acc[c] = fma(l, xs[p][c], acc[c]);
            }
#pragma unroll
            for (int c = 0; c < 2 * PGAS_MAX_NX; ++c) if (c < wk) K[(size_t)j * R2 + c] -= acc[c];
        }
        __syncthreads();
    }
    // now: mean[k][m] = K[M-1-m][k];  X[k][m] = (vt ? K : F)[M-1-m][nx + k]

    // ---- 4. Psi = eta2 - mean eta0   (src/BayesianInferrence.py:42)
    {
        double part[PGAS_MAX_NX * PGAS_MAX_NX];
#pragma unroll
        for (int e = 0; e < PGAS_MAX_NX * PGAS_MAX_NX; ++e) part[e] = 0.0;
        for (int mm = tid; mm < M; mm += DT)
            for (int r = 0; r < nx; ++r)
                for (int c = 0; c < nx; ++c)
                    part[r * PGAS_MAX_NX + c] = fma(K[(size_t)(M - 1 - mm) * R2 + r], eta0[(size_t)mm * nx + c], part[r * PGAS_MAX_NX + c]);
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int e = 0; e < PGAS_MAX_NX * PGAS_MAX_NX; ++e) {
            const double v = warp_sum(part[e]);
            if (lane == 0) red[warp][e] = v;
        }
    }
    __syncthreads();
    if (tid == 0) {
        // inverse-Wishart draw, src/PGAS.py:312-335
        double Psi[PGAS_MAX_NX][PGAS_MAX_NX], Cr[PGAS_MAX_NX][PGAS_MAX_NX], Li[PGAS_MAX_NX][PGAS_MAX_NX];
        double Tm[PGAS_MAX_NX][PGAS_MAX_NX], Cm[PGAS_MAX_NX][PGAS_MAX_NX], Ci[PGAS_MAX_NX][PGAS_MAX_NX];
        for (int r = 0; r < nx; ++r)
            for (int c = 0; c < nx; ++c) {
                double s = 0.0;
                for (int w = 0; w < DT / 32; ++w) s += red[w][r * PGAS_MAX_NX + c];
                Psi[r][c] = eta2[r * nx + c] - s;
            }
        // chol_row = chol(Psi)  (:317)
        for (int i = 0; i < nx; ++i) for (int j = 0; j < nx; ++j) { Cr[i][j] = 0.0; Li[i][j] = 0.0; Tm[i][j] = 0.0; Ci[i][j] = 0.0; }
        for (int j = 0; j < nx; ++j) {
            double d = Psi[j][j];
            for (int k = 0; k < j; ++k) d -= Cr[j][k] * Cr[j][k];
            if (!(d > 0.0) && s_status == 0) s_status = -(j + 1);
            d = sqrt(d);
            Cr[j][j] = d;
            for (int i = j + 1; i < nx; ++i) {
                double v = Psi[i][j];
                for (int k = 0; k < j; ++k) v -= Cr[i][k] * Cr[j][k];
                Cr[i][j] = v / d;
            }
        }
        // L = chol_row^-1  (:319)
        for (int j = 0; j < nx; ++j) {
            Li[j][j] = 1.0 / Cr[j][j];
            for (int i = j + 1; i < nx; ++i) {
                double v = 0.0;
                for (int k = j; k < i; ++k) v -= Cr[i][k] * Li[k][j];
                Li[i][j] = v / Cr[i][i];
            }
        }
        // T = tril(normals,-1) + diag(sqrt(chi2(nu - i)))  (:323-329)
        for (int i = 0; i < nx; ++i) {
            double c2;
            if (a.rng_mode == 1) c2 = a.chi2[(size_t)chain * nx + i];
            else c2 = philox_chisquare(a.seed, a.chain_base + chain, a.iteration, (unsigned)i, a.eta3 - (double)i);
            Tm[i][i] = sqrt(c2);
            for (int j = 0; j < i; ++j) {
                if (a.rng_mode == 1) Tm[i][j] = a.G[((size_t)chain * nx + i) * nx + j];
                else {
                    const unsigned flat = (unsigned)(i * nx + j);
                    double za, zb;
                    philox_normal2(a.seed, PURPOSE_DRAW_G, a.chain_base + chain, a.iteration, 0u, flat >> 1, za, zb);
                    Tm[i][j] = (flat & 1) ? zb : za;
                }
            }
        }
        // C = L T (:332);  S_chol = C^-T (:334);  S = S_chol S_chol^T (:335)
        for (int i = 0; i < nx; ++i)
            for (int j = 0; j < nx; ++j) {
                double s = 0.0;
                for (int k = 0; k < nx; ++k) s += Li[i][k] * Tm[k][j];
                Cm[i][j] = s;
            }
        for (int j = 0; j < nx; ++j) {                 // Ci = C^-1 (C lower triangular)
            Ci[j][j] = 1.0 / Cm[j][j];
            for (int i = j + 1; i < nx; ++i) {
                double v = 0.0;
                for (int k = j; k < i; ++k) v -= Cm[i][k] * Ci[k][j];
                Ci[i][j] = v / Cm[i][i];
            }
        }
        for (int i = 0; i < nx; ++i)
            for (int j = 0; j < nx; ++j) Sc[i][j] = Ci[j][i];        // S_chol = C^-T (upper triangular)
        for (int i = 0; i < nx; ++i)
            for (int j = 0; j < nx; ++j) {
                double s = 0.0;
                for (int k = 0; k < nx; ++k) s += Sc[i][k] * Sc[j][k];
                a.S[((size_t)chain * nx + i) * nx + j] = s;
            }
        if (a.status) a.status[chain] = s_status;
    }
    __syncthreads();
    // ---- 5. A = mean + S_chol (Nrm V_chol)   (:338-341)
    const double* Xs = vt ? K : F;
    for (int e = tid; e < nx * M; e += DT) {
        const int k = e / M, mm = e % M;
        double v = K[(size_t)(M - 1 - mm) * R2 + k];
        for (int j = 0; j < nx; ++j) v = fma(Sc[k][j], Xs[(size_t)(M - 1 - mm) * R2 + nx + j], v);
        Aout[(size_t)k * M + mm] = v;
    }
}


// =================================================================================== multi-CTA blocked factorisation (large M)
// Right-looking blocked Cholesky of the index-reversed eta1 over ALL SMs, one launch pair per 64-column panel:
//   chol_panel_kernel   every CTA factorises the 64x64 diagonal block redundantly in shared memory (left-looking, one barrier per
//                       column) and solves its own 128 rows of the panel, one thread per row with the row in registers;
//   chol_update_kernel  trailing update B[i,j] -= P_i P_j^T on lower-triangular 64x64 tiles with FP64 DMMA (mma.sync.m8n8k4.f64).
// The right-hand sides of the forward solves (eta0 and, unless PGAS_FLAG_VCHOL_TRANSPOSE, Nrm) ride along as 2 n_x extra ROWS of the
// matrix: after the last panel row M + c holds (L'^-1 rhs_c)^T, so no separate forward substitution exists.
constexpr int CNB = 64;       // panel width = tile edge
constexpr int CPT = 128;      // threads of the panel / update kernels
constexpr int LDK = CNB + 4;  // row length of a staged panel tile: == 4 (mod 16) doubles -> fragment loads hit every bank twice

__global__ void __launch_bounds__(256) chol_prep_kernel(const __grid_constant__ DrawArgs a) {
    const int chain = blockIdx.y, M = a.M, nx = a.nx, R2 = 2 * nx;
    const bool vt = (a.flags & PGAS_FLAG_VCHOL_TRANSPOSE) != 0;
    const double* eta0 = a.eta0 + (size_t)chain * a.eta_stride0;
    const double* eta1 = a.eta1 + (size_t)chain * a.eta_stride1;
    double* B = a.wsB + (size_t)chain * (M + R2) * M;
    double* K = a.wsK + (size_t)chain * M * R2;
    const size_t total = (size_t)(M + R2) * M;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
        const int i = (int)(e / M), j = (int)(e % M);
        if (i < M) {
            if (j <= i) B[e] = eta1[(size_t)(M - 1 - i) * M + (M - 1 - j)];
        } else {
            const int c = i - M;
            double v = 0.0;
            if (c < nx) v = eta0[(size_t)(M - 1 - j) * nx + c];
            else {
                const int k = c - nx;
                double z;
                if (a.rng_mode == 1) z = a.Nrm[((size_t)chain * nx + k) * M + (M - 1 - j)];
                else {
                    const unsigned flat = (unsigned)(k * M + (M - 1 - j));
                    double za, zb;
                    philox_normal2(a.seed, PURPOSE_DRAW_N, a.chain_base + chain, a.iteration, 0u, flat >> 1, za, zb);
                    z = (flat & 1) ? zb : za;
                }
                if (vt) K[(size_t)j * R2 + c] = z; else v = z;
            }
            B[e] = v;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.status) a.status[chain] = 0;
}

__device__ __forceinline__ void bar_sync_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__global__ void __launch_bounds__(CPT) chol_panel_kernel(const __grid_constant__ DrawArgs a, int kb) {
    const int chain = blockIdx.y, tid = threadIdx.x, M = a.M, R2 = 2 * a.nx;
    double* B = a.wsB + (size_t)chain * (M + R2) * M;
    const int nb = min(CNB, M - kb);
    constexpr int LDT = CNB + 2;                       // even row length: 16-byte aligned rows for the paired broadcast loads
    __shared__ __align__(16) double LT[CNB * LDT];     // staging of the block, then the TRANSPOSED factor: LT[c][p] = L[p][c]
    __shared__ __align__(16) double colbuf[2][CNB];    // column j of the factor, double-buffered by parity
    __shared__ double rdiag[CNB], pdiag[CNB];
    for (int e = tid; e < CNB * CNB; e += CPT) {
        const int r = e / CNB, c = e % CNB;
        LT[r * LDT + c] = (r < nb && c <= r) ? B[(size_t)(kb + r) * M + kb + c] : (r == c ? 1.0 : 0.0);
    }
    // this thread's panel row (below the diagonal block, incl. the right-hand-side rows) travels while the block is factorised
    const int i = kb + nb + blockIdx.x * CPT + tid;
    double x[CNB];
    if (i < M + R2) {
        const double* row = B + (size_t)i * M + kb;
#pragma unroll
        for (int c = 0; c < CNB; ++c) x[c] = (c < nb) ? row[c] : 0.0;
    }
    __syncthreads();
    // ---- the diagonal block, redundantly in every CTA: right-looking, row r in the registers of thread r (threads 0..63), one column per
    //      iteration: [A] l_rj = a_rj / l_jj into colbuf, barrier, [B] a_rc -= l_rj l_cj for j < c <= r (paired broadcast loads of the
    //      column), and the owner of row j + 1 publishes the next pivot, barrier.  Two named barriers of 64 threads per column.
    if (tid < CNB) {
        const int r = tid;
        double ar[CNB];
#pragma unroll
        for (int c = 0; c < CNB; ++c) ar[c] = LT[r * LDT + c];
        int bad = 0;
        if (r == 0) {
            const double d = ar[0];
            if (!(d > 0.0)) bad = kb + 1;
            const double ri = rsqrt(d);
            rdiag[0] = ri; pdiag[0] = d * ri;
        }
        bar_sync_named(1, CNB);
#pragma unroll
        for (int j = 0; j < CNB; ++j) {
            double* cb = colbuf[j & 1];
            const double l = (r > j) ? ar[j] * rdiag[j] : 0.0;
            ar[j] = (r == j) ? pdiag[j] : l;
            cb[r] = l;
            bar_sync_named(1, CNB);
            if (r > j) {
#pragma unroll
                for (int c = (j + 1) & ~1; c < CNB; c += 2) {
                    const double2 lc = *reinterpret_cast<const double2*>(cb + c);
                    if (c > j) ar[c] = fma(-l, lc.x, ar[c]);
                    ar[c + 1] = fma(-l, lc.y, ar[c + 1]);
                }
            }
            if (j + 1 < CNB && r == j + 1) {
                const double d = ar[j + 1];
                if (!(d > 0.0) && !bad && j + 1 < nb) bad = kb + j + 2;
                const double ri = rsqrt(d);
                rdiag[j + 1] = ri; pdiag[j + 1] = d * ri;
            }
            bar_sync_named(1, CNB);
        }
        if (bad && blockIdx.x == 0 && a.status) atomicCAS(&a.status[chain], 0, bad);     // first non-positive pivot of the chain (panels run in order; one owner per column)
        // transposed factor for the panel solve; CTA 0 also writes the block back
#pragma unroll
        for (int c = 0; c < CNB; ++c) LT[c * LDT + r] = (c <= r) ? ar[c] : 0.0;
        if (blockIdx.x == 0 && r < nb) {
            double* row = B + (size_t)(kb + r) * M + kb;
#pragma unroll
            for (int c = 0; c < CNB; ++c) if (c <= r) row[c] = ar[c];
        }
    }
    __syncthreads();
    // ---- panel rows: X L^T = B[i, kb:kb+nb], row i in registers, right-looking: the updates of a column are independent FMAs fed by
    //      paired broadcast loads of the transposed factor
    if (i < M + R2) {
#pragma unroll
        for (int c = 0; c < CNB; ++c) {
            const double xc = x[c] * rdiag[c];
            x[c] = xc;
#pragma unroll
            for (int p = (c + 1) & ~1; p < CNB; p += 2) {
                const double2 lp = *reinterpret_cast<const double2*>(&LT[c * LDT + p]);
                if (p > c) x[p] = fma(-xc, lp.x, x[p]);
                x[p + 1] = fma(-xc, lp.y, x[p + 1]);
            }
        }
        double* row = B + (size_t)i * M + kb;
#pragma unroll
        for (int c = 0; c < CNB; ++c) if (c < nb) row[c] = x[c];
    }
}

__global__ void __launch_bounds__(CPT) chol_update_kernel(const __grid_constant__ DrawArgs a, int kb) {
    const int chain = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, M = a.M, R2 = 2 * a.nx;
    double* B = a.wsB + (size_t)chain * (M + R2) * M;
    const int r0 = kb + CNB;                                   // first trailing row / column
    const int ntc = (M - r0 + CNB - 1) / CNB;                  // column tiles (columns < M)
    int bi = 0, rem = blockIdx.x;
    while (rem > bi) { rem -= bi + 1; ++bi; }
    const int bj = rem;
    if (bj >= ntc) return;                                     // right-hand-side rows have no columns of their own
    extern __shared__ __align__(16) double psm[];
    double* Pi = psm;                                          // [CNB][LDK]: rows of tile bi, panel columns
    double* Pj = (bi == bj) ? Pi : psm + CNB * LDK;
    for (int e = tid; e < CNB * (CNB / 2); e += CPT) {
        const int rr = e / (CNB / 2), c2 = e % (CNB / 2);
        const int gi = r0 + bi * CNB + rr, gj = r0 + bj * CNB + rr;
        double2 vi = make_double2(0.0, 0.0), vj = vi;
        if (gi < M + R2) vi = make_double2(B[(size_t)gi * M + kb + 2 * c2], B[(size_t)gi * M + kb + 2 * c2 + 1]);
        *reinterpret_cast<double2*>(&Pi[rr * LDK + 2 * c2]) = vi;
        if (bi != bj) {
            if (gj < M) vj = make_double2(B[(size_t)gj * M + kb + 2 * c2], B[(size_t)gj * M + kb + 2 * c2 + 1]);
            *reinterpret_cast<double2*>(&Pj[rr * LDK + 2 * c2]) = vj;
        }
    }
    __syncthreads();
    const int wi = warp >> 1, wj = warp & 1;
    double c0[4][4], c1[4][4];
#pragma unroll
    for (int fi = 0; fi < 4; ++fi)
#pragma unroll
        for (int fj = 0; fj < 4; ++fj) { c0[fi][fj] = 0.0; c1[fi][fj] = 0.0; }
    const int aoff = (wi * 32 + (lane >> 2)) * LDK + (lane & 3), boff = (wj * 32 + (lane >> 2)) * LDK + (lane & 3);
#pragma unroll 1
    for (int kk = 0; kk < CNB; kk += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            af[f] = Pi[aoff + 8 * f * LDK + kk];
            bf[f] = Pj[boff + 8 * f * LDK + kk];
        }
#pragma unroll
        for (int fi = 0; fi < 4; ++fi)
#pragma unroll
            for (int fj = 0; fj < 4; ++fj)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c0[fi][fj]), "+d"(c1[fi][fj]) : "d"(af[fi]), "d"(bf[fj]));
    }
#pragma unroll
    for (int fi = 0; fi < 4; ++fi)
#pragma unroll
        for (int fj = 0; fj < 4; ++fj) {
            const int gi = r0 + bi * CNB + wi * 32 + fi * 8 + (lane >> 2), gj = r0 + bj * CNB + wj * 32 + fj * 8 + 2 * (lane & 3);
            if (gi < M + R2) {
                if (gj < M && gj <= gi) B[(size_t)gi * M + gj] -= c0[fi][fj];
                if (gj + 1 < M && gj + 1 <= gi) B[(size_t)gi * M + gj + 1] -= c1[fi][fj];
            }
        }
}

static size_t draw_ws_per_chain(int M, int nx) {
    return sizeof(double) * ((size_t)(M + 2 * nx) * M + 2 * (size_t)M * 2 * nx);     // (augmented) factor + forward / backward solutions
}

extern "C" size_t pgas_mniw_draw_workspace_bytes(int32_t M, int32_t n_x, int32_t n_chains) {
    return draw_ws_per_chain(M, n_x) * (size_t)n_chains + 256;
}

int pgas_launch_mniw_draw(const double* eta0, const double* eta1, const double* eta2, double eta3, bool shared_eta, int M, int nx,
                          int n_chains, const pgas_rng* rng, int flags, double* A, double* S, int* status, void* ws, size_t ws_bytes,
                          cudaStream_t st) {
    if (!eta0 || !eta1 || !eta2 || !A || !S || !rng || !ws) PGAS_FAIL(-1, "pgas_mniw_draw_f64: null argument");
    if (M < 1 || nx < 1 || nx > PGAS_MAX_NX || n_chains < 1) PGAS_FAIL(-2, "bad shape (M=%d n_x=%d n_chains=%d)", M, nx, n_chains);
    if (ws_bytes < pgas_mniw_draw_workspace_bytes(M, nx, n_chains)) PGAS_FAIL(-5, "workspace too small for the MNIW draw");
    if (rng->mode == 1 && (!rng->chi2 || !rng->G || !rng->Nrm)) PGAS_FAIL(-1, "injected rng mode needs chi2, G and Nrm");
    DrawArgs a;
    memset(&a, 0, sizeof(a));
    a.M = M; a.nx = nx; a.n_chains = n_chains; a.flags = flags; a.eta3 = eta3;
    a.eta0 = eta0; a.eta1 = eta1; a.eta2 = eta2;
    a.eta_stride0 = shared_eta ? 0 : (long long)M * nx;
    a.eta_stride1 = shared_eta ? 0 : (long long)M * M;
    a.eta_stride2 = shared_eta ? 0 : (long long)nx * nx;
    a.rng_mode = rng->mode; a.seed = rng->seed; a.chain_base = rng->chain_base; a.iteration = rng->iteration;
    a.chi2 = rng->chi2; a.G = rng->G; a.Nrm = rng->Nrm;
    a.A = A; a.S = S; a.status = status;
    char* w = (char*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    a.wsB = (double*)w;
    a.wsF = a.wsB + (size_t)n_chains * M * M;
    a.wsK = a.wsF + (size_t)n_chains * M * 2 * nx;
    bool blocked = M >= 192;                        // below, one CTA per chain finishes before the launch sequence would
    if (const char* e = getenv("PGAS_DRAW_BLOCKED")) blocked = atoi(e) != 0;       // developer override
    if (!blocked) {
        mniw_draw_kernel<false><<<n_chains, DT, 0, st>>>(a);
        PGAS_KERNEL_CHECK();
        return 0;
    }
    // augmented matrix (M + 2 n_x rows) in the space of wsB + wsF of the one-CTA form; wsF moves behind it
    const int R2 = 2 * nx;
    a.wsF = a.wsB + (size_t)n_chains * (M + R2) * M;
    a.wsK = a.wsF + (size_t)n_chains * M * R2;
    {
        const unsigned gx = (unsigned)std::min<size_t>(((size_t)(M + R2) * M + 2047) / 2048, 1024);
        chol_prep_kernel<<<dim3(gx, (unsigned)n_chains), 256, 0, st>>>(a);
        PGAS_KERNEL_CHECK();
    }
    const size_t usm = sizeof(double) * 2 * CNB * LDK;
    PGAS_CUDA(cudaFuncSetAttribute(chol_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)usm));
    for (int kb = 0; kb < M; kb += CNB) {
        const int nb = std::min(CNB, M - kb);
        const int below = M + R2 - (kb + nb);
        chol_panel_kernel<<<dim3((unsigned)((below + CPT - 1) / CPT), (unsigned)n_chains), CPT, 0, st>>>(a, kb);
        PGAS_KERNEL_CHECK();
        const int r0 = kb + CNB;
        if (r0 < M) {
            const int ntr = (M + R2 - r0 + CNB - 1) / CNB;
            chol_update_kernel<<<dim3((unsigned)(ntr * (ntr + 1) / 2), (unsigned)n_chains), CPT, usm, st>>>(a, kb);
            PGAS_KERNEL_CHECK();
        }
    }
    mniw_draw_kernel<true><<<n_chains, DT, 0, st>>>(a);
    PGAS_KERNEL_CHECK();
    return 0;
}

extern "C" int pgas_mniw_draw_f64(const double* eta0, const double* eta1, const double* eta2, double eta3, int32_t M, int32_t n_x,
                                  int32_t n_chains, const pgas_rng* rng, int32_t flags, double* A_out, double* S_out,
                                  int32_t* status_out, void* workspace, size_t workspace_bytes, void* stream) {
    return pgas_launch_mniw_draw(eta0, eta1, eta2, eta3, false, M, n_x, n_chains, rng, flags, A_out, S_out, status_out, workspace,
                                 workspace_bytes, (cudaStream_t)stream);
}
