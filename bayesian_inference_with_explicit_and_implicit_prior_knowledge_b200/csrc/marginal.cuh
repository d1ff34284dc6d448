// marginal.cuh — device model, workspace layout and argument block of the marginalised particle
// filters (marginal.cu): Algorithm1 / Algorithm3 / Algorithm2 of the reference
// (src/Algorithm1.py, src/Algorithm3.py, src/Algorithm2.py), SURVEY.md 8a group B.
#pragma once
#include "common.cuh"

constexpr int MG_GP = PGAS_MAX_GP;
constexpr int MG_NX = PGAS_MAX_NX;
constexpr int MG_NY = PGAS_MAX_NY;
constexpr int MG_D = PGAS_MAX_D;
constexpr int MG_MAX_M = 128;          // 4 register rows per lane in the triangular solve

enum { PURPOSE_XI0 = 5, PURPOSE_TVAR = 6 /* + GP index */ };

// Model plug-in (pgas_b200.h: pgas_marg_program): a postfix expression program in the model's arena; len = 0: use the tables
struct MargProg {
    const int* ops;
    const double* consts;
    int len, n_out;
};

struct MargGP {
    int M, D, link, npk;               // npk = M (M + 1) / 2 packed lower-triangular entries
    double center[MG_D], L[MG_D], sqrt_invL[MG_D];
    const double* sqrt_eig;            // (M, D)
    const double* gp_in;               // (T, D, n_x + 1)
    const double* gp_post;             // (T, D, 2)
    const double* p0;                  // (M)    prior eta0
    const double* p1;                  // (npk)  prior eta1, packed lower triangle of its symmetric part
    double p2, p3;
    double xi_mean, xi_sd;
    MargProg prog;                     // GP-input map as a program over (state, inputs[t])
};

struct MargDev {
    int n_x, n_y, G, T, out_link, deterministic;
    MargGP gp[MG_GP];
    const double* trans;               // (T, n_x, n_x + G + 1)
    const double* outp;                // (T, n_y, n_x + G + 1)
    const double* obs;                 // (T, n_y)
    double Qc[MG_NX][MG_NX];           // chol(Q) lower (zero when deterministic)
    double Qw[MG_NX][MG_NX];           // inverse of chol(Q): e = Qw (x - mean)
    double Q_logc;                     // -n_x/2 log(2 pi) - sum log diag chol(Q)
    double Rw[MG_NY][MG_NY], R_logc;
    double m0[MG_NX], P0c[MG_NX][MG_NX];
    // model plug-in: any_prog selects the interpreting instantiation of the sweep kernel
    int n_u, any_prog;
    const double* inputs;              // (T, n_u)
    MargProg tprog, oprog;             // transition / output model over ([state; xi], inputs[t])
};

struct pgas_marg_model {
    MargDev dev;
    void* arena;
    size_t arena_bytes;
};

// per-particle strides (doubles), padded to even so that 16-byte asynchronous copies stay aligned
__host__ __device__ inline int mg_npkp(int M) { return (M * (M + 1) / 2 + 1) & ~1; }          // packed M x M lower triangle
__host__ __device__ inline int mg_naugp(int M) { return (M * (M + 1) / 2 + M + 1 + 1) & ~1; }  // augmented (M+1) x (M+1) factor

// Per-chain, per-parity block of the per-particle workspace (element offsets in doubles).  Two
// parities: pass t writes parity t & 1 and gathers, by ancestor, from parity (t - 1) & 1.
struct MargWs {
    size_t auxx, ellaux, lwaux, lwanc;
    size_t yv[MG_GP], psi[MG_GP], Lp[MG_GP], LB[MG_GP], T1p[MG_GP], T0[MG_GP], T2[MG_GP], T3[MG_GP];
    size_t parity_stride, chain_stride;
};

__host__ __device__ inline MargWs marg_ws_layout(const MargDev& m, int N) {
    MargWs w;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~(size_t)3; return r; };
    w.auxx = take((size_t)N * m.n_x);
    w.ellaux = take(N);
    w.lwaux = take(N);
    w.lwanc = take(N);
    for (int g = 0; g < MG_GP; ++g) {
        const bool on = g < m.G;
        const size_t M = on ? m.gp[g].M : 0, npk = on ? m.gp[g].npk : 0;
        w.yv[g] = take((size_t)N * M);
        w.psi[g] = take(on ? N : 0);
        // Lp: Algorithm1 keeps the M x M factor (inverse diagonal) here; Algorithm3 keeps the AUGMENTED factor of
        // [[eta1, eta0], [eta0^T, eta2]] (M+1 rows, true diagonal) of prior + statistics, and LB the one of
        // prior + remaining reference statistics + statistics (rank-1 up/down-dated, marginal.cu)
        w.Lp[g] = take((size_t)N * (on ? mg_naugp((int)M) : 0));
        w.LB[g] = take((size_t)N * (on ? mg_naugp((int)M) : 0));
        w.T1p[g] = take((size_t)N * (on ? mg_npkp((int)M) : 0));
        w.T0[g] = take((size_t)N * M);
        w.T2[g] = take(on ? N : 0);
        w.T3[g] = take(on ? N : 0);
    }
    w.parity_stride = o;
    w.chain_stride = 2 * o;
    return w;
}

// prior + remaining reference statistics after step t (src/Algorithm3.py:235-246, :163-174), one row
// per time step: PR1 packed (T, npk), PR0 (T, M), PR2 (T), PR3 (T); per chain.
struct MargRefTab {
    const double* RPHI[MG_GP];         // (T, M) basis of the reference trajectory at every step
    const double* PR0[MG_GP];
    const double* PR1[MG_GP];
    const double* PR2[MG_GP];
    const double* PR3[MG_GP];
};

struct MargArgs {
    MargDev m;
    int N, n_chains, CS, NW;           // CTAs per chain (a hardware cluster, or a cooperative group of CTAs), warps per CTA
    int sw_barrier;                    // 1: CTAs of a chain synchronise through a global counter (cooperative launch, CS up to 64)
    unsigned* bar_ctr;                 // (n_chains) counters of the software barrier, zero on entry
    int mode;                          // 0 = Algorithm1 (filter), 1 = Algorithm3 (conditional)
    double lambda;                     // forgetting factor (1 in mode 1)
    size_t warp_doubles, cta_doubles;  // shared-memory carve-up (warp_doubles depends on the mode)
    // reference trajectory (mode 1)
    const double* ref_x;  long long ref_x_stride;                       // (n_chains, T, n_x)
    const double* ref_xi; long long ref_xi_stride, ref_xi_gstride;      // xi[c][g][t]
    MargRefTab tab;                    // chain stride = T * (row length)
    // traces
    double* state_trace;               // (n_chains, T, N, n_x)
    double* xi_trace;                  // (n_chains, G, T, N)
    double* logw_trace;                // (n_chains, T, N)
    int* anc_trace;                    // (n_chains, T-1, N)
    double* sst[4 * MG_GP];            // optional weighted statistics trace (mode 0)
    double* ws;                        // per-particle workspace (n_chains * chain_stride doubles)
    int* status;                       // (n_chains)
    // variates
    int rng_mode;
    unsigned long long seed;
    unsigned chain_base, iteration;
    const double *Z, *ZXI0, *U, *TS;
};
