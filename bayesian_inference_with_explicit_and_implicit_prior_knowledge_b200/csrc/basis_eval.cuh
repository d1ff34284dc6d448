// basis_eval.cuh — Hilbert-space GP basis contraction mu = Theta phi(z) for one particle.
//
// Replaces, per particle, the reference's vmap(basis_fcn) + einsum("kj,ij->ik")
// (src/BasisFunctions.py:77-80, src/PGAS.py:52-55 and :67-70) WITHOUT forming phi:
//   phi_m(z) = norm * prod_d sin(pi f_{m,d} t_d),  t_d = (z_d - c_d + L_d) / (2 L_d)
// and the frequencies f sit on a lattice f_start + p*f_step, so
//   mu_k(p) = sum_rows lead[p,row] * sum_j Theta'[row,j,k] * s_last[p,j]
// i.e. a dense (particles x positions) x (positions x rows*n_x) product followed by a row
// scaling — run on the FP64 tensor pipe (DMMA m8n8k4) with the sines built by the 3-term
// recurrence s[j+1] = 2cos(pi f_step t) s[j] - s[j-1] (one sincospi per dimension per particle
// instead of the reference's M*D libm sines).  An earlier sparse DFMA walk over the selected
// lattice points executed ~40 % fewer flops but spent 80 % of its issue slots on control flow
// (profiles/r01_sweep_dfma_walk.md); the dense tile form is branch-free.
#pragma once
#include "common.cuh"

// sin/cos recurrence seed for one dimension: cur = sin(pi f_start t), prev = sin(pi (f_start-f_step) t),
// twoc = 2 cos(pi f_step t)
__device__ __forceinline__ void sine_seed(double t, int f_start, int f_step, double& cur, double& prev, double& twoc) {
    double ss, cs;
    sincospi_bf((double)f_step * t, ss, cs);
    twoc = 2.0 * cs;
    if (f_start == f_step) {
        cur = ss;
        prev = 0.0;
    } else {
        double sa, ca;
        sincospi_bf((double)f_start * t, sa, ca);
        cur = sa;
        prev = sa * cs - ca * ss;
    }
}

#ifndef PGAS_FINE_TICKS
#define PGAS_FINE_TICKS 0
#endif
#if PGAS_FINE_TICKS
__device__ long long g_fine[64];
#define PGAS_FTICK(K) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_fine[K] = clock64(); } while (0)
#else
#define PGAS_FTICK(K) do { } while (0)
#endif

constexpr int QUART = 8;        // particles per shared sine tile (one DMMA fragment row group)
constexpr int TILE_PS = 8;      // tile row pitch in doubles: a 4x8 fragment read covers 256 contiguous bytes
constexpr int NTB = 5;          // column tiles accumulated per pass (accumulators: NTB x 2 doubles)

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// doubles of shared memory one warp needs for its sine tile
__host__ __device__ __forceinline__ int sine_tile_doubles(const DevModel& m) {
    int rows = m.jmax + 4;                                   // +4: the A prefetch of the pipelined loop reads one step ahead
    for (int d = 0; d + 1 < m.D; ++d) rows += (m.npos_d[d] + 3) & ~3;
    return rows * TILE_PS;
}

// Warp-cooperative auxiliary mean for the warp's 32 particles (lane i owns particle il0 + i):
//   mu_k(p) = sum_r lead[p,r] (S B)[p, r n_x + k]      (see common.cuh)
// Four passes of 8 particles.  In a pass all 32 lanes build the sine tables of those 8 particles in
// the warp's shared tile: lane L serves particle 8h + (L & 7), dimension group (L >> 3) & 1 (last /
// leading), position parity L >> 4, each running the 3-term recurrence with stride 2,
//   s[j+2] = (2cos(2 theta)) s[j] - s[j-2],   2cos(2 theta) = (2cos theta)^2 - 2,
// from seeds fetched by shuffle from the owner lane.  The contraction over the last dimension runs
// on the FP64 tensor pipe (mma.sync m8n8k4: A = sines, B = Theta' fragments from shared memory);
// the per-row scaling and a 4-lane reduction finish it.  Row leaders (lane%4 == 0) write mu to
// mus[k*P + particle]; after the closing __syncwarp every owner can read its own entry.
// constants of the contraction, derived from the model ONCE per kernel and kept in registers (reading
// kernel parameters with run-time indices inside the time loop costs issue slots on every step)
template <int D>
struct EvalCtx {
    int NTNP, f_start, f_step;
    int dim_off[D], npos[D];       // tile layout: last dimension first (jmax rows), then the leading dimensions
    __device__ __forceinline__ void init(const DevModel& m) {
        NTNP = m.NTNP; f_start = m.f_start; f_step = m.f_step;
        dim_off[D - 1] = 0;
        int o = m.jmax + 4;
#pragma unroll
        for (int d = 0; d + 1 < D; ++d) { dim_off[d] = o; o += (m.npos_d[d] + 3) & ~3; }
#pragma unroll
        for (int d = 0; d < D; ++d) npos[d] = (d == D - 1) ? m.jmax : ((m.npos_d[d] + 3) & ~3);
    }
};

// kcb[b] = number of position steps whose non-zero column tiles reach into column block b (rows are sorted
// by decreasing length, so the non-zero (step, block) pairs of a block are a prefix of the steps): the
// contraction runs DENSE, branch-free DMMAs over exactly those steps.
template <int NX, int D>
__device__ __forceinline__ void eval_mu_warp(const EvalCtx<D>& cx, const double* __restrict__ bfrag, const int* __restrict__ rowpos,
                                             const int* __restrict__ kcb, double* __restrict__ tile, const double tz[D],
                                             int lane, double* __restrict__ mus, int P, int il0) {
    const int q = lane & 3, r = lane >> 2;
    const int NTNP = cx.NTNP;
    PGAS_FTICK(1);
    double cur0[D], prev0[D], twoc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) sine_seed(tz[d], cx.f_start, cx.f_step, cur0[d], prev0[d], twoc[d]);
    PGAS_FTICK(2);
    const int dsel = (lane >> 3) & 1, par = lane >> 4;
    const int* dim_off = cx.dim_off;
    const int* npos = cx.npos;

    for (int h = 0; h < 4; ++h) {
        // ---- sines of this pass's 8 particles -> tile[dim][pos][8].  One instruction stream for all lanes:
        //      lane group g = (lane>>3)&1 handles dimension D-1-g-2s in slot s (selected by data, not by
        //      branches, so the two groups do not serialise), position parity = lane>>4.
        {
            const int src = QUART * h + (lane & 7);
#pragma unroll
            for (int s = 0; 2 * s < D; ++s) {
                const int d0 = D - 1 - 2 * s, d1 = D - 2 - 2 * s;       // dimension of group 0 / group 1 in this slot
                const int dA = d0, dB = d1 >= 0 ? d1 : d0;
                const bool active = (dsel == 0) || (d1 >= 0);
                // the seeds live in the OWNER lane (src): fetch both groups' dimensions, select by this lane's group
                const double cA = __shfl_sync(0xffffffffu, cur0[dA], src), cB = __shfl_sync(0xffffffffu, cur0[dB], src);
                const double pA = __shfl_sync(0xffffffffu, prev0[dA], src), pB = __shfl_sync(0xffffffffu, prev0[dB], src);
                const double tA = __shfl_sync(0xffffffffu, twoc[dA], src), tB = __shfl_sync(0xffffffffu, twoc[dB], src);
                const double c = dsel ? cB : cA, pv = dsel ? pB : pA, tc = dsel ? tB : tA;
                const int np = active ? (dsel ? npos[dB] : npos[dA]) : 0;
                const int npmax = max(npos[dA], npos[dB]);
                double* col = tile + (size_t)(dsel ? dim_off[dB] : dim_off[dA]) * TILE_PS + (lane & 7) + (size_t)par * TILE_PS;
                const double tc2 = fma(tc, tc, -2.0);
                double cur = par ? fma(tc, c, -pv) : c;            // s[1] : s[0]
                double prev = par ? pv : fma(tc, pv, -c);          // s[-1] : s[-2]
                // two positions of this lane's parity per iteration (np is a multiple of 4: tiles are padded)
                for (int p = 0; p < npmax; p += 4) {
                    const double n1 = fma(tc2, cur, -prev);
                    const double n2 = fma(tc2, n1, -cur);
                    if (p < np) {
                        col[0] = cur;
                        col[2 * TILE_PS] = n1;
                    }
                    col += 4 * TILE_PS;
                    prev = n1; cur = n2;
                }
            }
        }
        __syncwarp();
        PGAS_FTICK(3 + 4 * h);
        double mup[NX];
#pragma unroll
        for (int k = 0; k < NX; ++k) mup[k] = 0.0;

        for (int nb = 0; nb < NTNP; nb += NTB) {
            double acc[NTB][2];
#pragma unroll
            for (int j = 0; j < NTB; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
            // software-pipelined: the fragments and the tile count of position step ks+1 are loaded while the
            // DMMAs of step ks issue (the count table is padded by one entry, the B array by one step of zeros)
            const double* ap = tile + (size_t)q * TILE_PS + r;
            const double* bp = bfrag + (size_t)nb * 32 + lane;
            const size_t bstep = (size_t)NTNP * 32;
            double a0 = ap[0], bv[NTB];
#pragma unroll
            for (int j = 0; j < NTB; ++j) bv[j] = bp[j * 32];
            const int kc = kcb[nb / NTB];
            for (int ks = 0; ks < kc; ++ks) {
                ap += 4 * TILE_PS;
                bp += bstep;
                const double a_n = ap[0];
                double b_n[NTB];
#pragma unroll
                for (int j = 0; j < NTB; ++j) b_n[j] = bp[j * 32];
#pragma unroll
                for (int j = 0; j < NTB; ++j) dmma_m8n8k4(acc[j][0], acc[j][1], a0, bv[j]);
                a0 = a_n;
#pragma unroll
                for (int j = 0; j < NTB; ++j) bv[j] = b_n[j];
            }
            PGAS_FTICK(4 + 4 * h);
            // ---- per-row scaling by the leading-dimension sines (padding rows have zero accumulators)
#pragma unroll
            for (int j = 0; j < NTB; ++j) {
                const int c0 = 8 * (nb + j) + 2 * q;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (NX == 2 && e == 1) break;               // both columns of the pair belong to one row
                    const int row = (c0 + e) / NX;
                    double lead = 1.0;
#pragma unroll
                    for (int d = 0; d + 1 < D; ++d) lead *= tile[(size_t)(dim_off[d] + rowpos[row * MAX_LEAD + d]) * TILE_PS + r];
                    if constexpr (NX == 2) {
                        mup[0] = fma(lead, acc[j][0], mup[0]);
                        mup[1] = fma(lead, acc[j][1], mup[1]);
                    } else {
                        const int k = (c0 + e) - row * NX;
#pragma unroll
                        for (int kk = 0; kk < NX; ++kk) mup[kk] = fma((kk == k) ? lead : 0.0, acc[j][e], mup[kk]);
                    }
                }
            }
        }
        PGAS_FTICK(5 + 4 * h);
        // ---- reduce over the 4 lanes of a fragment row, row leaders publish
#pragma unroll
        for (int k = 0; k < NX; ++k) {
            double v = mup[k];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            const int il = il0 + QUART * h + r;
            if (q == 0 && il < P) mus[(size_t)k * P + il] = v;
        }
        __syncwarp();
        PGAS_FTICK(6 + 4 * h);
    }
}

// ---------------------------------------------------------------------------------- expression programs (model plug-in)
// The GP-input map of a user callable outside the affine / slip-angle families, traced on the host into a postfix program
// (include/pgas_b200.h: PGAS_MAP_PROGRAM; models.py: Sym).  Interpreted per particle: the program is the same for every
// thread (uniform control flow), the operand stack lives in local memory.  Kept out of line so that the compiled-in families
// do not pay registers for it.
static __device__ __noinline__ void pgas_map_program(const DevModel* m, const double* x, const double* u, double* z) {
    double st[PGAS_PROG_STACK];
    int sp = 0;
    const int n = m->prog_len;
    for (int pc = 0; pc < n; ++pc) {
        const int ins = m->prog_op[pc], op = ins & 0xff, arg = ins >> 8;
        if (op <= PGAS_OP_PUSH_C) {
            st[sp & (PGAS_PROG_STACK - 1)] = (op == PGAS_OP_PUSH_X) ? x[arg] : (op == PGAS_OP_PUSH_U) ? u[arg] : m->prog_const[arg];
            ++sp;
        } else if (op <= PGAS_OP_DIV || op >= PGAS_OP_POW) {
            const double b = st[(--sp) & (PGAS_PROG_STACK - 1)], a = st[(sp - 1) & (PGAS_PROG_STACK - 1)];
            double r;
            switch (op) {
                case PGAS_OP_ADD: r = a + b; break;
                case PGAS_OP_SUB: r = a - b; break;
                case PGAS_OP_MUL: r = a * b; break;
                case PGAS_OP_DIV: r = a / b; break;
                case PGAS_OP_POW: r = pow(a, b); break;
                default: r = atan2(a, b); break;
            }
            st[(sp - 1) & (PGAS_PROG_STACK - 1)] = r;
        } else {
            const double a = st[(sp - 1) & (PGAS_PROG_STACK - 1)];
            double r;
            switch (op) {
                case PGAS_OP_NEG: r = -a; break;
                case PGAS_OP_SIN: r = sin(a); break;
                case PGAS_OP_COS: r = cos(a); break;
                case PGAS_OP_TAN: r = tan(a); break;
                case PGAS_OP_TANH: r = tanh(a); break;
                case PGAS_OP_ATAN: r = atan(a); break;
                case PGAS_OP_EXP: r = exp(a); break;
                case PGAS_OP_LOG: r = log(a); break;
                case PGAS_OP_SQRT: r = sqrt(a); break;
                default: r = fabs(a); break;
            }
            st[(sp - 1) & (PGAS_PROG_STACK - 1)] = r;
        }
    }
    for (int d = 0; d < m->D; ++d) z[d] = st[d];
}

// likelihood_fcn as an expression program (model plug-in; pgas_b200.h: lik_prog_*): log-density of observation y given state x and
// input u.  A SEPARATE interpreter on purpose: the GP-input interpreter above is reachable from the state kernel, whose register
// allocation at the 128-register cap reacts to every change of that call site (a shared, more general interpreter cost the state
// kernel 0.7 % at M = 256 and 3 % at M = 1024); this one is only reachable from the fused sweep kernel.
static __device__ __noinline__ double pgas_lik_program(const DevModel* m, const double* x, const double* u, const double* y) {
    double st[PGAS_PROG_STACK];
    int sp = 0;
    const int pc0 = m->lik_off, n = m->lik_len, coff = m->lik_coff;
    for (int pc = pc0; pc < pc0 + n; ++pc) {
        const int ins = m->prog_op[pc], op = ins & 0xff, arg = ins >> 8;
        if (op <= PGAS_OP_PUSH_C || op == PGAS_OP_PUSH_Y) {
            st[sp & (PGAS_PROG_STACK - 1)] = (op == PGAS_OP_PUSH_X) ? x[arg] : (op == PGAS_OP_PUSH_U) ? u[arg] : (op == PGAS_OP_PUSH_Y) ? y[arg]
                                                                                                                : m->prog_const[coff + arg];
            ++sp;
        } else if (op <= PGAS_OP_DIV || op >= PGAS_OP_POW) {
            const double b = st[(--sp) & (PGAS_PROG_STACK - 1)], a = st[(sp - 1) & (PGAS_PROG_STACK - 1)];
            double r;
            switch (op) {
                case PGAS_OP_ADD: r = a + b; break;
                case PGAS_OP_SUB: r = a - b; break;
                case PGAS_OP_MUL: r = a * b; break;
                case PGAS_OP_DIV: r = a / b; break;
                case PGAS_OP_POW: r = pow(a, b); break;
                default: r = atan2(a, b); break;
            }
            st[(sp - 1) & (PGAS_PROG_STACK - 1)] = r;
        } else {
            const double a = st[(sp - 1) & (PGAS_PROG_STACK - 1)];
            double r;
            switch (op) {
                case PGAS_OP_NEG: r = -a; break;
                case PGAS_OP_SIN: r = sin(a); break;
                case PGAS_OP_COS: r = cos(a); break;
                case PGAS_OP_TAN: r = tan(a); break;
                case PGAS_OP_TANH: r = tanh(a); break;
                case PGAS_OP_ATAN: r = atan(a); break;
                case PGAS_OP_EXP: r = exp(a); break;
                case PGAS_OP_LOG: r = log(a); break;
                case PGAS_OP_SQRT: r = sqrt(a); break;
                default: r = fabs(a); break;
            }
            st[(sp - 1) & (PGAS_PROG_STACK - 1)] = r;
        }
    }
    return st[0];
}

// GP-input map z = g(state, input) of any family on zero-padded arrays x[PGAS_MAX_NX], u[PGAS_MAX_NU] -> z[PGAS_MAX_D]
// (src/PGAS.py:52-54: basis_fcn(state, inputs[t])); used where the map is not on a hot path (statistics, basis evaluation)
__device__ __forceinline__ void gp_map_any(const DevModel& m, const double* x, const double* u, double* z) {
    if (m.map_kind == PGAS_MAP_PROGRAM) {
        pgas_map_program(&m, x, u, z);
    } else if (m.map_kind == PGAS_MAP_VEHICLE_SLIP) {
        // src/Vehicle.py:50-57: alpha_f = delta - atan((v_y + psi_dot l_f)/v_x), alpha_r = -atan((v_y - psi_dot l_r)/v_x)
        z[0] = u[0] - atan((x[1] + x[0] * m.slip_lf) / u[1]);
        z[1] = -atan((x[1] - x[0] * m.slip_lr) / u[1]);
        z[2] = 0.0;
    } else {
        for (int d = 0; d < m.D; ++d) {
            double acc = m.bz[d];
            for (int k = 0; k < m.n_x; ++k) acc = fma(m.Az[d][k], x[k], acc);
            for (int k = 0; k < m.n_u; ++k) acc = fma(m.Az[d][m.n_x + k], u[k], acc);
            z[d] = acc;
        }
    }
}


// GP-input map with the model constants hoisted into registers once per kernel (kernel-parameter reads
// with run-time indices cost ~100 cycles each on the critical path of every step)
template <int NX, int D>
struct MapRegs {
    double A[D][NX], off[D], sc[D], lf, lr;
    int kind;
    const DevModel* mp;
    __device__ __forceinline__ void init(const DevModel& m) {
        kind = m.map_kind; lf = m.slip_lf; lr = m.slip_lr; mp = &m;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            off[d] = m.L[d] - m.center[d];
            sc[d] = m.inv2L[d];
#pragma unroll
            for (int k = 0; k < NX; ++k) A[d][k] = m.Az[d][k];
        }
    }
    // cz[d] = bz[d] + sum_k Az[d][n_x+k] u[k] is the per-step constant part (StepConst)
    __device__ __forceinline__ void apply(const double x[NX], const double* __restrict__ cz, const double* __restrict__ u, double tz[D]) const {
        double z[D];
        if (kind == PGAS_MAP_PROGRAM) {                     // model plug-in: temporaries keep the address-taking out of the other paths
            double xl[PGAS_MAX_NX], ul[PGAS_MAX_NU], zl[PGAS_MAX_D];
#pragma unroll
            for (int k = 0; k < PGAS_MAX_NX; ++k) xl[k] = (k < NX) ? x[k < NX ? k : 0] : 0.0;
#pragma unroll
            for (int k = 0; k < PGAS_MAX_NU; ++k) ul[k] = u[k];
            pgas_map_program(mp, xl, ul, zl);
#pragma unroll
            for (int d = 0; d < D; ++d) z[d] = zl[d];
        } else if (kind == PGAS_MAP_VEHICLE_SLIP) {
            const double x0 = x[0], x1 = x[NX > 1 ? 1 : 0];
            z[0] = u[0] - atan((x1 + x0 * lf) / u[1]);
            if constexpr (D >= 2) z[1] = -atan((x1 - x0 * lr) / u[1]);
            if constexpr (D >= 3) z[2] = 0.0;
        } else {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                double acc = cz[d];
#pragma unroll
                for (int k = 0; k < NX; ++k) acc = fma(A[d][k], x[k], acc);
                z[d] = acc;
            }
        }
#pragma unroll
        for (int d = 0; d < D; ++d) tz[d] = (z[d] + off[d]) * sc[d];
    }
};

// GP-input map (state, input) -> normalised lattice coordinate t_d = (z_d - c_d + L_d)/(2 L_d)
template <int NX, int D>
__device__ __forceinline__ void gp_input(const DevModel& m, const double x[NX], const double* __restrict__ u, double tz[D]) {
    double z[D];
    if (m.map_kind == PGAS_MAP_PROGRAM) {
        double xl[PGAS_MAX_NX], ul[PGAS_MAX_NU], zl[PGAS_MAX_D];
#pragma unroll
        for (int k = 0; k < PGAS_MAX_NX; ++k) xl[k] = (k < NX) ? x[k < NX ? k : 0] : 0.0;
#pragma unroll
        for (int k = 0; k < PGAS_MAX_NU; ++k) ul[k] = (k < m.n_u) ? u[k] : 0.0;
        pgas_map_program(&m, xl, ul, zl);
#pragma unroll
        for (int d = 0; d < D; ++d) z[d] = zl[d];
    } else if (m.map_kind == PGAS_MAP_VEHICLE_SLIP) {
        // src/Vehicle.py:50-57: alpha_f = delta - atan((v_y + psi_dot l_f)/v_x), alpha_r = -atan((v_y - psi_dot l_r)/v_x)
        const double x0 = x[0], x1 = (NX > 1) ? x[NX > 1 ? 1 : 0] : 0.0;
        const double af = u[0] - atan((x1 + x0 * m.slip_lf) / u[1]);
        const double ar = -atan((x1 - x0 * m.slip_lr) / u[1]);
        z[0] = af;
        if constexpr (D >= 2) z[1] = ar;
        if constexpr (D >= 3) z[2] = 0.0;
    } else {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            double acc = m.bz[d];
#pragma unroll
            for (int k = 0; k < NX; ++k) acc = fma(m.Az[d][k], x[k], acc);
            for (int k = 0; k < m.n_u; ++k) acc = fma(m.Az[d][NX + k], u[k], acc);
            z[d] = acc;
        }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) tz[d] = (z[d] - m.center[d] + m.L[d]) * m.inv2L[d];
}
