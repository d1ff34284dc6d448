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

constexpr int QUART = 8;        // particles per shared sine tile (one DMMA fragment row group)
constexpr int TILE_PS = 8;      // tile row pitch in doubles: a 4x8 fragment read covers 256 contiguous bytes
constexpr int NTB = 5;          // column tiles accumulated per pass (accumulators: NTB x 2 doubles)

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// doubles of shared memory one warp needs for its sine tile
__host__ __device__ __forceinline__ int sine_tile_doubles(const DevModel& m) {
    int rows = m.jmax;
    for (int d = 0; d + 1 < m.D; ++d) rows += m.npos_d[d];
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
template <int NX, int D>
__device__ __forceinline__ void eval_mu_warp(const DevModel& m, const double* __restrict__ bfrag, const int* __restrict__ rowpos,
                                             double* __restrict__ tile, const double tz[D],
                                             int lane, double* __restrict__ mus, int P, int il0) {
    const int q = lane & 3, r = lane >> 2;
    const int KS = m.KS, NTNP = m.NTNP;
    // tile layout: last dimension first (jmax rows), then the leading dimensions
    int dim_off[D];
    dim_off[D - 1] = 0;
    {
        int o = m.jmax;
#pragma unroll
        for (int d = 0; d + 1 < D; ++d) { dim_off[d] = o; o += m.npos_d[d]; }
    }
    double cur0[D], prev0[D], twoc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) sine_seed(tz[d], m.f_start, m.f_step, cur0[d], prev0[d], twoc[d]);
    const int dsel = (lane >> 3) & 1, par = lane >> 4;

    for (int h = 0; h < 4; ++h) {
        // ---- sines of this pass's 8 particles -> tile[dim][pos][8]
        {
            const int src = QUART * h + (lane & 7);
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const double c = __shfl_sync(0xffffffffu, cur0[d], src), pv = __shfl_sync(0xffffffffu, prev0[d], src);
                const double tc = __shfl_sync(0xffffffffu, twoc[d], src);
                const bool mine = (D == 1) ? (dsel == 0) : ((d == D - 1) == (dsel == 0));
                if (mine) {
                    const int np = (d == D - 1) ? m.jmax : m.npos_d[d];
                    double* col = tile + (size_t)dim_off[d] * TILE_PS + (lane & 7);
                    const double tc2 = fma(tc, tc, -2.0);
                    double cur = par ? fma(tc, c, -pv) : c;            // s[1] : s[0]
                    double prev = par ? pv : fma(tc, pv, -c);          // s[-1] : s[-2]
                    for (int p = par; p < np; p += 2) {
                        col[(size_t)p * TILE_PS] = cur;
                        const double n = fma(tc2, cur, -prev);
                        prev = cur; cur = n;
                    }
                }
            }
        }
        __syncwarp();
        double mup[NX];
#pragma unroll
        for (int k = 0; k < NX; ++k) mup[k] = 0.0;

        for (int nb = 0; nb < NTNP; nb += NTB) {
            double acc[NTB][2];
#pragma unroll
            for (int j = 0; j < NTB; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
            for (int ks = 0; ks < KS; ++ks) {
                const int ntc = m.ntcount[ks] - nb;           // kernel-parameter data: warp-uniform
                if (ntc <= 0) continue;
                const double a0 = tile[(size_t)(4 * ks + q) * TILE_PS + r];
                const double* bp = bfrag + ((size_t)ks * NTNP + nb) * 32 + lane;
                double bv[NTB];
#pragma unroll
                for (int j = 0; j < NTB; ++j) bv[j] = bp[j * 32];   // tiles past ntcount hold zeros: always loadable
                switch (ntc) {                                   // warp-uniform: straight-line DMMAs, no predication
                    default: dmma_m8n8k4(acc[4][0], acc[4][1], a0, bv[4]);
                    case 4: dmma_m8n8k4(acc[3][0], acc[3][1], a0, bv[3]);
                    case 3: dmma_m8n8k4(acc[2][0], acc[2][1], a0, bv[2]);
                    case 2: dmma_m8n8k4(acc[1][0], acc[1][1], a0, bv[1]);
                    case 1: dmma_m8n8k4(acc[0][0], acc[0][1], a0, bv[0]);
                }
            }
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
    }
}

// GP-input map (state, input) -> normalised lattice coordinate t_d = (z_d - c_d + L_d)/(2 L_d)
template <int NX, int D>
__device__ __forceinline__ void gp_input(const DevModel& m, const double x[NX], const double* __restrict__ u, double tz[D]) {
    double z[D];
    if (m.map_kind == PGAS_MAP_VEHICLE_SLIP) {
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
