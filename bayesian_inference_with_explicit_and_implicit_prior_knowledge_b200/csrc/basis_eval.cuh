// basis_eval.cuh — Hilbert-space GP basis contraction mu = Theta phi(z) for one particle.
//
// Replaces, per particle, the reference's vmap(basis_fcn) + einsum("kj,ij->ik")
// (src/BasisFunctions.py:77-80, src/PGAS.py:52-55 and :67-70) WITHOUT forming phi:
//   phi_m(z) = norm * prod_d sin(pi f_{m,d} t_d),  t_d = (z_d - c_d + L_d) / (2 L_d)
// and the frequencies f sit on a lattice f_start + p*f_step, so
//   mu_k = norm * sum_rows lead(row) * sum_{p < len(row)} Theta'[row,p,k] * s_last[p]
// with s_last[p] = sin(pi (f_start + p f_step) t_last) kept in REGISTERS (statically indexed,
// built by the 3-term recurrence s[p+1] = 2cos(pi f_step t) s[p] - s[p-1]) and the leading
// dimensions' sines carried as running recurrences across rows (common.cuh).  Rows are walked
// as a flat list of 4-position chunks whose Theta' values are prefetched one chunk ahead, so
// the shared-memory latency is off the DFMA dependency chain.
// Cost per particle: n_packed*NX DFMA + JMAX + ~6 per row, versus M*(D + 2 NX) flops for the
// phi-then-einsum formulation and M*D libm sines in the reference.
#pragma once
#include "common.cuh"

// sin/cos recurrence seed for one dimension: cur = sin(pi f_start t), prev = sin(pi (f_start-f_step) t),
// twoc = 2 cos(pi f_step t)
__device__ __forceinline__ void sine_seed(double t, int f_start, int f_step, double& cur, double& prev, double& twoc) {
    double ss, cs;
    sincospi((double)f_step * t, &ss, &cs);
    twoc = 2.0 * cs;
    if (f_start == f_step) {
        cur = ss;
        prev = 0.0;
    } else {
        double sa, ca;
        sincospi((double)f_start * t, &sa, &ca);
        cur = sa;
        prev = sa * cs - ca * ss;
    }
}

// Theta' of one chunk: CHUNK positions x NX outputs, contiguous in shared memory, all lanes read
// the same address (broadcast).
template <int NX>
struct ThetaChunk {
    double v[CHUNK][NX];
    __device__ __forceinline__ void load(const double* __restrict__ th, int c) {
        constexpr int ND = CHUNK * NX;
        const double* p = th + (size_t)c * ND;
        if constexpr (ND % 2 == 0) {
            const double2* p2 = reinterpret_cast<const double2*>(p);
#pragma unroll
            for (int i = 0; i < ND / 2; ++i) {
                const double2 t = p2[i];
                v[(2 * i) / NX][(2 * i) % NX] = t.x;
                v[(2 * i + 1) / NX][(2 * i + 1) % NX] = t.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < ND; ++i) v[i / NX][i % NX] = p[i];
        }
    }
};

// 2*CHUNK*NX... the chunk's FMAs against the statically indexed sine block B (positions 4B..4B+3);
// positions alternate between two accumulator sets -> 2*NX independent DFMA chains.
#define PGAS_CHUNK_CASE(B)                                                                   \
    case (B):                                                                                \
        if constexpr ((B) * CHUNK < JMAX) {                                                  \
            _Pragma("unroll") for (int i = 0; i < CHUNK; ++i) {                              \
                _Pragma("unroll") for (int k = 0; k < NX; ++k) {                             \
                    if (i & 1) acc1[k] = fma(tc.v[i][k], s[(B) * CHUNK + i], acc1[k]);       \
                    else       acc0[k] = fma(tc.v[i][k], s[(B) * CHUNK + i], acc0[k]);       \
                }                                                                            \
            }                                                                                \
        }                                                                                    \
        break;

// th: shared-memory Theta' in chunk order [chunk][CHUNK][NX] (already scaled by norm);
// meta: shared-memory chunk metadata (common.cuh).
template <int NX, int D, int JMAX>
__device__ __forceinline__ void eval_mu(const double* __restrict__ th, const int* __restrict__ meta, int n_chunks, int f_start,
                                        int f_step, const double tz[D], double mu[NX]) {
    static_assert(JMAX % CHUNK == 0 && JMAX <= 40, "JMAX must be a multiple of CHUNK and <= 40");
    double s[JMAX];
    {
        double cur, prev, twoc;
        sine_seed(tz[D - 1], f_start, f_step, cur, prev, twoc);
        s[0] = cur;
        s[1] = fma(twoc, cur, -prev);
#pragma unroll
        for (int p = 2; p < JMAX; ++p) s[p] = fma(twoc, s[p - 1], -s[p - 2]);
    }
    // leading dimensions: fast = D-2, slow = D-3
    double fcur = 1.0, fprev = 0.0, ftwoc = 0.0, fcur0 = 1.0, fprev0 = 0.0;
    double scur = 1.0, sprev = 0.0, stwoc = 0.0;
    if constexpr (D >= 2) {
        sine_seed(tz[D - 2], f_start, f_step, fcur, fprev, ftwoc);
        fcur0 = fcur; fprev0 = fprev;
    }
    if constexpr (D >= 3) sine_seed(tz[D - 3], f_start, f_step, scur, sprev, stwoc);
    double acc0[NX], acc1[NX];
#pragma unroll
    for (int k = 0; k < NX; ++k) { mu[k] = 0.0; acc0[k] = 0.0; acc1[k] = 0.0; }

    // one chunk: FMAs against its sine block, then (row end) fold into mu and step the leading sines
    auto body = [&](const ThetaChunk<NX>& tc, int mt) {
        switch (mt & 0xff) {
            PGAS_CHUNK_CASE(0) PGAS_CHUNK_CASE(1) PGAS_CHUNK_CASE(2) PGAS_CHUNK_CASE(3) PGAS_CHUNK_CASE(4)
            PGAS_CHUNK_CASE(5) PGAS_CHUNK_CASE(6) PGAS_CHUNK_CASE(7) PGAS_CHUNK_CASE(8) PGAS_CHUNK_CASE(9)
            default: break;
        }
        if (mt & META_ROW_END) {
            double lead = 1.0;
            if constexpr (D == 2) lead = fcur;
            if constexpr (D >= 3) lead = fcur * scur;
#pragma unroll
            for (int k = 0; k < NX; ++k) {
                mu[k] = fma(lead, acc0[k] + acc1[k], mu[k]);
                acc0[k] = 0.0;
                acc1[k] = 0.0;
            }
            if constexpr (D >= 2) {
                const int adv = (mt >> META_ADV_SHIFT) & 3;
                if (adv == ROW_ADV_FAST) {
                    const double n = fma(ftwoc, fcur, -fprev);
                    fprev = fcur; fcur = n;
                } else if (adv == ROW_ADV_SLOW) {
                    if constexpr (D >= 3) {
                        const double n = fma(stwoc, scur, -sprev);
                        sprev = scur; scur = n;
                    }
                    fcur = fcur0; fprev = fprev0;
                }
            }
        }
    };

    // software-pipelined walk over the chunk list: chunk c+1's Theta' is loaded while chunk c is
    // multiplied (two register buffers, loop unrolled by two so no register copies are needed)
    ThetaChunk<NX> ta, tb;
    int ma, mb = 0;
    ta.load(th, 0);
    ma = meta[0];
    int c = 0;
    for (; c + 1 < n_chunks; c += 2) {
        tb.load(th, c + 1);
        mb = meta[c + 1];
        body(ta, ma);
        if (c + 2 < n_chunks) {
            ta.load(th, c + 2);
            ma = meta[c + 2];
        }
        body(tb, mb);
    }
    if (c < n_chunks) body(ta, ma);
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
