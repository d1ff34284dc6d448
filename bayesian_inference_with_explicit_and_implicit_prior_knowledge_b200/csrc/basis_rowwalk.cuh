// basis_rowwalk.cuh — mu = Theta phi(z) for two- and three-dimensional Hilbert bases, one or two particles per thread, on the
// FP64 FMA pipe.
//
// Replaces vmap(basis_fcn) + einsum("kj,ij->ik") (reference src/BasisFunctions.py:77-80, src/PGAS.py:52-55,
// :67-70) without forming phi:
//   mu_k = sum_i s0_i * ( sum_j Theta'[k, m(i,j)] s1_j ),   s0_i = sin(pi f_i t_0), s1_j = sin(pi f_j t_1)
// with both sine families produced by the three-term recurrence (one sincospi per dimension and particle
// instead of M*D libm sines).  A thread owns PP particles and walks the lattice in blocks of RW_RB first-
// dimension positions (common.cuh): every Theta' pair is fetched ONCE per warp with a broadcast LDS.128 and
// feeds PP * n_x independent DFMAs — no shared-memory staging of sines, no shuffles, no warp-level
// synchronisation.  On B200 the FP64 tensor pipe (DMMA m8n8k4: 512 flops / 16 pipe cycles) and the FP64
// FMA pipe (64 flops / 2 cycles) have the same flop rate; the tile form (basis_eval.cuh) spends its issue
// slots on fragment traffic, this form spends them on the FMAs themselves and is bound by the FP64 pipe
// (profiles/r01_sweep_history.md).
#pragma once
#include "basis_eval.cuh"

#ifndef PGAS_RW_SHORT_ROLLED
#define PGAS_RW_SHORT_ROLLED 0  // 1: do not unroll the short staircase segments (compile-time knob)
#endif
#ifndef PGAS_RW_UNROLL
#define PGAS_RW_UNROLL 2        // positions per loop body of the row walk (compile-time knob)
#endif
// One Theta' pair = the n_x coefficients of one (row, position): a 16-byte broadcast LDS.128 for n_x = 2.
template <int NX>
__device__ __forceinline__ void rw_load(const double* __restrict__ p, double (&w)[NX]) {
    if constexpr (NX == 2) {
        const double2 v = *reinterpret_cast<const double2*>(p);
        w[0] = v.x; w[1] = v.y;
    } else {
#pragma unroll
        for (int k = 0; k < NX; ++k) w[k] = p[k];
    }
}

// One walked position with the first R rows of the block active: R Theta' pairs, each feeding PP * NX DFMAs with the position's sine
// `sn`.  The pairs travel through a register ring `w`: on entry w[i] holds the pair of THIS position; a pair is re-loaded for the
// NEXT position (`nx`) right after its DFMAs, i.e. (R-1) pairs = 4 (R-1) PP DFMAs ahead of its use, so that the ~25-cycle LDS
// latency is covered inside ONE warp (ptxas left one pair of look-ahead: a warp walking alone reached a quarter of the pipe rate,
// profiles/r02_state_kernel_summary.md).  FIRST: the products start the accumulators (no zeroing).
template <int NX, int PP, int R, bool FIRST = false>
__device__ __forceinline__ void rw_position(const double* __restrict__ nx, double (&acc)[PP][RW_RB][NX], const double (&sn)[PP], double (&w)[RW_RB][NX]) {
#pragma unroll
    for (int i = 0; i < R; ++i) {
#pragma unroll
        for (int p = 0; p < PP; ++p)
#pragma unroll
            for (int k = 0; k < NX; ++k) acc[p][i][k] = FIRST ? w[i][k] * sn[p] : fma(w[i][k], sn[p], acc[p][i][k]);
        rw_load<NX>(nx + i * NX, w[i]);
    }
}

// n positions with R active rows.  Positions go in pairs with the sine recurrence in place (c, pv alternate as the current value:
// no register rotation) and the loop bounded by the read pointer (one uniform add, one compare, one branch per pair); an odd
// position at the end rotates.  On exit w[i] holds the pair at the new th[i] for i < R, which is what the next segment (R - 1
// rows) expects; the over-read past the last block touches the bytes that follow Theta' in shared memory.
template <int NX, int PP, int R, bool EVEN = false>
__device__ __forceinline__ void rw_segment(const double* __restrict__& th, int n, double (&acc)[PP][RW_RB][NX], double (&c)[PP],
                                           double (&pv)[PP], const double (&b_2c)[PP], double (&w)[RW_RB][NX]) {
    const double* __restrict__ q = th;
    const double* const e2 = q + (n >> 1) * (2 * R * NX);
#pragma unroll 1
    while (q != e2) {
        rw_position<NX, PP, R>(q + R * NX, acc, c, w);
#pragma unroll
        for (int p = 0; p < PP; ++p) pv[p] = fma(b_2c[p], c[p], -pv[p]);
        rw_position<NX, PP, R>(q + 2 * R * NX, acc, pv, w);
#pragma unroll
        for (int p = 0; p < PP; ++p) c[p] = fma(b_2c[p], pv[p], -c[p]);
        q += 2 * R * NX;
    }
    if (!EVEN && (n & 1)) {
        q += R * NX;
        rw_position<NX, PP, R>(q, acc, c, w);
#pragma unroll
        for (int p = 0; p < PP; ++p) {
            const double n2 = fma(b_2c[p], c[p], -pv[p]);
            pv[p] = c[p];
            c[p] = n2;
        }
    }
    th = q;
}

// the sine after `cur` as an instruction the compiler can neither hoist out of the block loop nor duplicate: left to itself it keeps
// the (block-invariant) second sine in registers across the walk and copies the recurrence state in on three paths (24 moves per block)
__device__ __forceinline__ double rw_next_sine(double two_c, double cur, double prev) {
    double r;
    asm volatile("{\n\t.reg .f64 t;\n\tneg.f64 t, %3;\n\tfma.rn.f64 %0, %1, %2, t;\n\t}" : "=d"(r) : "d"(two_c), "d"(cur), "d"(prev));
    return r;
}

// One SLICE of the walk: nblk blocks of RW_RB rows x positions, mu += sum_i a_i (sum_j Theta'[.., i, j] b_j) with the row sines a
// and the position sines b given by their recurrence seeds.  `w` is the Theta' register ring (rw_position), `th` the read pointer.
template <int NX, int PP>
__device__ __forceinline__ void rowwalk_slice(const double* __restrict__& th, const int* __restrict__ blen, int nblk, double (&w)[RW_RB][NX],
                                              double (&a_cur)[PP], double (&a_prev)[PP], const double (&a_2c)[PP],
                                              const double (&b_cur)[PP], const double (&b_prev)[PP], const double (&b_2c)[PP],
                                              double (&mu)[PP][NX]) {
    for (int b = 0; b < nblk; ++b) {
        double acc[PP][RW_RB][NX];
        double c[PP], pv[PP];
        const int L = blen[b];
        // The first position of a block has all RW_RB rows active (layout guarantee, model.cu): its products START the accumulators
        // (ptxas had placed 2 x 16 CS2R per block, 5 % of the kernel's instructions).  The block opens with one position if its
        // count of four-row positions is odd, with two if it is even: the four-row loop then runs whole pairs only, and the
        // recurrence state it starts from comes out of DFMAs (two-position opening) or one copy (one-position opening).
        th += RW_RB * NX;
        rw_position<NX, PP, RW_RB, true>(th, acc, b_cur, w);
        if (L & 1) {
#pragma unroll
            for (int p = 0; p < PP; ++p) {
                double s0 = b_cur[p];
                asm volatile("" : "+d"(s0));
                pv[p] = s0;
                c[p] = rw_next_sine(b_2c[p], s0, b_prev[p]);
            }
        } else {
            th += RW_RB * NX;
#pragma unroll
            for (int p = 0; p < PP; ++p) pv[p] = rw_next_sine(b_2c[p], b_cur[p], b_prev[p]);
            rw_position<NX, PP, RW_RB>(th, acc, pv, w);
#pragma unroll
            for (int p = 0; p < PP; ++p) c[p] = fma(b_2c[p], pv[p], -b_cur[p]);
        }
        // positions with 4, 3, 2, 1 active rows (packed byte counts, common.cuh): only selected lattice entries are walked
        rw_segment<NX, PP, 4, true>(th, ((L & 255) - 1) & ~1, acc, c, pv, b_2c, w);
        rw_segment<NX, PP, 3>(th, (L >> 8) & 255, acc, c, pv, b_2c, w);
        rw_segment<NX, PP, 2>(th, (L >> 16) & 255, acc, c, pv, b_2c, w);
        rw_segment<NX, PP, 1>(th, (L >> 24) & 255, acc, c, pv, b_2c, w);
        // the next block starts with all RW_RB rows: pairs 1.. of its first position (pair 0 is in the ring already) are
        // requested before this block's accumulators are folded
#pragma unroll
        for (int i = 1; i < RW_RB; ++i) rw_load<NX>(th + i * NX, w[i]);
#pragma unroll
        for (int i = 0; i < RW_RB; ++i)
#pragma unroll
            for (int p = 0; p < PP; ++p) {
#pragma unroll
                for (int k = 0; k < NX; ++k) mu[p][k] = fma(a_cur[p], acc[p][i][k], mu[p][k]);
                const double n = fma(a_2c[p], a_cur[p], -a_prev[p]);
                a_prev[p] = a_cur[p];
                a_cur[p] = n;
            }
    }
}

// D = 2: one slice (rows = first dimension, positions = second).  D = 3: one slice per first-dimension position i0,
//   mu_k = sum_i0 s0_i0 * [ sum_i1 s1_i1 ( sum_j Theta'[k, m(i0,i1,j)] s2_j ) ],
// rows = second dimension (its recurrence restarts in every slice), positions = third; slice_nblk[i0] blocks per slice.
template <int NX, int PP, int D>
__device__ __forceinline__ void rowwalk_mu(const double* __restrict__ bd, const int* __restrict__ blen, const int* __restrict__ slice_nblk,
                                           int nslice, int f_start, int f_step, const double (&tz)[PP][D], double (&mu)[PP][NX]) {
    double a_cur[PP], a_prev[PP], a_2c[PP], b_cur[PP], b_prev[PP], b_2c[PP];
    const double* __restrict__ th = bd;
    double w[RW_RB][NX];
#pragma unroll
    for (int i = 0; i < RW_RB; ++i) rw_load<NX>(th + i * NX, w[i]);
#pragma unroll
    for (int p = 0; p < PP; ++p) {
        sine_seed(tz[p][D - 2], f_start, f_step, a_cur[p], a_prev[p], a_2c[p]);
        sine_seed(tz[p][D - 1], f_start, f_step, b_cur[p], b_prev[p], b_2c[p]);
#pragma unroll
        for (int k = 0; k < NX; ++k) mu[p][k] = 0.0;
    }
    if constexpr (D == 2) {
        rowwalk_slice<NX, PP>(th, blen, slice_nblk[0], w, a_cur, a_prev, a_2c, b_cur, b_prev, b_2c, mu);
    } else {
        double s_cur[PP], s_prev[PP], s_2c[PP];
#pragma unroll
        for (int p = 0; p < PP; ++p) sine_seed(tz[p][0], f_start, f_step, s_cur[p], s_prev[p], s_2c[p]);
        int boff = 0;
        for (int sl = 0; sl < nslice; ++sl) {
            double ac[PP], ap[PP], part[PP][NX];
#pragma unroll
            for (int p = 0; p < PP; ++p) {
                ac[p] = a_cur[p]; ap[p] = a_prev[p];
#pragma unroll
                for (int k = 0; k < NX; ++k) part[p][k] = 0.0;
            }
            const int nb = slice_nblk[sl];
            rowwalk_slice<NX, PP>(th, blen + boff, nb, w, ac, ap, a_2c, b_cur, b_prev, b_2c, part);
            boff += nb;
#pragma unroll
            for (int p = 0; p < PP; ++p) {
#pragma unroll
                for (int k = 0; k < NX; ++k) mu[p][k] = fma(s_cur[p], part[p][k], mu[p][k]);
                const double n = fma(s_2c[p], s_cur[p], -s_prev[p]);
                s_prev[p] = s_cur[p];
                s_cur[p] = n;
            }
        }
    }
}


// Three-dimensional bases, FEW particles (the EMPS PGAS baseline: N = 200, one chain — src/EMPS.py:100-123, :240-255): G lanes share a
// particle, lane `sub` walks the slices sub, sub + G, ... (their own read pointer and register ring: loads are no longer warp-uniform)
// and a butterfly sums the lanes' partial means — commutative additions, so all G lanes end with the same bits.  With one particle per
// thread a 200-particle launch is 7 warps on 4 SMs and a step costs the latency of one thread walking all M entries (10.7 us at
// M = 729, profiles/r02_state_kernel_summary.md).  slice_off / slice_blk: first Theta' slot and first block of every slice; blen in
// shared memory (per-lane indices).
template <int NX, int G>
__device__ __forceinline__ void rowwalk_mu_lanes(const double* __restrict__ bd, const int* __restrict__ blen, const int* __restrict__ slice_nblk,
                                                 const int* __restrict__ slice_off, const int* __restrict__ slice_blk, int nslice, int sub,
                                                 int f_start, int f_step, const double (&tz)[1][3], double (&mu)[1][NX]) {
    static_assert(G == 2 || G == 4 || G == 8 || G == 16, "lanes per particle: a power of two inside a warp");
    double a_cur[1], a_prev[1], a_2c[1], b_cur[1], b_prev[1], b_2c[1], s_cur, s_prev, s_2c;
    sine_seed(tz[0][1], f_start, f_step, a_cur[0], a_prev[0], a_2c[0]);
    sine_seed(tz[0][2], f_start, f_step, b_cur[0], b_prev[0], b_2c[0]);
    sine_seed(tz[0][0], f_start, f_step, s_cur, s_prev, s_2c);
#pragma unroll
    for (int k = 0; k < NX; ++k) mu[0][k] = 0.0;
    for (int q = 0; q < sub; ++q) {                       // first-dimension sine at this lane's first slice
        const double n = fma(s_2c, s_cur, -s_prev);
        s_prev = s_cur; s_cur = n;
    }
    for (int sl = sub; sl < nslice; sl += G) {
        const double* __restrict__ th = bd + slice_off[sl];
        double w[RW_RB][NX], ac[1] = {a_cur[0]}, ap[1] = {a_prev[0]}, part[1][NX];
#pragma unroll
        for (int i = 0; i < RW_RB; ++i) rw_load<NX>(th + i * NX, w[i]);
#pragma unroll
        for (int k = 0; k < NX; ++k) part[0][k] = 0.0;
        rowwalk_slice<NX, 1>(th, blen + slice_blk[sl], slice_nblk[sl], w, ac, ap, a_2c, b_cur, b_prev, b_2c, part);
#pragma unroll
        for (int k = 0; k < NX; ++k) mu[0][k] = fma(s_cur, part[0][k], mu[0][k]);
#pragma unroll
        for (int q = 0; q < G; ++q) {
            const double n = fma(s_2c, s_cur, -s_prev);
            s_prev = s_cur; s_cur = n;
        }
    }
#pragma unroll
    for (int o = 1; o < G; o <<= 1)
#pragma unroll
        for (int k = 0; k < NX; ++k) mu[0][k] += __shfl_xor_sync(0xffffffffu, mu[0][k], o);
}
