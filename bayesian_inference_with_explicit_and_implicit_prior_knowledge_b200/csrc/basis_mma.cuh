// basis_mma.cuh — mu = Theta phi(z) for two-dimensional Hilbert bases with n_x = 2 on the FP64 TENSOR pipe, particles on the
// N dimension of the tile.
//
// Same factorisation as the row walk (basis_rowwalk.cuh; reference src/BasisFunctions.py:77-80, src/PGAS.py:52-55, :67-70),
//   mu_k(p) = sum_i a_i(p) * T[(i,k)][p],     T[(i,k)][p] = sum_j Theta'[k, m(i,j)] b_j(p),
// but the inner sums of a block of RW_RB = 4 rows are ONE m8n8k4 tile product per k-step of four walked positions and eight
// particles:   D[(i,k) : 8 rows][p : 8 particles] += A[(i,k)][j : 4] * B[j][p]
//   A = Theta' in fragment order (shared memory, one LDS.64 per lane and k-step, shared by all particle groups),
//   B = the sines b_j(p): lane L serves particle L / 4 of the group at walked positions j = 4 s + L % 4 with the stride-4
//       recurrence b_{j+4} = 2 cos(4 theta) b_j - b_{j-4}  (one DFMA per DMMA),
//   D = two accumulators per lane.
// A DMMA.8x8x4 occupies the FP64 pipe as long as the 8 DFMAs it replaces (the two share one pipe on B200,
// profiles/r02_microbench.md) but takes ONE issue slot instead of eight: the row walk is co-limited by the issue port
// (3088 instructions against 3208 pipe cycles per warp-step, profiles/r02_state_kernel_summary.md), this form is not.
// What it costs: the lattice rows of a block are padded to the block's longest row rounded up to four positions (80 instead
// of 72 positions at M = 256), and the row sines a_i(p) live with the particle's OWNER thread, so the tile results travel
// through shared memory once per block (8 x 32 doubles per half warp).
//
// Layout inside a warp: lane l owns particles l (slot 0) and 32 + l (slot 1) of the warp's 64; group g (8 particles) =
// particles 8 g .. 8 g + 7, i.e. slot g / 4.  The owner publishes, per particle, the recurrence seeds of the four phases
// (b_{-4..-1}, b_{0..3}) and 2 cos(4 theta) in shared memory; the serving lanes read theirs at the start of every block.
#pragma once
#include "basis_eval.cuh"

constexpr int MMA_SEED_LD = 9;       // doubles per particle in the seed table (odd: conflict-free column reads)
constexpr int MMA_TB_LD = 40;        // row pitch of the tile buffer: 64 bytes mod 128 -> the 16-byte fragment stores hit every bank group evenly
constexpr int MMA_WARP_DOUBLES = 64 * MMA_SEED_LD + 8 * MMA_TB_LD;

// seeds of the walked dimension for one particle: b_r = sin(pi (f_start + r f_step) t), r = -4 .. 3, and 2 cos(4 pi f_step t)
__device__ __forceinline__ void mma_seeds(double t, int f_start, int f_step, double* __restrict__ out) {
    double st, ct, sa, ca;
    sincospi_bf((double)f_step * t, st, ct);
    if (f_start == f_step) { sa = st; ca = ct; }
    else sincospi_bf((double)f_start * t, sa, ca);
    const double tc = 2.0 * ct;
    const double s0 = sa, s1 = fma(sa, ct, ca * st), sm1 = fma(sa, ct, -ca * st);
    const double s2 = fma(tc, s1, -s0), s3 = fma(tc, s2, -s1);
    const double sm2 = fma(tc, sm1, -s0), sm3 = fma(tc, sm2, -sm1), sm4 = fma(tc, sm3, -sm2);
    const double c2 = fma(tc, ct, -1.0), c4 = fma(2.0 * c2, c2, -1.0);
    out[0] = sm4; out[1] = sm3; out[2] = sm2; out[3] = sm1;
    out[4] = s0; out[5] = s1; out[6] = s2; out[7] = s3;
    out[8] = 2.0 * c4;
}

// frag: Theta' in A-fragment order [block][k-step][32]; ks: k-steps per block; wbuf: this warp's MMA_WARP_DOUBLES doubles;
// tz[p][d]: normalised GP inputs of the lane's two particles; mu[p][k]: result.  All 32 lanes must call.
template <int NX>
__device__ __forceinline__ void mma_mu(const double* __restrict__ frag, const unsigned char* __restrict__ ks, int nblk, double* __restrict__ wbuf,
                                       int lane, int f_start, int f_step, const double (&tz)[2][2], double (&mu)[2][NX]) {
    static_assert(NX == 2, "rows of the tile are (row of the block, state component) pairs: 4 x 2");
    double* seeds = wbuf;
    double* tb = wbuf + 64 * MMA_SEED_LD;
    double a_cur[2], a_prev[2], a_2c[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        mma_seeds(tz[p][1], f_start, f_step, seeds + (p * 32 + lane) * MMA_SEED_LD);
        sine_seed(tz[p][0], f_start, f_step, a_cur[p], a_prev[p], a_2c[p]);
        mu[p][0] = 0.0; mu[p][1] = 0.0;
    }
    __syncwarp();
    const int r = lane & 3, pi = lane >> 2;
    const double* __restrict__ fr = frag + lane;
    for (int b = 0; b < nblk; ++b) {
        const int n = ks[b];
#pragma unroll
        for (int h = 0; h < 2; ++h) {                                        // half = particle slot h: groups 4 h .. 4 h + 3
            double cur[4], prev[4], t4[4], c0[4], c1[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const double* sd = seeds + (h * 32 + g * 8 + pi) * MMA_SEED_LD;
                cur[g] = sd[4 + r]; prev[g] = sd[r]; t4[g] = sd[8];
                c0[g] = 0.0; c1[g] = 0.0;
            }
#pragma unroll 2
            for (int s = 0; s < n; ++s) {
                const double a = fr[s * 32];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    dmma_m8n8k4(c0[g], c1[g], a, cur[g]);
                    const double nx = fma(t4[g], cur[g], -prev[g]);
                    prev[g] = cur[g];
                    cur[g] = nx;
                }
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) *reinterpret_cast<double2*>(tb + pi * MMA_TB_LD + g * 8 + 2 * r) = make_double2(c0[g], c1[g]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < RW_RB; ++i) {
                const double t0 = tb[(2 * i) * MMA_TB_LD + lane], t1 = tb[(2 * i + 1) * MMA_TB_LD + lane];
                mu[h][0] = fma(a_cur[h], t0, mu[h][0]);
                mu[h][1] = fma(a_cur[h], t1, mu[h][1]);
                const double na = fma(a_2c[h], a_cur[h], -a_prev[h]);
                a_prev[h] = a_cur[h];
                a_cur[h] = na;
            }
            __syncwarp();
        }
        fr += n * 32;
    }
}
