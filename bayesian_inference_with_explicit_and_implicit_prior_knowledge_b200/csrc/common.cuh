// common.cuh — shared device/host helpers of libpgas_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include "../../include/pgas_b200.h"
#include "fastmath.cuh"

// ---------------------------------------------------------------------------------- errors
void pgas_set_error(const char* fmt, ...);
#define PGAS_FAIL(code, ...) do { pgas_set_error(__VA_ARGS__); return (code); } while (0)
#define PGAS_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
    pgas_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    return (int)e__; } } while (0)
// every kernel launch of the library goes through this (bench.py reports the count: pgas_launch_count)
extern long long g_pgas_launches;
#define PGAS_KERNEL_CHECK() do { __atomic_add_fetch(&g_pgas_launches, 1, __ATOMIC_RELAXED); PGAS_CUDA(cudaGetLastError()); } while (0)

// ---------------------------------------------------------------------------------- model
// Tensor-product layout of the Hilbert basis for the FP64 tensor-core contraction (DESIGN.md
// "Basis layout").  Basis functions sharing the lattice positions of the leading D-1 dimensions
// form a ROW r (R rows); inside a row the last dimension runs over positions j.  With
//   S[p, j]        = sin(pi f_j t_last(p))                       (particles x KP, KP = 4 KS)
//   B[j, r*n_x+k]  = norm * Theta[k, m(r, j)]   (0 where absent)  (KP x NCOL, NCOL = 8 NTN)
//   lead[p, r]     = prod_{d < D-1} sin(pi f_{pos_d(r)} t_d(p))
// the auxiliary mean is  mu_k(p) = sum_r lead[p, r] * (S B)[p, r*n_x+k]:  a dense (8 particles x
// 4 positions) x (4 positions x 8 columns) DMMA tile product per (ks, nt), followed by a per-row
// scaling and a 4-lane reduction.  B is stored in mma.sync m8n8k4 fragment order:
//   Bfrag[(ks*NTNP + nt)*32 + lane] = B[4 ks + lane%4][8 nt + lane/4]
// Rows are sorted by decreasing length, so the all-zero tiles outside the quarter-disc the
// lattice search selects are the trailing column tiles of each position step: ntcount[ks].
constexpr int MAX_LEAD = PGAS_MAX_D - 1;

// Row-walk layout of a TWO-dimensional basis for the thread-per-particle FP64-FMA contraction
// (basis_rowwalk.cuh): first-dimension lattice positions are grouped into blocks of RW_RB; inside a block
// the last-dimension positions j = 0 .. blen-1 are walked once while RW_RB x n_x accumulators collect
//   T[i][k] = sum_j Theta'[k, m(i, j)] sin(pi f_j t_last),
// then mu_k += sin(pi f_i t_first) T[i][k].  At position j only the first act(j) rows of the block are stored and
// walked (act = 1 + last row with a selected entry at a position >= j; non-increasing in j), so the padding of the
// staircase lattice is not computed.  Storage order: [block][j][i < act(j)][k], zero where the lattice search did not
// select (i, j); rw_blen[block] packs the number of positions with 4, 3, 2, 1 active rows into bytes 0..3.
constexpr int RW_RB = 4;
constexpr int RW_MAXBLK = 64;
constexpr int RW_MAXSLICE = 16;      // three-dimensional bases: first-dimension positions (one slice of blocks each)

struct DevModel {
    int n_x, n_y, n_u, D, M, T;
    int R, NTN, KS, n_packed;     // rows, column tiles (8 wide), position steps (4 deep), KS*NTNP*32 fragment slots
    int jmax;                     // positions of the last dimension incl. padding (= 4 KS)
    int npos_d[PGAS_MAX_D];       // lattice positions used per dimension (max position + 1)
    int NTNP;                     // NTN rounded up to a multiple of the accumulator block (zero tiles)
    int ntcount[16];              // per position step: column tiles up to the last non-zero one
    int f_start, f_step;
    int npos;                     // max lattice position over all dimensions + 1
    int map_kind, flags;
    double center[PGAS_MAX_D], L[PGAS_MAX_D], inv2L[PGAS_MAX_D];   // t_d = (z_d - center_d + L_d) * inv2L_d
    double norm;                                    // prod_d L_d^-1/2
    double Az[PGAS_MAX_D][PGAS_MAX_NX + PGAS_MAX_NU], bz[PGAS_MAX_D];
    double slip_lf, slip_lr;
    int prog_len, prog_op[PGAS_MAX_PROG];           // PGAS_MAP_PROGRAM: postfix expression program of the GP-input map (pgas_b200.h)
    double prog_const[PGAS_MAX_PROG];
    double H[PGAS_MAX_NY][PGAS_MAX_NX], h0[PGAS_MAX_NY];
    double Rw[PGAS_MAX_NY][PGAS_MAX_NY];            // inverse of chol(R) (lower): e = Rw (y - mean)
    double R_logc;                                  // -n_y/2 log(2 pi) - sum log diag chol(R)
    double m0[PGAS_MAX_NX], P0c[PGAS_MAX_NX][PGAS_MAX_NX];   // chol(P0) lower
    int rw_ok, rw_nblk, rw_slots;                   // row-walk layout (D == 2): available, blocks, doubles
    int rw_blen[RW_MAXBLK];                         // positions with 4 | 3 | 2 | 1 active rows per block (one byte each)
    int rw_nslice, rw_slice_nblk[RW_MAXSLICE];      // D = 3: blocks per first-dimension position (D = 2: one slice holding all blocks)
    int rw_slice_off[RW_MAXSLICE], rw_slice_blk[RW_MAXSLICE];   // first Theta' slot (doubles) and first block of every slice
    int mma_ok, mma_nblk, mma_slots;                // DMMA layout of a two-dimensional basis with n_x = 2 (basis_mma.cuh): available, blocks, doubles
    unsigned char mma_ks[RW_MAXBLK];                // k-steps (4 walked positions each) per block of RW_RB rows
    const int* rw_perm;     // [rw_slots] slot -> m * 4 + k of the Theta entry it holds, or -1 (zero)
    const int* mma_perm;    // [mma_slots] A-fragment slot -> m * 4 + k, or -1 (zero)
    const int* perm;        // [n_packed] fragment slot -> m * 4 + k of the Theta entry it holds, or -1 (zero)
    const int* row_pos;     // [8*NTN / n_x rounded up][MAX_LEAD] leading-dimension positions of each row (0 for padding rows)
    const int* freq;        // [M*D] integer frequencies, reference order
    const int* lat_perm;    // [M] basis functions in lattice order (positions sorted lexicographically, last dimension fastest)
    const double* obs;      // (T,n_y)
    const double* inputs;   // (T,n_u)
    // likelihood program (model plug-in): instructions prog_op[lik_off .. lik_off + lik_len), constants from prog_const[lik_coff].
    // Kept at the END of the struct: the kernels address DevModel through the constant bank, and the state kernel's code should
    // not move when a field is added.
    int lik_off, lik_len, lik_coff;
};

struct pgas_model {
    DevModel dev;
    void* arena;            // one device allocation holding all tables and data
    size_t arena_bytes;
};

// ---------------------------------------------------------------------------------- Philox
// Philox-4x32-10 (Salmon et al., SC'11): counter-based, no state; the counter layout is the
// library's RNG contract (pgas_b200.h: pgas_rng).
enum { PURPOSE_STATE = 0, PURPOSE_STEP_U = 1, PURPOSE_DRAW_G = 2, PURPOSE_DRAW_N = 3, PURPOSE_DRAW_CHI = 4 };

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform in [0,1) from two 32-bit words
__host__ __device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    uint64_t k = ((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6);
    return (double)k * (1.0 / 9007199254740992.0);
}

// two uniforms in [0,1) for (purpose, chain, iteration, time, index)
__device__ __forceinline__ void philox_uniform2(uint64_t seed, uint32_t purpose, uint32_t chain, uint32_t iter,
                                                uint32_t t, uint32_t i, double& ua, double& ub) {
    uint32_t o[4];
    philox4x32_10(i, t, iter, (purpose << 24) | (chain & 0xFFFFFFu), (uint32_t)seed, (uint32_t)(seed >> 32), o);
    ua = u53(o[0], o[1]);
    ub = u53(o[2], o[3]);
}

// Box-Muller: two independent standard normals from one Philox block.
// u1 in (0,1] (so log is finite), angle 2*pi*u2 evaluated with sincospi.
__device__ __forceinline__ void philox_normal2(uint64_t seed, uint32_t purpose, uint32_t chain, uint32_t iter,
                                               uint32_t t, uint32_t i, double& za, double& zb) {
    double ua, ub;
    philox_uniform2(seed, purpose, chain, iter, t, i, ua, ub);
    const double r = sqrt_bf(fmax(-2.0 * log_unit_bf(ua + (1.0 / 9007199254740992.0)), 1e-30));
    double s, c;
    sincospi_bf(2.0 * ub, s, c);
    za = r * c;
    zb = r * s;
}

// ---------------------------------------------------------------------------------- warp helpers
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// inclusive warp scan (Kogge-Stone)
__device__ __forceinline__ double warp_scan_incl(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}
