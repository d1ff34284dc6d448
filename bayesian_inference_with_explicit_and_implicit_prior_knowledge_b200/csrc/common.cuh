// common.cuh — shared device/host helpers of libpgas_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include "../../include/pgas_b200.h"

// ---------------------------------------------------------------------------------- errors
void pgas_set_error(const char* fmt, ...);
#define PGAS_FAIL(code, ...) do { pgas_set_error(__VA_ARGS__); return (code); } while (0)
#define PGAS_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
    pgas_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    return (int)e__; } } while (0)
#define PGAS_KERNEL_CHECK() PGAS_CUDA(cudaGetLastError())

// ---------------------------------------------------------------------------------- model
// Packed tensor-product layout of the Hilbert basis (see DESIGN.md "Basis layout"):
// basis functions are grouped into ROWS sharing the lattice positions of the leading D-1
// dimensions; inside a row the last dimension runs over positions 0..len-1 (holes and the tail
// up to a multiple of CHUNK are padded, their Theta entry is 0).
//   mu_k = norm * sum_rows lead(row) * sum_p Theta'[row,p,k] s_last[p]
// A row is cut into CHUNKs of 4 consecutive last-dimension positions; the kernel walks the flat
// chunk list.  Per chunk one int of metadata: bits 0-7 = position block (positions 4b..4b+3),
// bit 8 = last chunk of its row, bits 9-10 = how the NEXT row's leading sines follow from this
// row's: consecutive rows differ by ONE unit advance of the faster leading dimension or by
// "advance the slower dimension and reset the faster to 0", so the leading sines are carried as
// 3-term recurrences instead of tables.
constexpr int CHUNK = 4;
enum { ROW_ADV_NONE = 0, ROW_ADV_FAST = 1, ROW_ADV_SLOW = 2 };
constexpr int META_ROW_END = 0x100;
constexpr int META_ADV_SHIFT = 9;

struct DevModel {
    int n_x, n_y, n_u, D, M, T;
    int n_chunks, jmax, n_packed; // jmax = max padded row length (multiple of CHUNK), n_packed = n_chunks*CHUNK
    int f_start, f_step;
    int npos;                     // max lattice position over all dimensions + 1
    int map_kind, flags;
    double center[PGAS_MAX_D], L[PGAS_MAX_D], inv2L[PGAS_MAX_D];   // t_d = (z_d - center_d + L_d) * inv2L_d
    double norm;                                    // prod_d L_d^-1/2
    double Az[PGAS_MAX_D][PGAS_MAX_NX + PGAS_MAX_NU], bz[PGAS_MAX_D];
    double slip_lf, slip_lr;
    double H[PGAS_MAX_NY][PGAS_MAX_NX], h0[PGAS_MAX_NY];
    double Rw[PGAS_MAX_NY][PGAS_MAX_NY];            // inverse of chol(R) (lower): e = Rw (y - mean)
    double R_logc;                                  // -n_y/2 log(2 pi) - sum log diag chol(R)
    double m0[PGAS_MAX_NX], P0c[PGAS_MAX_NX][PGAS_MAX_NX];   // chol(P0) lower
    const int* chunk_meta;  // [n_chunks]
    const int* perm;        // [n_packed] packed slot -> basis index m, or -1 (padding)
    const int* freq;        // [M*D] integer frequencies, reference order
    const double* obs;      // (T,n_y)
    const double* inputs;   // (T,n_u)
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
    double r = sqrt(-2.0 * log(ua + (1.0 / 9007199254740992.0)));
    double s, c;
    sincospi(2.0 * ub, &s, &c);
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
