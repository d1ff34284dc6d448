// resample_math.cuh — the pieces of systematic resampling (reference src/Filtering.py:6-37) that the sweep kernels share:
// CDF values, stratified points exactly as the reference rounds them, and the fixed-radix search of a padded CDF.
#pragma once
#include "common.cuh"
#include "fastmath.cuh"

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }

// value of the resampling CDF at a local inclusive prefix `p` (same expression for the segment
// boundaries and the interior so they agree bit for bit)
__device__ __forceinline__ double cdf_value(double p, double f, double g, double rs) {
    return __dmul_rn(__fma_rn(p, f, g), rs);
}
__device__ __forceinline__ double clip01(double v) { return fmin(fmax(v, 0.0), 1.0); }

// U_j = (u + j) / N exactly as src/Filtering.py:28 evaluates it (correctly rounded quotient)
__device__ __forceinline__ double strat_point(double u, int j, double dN, double rN) {
    return div_by_count(__dadd_rn(u, (double)j), dN, rN);
}

// smallest j in [0,N] with U_j > b  (U_j is non-decreasing in j): four candidates around the
// arithmetic guess are tested in parallel; the loops only run if the guess was off by more
__device__ __forceinline__ int first_point_above(double b, double u, int N, double dN, double rN) {
    const double g = floor(fma(b, dN, -u));
    int j = (g < 1.0) ? 0 : (g > (double)N ? N : (int)g - 1);
    const bool c0 = strat_point(u, j, dN, rN) > b, c1 = strat_point(u, j + 1, dN, rN) > b;
    const bool c2 = strat_point(u, j + 2, dN, rN) > b, c3 = strat_point(u, j + 3, dN, rN) > b;
    j = c0 ? j : (c1 ? j + 1 : (c2 ? j + 2 : (c3 ? j + 3 : j + 4)));
    j = min(j, N);
    while (j > 0 && strat_point(u, j - 1, dN, rN) > b) --j;
    while (j < N && !(strat_point(u, j, dN, rN) > b)) ++j;
    return j;
}

// number of elements of the non-decreasing array w that are < x.  w is padded with +inf to a
// multiple of 256 entries (nblk blocks), so the search is three fixed-radix levels whose loads and
// compares are all independent: block (<= 16 probes), 16 probes of stride 16, 16 neighbours.
__device__ __forceinline__ int count_below_padded(const double* __restrict__ w, int nblk, double x) {
    int base = 0;
    if (nblk > 1) {
        int c = 0;
        for (int b = 0; b < nblk; ++b) c += (w[b * 256 + 255] < x) ? 1 : 0;
        base = min(c, nblk - 1) * 256;
    }
    const double* w1 = w + base;
    int c1 = 0;
#pragma unroll
    for (int g = 0; g < 16; ++g) c1 += (w1[16 * g + 15] < x) ? 1 : 0;
    c1 = min(c1, 15);
    const double2* w2 = reinterpret_cast<const double2*>(w1 + 16 * c1);
    int c2 = 0;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const double2 v = w2[g];
        c2 += ((v.x < x) ? 1 : 0) + ((v.y < x) ? 1 : 0);
    }
    return base + 16 * c1 + c2;
}


// the same count by a branch-free binary search (w sorted, padded with +inf up to n_pad entries): ~13 dependent probes
// instead of 48 independent ones — a third of the instructions and of the shared-memory traffic.  The resampling kernels of
// the split sweep share their SMs with the FP64-bound state kernel, where issue slots, not latency, are what a search costs.
__device__ __forceinline__ int count_below_binary(const double* __restrict__ w, int n_pad, double x) {
    int base = 0, len = n_pad;
    while (len > 1) {
        const int half = len >> 1;
        base = (w[base + half - 1] < x) ? base + half : base;
        len -= half;
    }
    return base + ((w[base] < x) ? 1 : 0);
}
