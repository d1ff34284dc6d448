// fastmath.cuh — branch-free float64 elementary functions for the sweep kernel.
//
// The sweep runs at 2 warps per SM sub-partition (one particle per thread, ~220 particles per
// SM at the BASELINE shapes), so its step time is set by DEPENDENT-instruction latency (DFMA:
// 8 cycles measured on B200), not by issue rate.  libdevice's sincospi/exp/log/sqrt/div carry
// slow-path branches and calls that stop ptxas from interleaving independent evaluations; these
// versions are straight-line (Horner chains the scheduler can interleave across the D basis
// dimensions, the two softmax numerators, and the Box-Muller pieces).  Accuracy ~1 ulp
// (checked against long double in tests/test_abi_and_host.py's NumPy mirror and on the GPU).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Polynomial coefficients live in the constant bank: ptxas reads two of them with one LDCU.128 and feeds them to DFMA as
// uniform-register operands, instead of materialising every 64-bit literal with a UMOV pair (two issue slots per
// coefficient — 10 % of the state kernel's instruction stream, profiles/r01_state_kernel_summary.md).
static __constant__ double FM_SIN[10] = {-2.2948428997269873e-08, 7.952054001475513e-07, -2.1915353447830217e-05, 0.00046630280576761255,
                                         -0.0073704309457143504, 0.08214588661112823, -0.5992645293207921, 2.5501640398773455,
                                         -5.16771278004997, 3.141592653589793};
static __constant__ double FM_COS[10] = {3.604730797462501e-09, -1.3878952462213771e-07, 4.303069587032947e-06, -0.0001046381049248457,
                                         0.0019295743094039231, -0.02580689139001406, 0.2353306303588932, -1.3352627688545895,
                                         4.0587121264167685, -4.934802200544679};
static __constant__ double FM_EXP[14] = {1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07,
                                         2.7557319223985893e-06, 2.48015873015873e-05, 0.0001984126984126984, 0.001388888888888889,
                                         0.008333333333333333, 0.041666666666666664, 0.16666666666666666, 0.5, 1.0, 1.0};
static __constant__ double FM_LOG[11] = {1.0 / 23.0, 1.0 / 21.0, 1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0, 1.0 / 9.0,
                                         1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0};

// sin(pi x), cos(pi x): reduce to r = x - n/2, |r| <= 1/4 (exact), Taylor in t = r^2
__device__ __forceinline__ void sincospi_bf(double x, double& s_out, double& c_out) {
    const double n = rint(x + x);
    const double r = fma(-0.5, n, x);
    const double t = r * r;
    double ps = FM_SIN[0], pc = FM_COS[0];
#pragma unroll
    for (int i = 1; i < 10; ++i) {
        ps = fma(ps, t, FM_SIN[i]);
        pc = fma(pc, t, FM_COS[i]);
    }
    const double s = r * ps, c = fma(pc, t, 1.0);
    const int q = (int)__double2ll_rn(n) & 3;
    const double a = (q & 1) ? c : s, b = (q & 1) ? s : c;
    s_out = (q & 2) ? -a : a;                    // q: 0 -> s, 1 -> c, 2 -> -s, 3 -> -c
    c_out = ((q + 1) & 2) ? -b : b;              // q: 0 -> c, 1 -> -s, 2 -> -c, 3 -> s
}

// exp(x) for x <= 0 (softmax numerators); 0 below the normal range, NaN propagates
__device__ __forceinline__ double exp_neg_bf(double x) {
    const double n = rint(x * 1.4426950408889634);
    double r = fma(-n, 6.93147180369123816490e-01, x);
    r = fma(-n, 1.90821492927058770002e-10, r);
    double p = FM_EXP[0];
#pragma unroll
    for (int i = 1; i < 14; ++i) p = fma(p, r, FM_EXP[i]);
    const int e = (int)n;
    const double scale = __longlong_as_double((long long)(e + 1023) << 52);
    const double v = p * scale;
    return (x < -708.0) ? 0.0 : v;        // valid for x <= ~+700 as well (the shift may undershoot the max slightly)
}

// 1/d for normal positive d: float seed + two Newton steps (~1 ulp).  The seed is the bare MUFU.RCP: __frcp_rn carries a
// slow-path CALL that splits the caller's basic block, which stops ptxas from interleaving independent evaluations.
__device__ __forceinline__ double rcp_bf(double d) {
    float y0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"((float)d));
    double y = (double)y0;
    y = y * fma(-d, y, 2.0);
    y = y * fma(-d, y, 2.0);
    return fma(y, fma(-d, y, 1.0), y);
}

// log(u), u in (0, 1]
__device__ __forceinline__ double log_unit_bf(double u) {
    const long long bits = __double_as_longlong(u);
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    double m = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);   // [1,2)
    const bool big = m > 1.4142135623730951;
    m = big ? 0.5 * m : m;
    e += big ? 1 : 0;
    const double num = m - 1.0, den = m + 1.0;
    const double y = rcp_bf(den);
    double f = num * y;
    f = fma(fma(-den, f, num), y, f);
    const double t = f * f;
    double p = FM_LOG[0];
#pragma unroll
    for (int i = 1; i < 11; ++i) p = fma(p, t, FM_LOG[i]);
    const double lm = fma(2.0 * f * t, p, 2.0 * f);
    const double de = (double)e;
    return fma(de, 6.93147180369123816490e-01, fma(de, 1.90821492927058770002e-10, lm));
}

// sqrt(a), a > 0 normal: float rsqrt seed + Newton (~1 ulp)
__device__ __forceinline__ double sqrt_bf(double a) {
    float y0;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"((float)a));
    double y = (double)y0;
    y = y * fma(-0.5 * a, y * y, 1.5);
    y = y * fma(-0.5 * a, y * y, 1.5);
    double s = a * y;
    return fma(fma(-s, s, a), 0.5 * y, s);
}

// correctly rounded a / N given rN = RN(1/N) (Markstein): equals IEEE division for integer N < 2^24
__device__ __forceinline__ double div_by_count(double a, double dN, double rN) {
    const double q = a * rN;
    return fma(fma(-q, dN, a), rN, q);
}

// upper bound-ish of max over the warp for use as a softmax SHIFT: exact maximum of the values with
// their low 32 mantissa bits cleared (one 32-bit redux instead of a 5-level 64-bit shuffle tree).
// The result differs from the true maximum by < 2^-20 relative, which only rescales numerator and
// denominator of the softmax identically.  NaNs are ignored like fmax does.
__device__ __forceinline__ double warp_shift_max(double v) {
    const unsigned hi = (unsigned)__double2hiint(v);
    const bool isnan_ = v != v;
    // order-preserving map of the sign-magnitude high word to unsigned
    unsigned key = (hi & 0x80000000u) ? ~hi : (hi | 0x80000000u);
    key = isnan_ ? 0u : key;
    const unsigned best = __reduce_max_sync(0xffffffffu, key);
    const unsigned bh = (best & 0x80000000u) ? (best & 0x7fffffffu) : ~best;
    return __hiloint2double((int)bh, 0);
}
