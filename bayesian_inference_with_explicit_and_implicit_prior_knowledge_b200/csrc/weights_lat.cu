// weights_lat.cu — the weight recursion of the split sweep, LATENCY form (sm_100a): one thread-block cluster per chain,
// two particles per thread, one CTA barrier and two byte-counted hand-offs per step.
//
// Same mathematics as weights.cu (reference src/PGAS.py:92-127, :137-147; src/Filtering.py:6-37):
//   w_aux = softmax(l_aux + logw)                      a     = systematic_SISR(u_res, w_aux)
//   w_anc = softmax(l_aux + logw + h)                  a_N-1 = searchsorted(cumsum(w_anc), u_anc)
//   logw' = ll - l_aux[a]
// The recursion is a chain of short dependent phases, so what a step costs is its critical path and its instruction count per
// thread, not its flops: weights.cu's one-CTA form spends ~25 k cycles per step (eight particles per thread walked one after
// the other, three CTA barriers, a 13-probe CDF search per point).  This kernel is built for the path:
//   * PARTICLES ACROSS THE CLUSTER: C CTAs x NT threads own PPT consecutive particles each (N = 4096: 8 x 256 x 2);
//   * EXPONENTS AS INTEGERS: exp(lw) = p 2^n with n = rint(lw log2 e) kept as an integer; the softmax shift of a warp is the
//     integer maximum of n (one redux.sync), and every later rescaling between warps / CTAs is a multiplication by an exact power
//     of two — no second exponential anywhere, no rounding in the rescaling;
//   * TWO-LEVEL FOLD IN FIXED POINT: the 8 warp sums of a CTA, then the C CTA sums of the chain, are rescaled to the largest
//     exponent (exact) and converted to 64-bit fixed point (2^-48 of the largest unit); lane w of every warp handles record w
//     and a three-level shuffle scan of INTEGERS gives the offsets — integer sums are associative, so every warp of the chain
//     obtains bit-identical, monotone unit boundaries whatever the order of the additions (floating-point scans do not);
//   * NO SEARCH: particle k computes the number of stratified points at or below its CDF value ARITHMETICALLY,
//     c_k = #{j : U_j <= W_k} = floor(W_k N - u) + 1 (exactly-rounded fallback when W_k N - u is within 1e-9 of an integer),
//     and writes its index over the points j in [c_{k-1}, c_k) — its offspring; the point's owner receives l_aux[k].  The counts
//     are clamped into [c(boundary below), c(boundary above)] of the particle's warp and made monotone inside the warp, so they
//     telescope to exactly N whatever the rounding of the prefix sums does;
//   * NO CLUSTER BARRIER, NO FENCE: both exchanges of a step (CTA sums, all-gathered; the scattered l_aux values) are st.async
//     stores into the peer's shared memory that complete a transaction count on the peer's mbarrier — the consumer waits for
//     BYTES, not for CTAs (a release/acquire cluster barrier costs a MEMBAR.ALL.GPU and an L1 invalidate each time,
//     profiles/r01_weights_kernel_summary.md).  Every point has exactly one ancestor, so a CTA expects exactly 8 bytes per point
//     it owns; every CTA sends one 32-byte record to every CTA and every warp one 16-byte "reference ancestor" record to the CTA
//     that owns particle N-1 — the expected byte counts are data-independent.
#include <algorithm>
#include <stdlib.h>
#include "sweep_args.cuh"
#include "resample_math.cuh"

#ifndef PGAS_WL_NT
#define PGAS_WL_NT 256
#endif
#ifndef PGAS_WL_PPT
#define PGAS_WL_PPT 2
#endif
constexpr int WL_NT = PGAS_WL_NT, WL_PPT = PGAS_WL_PPT, WL_MAXC = 16;
constexpr int WL_INLINE = 2;            // offspring written by the particle's own thread; more go through the warp
constexpr int WL_NOCAND = 1 << 30;

__device__ __forceinline__ uint32_t wl_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t wl_mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void wl_st_async2(uint32_t raddr, double x, double y, uint32_t rmbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];"
                 ::"r"(raddr), "d"(x), "d"(y), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void wl_st_async1(uint32_t raddr, double x, uint32_t rmbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(raddr), "d"(x), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void wl_expect(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wl_wait(uint32_t mbar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();      // watchdog (seconds): a byte count that never completes is a bug, not a wait
    } while (!ok);
}

// exact count near a stratified point: rare (|W N - u - integer| < 1e-9), kept out of line so that it does not bloat the step
__device__ __noinline__ int wl_points_exact(double W, double u, int N, double dN, double rN) { return first_point_above(W, u, N, dN, rN); }

// number of stratified points U_j = (u + j) / N, j in [0, N), that are <= W  (U_j evaluated as src/Filtering.py:28 does):
// the arithmetic guess and whether W N - u lies within 1e-9 of a point (then the caller decides with the exact quotient)
__device__ __forceinline__ int wl_points_guess(double W, double u, double dN, int N, bool& near) {
    const double t = fma(W, dN, -u);                                        // W in [0, 1]: t in [-1, N]
    near = fabs(t - rint(t)) < 1e-9;
    return min(max(__double2int_rd(t) + 1, 0), N);
}
// min(v, 1) for a non-negative, non-NaN v (the CDF values: sums of non-negative numerators times a positive reciprocal)
__device__ __forceinline__ double wl_min1(double v) { return v > 1.0 ? 1.0 : v; }

constexpr int WL_KNONE = -(1 << 29);    // exponent of an empty unit (all weights zero)
constexpr int WL_KNAN = 1 << 29;        // exponent code of a unit that saw a NaN log-weight: wins every maximum, the step falls back to uniform weights
constexpr int WL_QBITS2 = 48, WL_QBITS4 = 46;            // fixed-point fraction bits of the unit sums (values < 2^8 per warp of 128 particles, < 2^16 per chain of 16 CTAs x 16 warps: no overflow in 63 bits)

// exp(x) = p * 2^n with |log p| <= ln2 / 2 for -1e9 < x < 1.4e9 (n fits an int); x = -inf or below -1e9 -> (0, WL_KNONE); NaN -> (0, WL_KNAN);
// +inf or above 1.4e9: the conversion of n saturates above WL_KNAN, i.e. the step falls back to uniform weights like a NaN does
__device__ __forceinline__ void wl_exp_parts(double x, double& p, int& n) {
    const bool dead = !(x > -1e9);                                          // -inf, hugely negative, or NaN (the comparison is false)
    const bool isnan_ = x != x;
    const double xs = (dead || isnan_) ? 0.0 : x;
    const double nd = rint(xs * 1.4426950408889634);
    double r = fma(-nd, 6.93147180369123816490e-01, xs);
    r = fma(-nd, 1.90821492927058770002e-10, r);
    double q = FM_EXP[0];
#pragma unroll
    for (int i = 1; i < 14; ++i) q = fma(q, r, FM_EXP[i]);
    p = dead ? 0.0 : q;                                                     // `dead` is true for NaN as well
    n = isnan_ ? WL_KNAN : (dead ? WL_KNONE : (int)nd);
}
// 2^d for d < 1024 (exactly 0 at and below the bottom of the normal range): an exact scale factor, three integer instructions
__device__ __forceinline__ double wl_pow2(int d) { return __hiloint2double(max(d + 1023, 0) << 20, 0); }
__device__ __forceinline__ long long wl_scan3(long long v, int lane, int levels) {     // inclusive scan over the first 2^levels lanes
    for (int o = 1, l = 0; l < levels; o <<= 1, ++l) {
        const long long n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// Offspring ranges of the PPT consecutive particles of one lane: particle g0 + u gets the stratified points [c_{u-1}, cend[u]) with
// c_{-1} = the returned value.  W: the particles' CDF values; lo / hi: the CDF at the two boundaries of the lane's UNIT (the 32 PPT
// particles of its warp), which the neighbouring units evaluate from the same integers, bit for bit.  Counts are made monotone inside
// the unit and clamped between c(lo) and c(hi); the unit's last particle ends at c(hi): the ranges of a chain partition [0, N)
// whatever the rounding of the prefix sums does (src/Filtering.py:28-35: idx_j = #{k : W_k < U_j}, clipped to N - 1).
template <int PPT>
__device__ __forceinline__ int wl_offspring_ranges(const double (&W)[PPT], double lo, double hi, bool first_unit, bool last_unit, bool empty_unit,
                                                   int g0, int N, double u, double dN, double rN, int lane, int (&cend)[PPT]) {
    bool n_lo, n_hi, n_w[PPT];
    int c_lo = wl_points_guess(lo, u, dN, N, n_lo), c_hi = wl_points_guess(hi, u, dN, N, n_hi);
#pragma unroll
    for (int q = 0; q < PPT; ++q) cend[q] = wl_points_guess(W[q], u, dN, N, n_w[q]);
    bool any_near = n_lo || n_hi;
#pragma unroll
    for (int q = 0; q < PPT; ++q) any_near = any_near || n_w[q];
    if (any_near) {                                                         // rare: a CDF value within 1e-9 / N of a stratified point
        if (n_lo) c_lo = wl_points_exact(lo, u, N, dN, rN);
        if (n_hi) c_hi = wl_points_exact(hi, u, N, dN, rN);
#pragma unroll
        for (int q = 0; q < PPT; ++q)
            if (n_w[q]) cend[q] = wl_points_exact(W[q], u, N, dN, rN);
    }
    c_lo = first_unit ? 0 : c_lo;
    c_hi = last_unit ? N : c_hi;                                            // the last particle takes what is left (Filtering.py:35)
    if (empty_unit) c_lo = N;
    c_hi = max(c_hi, c_lo);
#pragma unroll
    for (int q = 1; q < PPT; ++q) cend[q] = max(cend[q], cend[q - 1]);
    // monotone inside the warp (running maximum over the lanes — only if some lane is out of order, which takes a rounding glitch
    // of the scan), clamped between the unit's boundaries; the unit's last particle ends at the upper boundary
    int rm = cend[PPT - 1];
    int cp = __shfl_up_sync(0xffffffffu, rm, 1);
    if (__any_sync(0xffffffffu, lane > 0 && cend[0] < cp)) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, rm, o);
            if (lane >= o) rm = max(rm, v);
        }
        cp = __shfl_up_sync(0xffffffffu, rm, 1);
    }
    cp = lane ? min(max(cp, c_lo), c_hi) : c_lo;
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        cend[q] = min(max(max(cend[q], cp), c_lo), c_hi);
        if ((lane == 31 && q == PPT - 1) || g0 + q >= N - 1) cend[q] = c_hi;
    }
    return cp;
}

// CT: compile-time bound of the cluster size (1 | 2 | 4 | 8 | 16; records of CTAs c >= C stay empty), so that both folds unroll
// eight particles per thread (N > 8192): capped at 128 registers (a few spills) so that a state CTA fits on the SM next to it — configs[4]:
// 301 ms per iteration against 342 uncapped (251 registers: the cluster's SMs are lost to the state kernel while a chunk runs); two
// particles per thread: uncapped (146 registers) measured 13.2 against 13.8 ms per iteration at 8 chains
template <int NT, int PPT, int CT>
__global__ void __launch_bounds__(NT, PPT >= 8 ? 2 : 1) csmc_weights_lat_kernel(const __grid_constant__ SweepArgs a) {
    constexpr int NW = NT / 32, P = NT * PPT, WL_QBITS = (PPT * NT <= 512) ? WL_QBITS2 : WL_QBITS4 - (PPT * NT > 1024 ? 2 : 0);
    static_assert(PPT == 1 || PPT == 2 || PPT == 4 || PPT == 8, "1, 2, 4 or 8 consecutive particles per thread");
    const int C = a.C, N = a.N;
    uint32_t rank_u;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank_u));
    const int rank = (int)rank_u;
    const int chain = blockIdx.x / C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int base = rank * P, il0 = tid * PPT, g0 = base + il0;             // g0: global index of this thread's first particle
    const int nvalid = max(0, min(PPT, N - g0));
    const int Pc = max(0, min(P, N - base));
    const int c_last = (N - 1) / P;                                          // CTA owning particle N-1
    const int U = C * NW, mine = rank * NW + warp;
    const int first_of_unit = g0 - lane * PPT;                               // global index of the warp's first particle
    const double dN = (double)N, rN = 1.0 / (double)N;
    const bool vec = PPT >= 2 && (N & 1) == 0;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* wrec = reinterpret_cast<double*>(smem_raw);                      // [2][NW][4]: (sum1, sum2, k1, k2) of every warp of this CTA
    double* crec = wrec + 2 * NW * 4;                                        // [2][WL_MAXC][4]: (sum1, sum2, k1, k2) of every CTA of the chain
    double* lauxg = crec + 2 * WL_MAXC * 4;                                  // [2][P]: l_aux[a_j] for the points j this CTA owns
    double* refmsg = lauxg + (size_t)2 * P;                                  // [2][U][2]: (candidate reference ancestor, its l_aux) per warp
    double* su = refmsg + (size_t)2 * U * 2;                                 // [2][2]: (u_res, u_anc) by step parity
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(su + 4);                        // CTA records x 2 parities, scatter x 2 parities
    auto step_uniforms = [&](int t) {                                        // the step's two uniforms (src/PGAS.py:105, :121), one thread per CTA
        double* dst = su + 2 * (t & 1);
        if (a.rng_mode == 1) {
            const double* up = a.U + ((size_t)chain * a.var_rows + (t - a.row_off)) * 2;
            dst[0] = up[0]; dst[1] = up[1];
        } else {
            philox_uniform2(a.seed, PURPOSE_STEP_U, a.chain_base + chain, a.iteration, (unsigned)t, 0u, dst[0], dst[1]);
        }
    };
    if (tid == 32 % NT) step_uniforms(a.t_begin);

    for (int i = tid; i < 2 * WL_MAXC; i += NT) {                            // records of absent CTAs: empty
        *reinterpret_cast<double2*>(crec + i * 4) = make_double2(0.0, 0.0);
        *reinterpret_cast<int2*>(crec + i * 4 + 2) = make_int2(WL_KNONE, WL_KNONE);
    }
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(wl_smem(mbar + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    double logw[PPT];
#pragma unroll
    for (int u = 0; u < PPT; ++u) logw[u] = (u < nvalid && a.init_logw) ? a.init_logw[(size_t)chain * N + g0 + u] : 0.0;
    cluster_arrive();
    cluster_wait();

    const uint32_t s_crec = wl_smem(crec), s_lauxg = wl_smem(lauxg), s_ref = wl_smem(refmsg), s_mbar = wl_smem(mbar);
    const uint32_t r_crec = wl_mapa(s_crec, lane < C ? lane : 0), r_mbar_rec = wl_mapa(s_mbar, lane < C ? lane : 0);   // lane c serves CTA c
    const uint32_t r_ref = wl_mapa(s_ref, c_last), r_mbar_last = wl_mapa(s_mbar, c_last);

    // row pointers of this thread's particles, advanced by N per step
    const size_t prow0 = ((size_t)chain * a.pre_rows + (size_t)(a.t_begin - a.pre_off)) * N + g0;
    const double *pla = a.pre_la + prow0, *plr = a.pre_lr + prow0, *pll = a.pre_ll + prow0;
    auto load_rows = [&](double (&la)[PPT], double (&lr)[PPT], double (&ll)[PPT]) {
        if (vec && nvalid == PPT) {
#pragma unroll
            for (int u = 0; u + 1 < PPT; u += 2) {
                const double2 x = *reinterpret_cast<const double2*>(pla + u), y = *reinterpret_cast<const double2*>(plr + u),
                              z = *reinterpret_cast<const double2*>(pll + u);
                la[u] = x.x; la[u + 1] = x.y; lr[u] = y.x; lr[u + 1] = y.y; ll[u] = z.x; ll[u + 1] = z.y;
            }
        } else {
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                la[u] = (u < nvalid) ? pla[u] : 0.0;
                lr[u] = (u < nvalid) ? plr[u] : 0.0;
                ll[u] = (u < nvalid) ? pll[u] : 0.0;
            }
        }
        pla += N; plr += N; pll += N;
    };
    double nla[PPT], nlr[PPT], nll[PPT];
    load_rows(nla, nlr, nll);

    int* anc_row = a.anc_trace + ((size_t)chain * a.anc_rows + (a.t_begin - 1 - a.row_off + a.anc_shift)) * N;   // advanced by N per step
#define WL_TICK(K) do { if (a.dbg && blockIdx.x < 2 && tid == 96) a.dbg[((size_t)(t - a.t_begin) * 2 + blockIdx.x) * 8 + (K)] = clock64(); } while (0)
    for (int t = a.t_begin; t < a.t_end; ++t) {
        WL_TICK(0);
        const int par = t & 1;
        const uint32_t ph = (uint32_t)((t - a.t_begin) >> 1) & 1u;
        if (tid == 0) {                                                     // arm this step's two transaction barriers
            wl_expect(s_mbar + 8 * par, (uint32_t)C * 32u);
            wl_expect(s_mbar + 8 * (2 + par), (uint32_t)Pc * 8u + (rank == c_last ? (uint32_t)U * 16u : 0u));
        }
        double la[PPT], lr[PPT], ll[PPT];
#pragma unroll
        for (int u = 0; u < PPT; ++u) { la[u] = nla[u]; lr[u] = nlr[u]; ll[u] = nll[u]; }
        if (t + 1 < a.t_end) load_rows(nla, nlr, nll);                      // in flight for a whole step
        if (tid == 32 % NT && t + 1 < a.t_end) step_uniforms(t + 1);            // read after the next step's CTA barrier

        // ---- A: softmax numerators p 2^(n - k_warp), in-warp inclusive prefixes, warp record -> shared memory of the CTA
        double s1[PPT], s2[PPT];
        int kw1, kw2;
        {
            double p1[PPT], p2[PPT];
            int n1[PPT], n2[PPT];
            int m1 = WL_KNONE, m2 = WL_KNONE;
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                const double lwa = (u < nvalid) ? la[u] + logw[u] : -INFINITY;
                const double lwr = (u < nvalid) ? lwa + lr[u] : -INFINITY;
                wl_exp_parts(lwa, p1[u], n1[u]);
                wl_exp_parts(lwr, p2[u], n2[u]);
                m1 = max(m1, n1[u]); m2 = max(m2, n2[u]);
            }
            kw1 = __reduce_max_sync(0xffffffffu, m1);
            kw2 = __reduce_max_sync(0xffffffffu, m2);
            double r1 = 0.0, r2 = 0.0;
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                r1 = __dadd_rn(r1, __dmul_rn(p1[u], wl_pow2(n1[u] - kw1)));
                r2 = __dadd_rn(r2, __dmul_rn(p2[u], wl_pow2(n2[u] - kw2)));
                s1[u] = r1; s2[u] = r2;
            }
            const double i1 = warp_scan_incl(r1, lane), i2 = warp_scan_incl(r2, lane);
            double x1 = __shfl_up_sync(0xffffffffu, i1, 1), x2 = __shfl_up_sync(0xffffffffu, i2, 1);
            x1 = lane ? x1 : 0.0; x2 = lane ? x2 : 0.0;
#pragma unroll
            for (int u = 0; u < PPT; ++u) { s1[u] = __dadd_rn(x1, s1[u]); s2[u] = __dadd_rn(x2, s2[u]); }
            if (lane == 31) {
                double* wr = wrec + (par * NW + warp) * 4;
                *reinterpret_cast<double2*>(wr) = make_double2(i1, i2);
                *reinterpret_cast<int2*>(wr + 2) = make_int2(kw1, kw2);
            }
        }
        WL_TICK(1);
        __syncthreads();
        const double ures = su[2 * par], uanc = su[2 * par + 1];

        // ---- X1: fold of this CTA's NW warp records.  The warp sums are rescaled to the CTA's largest exponent (exact) and turned
        //      into 64-bit FIXED-POINT numbers (2^-48 of the largest unit, 2^-46 at four particles per thread: below 1e-12 of the CDF, far below the 1e-12 tie tolerance):
        //      integer sums are associative, so the shuffle scan below gives every warp of the chain bit-identical, monotone
        //      unit boundaries whatever the order of the additions.  Lane w < NW of every warp handles record w.
        long long qw1, qn1, qw2, qn2;                                       // this warp's start / end offset inside the CTA (CTA scale)
        int K1, K2;
        constexpr int LNW = NW <= 1 ? 0 : NW <= 2 ? 1 : NW <= 4 ? 2 : NW <= 8 ? 3 : NW <= 16 ? 4 : 5;
        {
            const double* wr = wrec + (par * NW + min(lane, NW - 1)) * 4;
            const double2 sv = *reinterpret_cast<const double2*>(wr);
            const int2 kk = *reinterpret_cast<const int2*>(wr + 2);
            K1 = __reduce_max_sync(0xffffffffu, kk.x); K2 = __reduce_max_sync(0xffffffffu, kk.y);
            long long q1 = __double2ll_rn(__dmul_rn(sv.x, wl_pow2(kk.x - K1 + WL_QBITS)));
            long long q2 = __double2ll_rn(__dmul_rn(sv.y, wl_pow2(kk.y - K2 + WL_QBITS)));
            if (lane >= NW) { q1 = 0; q2 = 0; }
            const long long i1 = wl_scan3(q1, lane, LNW), i2 = wl_scan3(q2, lane, LNW);
            qn1 = __shfl_sync(0xffffffffu, i1, warp); qn2 = __shfl_sync(0xffffffffu, i2, warp);
            qw1 = qn1 - __shfl_sync(0xffffffffu, q1, warp); qw2 = qn2 - __shfl_sync(0xffffffffu, q2, warp);
            if (warp == 0) {                                                // lane c < C writes this CTA's record into CTA c
                const long long t1 = __shfl_sync(0xffffffffu, i1, NW - 1), t2 = __shfl_sync(0xffffffffu, i2, NW - 1);
                if (lane < C) {
                    const uint32_t dst = r_crec + (uint32_t)((par * WL_MAXC + rank) * 32), mb = r_mbar_rec + 8 * par;
                    wl_st_async2(dst, __longlong_as_double(t1), __longlong_as_double(t2), mb);
                    wl_st_async2(dst + 16, __hiloint2double(K2, K1), 0.0, mb);
                }
            }
        }
        WL_TICK(2);
        wl_wait(s_mbar + 8 * par, ph);
        WL_TICK(3);

        // ---- X2: fold of the chain's CTA records (lane c < CT handles record c): CTA sums shifted to the chain's largest exponent,
        //      integer scan; CDF boundaries of this thread's warp
        double S1, S2, lo1, hi1, lo2;                                       // reciprocal normalisers; first CDF at the warp's boundaries; second below
        double sc1, sc2, of1, of2;                                          // chain-level scale / offset of this warp's in-warp prefixes
        bool uni, nan2;
        {
            constexpr int LCT = CT <= 1 ? 0 : CT <= 2 ? 1 : CT <= 4 ? 2 : CT <= 8 ? 3 : 4;
            const double* cr = crec + (par * WL_MAXC + min(lane, CT - 1)) * 4;
            const double2 sv = *reinterpret_cast<const double2*>(cr);
            const int2 kk = *reinterpret_cast<const int2*>(cr + 2);
            const int KK1 = __reduce_max_sync(0xffffffffu, kk.x), KK2 = __reduce_max_sync(0xffffffffu, kk.y);
            long long q1 = __double_as_longlong(sv.x) >> min(KK1 - kk.x, 63), q2 = __double_as_longlong(sv.y) >> min(KK2 - kk.y, 63);
            if (lane >= CT) { q1 = 0; q2 = 0; }
            const long long i1 = wl_scan3(q1, lane, LCT), i2 = wl_scan3(q2, lane, LCT);
            const long long my1 = __shfl_sync(0xffffffffu, i1 - q1, rank), my2 = __shfl_sync(0xffffffffu, i2 - q2, rank);
            const long long T1 = __shfl_sync(0xffffffffu, i1, CT - 1), T2 = __shfl_sync(0xffffffffu, i2, CT - 1);
            const int sh1 = min(KK1 - K1, 63), sh2 = min(KK2 - K2, 63);    // this CTA's shift into the chain scale
            const double unitq = wl_pow2(-WL_QBITS);
            S1 = rcp_bf(__dmul_rn((double)T1, unitq)); S2 = rcp_bf(__dmul_rn((double)T2, unitq));
            uni = !(T1 > 0) || KK1 >= WL_KNAN;                              // all -inf or a NaN: uniform weights (src/Filtering.py:24-25)
            nan2 = KK2 >= WL_KNAN;
            sc1 = wl_pow2(kw1 - KK1); sc2 = wl_pow2(kw2 - KK2);             // exact powers of two
            of1 = __dmul_rn((double)(my1 + (qw1 >> sh1)), unitq); of2 = __dmul_rn((double)(my2 + (qw2 >> sh2)), unitq);
            lo1 = wl_min1(__dmul_rn(of1, S1));
            hi1 = wl_min1(__dmul_rn(__dmul_rn((double)(my1 + (qn1 >> sh1)), unitq), S1));
            lo2 = __dmul_rn(of2, S2);
        }

        // ---- B: CDF values, offspring counts, scatter of (ancestor index -> trace, l_aux -> owner of the point)
        {
            const int end_of_unit = min(first_of_unit + 32 * PPT, N);
            if (uni) { lo1 = div_by_count((double)min(first_of_unit, N), dN, rN); hi1 = div_by_count((double)end_of_unit, dN, rN); }
            double W[PPT];
            bool below[PPT];
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                W[u] = uni ? div_by_count((double)(g0 + u + 1), dN, rN) : wl_min1(__dmul_rn(__dadd_rn(of1, __dmul_rn(sc1, s1[u])), S1));
                below[u] = (u < nvalid) && !nan2 && (__dmul_rn(__dadd_rn(of2, __dmul_rn(sc2, s2[u])), S2) < uanc);
            }
            int cend[PPT];
            int cp = wl_offspring_ranges<PPT>(W, lo1, hi1, mine == 0, end_of_unit >= N, first_of_unit >= N, g0, N, ures, dN, rN, lane, cend);
            // reference ancestor (src/PGAS.py:118-124): the first particle whose second CDF value is not below u_anc
            bool bp = __shfl_up_sync(0xffffffffu, (int)below[PPT - 1], 1) != 0;
            bp = lane ? bp : (mine == 0 ? true : (!nan2 && lo2 < uanc));
            int cand = WL_NOCAND;
            double cand_la = 0.0;
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                if (u < nvalid && bp && !below[u] && cand == WL_NOCAND) { cand = g0 + u; cand_la = la[u]; }
                bp = below[u];
            }
            {   // one record per warp to the owner of particle N-1 (data-independent byte count)
                const unsigned has = __ballot_sync(0xffffffffu, cand != WL_NOCAND);
                const int src = has ? __ffs(has) - 1 : 0;
                const int ck = __shfl_sync(0xffffffffu, cand, src);
                const double cl = __shfl_sync(0xffffffffu, cand_la, src);
                if (lane == 0) wl_st_async2(r_ref + (uint32_t)((par * U + mine) * 16), (double)ck, cl, r_mbar_last + 8 * (2 + par));
            }
            // offspring of particle k: points [cp, cend)
            auto emit = [&](int j, int k, double lak) {
                const int cj = j / P;
                const uint32_t dst = wl_mapa(s_lauxg + (uint32_t)((par * P + (j - cj * P)) * 8), (uint32_t)cj);
                wl_st_async1(dst, lak, wl_mapa(s_mbar + 8 * (2 + par), (uint32_t)cj));
                if (j != N - 1) anc_row[j] = k;                             // the ancestor of particle N-1 is set by its owner (:127)
            };
            int start[PPT], cnt[PPT];
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                start[u] = cp;
                cnt[u] = (u < nvalid) ? cend[u] - cp : 0;
                cp = (u < nvalid) ? cend[u] : cp;
#pragma unroll
                for (int q = 0; q < WL_INLINE; ++q)
                    if (q < cnt[u]) emit(start[u] + q, g0 + u, la[u]);
            }
            // particles with many offspring: the whole warp writes their ranges
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                unsigned heavy = __ballot_sync(0xffffffffu, cnt[u] > WL_INLINE);
                while (heavy) {
                    const int src = __ffs(heavy) - 1;
                    heavy &= heavy - 1;
                    const int s0 = __shfl_sync(0xffffffffu, start[u], src), n = __shfl_sync(0xffffffffu, cnt[u], src);
                    const int kk = __shfl_sync(0xffffffffu, g0 + u, src);
                    const double lk = __shfl_sync(0xffffffffu, la[u], src);
                    for (int q = WL_INLINE + lane; q < n; q += 32) emit(s0 + q, kk, lk);
                }
            }
        }
        WL_TICK(4);
        wl_wait(s_mbar + 8 * (2 + par), ph);
        WL_TICK(5);

        // ---- C: new log-weights (src/PGAS.py:137-147); the owner of particle N-1 installs the reference ancestor
        {
            const double* lg = lauxg + (size_t)par * P + il0;
            double g[PPT];
#pragma unroll
            for (int u = 0; u < PPT; ++u) g[u] = (u < nvalid) ? lg[u] : 0.0;
            const int lastl = (N - 1) - c_last * P;                         // local index of particle N-1 in its CTA
            if (rank == c_last && warp == lastl / (32 * PPT)) {             // warp-uniform: the warp that owns particle N-1
                const double* rm = refmsg + (size_t)par * U * 2;
                int kmin = WL_NOCAND;
                double lmin = 0.0;
                for (int u = lane; u < U; u += 32) {
                    const double2 r = *reinterpret_cast<const double2*>(rm + (size_t)u * 2);
                    const int k = (int)r.x;
                    if (k < kmin) { kmin = k; lmin = r.y; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const int ok = __shfl_xor_sync(0xffffffffu, kmin, o);
                    const double ol = __shfl_xor_sync(0xffffffffu, lmin, o);
                    if (ok < kmin) { kmin = ok; lmin = ol; }
                }
                const int ul = lastl - il0;
                if (ul >= 0 && ul < PPT) {
                    // no particle reached u_anc: searchsorted returns N, the gather clamps to N-1 = this thread's own particle
                    anc_row[N - 1] = (kmin == WL_NOCAND) ? N : kmin;
#pragma unroll
                    for (int u = 0; u < PPT; ++u)
                        if (u == ul) g[u] = (kmin == WL_NOCAND) ? la[u] : lmin;
                }
            }
#pragma unroll
            for (int u = 0; u < PPT; ++u)
                if (u < nvalid) logw[u] = ll[u] - g[u];
        }
        WL_TICK(6);
        anc_row += N;
    }
    if (a.logw_last) {
#pragma unroll
        for (int u = 0; u < PPT; ++u)
            if (u < nvalid) a.logw_last[(size_t)chain * N + g0 + u] = logw[u];
    }
    cluster_arrive();                                                       // no CTA exits while peers may still write into it
    cluster_wait();
}

static size_t weights_lat_smem(int C, int ppt) {
    const size_t NW = WL_NT / 32, U = (size_t)C * NW, P = (size_t)WL_NT * ppt;
    return (2 * NW * 4 + 2 * WL_MAXC * 4 + 2 * P + 2 * U * 2 + 4) * sizeof(double) + 4 * sizeof(unsigned long long) + 16;
}

// particles per thread: the fewest of 2 | 4 | 8 with which a PORTABLE cluster (<= 8 CTAs of 256 threads) covers the chain — N <= 4096,
// 8192, 16384; clusters of 16 are co-resident only 7 at a time on B200 (configs[4], 16 chains: they ran in three waves, 404 ms per
// iteration) — beyond that 8 with up to 16 CTAs (N <= 32768).  PGAS_WL_PPT_MIN=2|4|8: developer override (tests: ragged tails, big units)
static int weights_lat_ppt(int N) {
    int lo = WL_PPT;
    if (const char* e = getenv("PGAS_WL_PPT_MIN")) { const int v = atoi(e); if (v == 2 * WL_PPT || v == 4 * WL_PPT) lo = v; }
    for (int ppt = lo; ppt < 4 * WL_PPT; ppt *= 2)
        if ((N + WL_NT * ppt - 1) / (WL_NT * ppt) <= 8) return ppt;
    return 4 * WL_PPT;
}

// Does a CTA of the latency form leave room on its SM for a 256-thread state CTA (128 registers = half of the register file)?  Only the
// eight-particle instantiation is capped at 128 registers; with two / four particles per thread the kernel takes 146+ registers
// (39 K of the SM's 64 K: faster by itself, and right next to the 64-thread state CTAs of launches that do not fill the GPU).
bool pgas_weights_lat_fits_beside_big_state(int N) { return weights_lat_ppt(N) >= 4 * WL_PPT; }

// cluster size of the latency form for N particles (0: not applicable)
int pgas_weights_lat_cluster(int N) {
    const int P = WL_NT * weights_lat_ppt(N);
    const int C = (N + P - 1) / P;
    return (N >= 64 && C <= WL_MAXC) ? C : 0;
}

template <int PPT>
static int weights_lat_launch(const SweepArgs& a, cudaStream_t stream) {
    auto kern = a.C <= 1 ? csmc_weights_lat_kernel<WL_NT, PPT, 1> : a.C <= 2 ? csmc_weights_lat_kernel<WL_NT, PPT, 2>
              : a.C <= 4 ? csmc_weights_lat_kernel<WL_NT, PPT, 4> : a.C <= 8 ? csmc_weights_lat_kernel<WL_NT, PPT, 8>
                                                                                 : csmc_weights_lat_kernel<WL_NT, PPT, 16>;
    size_t smem = weights_lat_smem(a.C, PPT);
    if (const char* e = getenv("PGAS_WL_SMEM_KB")) smem = std::max(smem, (size_t)atoi(e) * 1024);   // developer knob: a large request keeps other CTAs off the SM
    PGAS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (a.C > 8) PGAS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(a.C * a.n_chains), 1, 1);
    cfg.blockDim = dim3(WL_NT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PGAS_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    __atomic_add_fetch(&g_pgas_launches, 1, __ATOMIC_RELAXED);
    return 0;
}

int pgas_launch_weights_lat(const SweepArgs& a, cudaStream_t stream) {
    const int ppt = weights_lat_ppt(a.N);
    return ppt == WL_PPT ? weights_lat_launch<WL_PPT>(a, stream) : ppt == 2 * WL_PPT ? weights_lat_launch<2 * WL_PPT>(a, stream)
                                                                                   : weights_lat_launch<4 * WL_PPT>(a, stream);
}
