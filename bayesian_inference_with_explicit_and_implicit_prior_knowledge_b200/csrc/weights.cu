// weights.cu — the weight recursion of the split sweep (sm_100a): a dedicated resampling kernel.
//
// In the reference's semantics (src/PGAS.py:79-153) the states never see an ancestor index, so csmc_state_kernel
// (sweep.cu) runs ahead and leaves per particle and step  l_aux = log p(y_t | mu),  h = log N(x_ref,t; mu, Sigma)  and
// ll = log p(y_t | x_t).  What remains sequential in t is
//   w_aux = softmax(l_aux + logw)                      a    = systematic_SISR(u_res, w_aux)      (:92-106, Filtering.py:6-37)
//   w_anc = softmax(l_aux + logw + h)                  a_N-1 = searchsorted(cumsum(w_anc), u_anc) (:109-127)
//   logw' = ll - l_aux[a]                                                                       (:137-147)
// One cluster of C CTAs owns a chain for all steps of a launch.  Compared with csmc_sweep_kernel<PRE> (the general
// kernel, still used when a chain's CDF does not fit one CTA's shared memory) this kernel
//   * gives every thread PPT CONSECUTIVE particles: the log-weights live in registers for the whole launch, the
//     prefix sums are thread-serial + ONE warp scan per CDF (instead of PPT), a warp is one softmax unit;
//   * all-gathers the warp pairs of the whole cluster (DSMEM) and lets EVERY warp fold them redundantly with the same
//     shuffles: one fold level, no second barrier, no warp-0 serial section;
//   * REPLICATES the finished CDF and l_aux of the whole chain in every CTA of the cluster (DSMEM stores in phase B1):
//     each CTA then resamples its OWN P points against the full CDF — the work is balanced whatever the weights look
//     like (with owner-of-segment resampling a CTA holding most of the mass did most of the searches while its peer
//     waited: 11 k of 33 k cycles per step on skewed weights), the gathered l_aux[a_j] lands in the thread that needs
//     it, and the step ends without a barrier;
//   * searches the CDF once per thread and then walks forward: the PPT points of a thread are consecutive.
// Per step: two cluster barriers, no CTA barrier.
#include <cooperative_groups.h>
#include <algorithm>
#include <stdlib.h>
#include "sweep_args.cuh"
#include "resample_math.cuh"

namespace cg = cooperative_groups;

// threads per CTA and consecutive particles per thread (compile-time knobs for occupancy experiments)
#ifndef PGAS_WK_NT
#define PGAS_WK_NT 512
#endif
#ifndef PGAS_WK_PPT
#define PGAS_WK_PPT 4
#endif
constexpr int WK_NT = PGAS_WK_NT, WK_PPT = PGAS_WK_PPT, WK_MAXC = 8;

__device__ __forceinline__ void wk_load_step_u(const SweepArgs& a, int chain, int t, double* dst) {
    if (a.rng_mode == 1) {
        const double* up = a.U + ((size_t)chain * a.var_rows + (t - a.row_off)) * 2;
        dst[0] = up[0];
        dst[1] = up[1];
    } else {
        philox_uniform2(a.seed, PURPOSE_STEP_U, a.chain_base + chain, a.iteration, (unsigned)t, 0u, dst[0], dst[1]);
    }
}

// PPT consecutive doubles of a row; vectorised when the row offset is 16-byte aligned
template <int PPT>
__device__ __forceinline__ void wk_load_row(const double* __restrict__ p, int nvalid, bool vec, double (&out)[PPT]) {
    if (vec && nvalid == PPT) {
#pragma unroll
        for (int u = 0; u < PPT; u += 2) {
            const double2 v = *reinterpret_cast<const double2*>(p + u);
            out[u] = v.x;
            out[u + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int u = 0; u < PPT; ++u) out[u] = (u < nvalid) ? p[u] : 0.0;
    }
}

template <int NT, int PPT>
__global__ void __launch_bounds__(NT, 2) csmc_weights_kernel(const __grid_constant__ SweepArgs a) {
    constexpr int NW = NT / 32;
    static_assert(NW <= 32 && PPT % 2 == 0 && 16 % PPT == 0, "paired loads; whole threads per 128-byte line");
    cg::cluster_group cluster = cg::this_cluster();
    const int C = a.C, N = a.N, P = a.P;
    const int rank = (C > 1) ? (int)cluster.block_rank() : 0;
    const int chain = blockIdx.x / C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int base = rank * P;
    const int Pc = max(0, min(P, N - base));
    const int il0 = tid * PPT;
    const int nvalid = max(0, min(PPT, Pc - il0));
    const int nblkN = (N + 255) / 256;
    const double dN = (double)N, rN = 1.0 / (double)N;
    const bool vec = ((N & 1) == 0) && ((P & 1) == 0);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* b1f = reinterpret_cast<double*>(smem_raw);            // CDF of the whole chain, padded with +inf
    double* lauxf = b1f + (size_t)nblkN * 256 + 8;                // l_aux of the whole chain
    double* unit = lauxf + (((size_t)N + 1) & ~(size_t)1);        // per warp of the CLUSTER: (max1, sum1, max2, sum2), all-gathered
    double* su = unit + WK_MAXC * NW * 4;                              // (u_res, u_anc), double-buffered by step parity
    int* cnt = reinterpret_cast<int*>(su + 4);                    // particles of this CTA below u_anc, triple-buffered

    for (int r = tid; r < nblkN * 256 + 8; r += NT) b1f[r] = INFINITY;
    if (tid == 0) {
        cnt[0] = cnt[1] = cnt[2] = 0;
        wk_load_step_u(a, chain, a.t_begin, su + 2 * (a.t_begin & 1));
    }
    double logw[PPT];
#pragma unroll
    for (int u = 0; u < PPT; ++u) logw[u] = (u < nvalid && a.init_logw) ? a.init_logw[(size_t)chain * N + base + il0 + u] : 0.0;
    if (C > 1) cluster.sync(); else __syncthreads();

#define WK_TICK(K) do { if (a.dbg && blockIdx.x < 2 && tid == 96) a.dbg[((size_t)(t - a.t_begin) * 2 + blockIdx.x) * 8 + (K)] = clock64(); } while (0)
    for (int t = a.t_begin; t < a.t_end; ++t) {
        WK_TICK(0);
        if (tid == 0) {
            if (t + 1 < a.t_end) wk_load_step_u(a, chain, t + 1, su + 2 * ((t + 1) & 1));
            cnt[(t + 1) % 3] = 0;
        }
        // ---- A: first-stage log-weights, softmax numerators with a warp-local shift, prefix sums
        const size_t prow = ((size_t)chain * a.pre_rows + (size_t)(t - a.pre_off)) * N + base + il0;
        double la[PPT], lr[PPT];
        wk_load_row<PPT>(a.pre_la + prow, nvalid, vec, la);
        wk_load_row<PPT>(a.pre_lr + prow, nvalid, vec, lr);
        if (t + 1 < a.t_end && nvalid > 0 && (tid & (16 / PPT - 1)) == 0) {    // next step's rows towards L1 (one request per 128-byte line)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(a.pre_la + prow + N));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(a.pre_lr + prow + N));
        }
        if (nvalid > 0 && (tid & (16 / PPT - 1)) == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.pre_ll + prow));
        double s1[PPT], s2[PPT];
        double m1t = -INFINITY, m2t = -INFINITY;
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            const double lwa = (u < nvalid) ? la[u] + logw[u] : -INFINITY;
            const double lwr = (u < nvalid) ? lwa + lr[u] : -INFINITY;
            s1[u] = lwa;
            s2[u] = lwr;
            m1t = fmax(m1t, lwa);
            m2t = fmax(m2t, lwr);
        }
        const double m1w = warp_shift_max(m1t), m2w = warp_shift_max(m2t);
        // a warp whose log-weights are all -inf is an empty unit (numerators 0, not exp(-inf + inf) = NaN): jax.nn.softmax gives
        // such entries weight 0 as long as the chain has one finite entry
        const double sh1 = (m1w == -INFINITY) ? 0.0 : m1w, sh2 = (m2w == -INFINITY) ? 0.0 : m2w;
        {
            double r1 = 0.0, r2 = 0.0;
#pragma unroll
            for (int u = 0; u < PPT; ++u) {                        // thread-serial inclusive prefix of the numerators
                const double e1 = (u < nvalid) ? exp_neg_bf(s1[u] - sh1) : 0.0;
                const double e2 = (u < nvalid) ? exp_neg_bf(s2[u] - sh2) : 0.0;
                r1 = __dadd_rn(r1, e1);
                r2 = __dadd_rn(r2, e2);
                s1[u] = r1;
                s2[u] = r2;
            }
            const double i1 = warp_scan_incl(r1, lane), i2 = warp_scan_incl(r2, lane);
            double x1 = __shfl_up_sync(0xffffffffu, i1, 1), x2 = __shfl_up_sync(0xffffffffu, i2, 1);
            x1 = lane ? x1 : 0.0;
            x2 = lane ? x2 : 0.0;
#pragma unroll
            for (int u = 0; u < PPT; ++u) { s1[u] = __dadd_rn(x1, s1[u]); s2[u] = __dadd_rn(x2, s2[u]); }   // in-warp inclusive prefix
            if (lane == 31) {                                      // this warp's (max, sum) pairs -> every CTA of the cluster
                const bool any = warp * 32 * PPT < Pc;
                const double u0 = any ? m1w : -INFINITY, u1 = any ? i1 : 0.0, u2 = any ? m2w : -INFINITY, u3 = any ? i2 : 0.0;
                for (int c = 0; c < C; ++c) {
                    double* up = ((c == rank) ? unit : cluster.map_shared_rank(unit, c)) + (rank * NW + warp) * 4;
                    *reinterpret_cast<double2*>(up) = make_double2(u0, u1);
                    *reinterpret_cast<double2*>(up + 2) = make_double2(u2, u3);
                }
            }
        }
        WK_TICK(1);
        if (C > 1) { cluster_arrive(); cluster_wait(); } else __syncthreads();
        WK_TICK(2);

        // ---- X1: every warp folds the C * NW warp pairs of the chain (online-softmax rescaling) with the same shuffles —
        //      same bits in every warp of every CTA, no second barrier
        const double ures = su[2 * (t & 1)], uanc = su[2 * (t & 1) + 1];
        double fw1 = 0.0, gw1 = 0.0, fw2 = 0.0, gw2 = 0.0;          // rescale factor and exclusive offset of this thread's warp
        double S1, S2;                                             // reciprocals of the normalisers
        bool uni;                                                  // sum of the numerators not > 0 (all -inf, or a NaN): systematic_SISR
                                                                   // falls back to uniform weights (src/Filtering.py:24-25)
        {
            const int U = C * NW, mine = rank * NW + warp;
            double m1g = -INFINITY, m2g = -INFINITY;
            for (int u = lane; u < U; u += 32) { m1g = fmax(m1g, unit[u * 4]); m2g = fmax(m2g, unit[u * 4 + 2]); }
            m1g = warp_shift_max(m1g);                             // a shift, not the exact maximum: one 32-bit redux each
            m2g = warp_shift_max(m2g);
            double c1 = 0.0, c2 = 0.0;
            for (int u0 = 0; u0 < U; u0 += 32) {
                const int u = u0 + lane;
                const bool vu = u < U;
                const double mu1 = vu ? unit[u * 4] : -INFINITY, mu2 = vu ? unit[u * 4 + 2] : -INFINITY;
                const double f1 = (mu1 == -INFINITY) ? 0.0 : exp_neg_bf(mu1 - m1g);
                const double f2 = (mu2 == -INFINITY) ? 0.0 : exp_neg_bf(mu2 - m2g);
                const double v1 = vu ? __dmul_rn(unit[u * 4 + 1], f1) : 0.0, v2 = vu ? __dmul_rn(unit[u * 4 + 3], f2) : 0.0;
                const double i1 = warp_scan_incl(v1, lane), i2 = warp_scan_incl(v2, lane);
                double x1 = __shfl_up_sync(0xffffffffu, i1, 1), x2 = __shfl_up_sync(0xffffffffu, i2, 1);
                x1 = lane ? __dadd_rn(c1, x1) : c1;
                x2 = lane ? __dadd_rn(c2, x2) : c2;
                const int src = (mine - u0) & 31;
                const double tf1 = __shfl_sync(0xffffffffu, f1, src), tg1 = __shfl_sync(0xffffffffu, x1, src);
                const double tf2 = __shfl_sync(0xffffffffu, f2, src), tg2 = __shfl_sync(0xffffffffu, x2, src);
                if (mine >= u0 && mine < u0 + 32) { fw1 = tf1; gw1 = tg1; fw2 = tf2; gw2 = tg2; }
                c1 = __dadd_rn(c1, __shfl_sync(0xffffffffu, i1, 31));
                c2 = __dadd_rn(c2, __shfl_sync(0xffffffffu, i2, 31));
            }
            S1 = rcp_bf(c1);
            S2 = rcp_bf(c2);
            uni = !(c1 > 0.0);
        }

        WK_TICK(3);
        // ---- B1: finished CDF values and l_aux of this CTA's particles -> every CTA of the cluster
        {
            int mycnt = 0;
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                if (u < nvalid) {
                    // W = clip(cumsum(w / sum w), 0, 1)  (src/Filtering.py:23-32): chain-level inclusive prefix G_warp + s f_warp
                    const double p1 = __dadd_rn(gw1, __dmul_rn(s1[u], fw1));
                    const double p2 = __dadd_rn(gw2, __dmul_rn(s2[u], fw2));
                    s1[u] = uni ? div_by_count((double)(base + il0 + u + 1), dN, rN) : clip01(__dmul_rn(p1, S1));
                    // cumsum(softmax(lw_anc)) < u_anc  (src/PGAS.py:118-124), not clipped: the counts of all CTAs add up to
                    // searchsorted's result
                    mycnt += (__dmul_rn(p2, S2) < uanc) ? 1 : 0;
                }
            }
            for (int c = 0; c < C; ++c) {
                double* db = (c == rank) ? b1f : cluster.map_shared_rank(b1f, c);
                double* dl = (c == rank) ? lauxf : cluster.map_shared_rank(lauxf, c);
                if (vec && nvalid == PPT) {
#pragma unroll
                    for (int u = 0; u < PPT; u += 2) {
                        *reinterpret_cast<double2*>(db + base + il0 + u) = make_double2(s1[u], s1[u + 1]);
                        *reinterpret_cast<double2*>(dl + base + il0 + u) = make_double2(la[u], la[u + 1]);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < PPT; ++u)
                        if (u < nvalid) { db[base + il0 + u] = s1[u]; dl[base + il0 + u] = la[u]; }
                }
            }
            mycnt = __reduce_add_sync(0xffffffffu, mycnt);
            if (lane == 0 && mycnt) atomicAdd(&cnt[t % 3], mycnt);
        }
        double ll[PPT];
        wk_load_row<PPT>(a.pre_ll + prow, nvalid, vec, ll);      // in flight across the barrier
        WK_TICK(4);
        if (C > 1) { cluster_arrive(); cluster_wait(); } else __syncthreads();
        WK_TICK(5);

        // ---- B2 + C: this thread's PPT consecutive points against the full CDF (src/Filtering.py:28-35), the ancestor of
        //      the conditioned path (src/PGAS.py:118-127), new log-weights (:137-147)
        if (nvalid > 0) {
            int anc[PPT];
            int k = 0;
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                if (u < nvalid) {
                    const int j = base + il0 + u;
                    const double uj = strat_point(ures, j, dN, rN);
                    if (u == 0) {
                        k = count_below_binary(b1f, nblkN * 256, uj);
                    } else {                                       // CDF and points are sorted: walk forward from the previous answer
                        const int c4 = ((b1f[k] < uj) ? 1 : 0) + ((b1f[k + 1] < uj) ? 1 : 0) + ((b1f[k + 2] < uj) ? 1 : 0) +
                                       ((b1f[k + 3] < uj) ? 1 : 0);
                        k = (c4 == 4) ? count_below_binary(b1f, nblkN * 256, uj) : k + c4;
                    }
                    int kk = min(k, N - 1), av = kk;
                    if (j == N - 1) {                              // overwritten by the reference ancestor (:127)
                        int cv = 0;
                        for (int c = 0; c < C; ++c) cv += ((c == rank) ? cnt : cluster.map_shared_rank(cnt, c))[t % 3];
                        av = cv;                                   // may be N (cumsum never reached u_anc): the gather clamps
                        kk = min(cv, N - 1);
                    }
                    anc[u] = av;
                    logw[u] = ll[u] - lauxf[kk];
                }
            }
            int* anc_row = a.anc_trace + ((size_t)chain * a.anc_rows + (t - 1 - a.row_off + a.anc_shift)) * N + base + il0;
            if (nvalid == PPT && ((N & 3) == 0) && ((P & 3) == 0) && PPT % 4 == 0) {
#pragma unroll
                for (int u = 0; u < PPT; u += 4) *reinterpret_cast<int4*>(anc_row + u) = make_int4(anc[u], anc[u + 1], anc[u + 2], anc[u + 3]);
            } else {
#pragma unroll
                for (int u = 0; u < PPT; ++u)
                    if (u < nvalid) anc_row[u] = anc[u];
            }
            if (t + 1 == a.t_end && a.logw_last) {
#pragma unroll
                for (int u = 0; u < PPT; ++u)
                    if (u < nvalid) a.logw_last[(size_t)chain * N + base + il0 + u] = logw[u];
            }
        }
        WK_TICK(6);
        // no barrier here: the next step's phase A touches registers and `unit` only (its last readers passed the
        // barrier above); b1f / lauxf are rewritten after the next X1 cluster barrier, when every CTA has left B2
    }
    if (C > 1) cluster.sync();                                    // no CTA exits while peers may still read / push into its smem
}

// --------------------------------------------------------------------------------------------------------------------
// One CTA per chain (N <= H * NT * PPT): the CTA walks H slices of NT * PPT particles one after the other in every phase — the
// same arithmetic as a cluster of H CTAs (a slice is a softmax "rank": same units, same fold, same CDF expression), but the
// chain occupies ONE SM instead of H.  The resampling recursion is latency-bound, so running the slices back to back costs
// about what the DSMEM replication and the two release/acquire cluster barriers (MEMBAR.ALL.GPU + CCTL.IVALL each) cost the
// cluster form; what it buys is register space on H - 1 more SMs for the FP64-bound state kernel it shares the GPU with
// (a resident resampling CTA takes half of an SM's register file: only one state CTA fits next to it instead of two).
// Between the phases the per-particle prefixes wait in shared memory (registers: 64 per thread); three CTA barriers per step.
#ifndef PGAS_WK1_MINB
#define PGAS_WK1_MINB 2          // resident CTAs per SM the register allocation aims at (developer knob: 3 -> 40, 4 -> 32 registers)
#endif
template <int NT, int PPT, int H>
__global__ void __launch_bounds__(NT, PGAS_WK1_MINB) csmc_weights1_kernel(const __grid_constant__ SweepArgs a) {
    constexpr int NW = NT / 32, U = H * NW;
    static_assert(U <= 32, "one fold pass");
    const int N = a.N, P = NT * PPT;                              // slice length (the last slice may be ragged or empty)
    const int chain = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int il0 = tid * PPT;
    const int nblkN = (N + 255) / 256;
    const double dN = (double)N, rN = 1.0 / (double)N;
    const bool vec = (N & 1) == 0;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* b1f = reinterpret_cast<double*>(smem_raw);            // CDF of the chain, padded with +inf
    double* lauxf = b1f + (size_t)nblkN * 256 + 8;                // l_aux of the chain
    double* s1s = lauxf + (size_t)H * P;                          // in-warp inclusive prefixes of the two softmaxes
    double* s2s = s1s + (size_t)H * P;
    double* unit = s2s + (size_t)H * P;                           // per (slice, warp): (max1, sum1, max2, sum2)
    double* su = unit + U * 4;
    int* cnt = reinterpret_cast<int*>(su + 4);

    for (int r = tid; r < nblkN * 256 + 8; r += NT) b1f[r] = INFINITY;
    if (tid == 0) {
        cnt[0] = cnt[1] = cnt[2] = 0;
        wk_load_step_u(a, chain, a.t_begin, su + 2 * (a.t_begin & 1));
    }
    double logw[H][PPT];
    int nvalid[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
        nvalid[h] = max(0, min(PPT, N - h * P - il0));
#pragma unroll
        for (int u = 0; u < PPT; ++u)
            logw[h][u] = (u < nvalid[h] && a.init_logw) ? a.init_logw[(size_t)chain * N + h * P + il0 + u] : 0.0;
    }
    __syncthreads();

    for (int t = a.t_begin; t < a.t_end; ++t) {
        WK_TICK(0);
        if (tid == 0) {
            if (t + 1 < a.t_end) wk_load_step_u(a, chain, t + 1, su + 2 * ((t + 1) & 1));
            cnt[(t + 1) % 3] = 0;
        }
        const size_t prow0 = ((size_t)chain * a.pre_rows + (size_t)(t - a.pre_off)) * N + il0;
        // ---- A: per slice, first-stage log-weights, softmax numerators with a warp-local shift, prefix sums
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const int nv = nvalid[h];
            const size_t prow = prow0 + (size_t)h * P;
            double la[PPT], lr[PPT];
            wk_load_row<PPT>(a.pre_la + prow, nv, vec, la);
            wk_load_row<PPT>(a.pre_lr + prow, nv, vec, lr);
            if (nv > 0 && (tid & (16 / PPT - 1)) == 0) {          // no cluster barrier in this kernel: L1 prefetches survive
                if (t + 1 < a.t_end) {
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.pre_la + prow + N));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.pre_lr + prow + N));
                }
                asm volatile("prefetch.global.L1 [%0];" ::"l"(a.pre_ll + prow));
            }
            double s1[PPT], s2[PPT];
            double m1t = -INFINITY, m2t = -INFINITY;
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                const double lwa = (u < nv) ? la[u] + logw[h][u] : -INFINITY;
                const double lwr = (u < nv) ? lwa + lr[u] : -INFINITY;
                s1[u] = lwa;
                s2[u] = lwr;
                m1t = fmax(m1t, lwa);
                m2t = fmax(m2t, lwr);
            }
            const double m1w = warp_shift_max(m1t), m2w = warp_shift_max(m2t);
            const double sh1 = (m1w == -INFINITY) ? 0.0 : m1w, sh2 = (m2w == -INFINITY) ? 0.0 : m2w;     // all -inf: an empty unit, not NaN
            double r1 = 0.0, r2 = 0.0;
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                const double e1 = (u < nv) ? exp_neg_bf(s1[u] - sh1) : 0.0;
                const double e2 = (u < nv) ? exp_neg_bf(s2[u] - sh2) : 0.0;
                r1 = __dadd_rn(r1, e1);
                r2 = __dadd_rn(r2, e2);
                s1[u] = r1;
                s2[u] = r2;
            }
            const double i1 = warp_scan_incl(r1, lane), i2 = warp_scan_incl(r2, lane);
            double x1 = __shfl_up_sync(0xffffffffu, i1, 1), x2 = __shfl_up_sync(0xffffffffu, i2, 1);
            x1 = lane ? x1 : 0.0;
            x2 = lane ? x2 : 0.0;
            double* d1 = s1s + (size_t)h * P + il0;
            double* d2 = s2s + (size_t)h * P + il0;
            double* dl = lauxf + (size_t)h * P + il0;
#pragma unroll
            for (int u = 0; u < PPT; u += 2) {
                *reinterpret_cast<double2*>(d1 + u) = make_double2(__dadd_rn(x1, s1[u]), __dadd_rn(x1, s1[u + 1]));
                *reinterpret_cast<double2*>(d2 + u) = make_double2(__dadd_rn(x2, s2[u]), __dadd_rn(x2, s2[u + 1]));
                *reinterpret_cast<double2*>(dl + u) = make_double2(la[u], la[u + 1]);
            }
            if (lane == 31) {
                const bool any = h * P + warp * 32 * PPT < N;
                double* up = unit + (h * NW + warp) * 4;
                *reinterpret_cast<double2*>(up) = make_double2(any ? m1w : -INFINITY, any ? i1 : 0.0);
                *reinterpret_cast<double2*>(up + 2) = make_double2(any ? m2w : -INFINITY, any ? i2 : 0.0);
            }
        }
        WK_TICK(1);
        __syncthreads();
        WK_TICK(2);
        // ---- X1: every warp folds the H * NW (slice, warp) pairs
        const double ures = su[2 * (t & 1)], uanc = su[2 * (t & 1) + 1];
        double fw1[H], gw1[H], fw2[H], gw2[H], S1, S2;
        bool uni;                                                 // sum not > 0 (all -inf, or a NaN) -> uniform weights (src/Filtering.py:24-25)
        {
            const bool vu = lane < U;
            const double mu1 = vu ? unit[lane * 4] : -INFINITY, mu2 = vu ? unit[lane * 4 + 2] : -INFINITY;
            const double m1g = warp_shift_max(mu1), m2g = warp_shift_max(mu2);
            const double f1 = (mu1 == -INFINITY) ? 0.0 : exp_neg_bf(mu1 - m1g);
            const double f2 = (mu2 == -INFINITY) ? 0.0 : exp_neg_bf(mu2 - m2g);
            const double v1 = vu ? __dmul_rn(unit[lane * 4 + 1], f1) : 0.0, v2 = vu ? __dmul_rn(unit[lane * 4 + 3], f2) : 0.0;
            const double i1 = warp_scan_incl(v1, lane), i2 = warp_scan_incl(v2, lane);
            double x1 = __shfl_up_sync(0xffffffffu, i1, 1), x2 = __shfl_up_sync(0xffffffffu, i2, 1);
            x1 = lane ? x1 : 0.0;
            x2 = lane ? x2 : 0.0;
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const int src = h * NW + warp;
                fw1[h] = __shfl_sync(0xffffffffu, f1, src); gw1[h] = __shfl_sync(0xffffffffu, x1, src);
                fw2[h] = __shfl_sync(0xffffffffu, f2, src); gw2[h] = __shfl_sync(0xffffffffu, x2, src);
            }
            const double c1 = __shfl_sync(0xffffffffu, i1, 31);
            S1 = rcp_bf(c1);
            S2 = rcp_bf(__shfl_sync(0xffffffffu, i2, 31));
            uni = !(c1 > 0.0);
        }
        WK_TICK(3);
        // ---- B1: finished CDF values; particles below u_anc
        {
            int mycnt = 0;
#pragma unroll
            for (int h = 0; h < H; ++h) {
                if (nvalid[h] == PPT) {
                    double w[PPT];
#pragma unroll
                    for (int u = 0; u < PPT; u += 2) {
                        const double2 a1 = *reinterpret_cast<const double2*>(s1s + (size_t)h * P + il0 + u);
                        const double2 a2 = *reinterpret_cast<const double2*>(s2s + (size_t)h * P + il0 + u);
                        w[u] = clip01(__dmul_rn(__dadd_rn(gw1[h], __dmul_rn(a1.x, fw1[h])), S1));
                        w[u + 1] = clip01(__dmul_rn(__dadd_rn(gw1[h], __dmul_rn(a1.y, fw1[h])), S1));
                        if (uni) {
                            w[u] = div_by_count((double)(h * P + il0 + u + 1), dN, rN);
                            w[u + 1] = div_by_count((double)(h * P + il0 + u + 2), dN, rN);
                        }
                        mycnt += (__dmul_rn(__dadd_rn(gw2[h], __dmul_rn(a2.x, fw2[h])), S2) < uanc) ? 1 : 0;
                        mycnt += (__dmul_rn(__dadd_rn(gw2[h], __dmul_rn(a2.y, fw2[h])), S2) < uanc) ? 1 : 0;
                        *reinterpret_cast<double2*>(b1f + (size_t)h * P + il0 + u) = make_double2(w[u], w[u + 1]);
                    }
                } else {
                    for (int u = 0; u < nvalid[h]; ++u) {
                        const size_t e = (size_t)h * P + il0 + u;
                        b1f[e] = uni ? div_by_count((double)(e + 1), dN, rN) : clip01(__dmul_rn(__dadd_rn(gw1[h], __dmul_rn(s1s[e], fw1[h])), S1));
                        mycnt += (__dmul_rn(__dadd_rn(gw2[h], __dmul_rn(s2s[e], fw2[h])), S2) < uanc) ? 1 : 0;
                    }
                }
            }
            mycnt = __reduce_add_sync(0xffffffffu, mycnt);
            if (lane == 0 && mycnt) atomicAdd(&cnt[t % 3], mycnt);
        }
        double ll[H][PPT];                                        // in flight across the barrier
#pragma unroll
        for (int h = 0; h < H; ++h) wk_load_row<PPT>(a.pre_ll + prow0 + (size_t)h * P, nvalid[h], vec, ll[h]);
        WK_TICK(4);
        __syncthreads();
        WK_TICK(5);
        // ---- B2 + C: per slice, this thread's PPT consecutive points against the full CDF, reference ancestor, new log-weights
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const int nv = nvalid[h];
            if (nv > 0) {
                int anc[PPT];
                int k = 0;
#pragma unroll
                for (int u = 0; u < PPT; ++u) {
                    if (u < nv) {
                        const int j = h * P + il0 + u;
                        const double uj = strat_point(ures, j, dN, rN);
                        if (u == 0) {
                            k = count_below_binary(b1f, nblkN * 256, uj);
                        } else {
                            const int c4 = ((b1f[k] < uj) ? 1 : 0) + ((b1f[k + 1] < uj) ? 1 : 0) + ((b1f[k + 2] < uj) ? 1 : 0) +
                                           ((b1f[k + 3] < uj) ? 1 : 0);
                            k = (c4 == 4) ? count_below_binary(b1f, nblkN * 256, uj) : k + c4;
                        }
                        int kk = min(k, N - 1), av = kk;
                        if (j == N - 1) {                          // overwritten by the reference ancestor (src/PGAS.py:127)
                            av = cnt[t % 3];                       // may be N: the gather clamps
                            kk = min(av, N - 1);
                        }
                        anc[u] = av;
                        logw[h][u] = ll[h][u] - lauxf[kk];
                    }
                }
                int* anc_row = a.anc_trace + ((size_t)chain * a.anc_rows + (t - 1 - a.row_off + a.anc_shift)) * N + h * P + il0;
                if (nv == PPT && ((N & 3) == 0) && PPT % 4 == 0) {
#pragma unroll
                    for (int u = 0; u < PPT; u += 4) *reinterpret_cast<int4*>(anc_row + u) = make_int4(anc[u], anc[u + 1], anc[u + 2], anc[u + 3]);
                } else {
#pragma unroll
                    for (int u = 0; u < PPT; ++u)
                        if (u < nv) anc_row[u] = anc[u];
                }
                if (t + 1 == a.t_end && a.logw_last) {
#pragma unroll
                    for (int u = 0; u < PPT; ++u)
                        if (u < nv) a.logw_last[(size_t)chain * N + h * P + il0 + u] = logw[h][u];
                }
            }
        }
        WK_TICK(6);
        __syncthreads();                                          // lauxf / s1s / s2s / unit are rewritten by the next step's phase A
    }
}

template <int H>
static size_t weights1_smem_bytes(int N) {
    const size_t nblkN = ((size_t)N + 255) / 256, P = (size_t)WK_NT * WK_PPT;
    return (nblkN * 256 + 8 + 3 * H * P + H * (WK_NT / 32) * 4 + 4) * sizeof(double) + 4 * sizeof(int) + 16;
}

static size_t weights_smem_bytes(int N) {
    const size_t nblkN = ((size_t)N + 255) / 256;
    return (nblkN * 256 + 8 + (((size_t)N + 1) & ~(size_t)1) + WK_MAXC * (WK_NT / 32) * 4 + 4) * sizeof(double) + 4 * sizeof(int) + 16;
}

// cluster size of the dedicated kernel for N particles (0: not applicable -> csmc_sweep_kernel<PRE>).  Chains of up to
// 2 * NT * PPT particles run in ONE CTA (csmc_weights1_kernel, two slices), larger ones in a cluster (csmc_weights_kernel).
int pgas_weights_cluster(int N) {
    if (const char* e = getenv("PGAS_WEIGHTS_KERNEL")) { if (*e && atoi(e) == 0) return 0; }      // developer override
    if (N < 64 || weights_smem_bytes(N) > 200 * 1024) return 0;
    const bool one_cta = !(getenv("PGAS_WEIGHTS_KERNEL") && atoi(getenv("PGAS_WEIGHTS_KERNEL")) == 2);   // 2: force the cluster form
    if (one_cta && N <= 2 * WK_NT * WK_PPT) return 1;
    int C = 1;
    while (C * WK_NT * WK_PPT < N) C *= 2;
    return C <= WK_MAXC ? C : 0;
}

template <typename K>
static int weights_launch(K kern, const SweepArgs& a, int C, size_t smem, cudaStream_t stream) {
    PGAS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(C * a.n_chains), 1, 1);
    cfg.blockDim = dim3(WK_NT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PGAS_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    __atomic_add_fetch(&g_pgas_launches, 1, __ATOMIC_RELAXED);
    return 0;
}

int pgas_launch_weights(const SweepArgs& a, cudaStream_t stream) {
    if (a.C == 1 && a.N > WK_NT * WK_PPT) return weights_launch(csmc_weights1_kernel<WK_NT, WK_PPT, 2>, a, 1, weights1_smem_bytes<2>(a.N), stream);
    if (a.C == 1) return weights_launch(csmc_weights1_kernel<WK_NT, WK_PPT, 1>, a, 1, weights1_smem_bytes<1>(a.N), stream);
    return weights_launch(csmc_weights_kernel<WK_NT, WK_PPT>, a, a.C, weights_smem_bytes(a.N), stream);
}
