// sweep.cu — conditional-SMC sweep with ancestor sampling (sm_100a): the two schedules of condSequentialMonteCarlo.step / __call__
// (reference src/PGAS.py:79-153, :176-228) and their launcher.
//
// SPLIT FORM (default for two-dimensional bases; second half of this file): in the reference's semantics the states never see an
// ancestor index, so csmc_state_kernel propagates all particles ahead in launches of <= 16 steps (no synchronisation; particles in
// registers; traces and three log-densities per particle-step to HBM) while a resampling kernel runs the weight recursion
// chunk-wise behind it on a second stream: weights_lat.cu (cluster per chain, up to 32 chains per launch), weights.cu (one CTA per
// chain, many chains) or csmc_sweep_kernel<PRE> below (chains whose CDF does not fit one CTA).  About 290 launches per sweep at
// T = 2000.
//
// FUSED FORM (csmc_sweep_kernel<PRE = false>; textbook ancestor gather, D = 1, likelihood programs, single steps, no workspace): one CTA — or one
// thread-block cluster of C CTAs — owns one chain for all T steps of a launch; the particle set (state, log-weight, auxiliary mean,
// CDF) lives in shared memory / DSMEM and the only HBM traffic is the trace.  Each step, in one pass:
//   A  mu_i = Theta phi(x_{t-1}^i, u_t)  (basis_eval.cuh), l_aux, h;  warp-local softmax shift + exp + scans
//   X1 per-CTA (max, sum) pairs all-gathered through DSMEM  -> cluster barrier #1
//   B  global log-sum-exp / CDF offsets; systematic resampling by the owner of the CDF segment
//      (search in local shared memory), ancestor sampling of the reference particle;
//      l_aux[a_j] (and mu[a_j] in gather mode) pushed to the owner of j through DSMEM
//   X2 cluster barrier #2 — its latency is covered by the noise draw and the new state
//   C  x_t^i = mu + chol(Sigma) z,  logw_t^i = log p(y_t|x_t^i) - l_aux[a_i],  trace row written
#include <cooperative_groups.h>
#include <algorithm>
#include <vector>
#include <stdlib.h>
#include "basis_eval.cuh"
#include "basis_rowwalk.cuh"
#include "basis_mma.cuh"
#include "sweep_args.cuh"
#include "resample_math.cuh"

namespace cg = cooperative_groups;

// threads per CTA of the state kernel / the resampling kernel of the split form (compile-time knobs for occupancy experiments)
#ifndef PGAS_ST_NT
#define PGAS_ST_NT 256
#endif
#ifndef PGAS_PRE_NT
#define PGAS_PRE_NT 512
#endif

constexpr int MAXC = 16;         // largest (non-portable) cluster



// log N(y; H x + h0, R) with e = Rw (y - mean)   (src/StateSpaceModel.py:83-87 semantics, src/EMPS.py:250-252)
template <int NX, int NY>
__device__ __forceinline__ double gauss_loglik(const DevModel& m, const double* __restrict__ y, const double x[NX]) {
    double d[NY];
#pragma unroll
    for (int r = 0; r < NY; ++r) {
        double mean = m.h0[r];
#pragma unroll
        for (int k = 0; k < NX; ++k) mean = fma(m.H[r][k], x[k], mean);
        d[r] = y[r] - mean;
    }
    double q = 0.0;
#pragma unroll
    for (int r = 0; r < NY; ++r) {
        double e = 0.0;
#pragma unroll
        for (int c = 0; c <= r; ++c) e = fma(m.Rw[r][c], d[c], e);
        q = fma(e, e, q);
    }
    return m.R_logc - 0.5 * q;
}

// likelihood_fcn of any family (src/PGAS.py:93-100, :137-147: observations[time], state, inputs[time]): the Gaussian family above or
// an expression program (model plug-in).  Fused kernel only — the split form's state kernel is specialised for the Gaussian family
// and pgas_sweep_split_eligible sends models with a likelihood program here.
template <int NX, int NY>
__device__ __forceinline__ double loglik_any(const DevModel& m, int t, const double* y, const double x[NX]) {
    if (m.lik_len) return pgas_lik_program(&m, x, m.inputs + (size_t)t * m.n_u, y);
    return gauss_loglik<NX, NY>(m, y, x);
}

// log N(ref; mu, Sigma) with e = Sw (ref - mu)   (src/PGAS.py:109-116)
template <int NX>
__device__ __forceinline__ double gauss_logpdf_state(const double* __restrict__ Sw, double logc, const double* __restrict__ ref,
                                                     const double mu[NX]) {
    double d[NX];
#pragma unroll
    for (int k = 0; k < NX; ++k) d[k] = ref[k] - mu[k];
    double q = 0.0;
#pragma unroll
    for (int r = 0; r < NX; ++r) {
        double e = 0.0;
#pragma unroll
        for (int c = 0; c <= r; ++c) e = fma(Sw[r * NX + c], d[c], e);
        q = fma(e, e, q);
    }
    return logc - 0.5 * q;
}

// per-step constants staged in shared memory (double-buffered by step parity)
struct StepConst {
    double y[PGAS_MAX_NY];
    double u[PGAS_MAX_NU];
    double ref[PGAS_MAX_NX];
    double cz[PGAS_MAX_D];       // input-dependent constant part of the affine GP-input map
    double ures, uanc;
};

template <int NX>
__device__ __forceinline__ void draw_normals(const SweepArgs& a, int chain, int t, int i, double z[NX]) {
    if (a.rng_mode == 1) {
        const double* zp = a.Z + (((size_t)chain * a.var_rows + (t - a.row_off)) * a.N + i) * NX;
#pragma unroll
        for (int k = 0; k < NX; ++k) z[k] = zp[k];
    } else {
#pragma unroll
        for (int k = 0; k < NX; k += 2) {
            double za, zb;
            philox_normal2(a.seed, PURPOSE_STATE, a.chain_base + chain, a.iteration, (unsigned)t | ((unsigned)(k >> 1) << 28),
                           (unsigned)i, za, zb);
            z[k] = za;
            if (k + 1 < NX) z[k + 1] = zb;
        }
    }
}

__device__ __forceinline__ void load_step_const(const SweepArgs& a, int chain, int t, StepConst* sc) {
    const DevModel& m = a.m;
    const int tin = (m.flags & PGAS_FLAG_INPUT_PREV) ? t - 1 : t;     // quirk (ii), src/PGAS.py:52-54
    for (int r = 0; r < m.n_y; ++r) sc->y[r] = m.obs[(size_t)t * m.n_y + r];
    for (int k = 0; k < m.n_u; ++k) sc->u[k] = m.inputs[(size_t)tin * m.n_u + k];
    for (int d = 0; d < m.D; ++d) {
        double acc = m.bz[d];
        for (int k = 0; k < m.n_u; ++k) acc = fma(m.Az[d][m.n_x + k], sc->u[k], acc);
        sc->cz[d] = acc;
    }
    for (int k = 0; k < m.n_x; ++k) sc->ref[k] = a.ref ? a.ref[(size_t)chain * a.ref_stride + (size_t)(t - a.row_off) * m.n_x + k] : 0.0;
    if (a.rng_mode == 1) {
        const double* up = a.U + ((size_t)chain * a.var_rows + (t - a.row_off)) * 2;
        sc->ures = up[0];
        sc->uanc = up[1];
    } else {
        philox_uniform2(a.seed, PURPOSE_STEP_U, a.chain_base + chain, a.iteration, (unsigned)t, 0u, sc->ures, sc->uanc);
    }
}

template <int NX, int NY, int D, int NT, bool PRE>
__global__ void __launch_bounds__(NT, PRE ? 1024 / NT : (NT == 256 ? 2 : 1)) csmc_sweep_kernel(const __grid_constant__ SweepArgs a) {
    constexpr int NW = NT / 32;
    const DevModel& m = a.m;
    cg::cluster_group cluster = cg::this_cluster();
    const int C = a.C;
    const int rank = (C > 1) ? (int)cluster.block_rank() : 0;
    const int chain = blockIdx.x / C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = a.P, N = a.N;
    const int base = rank * P;
    const int Pc = max(0, min(P, N - base));          // valid particles of this CTA
    const int PPT = (P + NT - 1) / NT;
    const bool gather = (m.flags & PGAS_FLAG_ANCESTOR_GATHER) != 0;
    const double dN = (double)N, rN = 1.0 / (double)N;

    // ------------------------------------------------------------------ shared memory carve
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sp = reinterpret_cast<double*>(smem_raw);
    const bool rw = (D == 2) && m.rw_ok;                    // thread-per-particle FMA row walk (basis_rowwalk.cuh)
    double* bfrag = sp;         sp += PRE ? 0 : rw ? (size_t)((m.rw_slots + 1) & ~1)      // Theta' in row-walk order, or
                                         : m.n_packed + (size_t)m.NTNP * 32;    // in DMMA B-fragment order (+1 zero step: prefetch)
    double* tiles = sp;         sp += (PRE || rw) ? 0 : (size_t)NW * sine_tile_doubles(m);   // per-warp sine tiles (tile form only)
    double* xs = sp;            sp += PRE ? 0 : (size_t)NX * P;          // [k][i]
    double* mus = sp;           sp += PRE ? 0 : (size_t)NX * P;          // [k][i] auxiliary mean
    double* logw = sp;          sp += P;
    double* laux = sp;          sp += P;
    const int nblk = (P + 255) / 256;
    double* b1 = sp;            sp += nblk * 256;              // lw_aux -> prefix -> CDF (padded with +inf for the search)
    double* b2 = sp;            sp += P;                       // lw_anc -> prefix
    double* lauxg = sp;         sp += P;                       // l_aux[a_i], pushed by the CDF owner
    double* mug = sp;           sp += (gather && !PRE) ? (size_t)NX * P : 0;
    double* exch = sp;          sp += MAXC * 4;                // per-CTA (m1,s1,m2,s2), all-gathered
    double* gsum = sp;          sp += 2 * (MAXC + 1);          // exclusive CDF offsets G1[c], G2[c], c = 0..C
    double* fx = sp;            sp += 2 * MAXC + 2;            // rescale factors exp(m_c - M); then 1/S1, 1/S2
    double* unit = sp;          sp += (size_t)PPT * NW * 4;    // per (round, warp): (max1, sum1, max2, sum2)
    double* ufac = sp;          sp += (size_t)PPT * NW * 4;    // per (round, warp): (f1, G1, f2, G2)
    double* chol = sp;          sp += NX * NX;                 // chol(Sigma) lower
    double* sw = sp;            sp += NX * NX;                 // chol(Sigma)^-1
    double* slogc = sp;         sp += 2;
    StepConst* sc = reinterpret_cast<StepConst*>(sp);  sp += 2 * ((sizeof(StepConst) + 7) / 8);
    int* ip = reinterpret_cast<int*>(sp);
    int* rowpos = ip;           ip += ((8 * m.NTNP + NX - 1) / NX + 1) * MAX_LEAD;
    int* cnt = ip;              ip += 4;
    int* ntc_s = ip;            ip += 18;                      // per column block: position steps with non-zero tiles (basis_eval.cuh: kcb)
    int* rwlen = ip;            ip += RW_MAXBLK;               // row-walk block lengths

    // ------------------------------------------------------------------ prologue
    for (int r = tid; r < ((8 * m.NTNP + NX - 1) / NX + 1) * MAX_LEAD; r += NT) rowpos[r] = m.row_pos[r];
    for (int r = Pc + tid; r < nblk * 256; r += NT) b1[r] = INFINITY;
    if (tid < RW_MAXBLK) rwlen[tid] = m.rw_blen[tid];
    if (tid < 18) {
        int kc = 0;
        for (int ks = 0; ks < m.KS; ++ks) kc += (m.ntcount[ks] > tid * NTB) ? 1 : 0;
        ntc_s[tid] = kc;
    }
    EvalCtx<D> ecx;
    ecx.init(m);
    const int tile_doubles = sine_tile_doubles(m);
    MapRegs<NX, D> mapr;
    mapr.init(m);
    if constexpr (!PRE) {   // Theta' = norm * Theta scattered into B-fragment order
        const double* Th = a.Theta + (size_t)chain * NX * m.M;
        if (rw) {
            for (int s = tid; s < m.rw_slots; s += NT) {
                const int e = m.rw_perm[s];
                bfrag[s] = (e >= 0) ? m.norm * Th[(size_t)(e & 3) * m.M + (e >> 2)] : 0.0;
            }
        } else {
            for (int s = tid; s < m.n_packed + m.NTNP * 32; s += NT) {
                const int e = (s < m.n_packed) ? m.perm[s] : -1;
                bfrag[s] = (e >= 0) ? m.norm * Th[(size_t)(e & 3) * m.M + (e >> 2)] : 0.0;
            }
        }
    }
    if (PRE && tid == 0) {
        cnt[0] = 0;
        load_step_const(a, chain, a.t_begin, &sc[a.t_begin & 1]);
    }
    if (!PRE && tid == 0) {
        // chol(Sigma) and its inverse (src/PGAS.py:72-75 multivariate_normal; :109-116 logpdf)
        const double* S = a.Sigma + (size_t)chain * NX * NX;
        double Lc[NX][NX], Li[NX][NX];
        double logdet = 0.0;
        for (int i = 0; i < NX; ++i)
            for (int j = 0; j < NX; ++j) { Lc[i][j] = 0.0; Li[i][j] = 0.0; }
        for (int j = 0; j < NX; ++j) {
            double d = S[j * NX + j];
            for (int k = 0; k < j; ++k) d -= Lc[j][k] * Lc[j][k];
            d = sqrt(d);
            Lc[j][j] = d;
            logdet += log(d);
            for (int i = j + 1; i < NX; ++i) {
                double v = S[i * NX + j];
                for (int k = 0; k < j; ++k) v -= Lc[i][k] * Lc[j][k];
                Lc[i][j] = v / d;
            }
        }
        for (int j = 0; j < NX; ++j) {          // forward substitution for the inverse
            Li[j][j] = 1.0 / Lc[j][j];
            for (int i = j + 1; i < NX; ++i) {
                double v = 0.0;
                for (int k = j; k < i; ++k) v -= Lc[i][k] * Li[k][j];
                Li[i][j] = v / Lc[i][i];
            }
        }
        for (int i = 0; i < NX; ++i)
            for (int j = 0; j < NX; ++j) { chol[i * NX + j] = Lc[i][j]; sw[i * NX + j] = Li[i][j]; }
        slogc[0] = -0.5 * NX * 1.8378770664093453 - logdet;      // log(2 pi)
        cnt[0] = 0;
        load_step_const(a, chain, a.t_begin, &sc[a.t_begin & 1]);
    }
    // initial particles
    for (int q = 0; q < PPT; ++q) {
        const int il = q * NT + tid;
        if constexpr (PRE) {
            if (il < Pc) logw[il] = a.init_logw ? a.init_logw[(size_t)chain * N + base + il] : 0.0;
            continue;
        }
        if (il < Pc) {
            const int i = base + il;
            double x[NX];
            if (a.init_state) {
#pragma unroll
                for (int k = 0; k < NX; ++k) x[k] = a.init_state[((size_t)chain * N + i) * NX + k];
                logw[il] = a.init_logw ? a.init_logw[(size_t)chain * N + i] : 0.0;
            } else {
                // x_0 ~ N(m0, P0) (src/PGAS.py:167-172), particle N-1 = reference (:194); logw_0 = 0
                double z[NX];
                draw_normals<NX>(a, chain, 0, i, z);
#pragma unroll
                for (int r = 0; r < NX; ++r) {
                    double v = m.m0[r];
#pragma unroll
                    for (int c = 0; c <= r; ++c) v = fma(m.P0c[r][c], z[c], v);
                    x[r] = v;
                }
                if (i == N - 1) {
#pragma unroll
                    for (int k = 0; k < NX; ++k) x[k] = a.ref[(size_t)chain * a.ref_stride + k];
                }
                logw[il] = 0.0;
                double* out = a.state_trace + (((size_t)chain * a.trace_rows + 0) * N + i) * NX;
#pragma unroll
                for (int k = 0; k < NX; ++k) out[k] = x[k];
            }
#pragma unroll
            for (int k = 0; k < NX; ++k) xs[(size_t)k * P + il] = x[k];
        }
    }
    if (C > 1) cluster.sync(); else __syncthreads();

    // ------------------------------------------------------------------ time loop
    for (int t = a.t_begin; t < a.t_end; ++t) {
        const StepConst& k_t = sc[t & 1];
        StepConst nxt;                                       // prefetch of step t+1 (thread 0)
        if (tid == 0 && t + 1 < a.t_end) load_step_const(a, chain, t + 1, &nxt);

#define PGAS_TICK(K) do { if (a.dbg && blockIdx.x < 2 && tid == 96) a.dbg[((size_t)(t - a.t_begin) * 2 + blockIdx.x) * 8 + (K)] = clock64(); } while (0)
        PGAS_TICK(0);
        // ---- A: auxiliary mean (DMMA), first-stage log-weights (src/PGAS.py:89-101, :109-117), and the
        //      softmax numerators (:102,:118) with a WARP-local shift: exp(lw - max_warp) and its in-warp
        //      inclusive scan need no block-wide reduction; the (max, sum) pair of every warp is combined
        //      once per step by warp 0 (online-softmax rescaling), first across the CTA, then across the cluster.
        // softmax numerators of four particles of this thread: warp-local shift, exp, in-warp inclusive scan —
        // eight independent sequences interleaved; lane 31 publishes the (max, sum) pairs of the warp
        auto softmax4 = [&](int qc, const double (&lwa)[4], const double (&lwr)[4]) {
            double m1w[4], m2w[4], e1[4], e2[4], s1[4], s2[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { m1w[u] = warp_shift_max(lwa[u]); m2w[u] = warp_shift_max(lwr[u]); }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool v = (qc + u) * NT + tid < Pc && qc + u < PPT;
                // a warp whose log-weights are all -inf is an empty unit (numerators 0, not exp(-inf + inf) = NaN)
                e1[u] = v ? exp_neg_bf(lwa[u] - (m1w[u] == -INFINITY ? 0.0 : m1w[u])) : 0.0;
                e2[u] = v ? exp_neg_bf(lwr[u] - (m2w[u] == -INFINITY ? 0.0 : m2w[u])) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { s1[u] = e1[u]; s2[u] = e2[u]; }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {                          // eight interleaved Kogge-Stone scans
                double n1[4], n2[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { n1[u] = __shfl_up_sync(0xffffffffu, s1[u], o); n2[u] = __shfl_up_sync(0xffffffffu, s2[u], o); }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (lane >= o) { s1[u] += n1[u]; s2[u] += n2[u]; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int q = qc + u, il = q * NT + tid;
                if (q < PPT) {
                    if (il < Pc) { b1[il] = s1[u]; b2[il] = s2[u]; }
                    if (lane == 31) {
                        double* up = unit + (size_t)(q * NW + warp) * 4;
                        const bool any = q * NT + warp * 32 < Pc;
                        up[0] = any ? m1w[u] : -INFINITY; up[1] = any ? s1[u] : 0.0;
                        up[2] = any ? m2w[u] : -INFINITY; up[3] = any ? s2[u] : 0.0;
                    }
                }
            }
        };
        if constexpr (PRE) {
            // log-densities were left by the state kernel (sweep_split.cu): l_aux, h; only the weight recursion runs here
            const size_t prow = ((size_t)chain * a.pre_rows + (size_t)(t - a.pre_off)) * N + base;
            if (t + 1 < a.t_end) {
                // the rows of step t+1 were written long ago (often evicted to HBM): pull them towards L1 now, a step ahead
                const size_t nrow = prow + N;
                for (int q = 0; q < PPT; ++q) {
                    const int il = q * NT + tid;
                    if (il < Pc && (il & 15) == 0) {           // one request per 128-byte line
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(a.pre_la + nrow + il));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(a.pre_lr + nrow + il));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(a.pre_ll + nrow + il));
                    }
                }
            }
            for (int qc = 0; qc < PPT; qc += 4) {
                double lwa[4], lwr[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = qc + u, il = q * NT + tid;
                    lwa[u] = lwr[u] = -INFINITY;
                    if (q < PPT && il < Pc) {
                        const double la = a.pre_la[prow + il];
                        lwa[u] = la + logw[il];
                        lwr[u] = lwa[u] + a.pre_lr[prow + il];
                        laux[il] = la;
                    }
                }
                softmax4(qc, lwa, lwr);
            }
        }
        if constexpr (D == 2 && !PRE) {
            if (rw) {
                // Row-walk form.  A thread's particles are handled four at a time so that the dependent chains of
                // the weights (log-densities, softmax shift, exp, in-warp scan) of the four are interleaved: first
                // the contraction two particles at a time (each Theta' pair read once per warp feeds 2 n_x DFMAs),
                // then eight independent shift-max / exp / scan sequences.
                const int nblk_rw = m.rw_nblk;
                auto weights = [&](int il, const double (&mu)[NX], double& lwa, double& lwr) {
#pragma unroll
                    for (int k = 0; k < NX; ++k) mus[(size_t)k * P + il] = mu[k];
                    const double la = loglik_any<NX, NY>(m, t, k_t.y, mu);
                    lwa = la + logw[il];
                    lwr = lwa + gauss_logpdf_state<NX>(sw, slogc[0], k_t.ref, mu);
                    laux[il] = la;
                };
                for (int qc = 0; qc < PPT; qc += 4) {
                    double lwa[4], lwr[4];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int q0 = qc + 2 * h;
                        const int ila = q0 * NT + tid, ilb = ila + NT;
                        lwa[2 * h] = lwr[2 * h] = lwa[2 * h + 1] = lwr[2 * h + 1] = -INFINITY;
                        if (q0 < PPT && q0 * NT + warp * 32 < Pc) {            // warp-uniform
                            if (q0 + 1 < PPT) {
                                double xa[NX], xb[NX], ta[D], tb[D];
#pragma unroll
                                for (int k = 0; k < NX; ++k) {
                                    xa[k] = (ila < Pc) ? xs[(size_t)k * P + ila] : 0.0;
                                    xb[k] = (ilb < Pc) ? xs[(size_t)k * P + ilb] : 0.0;
                                }
                                mapr.apply(xa, k_t.cz, k_t.u, ta);
                                mapr.apply(xb, k_t.cz, k_t.u, tb);
                                const double tzz[2][2] = {{ta[0], ta[1]}, {tb[0], tb[1]}};
                                double mu2[2][NX];
                                rowwalk_mu<NX, 2, 2>(bfrag, rwlen, &nblk_rw, 1, ecx.f_start, ecx.f_step, tzz, mu2);
                                if (ila < Pc) weights(ila, mu2[0], lwa[2 * h], lwr[2 * h]);
                                if (ilb < Pc) weights(ilb, mu2[1], lwa[2 * h + 1], lwr[2 * h + 1]);
                            } else {
                                double xa[NX], ta[D];
#pragma unroll
                                for (int k = 0; k < NX; ++k) xa[k] = (ila < Pc) ? xs[(size_t)k * P + ila] : 0.0;
                                mapr.apply(xa, k_t.cz, k_t.u, ta);
                                const double tzz[1][2] = {{ta[0], ta[1]}};
                                double mu1[1][NX];
                                rowwalk_mu<NX, 1, 2>(bfrag, rwlen, &nblk_rw, 1, ecx.f_start, ecx.f_step, tzz, mu1);
                                if (ila < Pc) weights(ila, mu1[0], lwa[2 * h], lwr[2 * h]);
                            }
                        }
                    }
                    softmax4(qc, lwa, lwr);
                }
            }
        }
        for (int q = 0; q < ((rw || PRE) ? 0 : PPT); ++q) {
            const int il = q * NT + tid, il0 = q * NT + warp * 32;
            double lwa = -INFINITY, lwr = -INFINITY;
            if (il0 < Pc) {                                   // warp-uniform: the whole warp takes part in the DMMA
                double tz[D];
#pragma unroll
                for (int d = 0; d < D; ++d) tz[d] = 0.0;
                PGAS_FTICK(0);
                if (!rw) {
                    if (il < Pc) {
                        double x[NX];
#pragma unroll
                        for (int k = 0; k < NX; ++k) x[k] = xs[(size_t)k * P + il];
                        mapr.apply(x, k_t.cz, k_t.u, tz);
                    }
                    eval_mu_warp<NX, D>(ecx, bfrag, rowpos, ntc_s, tiles + (size_t)warp * tile_doubles, tz, lane, mus, P, il0);
                }
                if (il < Pc) {
                    double mu[NX];
#pragma unroll
                    for (int k = 0; k < NX; ++k) mu[k] = mus[(size_t)k * P + il];
                    const double la = loglik_any<NX, NY>(m, t, k_t.y, mu);
                    lwa = la + logw[il];
                    lwr = lwa + gauss_logpdf_state<NX>(sw, slogc[0], k_t.ref, mu);
                    laux[il] = la;
                }
                PGAS_FTICK(20);
                const double m1w = warp_shift_max(lwa), m2w = warp_shift_max(lwr);
                PGAS_FTICK(21);
                const double e1 = (il < Pc) ? exp_neg_bf(lwa - (m1w == -INFINITY ? 0.0 : m1w)) : 0.0;
                const double e2 = (il < Pc) ? exp_neg_bf(lwr - (m2w == -INFINITY ? 0.0 : m2w)) : 0.0;
                PGAS_FTICK(22);
                const double s1 = warp_scan_incl(e1, lane), s2 = warp_scan_incl(e2, lane);
                PGAS_FTICK(23);
                if (il < Pc) { b1[il] = s1; b2[il] = s2; }
                if (lane == 31) {
                    double* up = unit + (size_t)(q * NW + warp) * 4;
                    up[0] = m1w; up[1] = s1; up[2] = m2w; up[3] = s2;
                }
            } else if (lane == 31) {
                double* up = unit + (size_t)(q * NW + warp) * 4;
                up[0] = -INFINITY; up[1] = 0.0; up[2] = -INFINITY; up[3] = 0.0;
            }
        }
        PGAS_TICK(1);
        __syncthreads();
        PGAS_TICK(2);

        // ---- X1: warp 0 folds the warp pairs into the CTA pair; all-gather across the cluster; fold again
        const int c_last = (N - 1) / P;                       // CTA owning particle N-1 (later CTAs are empty)
        const int U = PPT * NW;
        double m1c = 0.0, m2c = 0.0, s1c = 0.0, s2c = 0.0;
        if (warp == 0) {
            m1c = -INFINITY; m2c = -INFINITY;
            for (int u = lane; u < U; u += 32) { m1c = fmax(m1c, unit[u * 4]); m2c = fmax(m2c, unit[u * 4 + 2]); }
            m1c = warp_max(m1c);
            m2c = warp_max(m2c);
            for (int u0 = 0; u0 < U; u0 += 32) {
                const int u = u0 + lane;
                const bool vu = u < U;
                const double mu1 = vu ? unit[u * 4] : -INFINITY, mu2 = vu ? unit[u * 4 + 2] : -INFINITY;
                const double f1 = (mu1 == -INFINITY) ? 0.0 : exp_neg_bf(mu1 - m1c);
                const double f2 = (mu2 == -INFINITY) ? 0.0 : exp_neg_bf(mu2 - m2c);
                const double v1 = vu ? __dmul_rn(unit[u * 4 + 1], f1) : 0.0, v2 = vu ? __dmul_rn(unit[u * 4 + 3], f2) : 0.0;
                // exclusive prefix in unit order with the same association the particles use (G + s f)
                double i1 = warp_scan_incl(v1, lane), i2 = warp_scan_incl(v2, lane);
                double x1 = __shfl_up_sync(0xffffffffu, i1, 1), x2 = __shfl_up_sync(0xffffffffu, i2, 1);
                x1 = (lane == 0) ? s1c : __dadd_rn(s1c, x1);
                x2 = (lane == 0) ? s2c : __dadd_rn(s2c, x2);
                if (vu) {
                    double* fp = ufac + (size_t)u * 4;
                    fp[0] = f1; fp[1] = x1; fp[2] = f2; fp[3] = x2;
                }
                // carry = exclusive prefix of the last lane + its value (what its last particle computes)
                s1c = __shfl_sync(0xffffffffu, __dadd_rn(x1, v1), 31);
                s2c = __shfl_sync(0xffffffffu, __dadd_rn(x2, v2), 31);
            }
            if (C > 1) {
                if (lane < C) {
                    double* dst = cluster.map_shared_rank(exch, lane) + rank * 4;
                    dst[0] = m1c; dst[1] = s1c; dst[2] = m2c; dst[3] = s2c;
                }
            } else if (lane == 0) {
                exch[0] = m1c; exch[1] = s1c; exch[2] = m2c; exch[3] = s2c;
            }
            __syncwarp();
        }
        if (C > 1) {
            cluster_arrive();
            cluster_wait();
        }
        if (warp == 0) {
            const bool vc = lane < C;
            const double mc1 = vc ? exch[lane * 4] : -INFINITY, mc2 = vc ? exch[lane * 4 + 2] : -INFINITY;
            const double M1 = warp_max(mc1), M2 = warp_max(mc2);
            const double f1 = (mc1 == -INFINITY) ? 0.0 : exp_neg_bf(mc1 - M1);
            const double f2 = (mc2 == -INFINITY) ? 0.0 : exp_neg_bf(mc2 - M2);
            const double v1 = vc ? exch[lane * 4 + 1] * f1 : 0.0, v2 = vc ? exch[lane * 4 + 3] * f2 : 0.0;
            // every CTA of the cluster runs the same shuffle tree on the same data -> identical bits
            const double i1 = warp_scan_incl(v1, lane), i2 = warp_scan_incl(v2, lane);
            const double r1 = rcp_bf(__shfl_sync(0xffffffffu, i1, 31)), r2 = rcp_bf(__shfl_sync(0xffffffffu, i2, 31));
            // CTA holding the reference ancestor: number of leading CTAs whose whole CDF segment lies below u_anc
            const unsigned below = __ballot_sync(0xffffffffu, lane <= c_last && __dmul_rn(i2, r2) < k_t.uanc);
            if (vc) {
                fx[lane * 2] = f1;
                fx[lane * 2 + 1] = f2;
                gsum[(lane + 1) * 2] = i1;
                gsum[(lane + 1) * 2 + 1] = i2;
            }
            const int uni_flag = !(__shfl_sync(0xffffffffu, i1, 31) > 0.0) ? 1 : 0;    // all -inf, or a NaN: uniform weights (src/Filtering.py:24-25)
            if (lane == 0) {
                gsum[0] = 0.0;
                gsum[1] = 0.0;
                fx[2 * MAXC] = r1;
                fx[2 * MAXC + 1] = r2;
                cnt[1] = __popc(below);
                cnt[2] = uni_flag;
            }
        }
        PGAS_TICK(3);
        if (tid == 0 && t + 1 < a.t_end) sc[(t + 1) & 1] = nxt;
        __syncthreads();
        PGAS_TICK(4);
        const double myg1 = gsum[rank * 2], myg2 = gsum[rank * 2 + 1], g1hi = gsum[(rank + 1) * 2];
        const double myf1 = fx[rank * 2], myf2 = fx[rank * 2 + 1];
        const double S1 = fx[2 * MAXC], S2 = fx[2 * MAXC + 1];      // reciprocals of the normalisers
        const int cstar = cnt[1];
        const bool uni = cnt[2] != 0;
        int mycnt = 0;
        for (int q = 0; q < PPT; ++q) {
            const int il = q * NT + tid;
            if (il < Pc) {
                const double* fp = ufac + (size_t)(q * NW + warp) * 4;
                // CTA-level inclusive prefix of this particle: G_unit + s_i f_unit
                const double p1 = __dadd_rn(fp[1], __dmul_rn(b1[il], fp[0]));
                const double p2 = __dadd_rn(fp[3], __dmul_rn(b2[il], fp[2]));
                // W = clip(cumsum(w / sum w), 0, 1)  (src/Filtering.py:23-32)
                b1[il] = uni ? div_by_count((double)(base + il + 1), dN, rN) : clip01(cdf_value(p1, myf1, myg1, S1));
                // cumsum(softmax(lw_anc)) < u_anc  (src/PGAS.py:118-124), not clipped
                if (rank == cstar) mycnt += (cdf_value(p2, myf2, myg2, S2) < k_t.uanc) ? 1 : 0;
            }
        }
        if (rank == cstar) {
            mycnt = __reduce_add_sync(0xffffffffu, mycnt);
            if (lane == 0 && mycnt) atomicAdd(&cnt[0], mycnt);
        }
        __syncthreads();

        PGAS_TICK(5);
        // ---- B2: systematic resampling (src/Filtering.py:28-35) by the owner of the CDF segment
        {
            const double blo = uni ? div_by_count((double)min(base, N), dN, rN) : clip01(__dmul_rn(myg1, S1));
            const double bhi = uni ? div_by_count((double)min(base + P, N), dN, rN) : clip01(__dmul_rn(g1hi, S1));
            const int jlo = (rank == 0) ? 0 : first_point_above(blo, k_t.ures, N, dN, rN);
            int jhi = (rank == c_last) ? N : first_point_above(bhi, k_t.ures, N, dN, rN);
            if (rank > c_last) jhi = jlo;                     // empty CTA
            int* anc_row = a.anc_trace + ((size_t)chain * a.anc_rows + (t - 1 - a.row_off + a.anc_shift)) * N;
            for (int j = jlo + tid; j < jhi; j += NT) {
                if (j == N - 1) continue;                     // overwritten by the reference ancestor (:127)
                const double uj = strat_point(k_t.ures, j, dN, rN);
                const int k = min(count_below_padded(b1, nblk, uj), Pc - 1);
                anc_row[j] = base + k;
                const int cj = j / P, jl = j - cj * P;
                double* dl = (C > 1 && cj != rank) ? cluster.map_shared_rank(lauxg, cj) : lauxg;
                dl[jl] = laux[k];
                if (gather && !PRE) {
                    double* dm = (C > 1 && cj != rank) ? cluster.map_shared_rank(mug, cj) : mug;
#pragma unroll
                    for (int kk = 0; kk < NX; ++kk) dm[(size_t)kk * P + jl] = mus[(size_t)kk * P + k];
                }
            }
            // ---- B3: ancestor of the reference particle (src/PGAS.py:118-127)
            if (tid == 0 && (rank == cstar || (cstar == c_last + 1 && rank == c_last))) {
                int ref_idx, k;
                if (rank == cstar) { k = min(cnt[0], Pc - 1); ref_idx = base + cnt[0]; }
                else { k = Pc - 1; ref_idx = N; }             // cumsum never reached u_anc: searchsorted returns N, gather clamps
                cnt[0] = 0;
                anc_row[N - 1] = ref_idx;
                const int cj = c_last, jl = (N - 1) - cj * P;
                double* dl = (C > 1 && cj != rank) ? cluster.map_shared_rank(lauxg, cj) : lauxg;
                dl[jl] = laux[k];
                if (gather && !PRE) {
                    double* dm = (C > 1 && cj != rank) ? cluster.map_shared_rank(mug, cj) : mug;
#pragma unroll
                    for (int kk = 0; kk < NX; ++kk) dm[(size_t)kk * P + jl] = mus[(size_t)kk * P + k];
                }
            }
        }

        PGAS_TICK(6);
        // ---- X2 + C: barrier #2 overlapped with the noise draw; new state, new log-weights
        if (C > 1) cluster_arrive(); else __syncthreads();
        const bool last_step = (t + 1 == a.t_end);
        if constexpr (PRE) {
            if (C > 1) cluster_wait();
            PGAS_TICK(7);
            const size_t prow = ((size_t)chain * a.pre_rows + (size_t)(t - a.pre_off)) * N + base;
            for (int q = 0; q < PPT; ++q) {
                const int il = q * NT + tid;
                if (il < Pc) {
                    const double lw = a.pre_ll[prow + il] - lauxg[il];             // src/PGAS.py:137-147
                    logw[il] = lw;
                    if (last_step && a.logw_last) a.logw_last[(size_t)chain * N + base + il] = lw;
                }
            }
        } else {
            constexpr int ZQ = 4;                                 // particles per thread whose noise is drawn under the barrier
            double zreg[ZQ][NX];
            // the draws of a thread's particles are independent chains (Philox -> log/sqrt/sincospi): issued
            // back to back without per-lane branches so the scheduler interleaves them (padding lanes draw too)
    #pragma unroll
            for (int q = 0; q < ZQ; ++q) {
                if (q < PPT) draw_normals<NX>(a, chain, t, min(base + q * NT + tid, N - 1), zreg[q]);
            }
            if (C > 1) cluster_wait();
            double* st_row = a.state_trace + (((size_t)chain * a.trace_rows + (t - a.row_off)) * N) * NX;
            const double* msrc = gather ? mug : mus;              // quirk (i): own particle unless gather mode
            auto finish = [&](int il, const double z[NX]) {
                const int i = base + il;
                double x[NX];
    #pragma unroll
                for (int r = 0; r < NX; ++r) {
                    double v = msrc[(size_t)r * P + il];
    #pragma unroll
                    for (int c = 0; c <= r; ++c) v = fma(chol[r * NX + c], z[c], v);
                    x[r] = v;
                }
                if (i == N - 1) {
    #pragma unroll
                    for (int k = 0; k < NX; ++k) x[k] = k_t.ref[k];               // src/PGAS.py:134
                }
                const double lw = loglik_any<NX, NY>(m, t, k_t.y, x) - lauxg[il];   // :137-147
                logw[il] = lw;
    #pragma unroll
                for (int k = 0; k < NX; ++k) { xs[(size_t)k * P + il] = x[k]; st_row[(size_t)i * NX + k] = x[k]; }
                if (last_step && a.logw_last) a.logw_last[(size_t)chain * N + i] = lw;
            };
    #pragma unroll
            for (int q = 0; q < ZQ; ++q) {
                const int il = q * NT + tid;
                if (q < PPT && il < Pc) finish(il, zreg[q]);
            }
            for (int q = ZQ; q < PPT; ++q) {
                const int il = q * NT + tid;
                if (il < Pc) {
                    double z[NX];
                    draw_normals<NX>(a, chain, t, base + il, z);
                    finish(il, z);
                }
            }
        }
        PGAS_TICK(7);
        // next step's A1 only touches this thread's own xs/logw entries and arrays whose previous
        // readers are fenced by the barriers above; b1/b2/laux/mus are rewritten after every
        // thread of the CTA has left B2, which the X2 barrier guarantees.
    }
    if (C > 1) cluster.sync();                                // no CTA exits while peers may still push into its smem
}

// ------------------------------------------------------------------------------------ launch
static size_t sweep_smem_bytes(const DevModel& m, int NX, int P, bool gather, int NT, bool pre = false) {
    const int NW = NT / 32;
    const bool rw = (m.D == 2) && m.rw_ok;
    size_t d = (pre ? 0 : (rw ? (size_t)((m.rw_slots + 1) & ~1) : (size_t)m.n_packed + (size_t)m.NTNP * 32 + (size_t)NW * sine_tile_doubles(m)) + (size_t)NX * P * 2) + (size_t)P * 4 + (size_t)((P + 255) / 256) * 256 +
               ((gather && !pre) ? (size_t)NX * P : 0) + MAXC * 4 + 2 * (MAXC + 1) + 2 * MAXC + 2 + (size_t)((P + NT - 1) / NT) * NW * 8 + 2 * NX * NX + 2 +
               2 * ((sizeof(StepConst) + 7) / 8);
    return d * 8 + (size_t)(((8 * m.NTNP + NX - 1) / NX + 1) * MAX_LEAD + 4 + 18 + RW_MAXBLK) * 4 + 32;
}

// threads per CTA: 512 (16 warps hide the dependent-FP64 latency best) when the per-warp sine tiles
// still fit shared memory, else 256
static int sweep_threads(const DevModel& m, int P) {
    if (const char* e = getenv("PGAS_SWEEP_THREADS")) { const int v = atoi(e); if (v == 256 || v == 512) return v; }   // developer override
    const bool gather = (m.flags & PGAS_FLAG_ANCESTOR_GATHER) != 0;
    return (sweep_smem_bytes(m, m.n_x, P, gather, 512) <= 227 * 1024 && P > 128) ? 512 : 256;
}

static int* g_query_clusters = nullptr;      // debug: when set, launch_variant reports occupancy instead of launching

template <int NX, int NY, int D, int NT, bool PRE = false>
static int launch_variant(const SweepArgs& a, size_t smem, cudaStream_t stream) {
    auto kern = csmc_sweep_kernel<NX, NY, D, NT, PRE>;
    PGAS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (a.C > 8) PGAS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(a.C * a.n_chains), 1, 1);
    cfg.blockDim = dim3(NT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (g_query_clusters) {
        PGAS_CUDA(cudaOccupancyMaxActiveClusters(g_query_clusters, kern, &cfg));
        return 0;
    }
    PGAS_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    __atomic_add_fetch(&g_pgas_launches, 1, __ATOMIC_RELAXED);
    return 0;
}

// resampling recursion on precomputed log-densities (split form); the model dimensions do not matter here
int pgas_launch_sweep_pre(const SweepArgs& a, cudaStream_t stream) {
    const DevModel& m = a.m;
    const size_t smem = sweep_smem_bytes(m, m.n_x, a.P, false, PGAS_PRE_NT, true);
    if (smem > 227 * 1024) PGAS_FAIL(-21, "resampling kernel needs %zu bytes of shared memory per CTA", smem);
    return launch_variant<2, 1, 2, PGAS_PRE_NT, true>(a, smem, stream);
}

int pgas_launch_sweep_fused(const SweepArgs& a, cudaStream_t stream) {
    const DevModel& m = a.m;
    const bool gather = (m.flags & PGAS_FLAG_ANCESTOR_GATHER) != 0;
    const int nt = sweep_threads(m, a.P);
    const size_t smem = sweep_smem_bytes(m, m.n_x, a.P, gather, nt);
    if (smem > 227 * 1024)
        PGAS_FAIL(-21, "sweep needs %zu bytes of shared memory per CTA (N=%d over a cluster of %d); use a larger cluster", smem, a.N, a.C);
#define PGAS_DISPATCH(NXv, NYv, Dv) \
    if (m.n_x == NXv && m.n_y == NYv && m.D == Dv)            \
        return nt == 512 ? launch_variant<NXv, NYv, Dv, 512>(a, smem, stream) : launch_variant<NXv, NYv, Dv, 256>(a, smem, stream);
    PGAS_DISPATCH(1, 1, 1)
    PGAS_DISPATCH(2, 1, 1)
    PGAS_DISPATCH(2, 1, 2)
    PGAS_DISPATCH(2, 2, 2)
    PGAS_DISPATCH(2, 1, 3)
    PGAS_DISPATCH(2, 2, 1)
#undef PGAS_DISPATCH
    PGAS_FAIL(-22, "no compiled sweep variant for n_x=%d n_y=%d D=%d", m.n_x, m.n_y, m.D);
}

size_t pgas_sweep_smem_for(const DevModel& m, int P) {
    return sweep_smem_bytes(m, m.n_x, P, (m.flags & PGAS_FLAG_ANCESTOR_GATHER) != 0, sweep_threads(m, P));
}

// choose the cluster size: the largest power of two such that (i) every chain's cluster is resident
// at once — the number of co-resident clusters measured on B200 with cudaOccupancyMaxActiveClusters is
// 148 / 74 / 33 / 15 / 7 for sizes 1 / 2 / 4 / 8 / 16 (GPC granularity; profiles/r01_microbench.md) —
// (ii) a CTA keeps at least 128 particles, and (iii) the per-CTA slice fits shared memory.
int pgas_choose_cluster(const DevModel& m, int N, int n_chains, int requested) {
    if (requested > 0) return requested;
    static const int sizes[5] = {16, 8, 4, 2, 1}, resident[5] = {7, 15, 33, 74, 148};
    int fallback = MAXC;
    for (int i = 0; i < 5; ++i) {
        const int C = sizes[i], P = (N + C - 1) / C;
        if (pgas_sweep_smem_for(m, P) > 227 * 1024) break;           // smaller clusters only need more
        fallback = C;
        if (n_chains <= resident[i] && (P >= 128 || C == 1)) return C;
    }
    return fallback;                                                 // more chains than SMs: smallest cluster that fits
}

// how many clusters of `cluster_size` CTAs of the sweep kernel the device can hold at once
extern "C" int pgas_debug_max_active_clusters(const pgas_model* model, int32_t N, int32_t cluster_size) {
    SweepArgs a;
    memset(&a, 0, sizeof(a));
    a.m = model->dev;
    a.N = N; a.n_chains = 1; a.C = cluster_size; a.P = (N + cluster_size - 1) / cluster_size;
    int n = -1;
    g_query_clusters = &n;
    const int rc = pgas_launch_sweep_fused(a, 0);
    g_query_clusters = nullptr;
    return rc ? -rc : n;
}

// developer aid: run a sweep with phase clocks of CTA 0 recorded into dbg (2 threads x 8 ticks per step)
extern "C" int pgas_debug_sweep_ticks(const pgas_model* model, int32_t N, int32_t n_chains, const double* ref, const double* Theta,
                                      const double* Sigma, double* state_trace, int32_t* anc_trace, double* logw_last,
                                      int32_t cluster_size, long long* dbg, void* stream) {
    SweepArgs a;
    memset(&a, 0, sizeof(a));
    a.m = model->dev;
    a.N = N; a.n_chains = n_chains; a.C = pgas_choose_cluster(a.m, N, n_chains, cluster_size); a.P = (N + a.C - 1) / a.C;
    a.t_begin = 1; a.t_end = a.m.T;
    a.ref_rows = a.m.T; a.trace_rows = a.m.T; a.anc_rows = a.m.T - 1; a.var_rows = a.m.T;
    a.ref = ref; a.ref_stride = (long long)a.m.T * a.m.n_x; a.Theta = Theta; a.Sigma = Sigma;
    a.state_trace = state_trace; a.anc_trace = anc_trace; a.logw_last = logw_last;
    a.rng_mode = 0; a.seed = 1234;
    a.dbg = dbg;
    return pgas_launch_sweep_fused(a, (cudaStream_t)stream);
}

#if PGAS_FINE_TICKS
extern "C" int pgas_debug_fine_ticks(long long* host64) {
    return (int)cudaMemcpyFromSymbol(host64, g_fine, sizeof(long long) * 64);
}
#endif


// ====================================================================================== split form
// In the reference's semantics the state recursion is independent of the resampling: _draw_states propagates
// particle i from particle i (src/PGAS.py:131-133), so x_t^i = Theta phi(x_{t-1}^i, u_t) + chol(Sigma) z_t^i never
// sees an ancestor index; only the WEIGHT recursion does (logw_t^i = log p(y_t|x_t^i) - l_aux[a_i], :137-147).
// The sweep therefore splits into
//   csmc_state_kernel      all particles, all steps of a chunk, no synchronisation at all: two particles per
//                          thread stay in registers; FP64-FMA row walk (basis_rowwalk.cuh), log-densities,
//                          Philox / Box-Muller, trace row; bound by the FP64 pipe;
//   csmc_sweep_kernel<PRE> the resampling recursion on the three log-densities the state kernel left per
//                          particle and step (24 B), latency-bound (barriers, scans, searches),
// run chunk-wise on two streams so that the FP64-bound kernel of chunk c+1 shares the SMs with the latency-
// bound kernel of chunk c.  With PGAS_FLAG_ANCESTOR_GATHER (textbook move) the state depends on the
// ancestors and the fused kernel is used.
struct StateArgs {
    SweepArgs a;
    double* x_carry;             // (n_chains, N, NX): state at step t0-1 on entry (unless first), at t1-1 on exit
    double *la, *lr, *ll;        // (n_chains, rows, N), row t - t0
    int t0, t1, rows, first, bpc;
    int chain0, nch;             // this launch covers chains chain0 .. chain0 + nch - 1
};

#ifndef PGAS_ST_PP
#define PGAS_ST_PP 2
#endif
constexpr int ST_NT_BIG = PGAS_ST_NT, ST_PP_BIG = PGAS_ST_PP;      // geometry of the state kernel when the launch fills the GPU
constexpr int ST_NT_SMALL = 64, ST_PP_SMALL = 1;                   // ... and when it does not
constexpr int ST_LANES = 8, ST_LANES_MAX_N = 1024;                 // lanes per particle / largest N of the lane-split form (three-dimensional bases)

// INJ: injected variates (tests) instead of the in-kernel Philox stream — a template parameter so that the noise of a thread's
// particles sits in ONE basic block (no run-time branch per draw) and ptxas interleaves their Philox / Box-Muller chains
// ST_NT threads per CTA, ST_PP particles per thread: <256, 2> when the chains of a launch fill the GPU (two particles share every
// Theta' pair: half the shared-memory traffic per DFMA, 128 registers, two CTAs per SM); <64, 1> when they do not (few chains:
// four times as many, smaller CTAs spread over all SMs, and a step's dependent chain per thread is half as long).
// D: dimension of the Hilbert basis (2, or 3: one more level of the row walk — the EMPS baseline of src/EMPS.py:101-123).
// MMA: the contraction on DMMA tiles (basis_mma.cuh) instead of the FMA row walk; needs ST_PP == 2, D == 2, NX == 2.
// G: lanes per particle (1, or 8 for a three-dimensional basis with few particles: rowwalk_mu_lanes, basis_rowwalk.cuh); all G lanes
// carry the particle's state and variates (same Philox counters, same bits), lane 0 of the group stores.
template <int NX, int NY, bool INJ, int ST_NT, int ST_PP, int D = 2, bool MMA = false, int G = 1>
__global__ void __launch_bounds__(ST_NT, ST_PP == 2 ? 512 / ST_NT : 8) csmc_state_kernel(const __grid_constant__ StateArgs s) {
    static_assert(!MMA || (ST_PP == 2 && D == 2 && NX == 2), "DMMA form: two particles per thread, two-dimensional basis, n_x = 2");
    static_assert(G == 1 || (ST_PP == 1 && D == 3 && !MMA), "lane-split form: one particle per lane group, three-dimensional basis");
    const SweepArgs& a = s.a;
    const DevModel& m = a.m;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* bd = reinterpret_cast<double*>(smem_raw);
    const int n_theta = MMA ? m.mma_slots : m.rw_slots;
    double* chol = bd + ((n_theta + 1) & ~1);
    double* sw = chol + NX * NX;
    double* slogc = sw + NX * NX;
    int* rwlen = reinterpret_cast<int*>(slogc + 2);
    __shared__ int rwslice[3 * RW_MAXSLICE];      // lane-split form: blocks, first slot, first block of every slice
    double* mma_w = reinterpret_cast<double*>(rwlen + RW_MAXBLK) + (size_t)(threadIdx.x >> 5) * MMA_WARP_DOUBLES;   // DMMA form: per-warp seeds + tile buffer
    const int tid = threadIdx.x, N = a.N;
    const int chain = s.chain0 + blockIdx.x / s.bpc, blk = blockIdx.x % s.bpc;
    {
        const double* Th = a.Theta + (size_t)chain * NX * m.M;
        const int* perm = MMA ? m.mma_perm : m.rw_perm;
        for (int e = tid; e < n_theta; e += ST_NT) {
            const int q = perm[e];
            bd[e] = (q >= 0) ? m.norm * Th[(size_t)(q & 3) * m.M + (q >> 2)] : 0.0;
        }
        if (tid < RW_MAXBLK) rwlen[tid] = MMA ? (int)m.mma_ks[tid] : m.rw_blen[tid];
        if (G > 1 && tid < RW_MAXSLICE) {
            rwslice[tid] = m.rw_slice_nblk[tid];
            rwslice[RW_MAXSLICE + tid] = m.rw_slice_off[tid];
            rwslice[2 * RW_MAXSLICE + tid] = m.rw_slice_blk[tid];
        }
        if (tid == 0) {
            const double* S = a.Sigma + (size_t)chain * NX * NX;
            double Lc[NX][NX], Li[NX][NX], logdet = 0.0;
            for (int i = 0; i < NX; ++i)
                for (int j = 0; j < NX; ++j) { Lc[i][j] = 0.0; Li[i][j] = 0.0; }
            for (int j = 0; j < NX; ++j) {
                double d = S[j * NX + j];
                for (int k = 0; k < j; ++k) d -= Lc[j][k] * Lc[j][k];
                d = sqrt(d);
                Lc[j][j] = d;
                logdet += log(d);
                for (int i = j + 1; i < NX; ++i) {
                    double v = S[i * NX + j];
                    for (int k = 0; k < j; ++k) v -= Lc[i][k] * Lc[j][k];
                    Lc[i][j] = v / d;
                }
            }
            for (int j = 0; j < NX; ++j) {
                Li[j][j] = 1.0 / Lc[j][j];
                for (int i = j + 1; i < NX; ++i) {
                    double v = 0.0;
                    for (int k = j; k < i; ++k) v -= Lc[i][k] * Li[k][j];
                    Li[i][j] = v / Lc[i][i];
                }
            }
            for (int i = 0; i < NX; ++i)
                for (int j = 0; j < NX; ++j) { chol[i * NX + j] = Lc[i][j]; sw[i * NX + j] = Li[i][j]; }
            slogc[0] = -0.5 * NX * 1.8378770664093453 - logdet;
        }
    }
    __syncthreads();
    MapRegs<NX, D> mapr;
    mapr.init(m);
    const int f_start = m.f_start, f_step = m.f_step;
    int ip[ST_PP];
    bool val[ST_PP], own[ST_PP];                  // val: this thread stores the particle's rows; own: the particle exists
    const int sub = (G > 1) ? (tid % G) : 0;
#pragma unroll
    for (int p = 0; p < ST_PP; ++p) {
        // DMMA form: a warp owns 64 consecutive particles, lane l the particles l and 32 + l of them
        ip[p] = (G > 1) ? blk * (ST_NT / G) + tid / G
                        : MMA ? blk * ST_PP * ST_NT + (tid >> 5) * 64 + p * 32 + (tid & 31) : blk * ST_PP * ST_NT + p * ST_NT + tid;
        own[p] = ip[p] < N;
        val[p] = own[p] && sub == 0;
    }
    double x[ST_PP][NX];
    const double* refc = a.ref + (size_t)chain * a.ref_stride;
#pragma unroll
    for (int p = 0; p < ST_PP; ++p) {
        if (s.first) {
            // x_0 ~ N(m0, P0) (src/PGAS.py:167-172), particle N-1 = reference (:194)
            double z[NX];
            draw_normals<NX>(a, chain, 0, min(ip[p], N - 1), z);
#pragma unroll
            for (int r = 0; r < NX; ++r) {
                double v = m.m0[r];
#pragma unroll
                for (int c = 0; c <= r; ++c) v = fma(m.P0c[r][c], z[c], v);
                x[p][r] = v;
            }
            if (ip[p] == N - 1) {
#pragma unroll
                for (int k = 0; k < NX; ++k) x[p][k] = refc[k];
            }
            if (val[p]) {
                double* out = a.state_trace + (((size_t)chain * a.trace_rows + 0) * N + ip[p]) * NX;
#pragma unroll
                for (int k = 0; k < NX; ++k) out[k] = x[p][k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < NX; ++k) x[p][k] = own[p] ? s.x_carry[((size_t)chain * N + ip[p]) * NX + k] : 0.0;
        }
    }
    for (int t = s.t0; t < s.t1; ++t) {
        // per-step constants (uniform addresses: broadcast loads)
        const int tin = (m.flags & PGAS_FLAG_INPUT_PREV) ? t - 1 : t;          // quirk (ii), src/PGAS.py:52-54
        double y[NY], u[PGAS_MAX_NU], cz[D], ref[NX];
#pragma unroll
        for (int r = 0; r < NY; ++r) y[r] = m.obs[(size_t)t * NY + r];
#pragma unroll
        for (int k = 0; k < PGAS_MAX_NU; ++k) u[k] = (k < m.n_u) ? m.inputs[(size_t)tin * m.n_u + k] : 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            double acc = m.bz[d];
#pragma unroll
            for (int k = 0; k < PGAS_MAX_NU; ++k) acc = (k < m.n_u) ? fma(m.Az[d][NX + k], u[k], acc) : acc;
            cz[d] = acc;
        }
#pragma unroll
        for (int k = 0; k < NX; ++k) ref[k] = refc[(size_t)t * NX + k];
        // the noise of step t does not depend on the state: drawn first, both particles' Philox / Box-Muller chains in one
        // straight-line block next to the sine seeds of the row walk (independent dependent-FP64 chains for the scheduler)
        double z[ST_PP][NX];
        if constexpr (INJ) {
#pragma unroll
            for (int p = 0; p < ST_PP; ++p) {
                const double* zp = a.Z + (((size_t)chain * a.var_rows + (t - a.row_off)) * N + min(ip[p], N - 1)) * NX;
#pragma unroll
                for (int k = 0; k < NX; ++k) z[p][k] = zp[k];
            }
        } else {
#pragma unroll
            for (int p = 0; p < ST_PP; ++p) {
#pragma unroll
                for (int k = 0; k < NX; k += 2) {
                    double za, zb;
                    philox_normal2(a.seed, PURPOSE_STATE, a.chain_base + chain, a.iteration, (unsigned)t | ((unsigned)(k >> 1) << 28),
                                   (unsigned)min(ip[p], N - 1), za, zb);
                    z[p][k] = za;
                    if (k + 1 < NX) z[p][k + 1] = zb;
                }
            }
        }
        double tzv[ST_PP][D];
#pragma unroll
        for (int p = 0; p < ST_PP; ++p) mapr.apply(x[p], cz, u, tzv[p]);
        double mu[ST_PP][NX];
        if constexpr (MMA) mma_mu<NX>(bd, m.mma_ks, m.mma_nblk, mma_w, tid & 31, f_start, f_step, tzv, mu);
        else if constexpr (G > 1) rowwalk_mu_lanes<NX, G>(bd, rwlen, rwslice, rwslice + RW_MAXSLICE, rwslice + 2 * RW_MAXSLICE, m.rw_nslice, sub, f_start, f_step, tzv, mu);
        else rowwalk_mu<NX, ST_PP, D>(bd, m.rw_blen, m.rw_slice_nblk, m.rw_nslice, f_start, f_step, tzv, mu);
        const size_t prow = ((size_t)chain * s.rows + (size_t)(t - s.t0)) * N;
        // log-densities and the new state of all particles of the thread, branch-free; the stores follow
        double la[ST_PP], lr[ST_PP], ll[ST_PP];
#pragma unroll
        for (int p = 0; p < ST_PP; ++p) {
            la[p] = gauss_loglik<NX, NY>(m, y, mu[p]);
            lr[p] = gauss_logpdf_state<NX>(sw, slogc[0], ref, mu[p]);
            const bool is_ref = ip[p] == N - 1;
#pragma unroll
            for (int r = 0; r < NX; ++r) {
                double v = mu[p][r];
#pragma unroll
                for (int c = 0; c <= r; ++c) v = fma(chol[r * NX + c], z[p][c], v);
                x[p][r] = is_ref ? ref[r] : v;                              // src/PGAS.py:134
            }
            ll[p] = gauss_loglik<NX, NY>(m, y, x[p]);
        }
#pragma unroll
        for (int p = 0; p < ST_PP; ++p) {
            if (val[p]) {
                double* out = a.state_trace + (((size_t)chain * a.trace_rows + t) * N + ip[p]) * NX;
                if constexpr (NX == 2) *reinterpret_cast<double2*>(out) = make_double2(x[p][0], x[p][1]);
                else {
#pragma unroll
                    for (int k = 0; k < NX; ++k) out[k] = x[p][k];
                }
                s.la[prow + ip[p]] = la[p];
                s.lr[prow + ip[p]] = lr[p];
                s.ll[prow + ip[p]] = ll[p];
            }
        }
    }
#pragma unroll
    for (int p = 0; p < ST_PP; ++p)
        if (val[p]) {
#pragma unroll
            for (int k = 0; k < NX; ++k) s.x_carry[((size_t)chain * N + ip[p]) * NX + k] = x[p][k];
        }
}

static int split_chunk_rows(const DevModel& m, int N, int n_chains) {
    // rows per chunk: enough work to amortise two launches, small enough that the 24 B per particle-step stay modest
    const size_t per_row = (size_t)n_chains * N * 3 * sizeof(double);
    size_t rows = (size_t)(768ull << 20) / std::max<size_t>(per_row, 1);      // <= 768 MB per buffer
    rows = std::min<size_t>(std::max<size_t>(rows, 8), 64);
    if (const char* e = getenv("PGAS_SPLIT_ROWS")) { const int v = atoi(e); if (v >= 8) rows = std::min<size_t>(rows, (size_t)v); }   // developer override
    return (int)std::min<size_t>(rows, (size_t)std::max(m.T - 1, 1));
}

size_t pgas_sweep_split_workspace(const DevModel& m, int N, int n_chains) {
    if (!((m.D == 2 || m.D == 3) && m.rw_ok) || (m.flags & PGAS_FLAG_ANCESTOR_GATHER)) return 256;
    const size_t rows = split_chunk_rows(m, N, n_chains);
    const size_t pre = 2 * rows * (size_t)n_chains * N * 3 * sizeof(double);
    const size_t carry = (size_t)n_chains * N * (m.n_x + 2) * sizeof(double);
    return pre + carry + 1024;
}

bool pgas_sweep_split_eligible(const SweepArgs& a) {
    const DevModel& m = a.m;
    if (!((m.D == 2 || m.D == 3) && m.rw_ok) || (m.flags & PGAS_FLAG_ANCESTOR_GATHER)) return false;
    if (m.lik_len) return false;                 // likelihood program (model plug-in): fused kernel
    if (!((m.n_x == 2 && m.n_y == 1) || (m.n_x == 2 && m.n_y == 2 && m.D == 2))) return false;
    if (a.init_state || a.init_logw || a.dbg || a.row_off != 0 || a.t_begin != 1 || a.t_end != m.T) return false;   // full sweeps only
    if (a.t_end - a.t_begin < 16 || !a.logw_last) return false;
    if (getenv("PGAS_SWEEP_FUSED")) return false;                             // developer override
    return a.ws && a.ws_bytes >= pgas_sweep_split_workspace(m, a.N, a.n_chains);
}

static long long* g_dbg_split_ticks = nullptr;     // developer aid (tools/ticks_split.py): phase clocks of the resampling kernel
extern "C" int pgas_debug_set_split_ticks(long long* dev_buf) { g_dbg_split_ticks = dev_buf; return 0; }

#ifndef PGAS_SPLIT_GROUPS
#define PGAS_SPLIT_GROUPS 2
#endif
#ifndef PGAS_STATE_MMA_DEFAULT
#define PGAS_STATE_MMA_DEFAULT 0    // 1: the DMMA form of the contraction is the default in the big geometry
#endif
#ifndef PGAS_LAT_MAX_CHAINS
#define PGAS_LAT_MAX_CHAINS 32      // up to this many chains per launch the latency form of the resampling kernel is the default (A/B on B200: profiles/r02_strong_scaling.md)
#endif
constexpr int SPLIT_GROUPS = PGAS_SPLIT_GROUPS;      // chain groups of the state kernel: each on its own stream, so that a group's next
                                     // launch starts as soon as ITS CTAs retire (no wave-quantisation tail across all chains)
struct SplitStreams {
    cudaStream_t aux = nullptr, auxg[SPLIT_GROUPS] = {};
    cudaEvent_t start = nullptr, k1[2][SPLIT_GROUPS] = {}, k2[2] = {nullptr, nullptr};
    int device = -1;
};
static thread_local SplitStreams g_split;

static int split_streams_init() {
    int dev = 0;
    PGAS_CUDA(cudaGetDevice(&dev));
    if (g_split.aux && g_split.device == dev) return 0;
    g_split = SplitStreams();
    g_split.device = dev;
    {   // the state kernel runs AHEAD of the resampling kernel; give the (latency-bound) resampling kernel on the caller's
        // stream the first pick of freed SM resources by putting the state kernel on the lowest-priority stream
        int lo = 0, hi = 0;
        PGAS_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        for (int g = 0; g < SPLIT_GROUPS; ++g) PGAS_CUDA(cudaStreamCreateWithPriority(&g_split.auxg[g], cudaStreamNonBlocking, lo));
        g_split.aux = g_split.auxg[0];
    }
    PGAS_CUDA(cudaEventCreateWithFlags(&g_split.start, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        for (int g = 0; g < SPLIT_GROUPS; ++g) PGAS_CUDA(cudaEventCreateWithFlags(&g_split.k1[i][g], cudaEventDisableTiming));
        PGAS_CUDA(cudaEventCreateWithFlags(&g_split.k2[i], cudaEventDisableTiming));
    }
    return 0;
}

// geometry of the state kernel: big CTAs (256 threads, 2 particles per thread) unless they would leave SMs idle (fewer than ~1.5 CTAs
// per SM slot pair); judged on the whole sweep, not on one chain group; PGAS_STATE_SMALL=0|1 forces one (developer override)
static bool state_geometry_small(int N, int n_chains) {
    int sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const int bpc_big = (N + ST_PP_BIG * ST_NT_BIG - 1) / (ST_PP_BIG * ST_NT_BIG);
    bool small = n_chains * bpc_big < (3 * sms) / 2;
    if (const char* e = getenv("PGAS_STATE_SMALL")) small = atoi(e) != 0;
    return small;
}

static int launch_state(const StateArgs& s_in, cudaStream_t st) {
    StateArgs s = s_in;
    const DevModel& m = s.a.m;
    const bool inj = s.a.rng_mode == 1;
    bool small = state_geometry_small(s.a.N, s.a.n_chains);
    // contraction on DMMA tiles (basis_mma.cuh) in the big geometry of two-dimensional bases; PGAS_STATE_MMA=0|1 forces
    bool mma = !small && m.D == 2 && m.mma_ok && ST_PP_BIG == 2 && PGAS_STATE_MMA_DEFAULT;
    if (const char* e = getenv("PGAS_STATE_MMA")) mma = atoi(e) != 0 && !small && m.D == 2 && m.mma_ok && ST_PP_BIG == 2;
    const int n_theta = mma ? m.mma_slots : m.rw_slots;
    const size_t smem = sizeof(double) * (((size_t)n_theta + 1) & ~(size_t)1) + sizeof(double) * (2 * m.n_x * m.n_x + 2) + sizeof(int) * RW_MAXBLK +
                        (mma ? sizeof(double) * (size_t)MMA_WARP_DOUBLES * (ST_NT_BIG / 32) : 0) + 32;
    // few particles on a three-dimensional basis: ST_LANES lanes share a particle (judged on N alone, so that a chain's bits do
    // not depend on how many chains run next to it); PGAS_STATE_LANES=0|1 forces
    bool lanes = m.D == 3 && s.a.N <= ST_LANES_MAX_N;
    if (const char* e = getenv("PGAS_STATE_LANES")) lanes = atoi(e) != 0 && m.D == 3;
    if (lanes) small = true;
    const int per = lanes ? ST_NT_SMALL / ST_LANES : small ? ST_PP_SMALL * ST_NT_SMALL : ST_PP_BIG * ST_NT_BIG;
    s.bpc = (s.a.N + per - 1) / per;
    const dim3 grid((unsigned)(s.nch * s.bpc));
#define PGAS_ST_LAUNCH(NYv, INJv, NTv, PPv, Dv, MMAv) do { \
        PGAS_CUDA(cudaFuncSetAttribute(csmc_state_kernel<2, NYv, INJv, NTv, PPv, Dv, MMAv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        csmc_state_kernel<2, NYv, INJv, NTv, PPv, Dv, MMAv><<<grid, NTv, smem, st>>>(s); } while (0)
#define PGAS_ST_GEOM(NYv, INJv, Dv) do { if (small) PGAS_ST_LAUNCH(NYv, INJv, ST_NT_SMALL, ST_PP_SMALL, Dv, false); \
        else if (mma && Dv == 2) PGAS_ST_LAUNCH(NYv, INJv, ST_NT_BIG, 2, 2, true); else PGAS_ST_LAUNCH(NYv, INJv, ST_NT_BIG, ST_PP_BIG, Dv, false); } while (0)
    if (m.D == 3 && lanes) {
        if (inj) { PGAS_CUDA(cudaFuncSetAttribute(csmc_state_kernel<2, 1, true, ST_NT_SMALL, 1, 3, false, ST_LANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                   csmc_state_kernel<2, 1, true, ST_NT_SMALL, 1, 3, false, ST_LANES><<<grid, ST_NT_SMALL, smem, st>>>(s); }
        else { PGAS_CUDA(cudaFuncSetAttribute(csmc_state_kernel<2, 1, false, ST_NT_SMALL, 1, 3, false, ST_LANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
               csmc_state_kernel<2, 1, false, ST_NT_SMALL, 1, 3, false, ST_LANES><<<grid, ST_NT_SMALL, smem, st>>>(s); }
    }
    else if (m.D == 3) { if (inj) PGAS_ST_GEOM(1, true, 3); else PGAS_ST_GEOM(1, false, 3); }
    else if (m.n_y == 1) { if (inj) PGAS_ST_GEOM(1, true, 2); else PGAS_ST_GEOM(1, false, 2); }
    else { if (inj) PGAS_ST_GEOM(2, true, 2); else PGAS_ST_GEOM(2, false, 2); }
#undef PGAS_ST_GEOM
#undef PGAS_ST_LAUNCH
    PGAS_KERNEL_CHECK();
    return 0;
}

int pgas_launch_sweep(const SweepArgs& a, cudaStream_t stream) {
    if (!pgas_sweep_split_eligible(a)) return pgas_launch_sweep_fused(a, stream);
    const DevModel& m = a.m;
    if (int rc = split_streams_init()) return rc;
    const int rows = split_chunk_rows(m, a.N, a.n_chains);
    char* base = (char*)(((uintptr_t)a.ws + 255) & ~(uintptr_t)255);
    const size_t buf = (size_t)rows * a.n_chains * a.N;            // doubles per array per buffer
    double* pre = (double*)base;
    double* x_carry = pre + 2 * 3 * buf;
    double* lw_carry = x_carry + (size_t)a.n_chains * a.N * m.n_x;
    const bool serial = getenv("PGAS_SPLIT_SERIAL") != nullptr;                // developer override: no overlap
    const int ngroups = (!serial && a.n_chains >= 2 * SPLIT_GROUPS && !getenv("PGAS_SPLIT_ONE_GROUP")) ? SPLIT_GROUPS : 1;
    cudaStream_t sg[SPLIT_GROUPS];
    for (int g = 0; g < SPLIT_GROUPS; ++g) sg[g] = serial ? stream : g_split.auxg[g];
    PGAS_CUDA(cudaEventRecord(g_split.start, stream));
    const bool timeline = getenv("PGAS_SPLIT_TIMELINE") != nullptr;
    std::vector<cudaEvent_t> tl_state, tl_w;
    cudaEvent_t tl_start = nullptr;
    if (timeline) { cudaEventCreate(&tl_start); cudaEventRecord(tl_start, stream); }
    for (int g = 0; g < ngroups; ++g)
        if (!serial) PGAS_CUDA(cudaStreamWaitEvent(sg[g], g_split.start, 0));  // inputs (Theta, Sigma, ref) are ready
    // chunk boundaries: a short first chunk (the resampling kernel can start early) and a short last chunk (little
    // resampling work is left when the state kernel has finished), `rows` in between
    const int edge = std::min(rows, 16);
    int c = 0;
    for (int t0 = a.t_begin, t1 = 0; t0 < a.t_end; t0 = t1, ++c) {
        const int left = a.t_end - t0, b = c & 1;
        int len;
        if (left <= edge) len = left;                          // last chunk
        else if (c == 0) len = edge;                           // first chunk
        else if (left <= rows + edge) len = left - edge;       // leave exactly `edge` rows for the last chunk
        else len = rows;
        t1 = t0 + len;
        double* la = pre + (size_t)b * 3 * buf;
        StateArgs s;
        s.a = a;
        s.x_carry = x_carry; s.la = la; s.lr = la + buf; s.ll = la + 2 * buf;
        s.t0 = t0; s.t1 = t1; s.rows = rows; s.first = (c == 0); s.bpc = 0;     // set by launch_state
        for (int g = 0; g < ngroups; ++g)
            if (c >= 2) PGAS_CUDA(cudaStreamWaitEvent(sg[g], g_split.k2[b], 0));   // buffer b was consumed by chunk c-2
        {   // short-lived state CTAs (sub-chunks) so that resampling CTAs find free slots quickly
            int sr = 16;
            if (const char* e = getenv("PGAS_SPLIT_STATE_ROWS")) { const int v = atoi(e); if (v >= 1) sr = v; }
            for (int ts = t0; ts < t1; ts += sr) {
                StateArgs q = s;
                q.t0 = ts; q.t1 = std::min(ts + sr, t1);
                q.la = s.la + (size_t)(ts - t0) * a.N; q.lr = s.lr + (size_t)(ts - t0) * a.N; q.ll = s.ll + (size_t)(ts - t0) * a.N;
                q.first = s.first && ts == t0;
                for (int g = 0; g < ngroups; ++g) {
                    q.chain0 = (int)((long long)a.n_chains * g / ngroups);
                    q.nch = (int)((long long)a.n_chains * (g + 1) / ngroups) - q.chain0;
                    if (int rc = launch_state(q, sg[g])) return rc;
                }
            }
        }
        for (int g = 0; g < ngroups; ++g) PGAS_CUDA(cudaEventRecord(g_split.k1[b][g], sg[g]));
        if (timeline) {                                                       // developer aid: PGAS_SPLIT_TIMELINE=1
            cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, sg[0]); tl_state.push_back(e);
        }
        SweepArgs r = a;
        r.dbg = g_dbg_split_ticks;
        r.t_begin = t0; r.t_end = t1;
        r.pre_la = s.la; r.pre_lr = s.lr; r.pre_ll = s.ll; r.pre_rows = rows; r.pre_off = t0;
        r.init_logw = (c == 0) ? nullptr : lw_carry;
        r.logw_last = (t1 == a.t_end) ? a.logw_last : lw_carry;
        r.ref = nullptr; r.Theta = nullptr; r.Sigma = nullptr; r.state_trace = nullptr;
        if (const char* e = getenv("PGAS_SPLIT_PRE_C")) {                      // developer override: cluster size of the resampling kernel
            const int v = atoi(e);
            if (v >= 1 && v <= 16) { r.C = v; r.P = (a.N + v - 1) / v; }
        }
        for (int g = 0; g < ngroups; ++g) PGAS_CUDA(cudaStreamWaitEvent(stream, g_split.k1[b][g], 0));
        const int wc = getenv("PGAS_SPLIT_PRE_C") ? 0 : pgas_weights_cluster(a.N);
        // few chains: the per-step latency of the recursion bounds the sweep -> latency form (weights_lat.cu); many chains: the
        // FP64 pipe does, and one CTA per chain (weights.cu) leaves the most room for the state kernel.  PGAS_WEIGHTS_KERNEL=3 / 1
        // force one or the other (developer override).
        const char* wk_env = getenv("PGAS_WEIGHTS_KERNEL");
        if (wk_env && !*wk_env) wk_env = nullptr;                                  // set but empty: not an override
        const int lc = getenv("PGAS_SPLIT_PRE_C") ? 0 : pgas_weights_lat_cluster(a.N);
        // default only with portable cluster sizes (N <= 16384: 2 | 4 | 8 particles per thread, weights_lat.cu): clusters of 16 are
        // co-resident 7 at a time on B200 — at configs[4] (16 chains of N = 16384) they measured 424 ms per iteration, the general
        // kernel 332-342, clusters of 8 with eight particles per thread 301
        // ... and only where its CTAs can share an SM with the state CTAs of this sweep: the two- / four-particle instantiations take
        // 146+ registers, next to which a 256-thread state CTA (half of the register file) does not fit — 28 to 32 chains of
        // N = 4096 measured 38.0 ms per iteration that way against 31.9 with one CTA per chain (2-GPU split of configs[3])
        const bool lat_coresident = pgas_weights_lat_fits_beside_big_state(a.N) || state_geometry_small(a.N, a.n_chains) ||
                                    (m.D == 3 && a.N <= ST_LANES_MAX_N);
        const bool use_lat = lc > 0 && (wk_env ? atoi(wk_env) == 3 : (a.n_chains <= PGAS_LAT_MAX_CHAINS && lc <= 8 && lat_coresident));
        if (use_lat) {
            r.C = lc; r.P = (a.N + lc - 1) / lc;
            if (int rc = pgas_launch_weights_lat(r, stream)) return rc;
        } else if (wc > 0) {                                                          // dedicated kernel (weights.cu): the chain's CDF fits one CTA
            r.C = wc; r.P = (a.N + wc - 1) / wc;
            if (int rc = pgas_launch_weights(r, stream)) return rc;
        } else if (int rc = pgas_launch_sweep_pre(r, stream)) return rc;
        PGAS_CUDA(cudaEventRecord(g_split.k2[b], stream));
        if (timeline) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, stream); tl_w.push_back(e); }
    }
    if (timeline) {
        cudaStreamSynchronize(stream);
        fprintf(stderr, "chunk: state group 0 done / resampling done [ms since sweep start]\n");
        for (size_t i = 0; i < tl_w.size(); ++i) {
            float ts = 0, tw = 0;
            cudaEventElapsedTime(&ts, tl_start, tl_state[i]);
            cudaEventElapsedTime(&tw, tl_start, tl_w[i]);
            fprintf(stderr, "  %2zu  %8.3f  %8.3f\n", i, ts, tw);
            cudaEventDestroy(tl_state[i]); cudaEventDestroy(tl_w[i]);
        }
        cudaEventDestroy(tl_start);
    }
    return 0;
}


// Measurement aid (bench.py roofline): the state kernel ALONE over all T-1 steps on `stream`, launched exactly as
// pgas_launch_sweep launches it (same sub-chunks, same buffers) but without the resampling kernel.
extern "C" int pgas_debug_state_kernel_f64(const pgas_model* model, int32_t N, int32_t n_chains, const double* ref, const double* Theta,
                                           const double* Sigma, const pgas_rng* rng, double* state_trace, void* workspace,
                                           size_t workspace_bytes, void* stream) {
    if (!model || !ref || !Theta || !Sigma || !rng || !state_trace || !workspace) PGAS_FAIL(-1, "pgas_debug_state_kernel_f64: null argument");
    SweepArgs a;
    memset(&a, 0, sizeof(a));
    a.m = model->dev;
    const DevModel& m = a.m;
    a.N = N; a.n_chains = n_chains; a.C = 1; a.P = N;
    a.t_begin = 1; a.t_end = m.T;
    a.ref_rows = m.T; a.trace_rows = m.T; a.anc_rows = m.T - 1; a.var_rows = m.T;
    a.ref = ref; a.ref_stride = (long long)m.T * m.n_x; a.Theta = Theta; a.Sigma = Sigma;
    a.state_trace = state_trace;
    a.rng_mode = rng->mode; a.seed = rng->seed; a.chain_base = rng->chain_base; a.iteration = rng->iteration; a.Z = rng->Z; a.U = rng->U;
    a.logw_last = (double*)workspace;                     // only to pass the eligibility test
    a.ws = workspace; a.ws_bytes = workspace_bytes;
    if (!pgas_sweep_split_eligible(a)) PGAS_FAIL(-2, "the split form does not apply to this model / workspace");
    const int rows = split_chunk_rows(m, N, n_chains);
    char* base = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    const size_t buf = (size_t)rows * n_chains * N;
    double* pre = (double*)base;
    double* x_carry = pre + 2 * 3 * buf;
    // the same chain groups on the same low-priority streams as pgas_launch_sweep
    if (split_streams_init()) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    const int ngroups = (n_chains >= 2 * SPLIT_GROUPS && !getenv("PGAS_SPLIT_ONE_GROUP")) ? SPLIT_GROUPS : 1;
    PGAS_CUDA(cudaEventRecord(g_split.start, st));
    for (int g = 0; g < ngroups; ++g) PGAS_CUDA(cudaStreamWaitEvent(g_split.auxg[g], g_split.start, 0));
    int c = 0;
    for (int t0 = a.t_begin; t0 < a.t_end; t0 += rows, ++c) {
        const int t1 = std::min(t0 + rows, a.t_end);
        double* la = pre + (size_t)(c & 1) * 3 * buf;
        for (int ts = t0; ts < t1; ts += 16) {
            StateArgs q;
            q.a = a;
            q.x_carry = x_carry;
            q.t0 = ts; q.t1 = std::min(ts + 16, t1); q.rows = rows;
            q.la = la + (size_t)(ts - t0) * N; q.lr = la + buf + (size_t)(ts - t0) * N; q.ll = la + 2 * buf + (size_t)(ts - t0) * N;
            q.first = (c == 0 && ts == t0); q.bpc = 0;
            for (int g = 0; g < ngroups; ++g) {
                q.chain0 = (int)((long long)n_chains * g / ngroups);
                q.nch = (int)((long long)n_chains * (g + 1) / ngroups) - q.chain0;
                if (int rc = launch_state(q, g_split.auxg[g])) return rc;
            }
        }
    }
    for (int g = 0; g < ngroups; ++g) {
        PGAS_CUDA(cudaEventRecord(g_split.k1[0][g], g_split.auxg[g]));
        PGAS_CUDA(cudaStreamWaitEvent(st, g_split.k1[0][g], 0));
    }
    return 0;
}
