// sweep.cu — persistent conditional-SMC sweep with ancestor sampling (sm_100a).
//
// Replaces condSequentialMonteCarlo.step / __call__ (reference src/PGAS.py:79-153, :176-228):
// one CTA — or one thread-block cluster of C CTAs for large N — owns one chain for ALL T steps.
// The particle set (state, log-weight, auxiliary mean, CDF) lives in shared memory / DSMEM; per
// step the only HBM traffic is the trace (state row + ancestor row).  Each step, in one pass:
//   A  mu_i = Theta phi(x_{t-1}^i, u_t)  (basis_eval.cuh), l_aux, h;  CTA max + exp + CTA scan
//   X1 per-CTA (max, sum) pairs all-gathered through DSMEM  -> cluster barrier #1
//   B  global log-sum-exp / CDF offsets; systematic resampling by the owner of the CDF segment
//      (binary search in local shared memory), ancestor sampling of the reference particle;
//      l_aux[a_j] (and mu[a_j] in gather mode) pushed to the owner of j through DSMEM
//   X2 cluster barrier #2 — its latency is covered by the noise draw and the new state
//   C  x_t^i = mu + chol(Sigma) z,  logw_t^i = log p(y_t|x_t^i) - l_aux[a_i],  trace row written
#include <cooperative_groups.h>
#include "basis_eval.cuh"
#include "sweep_args.cuh"

namespace cg = cooperative_groups;

constexpr int NT = 256;          // threads per CTA
constexpr int NW = NT / 32;
constexpr int MAXC = 16;         // largest (non-portable) cluster


__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }

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

// value of the resampling CDF at a local inclusive prefix `p` (same expression for the segment
// boundaries and the interior so they agree bit for bit)
__device__ __forceinline__ double cdf_value(double p, double f, double g, double rs) {
    return __dmul_rn(__fma_rn(p, f, g), rs);
}
__device__ __forceinline__ double clip01(double v) { return fmin(fmax(v, 0.0), 1.0); }

// U_j = (u + j) / N exactly as src/Filtering.py:28 evaluates it
__device__ __forceinline__ double strat_point(double u, int j, double dN) { return __ddiv_rn(__dadd_rn(u, (double)j), dN); }

// smallest j in [0,N] with U_j > b  (U_j is non-decreasing in j)
__device__ __forceinline__ int first_point_above(double b, double u, int N, double dN) {
    double g = floor(b * dN - u);
    int j = (g < 0.0) ? 0 : (g > (double)N ? N : (int)g);
    while (j > 0 && strat_point(u, j - 1, dN) > b) --j;
    while (j < N && !(strat_point(u, j, dN) > b)) ++j;
    return j;
}

// per-step constants staged in shared memory (double-buffered by step parity)
struct StepConst {
    double y[PGAS_MAX_NY];
    double u[PGAS_MAX_NU];
    double ref[PGAS_MAX_NX];
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
    for (int k = 0; k < m.n_x; ++k) sc->ref[k] = a.ref[(size_t)chain * a.ref_stride + (size_t)(t - a.row_off) * m.n_x + k];
    if (a.rng_mode == 1) {
        const double* up = a.U + ((size_t)chain * a.var_rows + (t - a.row_off)) * 2;
        sc->ures = up[0];
        sc->uanc = up[1];
    } else {
        philox_uniform2(a.seed, PURPOSE_STEP_U, a.chain_base + chain, a.iteration, (unsigned)t, 0u, sc->ures, sc->uanc);
    }
}

template <int NX, int NY, int D, int JMAX>
__global__ void __launch_bounds__(NT, 1) csmc_sweep_kernel(const __grid_constant__ SweepArgs a) {
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
    const double dN = (double)N;

    // ------------------------------------------------------------------ shared memory carve
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sp = reinterpret_cast<double*>(smem_raw);
    double* th = sp;            sp += ((size_t)m.n_packed * NX + 1) & ~(size_t)1;
    double* xs = sp;            sp += (size_t)NX * P;          // [k][i]
    double* mus = sp;           sp += (size_t)NX * P;          // [k][i] auxiliary mean
    double* logw = sp;          sp += P;
    double* laux = sp;          sp += P;
    double* b1 = sp;            sp += P;                       // lw_aux -> prefix -> CDF
    double* b2 = sp;            sp += P;                       // lw_anc -> prefix
    double* lauxg = sp;         sp += P;                       // l_aux[a_i], pushed by the CDF owner
    double* mug = sp;           sp += gather ? (size_t)NX * P : 0;
    double* exch = sp;          sp += MAXC * 4;                // per-CTA (m1,s1,m2,s2), all-gathered
    double* gsum = sp;          sp += 2 * (MAXC + 1);          // exclusive CDF offsets G1[c], G2[c], c = 0..C
    double* fx = sp;            sp += 2 * MAXC + 2;            // rescale factors exp(m_c - M); then 1/S1, 1/S2
    double* wtmax = sp;         sp += NW * 2;
    double* wtscan = sp;        sp += 2 * NW * 2;
    double* chol = sp;          sp += NX * NX;                 // chol(Sigma) lower
    double* sw = sp;            sp += NX * NX;                 // chol(Sigma)^-1
    double* slogc = sp;         sp += 2;
    StepConst* sc = reinterpret_cast<StepConst*>(sp);  sp += 2 * ((sizeof(StepConst) + 7) / 8);
    int* ip = reinterpret_cast<int*>(sp);
    int* meta = ip;             ip += m.n_chunks;
    int* cnt = ip;              ip += 2;

    // ------------------------------------------------------------------ prologue
    for (int r = tid; r < m.n_chunks; r += NT) meta[r] = m.chunk_meta[r];
    {   // Theta' = norm * Theta in packed order
        const double* Th = a.Theta + (size_t)chain * NX * m.M;
        for (int s = tid; s < m.n_packed; s += NT) {
            const int mm = m.perm[s];
#pragma unroll
            for (int k = 0; k < NX; ++k) th[(size_t)s * NX + k] = (mm >= 0) ? m.norm * Th[(size_t)k * m.M + mm] : 0.0;
        }
    }
    if (tid == 0) {
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

        // ---- A1: auxiliary mean, first-stage log-weights (src/PGAS.py:89-101, :109-117)
        double tmax1 = -INFINITY, tmax2 = -INFINITY;
        for (int q = 0; q < PPT; ++q) {
            const int il = q * NT + tid;
            if (il < Pc) {
                double x[NX], tz[D], mu[NX];
#pragma unroll
                for (int k = 0; k < NX; ++k) x[k] = xs[(size_t)k * P + il];
                gp_input<NX, D>(m, x, k_t.u, tz);
                eval_mu<NX, D, JMAX>(th, meta, m.n_chunks, m.f_start, m.f_step, tz, mu);
                const double la = gauss_loglik<NX, NY>(m, k_t.y, mu);
                const double lwa = la + logw[il];
                const double h = gauss_logpdf_state<NX>(sw, slogc[0], k_t.ref, mu);
                const double lwr = lwa + h;
#pragma unroll
                for (int k = 0; k < NX; ++k) mus[(size_t)k * P + il] = mu[k];
                laux[il] = la;
                b1[il] = lwa;
                b2[il] = lwr;
                tmax1 = fmax(tmax1, lwa);
                tmax2 = fmax(tmax2, lwr);
            }
        }
        tmax1 = warp_max(tmax1);
        tmax2 = warp_max(tmax2);
        if (lane == 0) { wtmax[warp * 2] = tmax1; wtmax[warp * 2 + 1] = tmax2; }
        __syncthreads();
        double m1c = wtmax[0], m2c = wtmax[1];
#pragma unroll
        for (int w = 1; w < NW; ++w) { m1c = fmax(m1c, wtmax[w * 2]); m2c = fmax(m2c, wtmax[w * 2 + 1]); }

        // ---- A2: exp and CTA-wide inclusive scan in particle order (softmax numerators, src/PGAS.py:102,118)
        double carry1 = 0.0, carry2 = 0.0;
        for (int q = 0; q < PPT; ++q) {
            const int il = q * NT + tid;
            const bool v = il < Pc;
            const double e1 = v ? exp(b1[il] - m1c) : 0.0;
            const double e2 = v ? exp(b2[il] - m2c) : 0.0;
            const double s1 = warp_scan_incl(e1, lane), s2 = warp_scan_incl(e2, lane);
            double* wt = wtscan + (q & 1) * NW * 2;
            if (lane == 31) { wt[warp * 2] = s1; wt[warp * 2 + 1] = s2; }
            __syncthreads();
            double run1 = carry1, run2 = carry2, my1 = 0.0, my2 = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                if (w == warp) { my1 = run1; my2 = run2; }
                run1 += wt[w * 2];
                run2 += wt[w * 2 + 1];
            }
            carry1 = run1;
            carry2 = run2;
            if (v) { b1[il] = my1 + s1; b2[il] = my2 + s2; }
        }

        // ---- X1: all-gather (m1c, s1c, m2c, s2c) across the cluster
        if (C > 1) {
            if (tid < C) {
                double* dst = cluster.map_shared_rank(exch, tid) + rank * 4;
                dst[0] = m1c; dst[1] = carry1; dst[2] = m2c; dst[3] = carry2;
            }
            cluster_arrive();
            cluster_wait();
        } else {
            if (tid == 0) { exch[0] = m1c; exch[1] = carry1; exch[2] = m2c; exch[3] = carry2; }
            __syncthreads();
        }

        // ---- B1: global normalisers and this CTA's CDF segment (warp 0 reduces the C pairs once)
        const int c_last = (N - 1) / P;                       // CTA owning particle N-1 (later CTAs are empty)
        if (warp == 0) {
            const bool vc = lane < C;
            const double mc1 = vc ? exch[lane * 4] : -INFINITY, mc2 = vc ? exch[lane * 4 + 2] : -INFINITY;
            const double M1 = warp_max(mc1), M2 = warp_max(mc2);
            if (vc) {
                fx[lane * 2] = (mc1 == -INFINITY) ? 0.0 : exp(mc1 - M1);
                fx[lane * 2 + 1] = (mc2 == -INFINITY) ? 0.0 : exp(mc2 - M2);
            }
            __syncwarp();
            if (lane == 0) {
                double g1 = 0.0, g2 = 0.0;
                for (int c = 0; c < C; ++c) {                 // fixed order: every CTA of the cluster gets identical bits
                    gsum[c * 2] = g1;
                    gsum[c * 2 + 1] = g2;
                    g1 = __fma_rn(exch[c * 4 + 1], fx[c * 2], g1);
                    g2 = __fma_rn(exch[c * 4 + 3], fx[c * 2 + 1], g2);
                }
                gsum[C * 2] = g1;
                gsum[C * 2 + 1] = g2;
                const double r1 = 1.0 / g1, r2 = 1.0 / g2;
                fx[2 * MAXC] = r1;
                fx[2 * MAXC + 1] = r2;
                // CTA holding the reference ancestor: number of leading CTAs whose whole CDF segment lies below u_anc
                int cs = 0;
                while (cs <= c_last && __dmul_rn(gsum[(cs + 1) * 2 + 1], r2) < k_t.uanc) ++cs;
                cnt[1] = cs;
            }
        }
        if (tid == 0 && t + 1 < a.t_end) sc[(t + 1) & 1] = nxt;
        __syncthreads();
        const double myg1 = gsum[rank * 2], myg2 = gsum[rank * 2 + 1], g1hi = gsum[(rank + 1) * 2];
        const double myf1 = fx[rank * 2], myf2 = fx[rank * 2 + 1];
        const double S1 = fx[2 * MAXC], S2 = fx[2 * MAXC + 1];      // reciprocals of the normalisers
        const int cstar = cnt[1];
        int mycnt = 0;
        for (int q = 0; q < PPT; ++q) {
            const int il = q * NT + tid;
            if (il < Pc) {
                // W = clip(cumsum(w / sum w), 0, 1)  (src/Filtering.py:23-32)
                b1[il] = clip01(cdf_value(b1[il], myf1, myg1, S1));
                // cumsum(softmax(lw_anc)) < u_anc  (src/PGAS.py:118-124), not clipped
                if (rank == cstar) mycnt += (cdf_value(b2[il], myf2, myg2, S2) < k_t.uanc) ? 1 : 0;
            }
        }
        if (rank == cstar) {
            mycnt = __reduce_add_sync(0xffffffffu, mycnt);
            if (lane == 0 && mycnt) atomicAdd(&cnt[0], mycnt);
        }
        __syncthreads();

        // ---- B2: systematic resampling (src/Filtering.py:28-35) by the owner of the CDF segment
        {
            const double blo = clip01(__dmul_rn(myg1, S1)), bhi = clip01(__dmul_rn(g1hi, S1));
            const int jlo = (rank == 0) ? 0 : first_point_above(blo, k_t.ures, N, dN);
            int jhi = (rank == c_last) ? N : first_point_above(bhi, k_t.ures, N, dN);
            if (rank > c_last) jhi = jlo;                     // empty CTA
            int* anc_row = a.anc_trace + ((size_t)chain * a.anc_rows + (t - 1 - a.row_off + a.anc_shift)) * N;
            for (int j = jlo + tid; j < jhi; j += NT) {
                if (j == N - 1) continue;                     // overwritten by the reference ancestor (:127)
                const double uj = strat_point(k_t.ures, j, dN);
                int lo = 0, hi = Pc;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (b1[mid] < uj) lo = mid + 1; else hi = mid;
                }
                const int k = min(lo, Pc - 1);
                anc_row[j] = base + k;
                const int cj = j / P, jl = j - cj * P;
                double* dl = (C > 1 && cj != rank) ? cluster.map_shared_rank(lauxg, cj) : lauxg;
                dl[jl] = laux[k];
                if (gather) {
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
                if (gather) {
                    double* dm = (C > 1 && cj != rank) ? cluster.map_shared_rank(mug, cj) : mug;
#pragma unroll
                    for (int kk = 0; kk < NX; ++kk) dm[(size_t)kk * P + jl] = mus[(size_t)kk * P + k];
                }
            }
        }

        // ---- X2 + C: barrier #2 overlapped with the noise draw; new state, new log-weights
        if (C > 1) cluster_arrive(); else __syncthreads();
        const bool last_step = (t + 1 == a.t_end);
        constexpr int ZQ = 4;                                 // particles per thread whose noise is drawn under the barrier
        double zreg[ZQ][NX];
#pragma unroll
        for (int q = 0; q < ZQ; ++q) {
            const int il = q * NT + tid;
            if (q < PPT && il < Pc) draw_normals<NX>(a, chain, t, base + il, zreg[q]);
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
            const double lw = gauss_loglik<NX, NY>(m, k_t.y, x) - lauxg[il];   // :137-147
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
        // next step's A1 only touches this thread's own xs/logw entries and arrays whose previous
        // readers are fenced by the barriers above; b1/b2/laux/mus are rewritten after every
        // thread of the CTA has left B2, which the X2 barrier guarantees.
    }
    if (C > 1) cluster.sync();                                // no CTA exits while peers may still push into its smem
}

// ------------------------------------------------------------------------------------ launch
static size_t sweep_smem_bytes(const DevModel& m, int NX, int P, bool gather) {
    size_t d = (((size_t)m.n_packed * NX + 1) & ~(size_t)1) + (size_t)NX * P * 2 + (size_t)P * 5 + (gather ? (size_t)NX * P : 0) +
               MAXC * 4 + 2 * (MAXC + 1) + 2 * MAXC + 2 + NW * 2 + 2 * NW * 2 + 2 * NX * NX + 2 + 2 * ((sizeof(StepConst) + 7) / 8);
    return d * 8 + (size_t)(m.n_chunks + 2) * 4 + 16;
}

template <int NX, int NY, int D, int JMAX>
static int launch_variant(const SweepArgs& a, size_t smem, cudaStream_t stream) {
    auto kern = csmc_sweep_kernel<NX, NY, D, JMAX>;
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
    PGAS_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    return 0;
}

template <int NX, int NY, int D>
static int launch_jmax(const SweepArgs& a, size_t smem, cudaStream_t stream) {
    const int j = a.m.jmax;
    if (j <= 12) return launch_variant<NX, NY, D, 12>(a, smem, stream);
    if (j <= 20) return launch_variant<NX, NY, D, 20>(a, smem, stream);
    if (j <= 40) return launch_variant<NX, NY, D, 40>(a, smem, stream);
    PGAS_FAIL(-20, "basis needs %d lattice positions in its last dimension; this build supports <= 40", j);
}

int pgas_launch_sweep(const SweepArgs& a, cudaStream_t stream) {
    const DevModel& m = a.m;
    const bool gather = (m.flags & PGAS_FLAG_ANCESTOR_GATHER) != 0;
    const size_t smem = sweep_smem_bytes(m, m.n_x, a.P, gather);
    if (smem > 227 * 1024)
        PGAS_FAIL(-21, "sweep needs %zu bytes of shared memory per CTA (N=%d over a cluster of %d); use a larger cluster", smem, a.N, a.C);
#define PGAS_DISPATCH(NXv, NYv, Dv) \
    if (m.n_x == NXv && m.n_y == NYv && m.D == Dv) return launch_jmax<NXv, NYv, Dv>(a, smem, stream);
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
    return sweep_smem_bytes(m, m.n_x, P, (m.flags & PGAS_FLAG_ANCESTOR_GATHER) != 0);
}

// choose the cluster size: the smallest power of two whose per-CTA particle slice fits shared
// memory, then grown while the whole launch still fits the GPU (chains * C <= #SM) so that few
// chains still spread over many SMs.
int pgas_choose_cluster(const DevModel& m, int N, int n_chains, int requested) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (requested > 0) return requested;
    int C = 1;
    while (C < MAXC && pgas_sweep_smem_for(m, (N + C - 1) / C) > 200 * 1024) C *= 2;
    while (C < MAXC && n_chains * C * 2 <= sms && (N + 2 * C - 1) / (2 * C) >= 64) C *= 2;
    return C;
}
