// marginal.cu — the marginalised particle filters on sm_100a (SURVEY.md 8a group B).
//
//   Algorithm1.step / __call__   src/Algorithm1.py:298-397, :399-492   (online marginalised APF)
//   Algorithm3.step / __call__   src/Algorithm3.py:43-197, :199-303    (marginalised cSMC + ancestor sampling)
//   Algorithm2.__call__          src/Algorithm2.py:106-187             (PGAS outer loop)
//   prior_mniw_mean / _2naturalPara_inv / _Predictive / _drawPred / _log_base_measure
//                                src/BayesianInferrence.py:35-124
//
// One persistent kernel runs all T steps of a chain: a thread-block cluster of CS CTAs owns the chain,
// ONE WARP owns a particle.  Each particle carries, per GP, the MNIW statistics (T0 (M), T1 (M x M,
// packed lower triangle), T2, T3) in an L2-resident ping-pong workspace; the only cluster-wide
// dependency of a step is the resampling, so a step costs one hardware cluster barrier.  Every CTA
// recomputes the (tiny) softmax / CDF from the same global log-weights with the same code, so all CTAs
// agree bit for bit on the ancestors without a second barrier.
//
// What the reference does with ~6 batched M^3 operations per particle and step is restructured around
// ONE packed Cholesky per (particle, GP) [two in Algorithm3], shared by every consumer:
//   * the factor L of eta1 = prior1 + lambda T1 is computed once, right after the statistics update, for
//     the NEXT step;  it is "augmented" by two extra rows, eta0^T and phi(aux state)^T, so that the same
//     left-looking sweep also yields  y = L^-1 eta0,  v = L^-1 phi_aux,  Psi = eta2 - y.y.  Then
//       prior_mniw_mean . phi_aux          = y . v                     (src/Algorithm1.py:211-231)
//       log_base_measure(prior + stats)    from sum log L_jj^2 and Psi (src/Algorithm3.py:101-106)
//   * children of the particle (next step) read L and y:  c = 1 + |L^-1 phi|^2,  m = y . L^-1 phi give
//     the Student-t predictive (src/Algorithm1.py:235-274) with one forward solve — no M x M inverse;
//   * Algorithm3's g_T needs chol(prior1 + ref1 + T1): a second factorisation with eta0 + ref0 as the
//     extra row (src/Algorithm3.py:95-100).
// The per-step "prior + remaining reference statistics" tables are built once per sweep by
// marg_refstats_kernel in the reference's own (sequential) order of subtraction.
#include <cooperative_groups.h>
#include <algorithm>
#include <vector>
#include "marginal.cuh"

namespace cg = cooperative_groups;

#define FULL 0xffffffffu

// ---------------------------------------------------------------------------------- small helpers
__device__ __forceinline__ double ldcg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }

// Barrier among the warps of one chain through a global counter (all CTAs are co-resident: cooperative launch).  Same shape as a
// cooperative-groups grid sync — publish with fence + atomic, spin, fence — but with WARPS as participants: a warp arrives as soon
// as its own particle is written and leaves as soon as the count is complete; no CTA barrier on either side, so a warp neither
// waits for its three neighbours before publishing nor for thread 0's poll after the last arrival.  __syncwarp orders the lanes'
// stores before lane 0's fence (cumulativity), the fence before the add.
__device__ __forceinline__ void mg_warp_barrier(unsigned* ctr, unsigned target, int lane) {
    __syncwarp();
    if (lane == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned v;
        do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < target);
        __threadfence();
    }
    __syncwarp();
}

__device__ __forceinline__ void mg_cluster_barrier(bool multi) {
    if (multi) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        __syncthreads();
    }
}

// asynchronous 16-byte global -> shared copies (L2 only: the sources are rewritten by other SMs every other step)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// The small model functions below run on every lane with the same operands.  Their loops are written over the compile-time
// maxima with a guard on the run-time size, so that they unroll completely and the little vectors (x, xi, z, e) stay in
// registers; with run-time trip counts they were indexed dynamically and lived in local memory (816 bytes of stack).

// Model plug-in (pgas_b200.h: pgas_marg_program): a user callable outside the coefficient-table families,
// traced on the host into a postfix program (StateSpaceModel.programs, models.py: Sym) and interpreted here — by every lane with
// the same operands, like the table forms.  v = [state; interface variables], u = inputs[t].  Out of line and only reachable from
// the PROG instantiations of the sweep kernel: the table-driven instantiations do not change by a single instruction.
static __device__ __noinline__ void mg_run_program(const int* __restrict__ ops, const double* __restrict__ consts, int len, int n_out,
                                                   const double* v, const double* u, double* out) {
    double st[PGAS_PROG_STACK];
    int sp = 0;
    for (int pc = 0; pc < len; ++pc) {
        const int ins = __ldg(ops + pc), op = ins & 0xff, arg = ins >> 8;
        if (op <= PGAS_OP_PUSH_C) {
            st[sp & (PGAS_PROG_STACK - 1)] = (op == PGAS_OP_PUSH_X) ? v[arg] : (op == PGAS_OP_PUSH_U) ? u[arg] : __ldg(consts + arg);
            ++sp;
        } else if (op <= PGAS_OP_DIV || op >= PGAS_OP_POW) {
            const double b = st[(--sp) & (PGAS_PROG_STACK - 1)], a = st[(sp - 1) & (PGAS_PROG_STACK - 1)];
            double r;
            switch (op) {
                case PGAS_OP_ADD: r = a + b; break;
                case PGAS_OP_SUB: r = a - b; break;
                case PGAS_OP_MUL: r = a * b; break;
                case PGAS_OP_DIV: r = a / b; break;
                case PGAS_OP_POW: r = pow(a, b); break;
                default: r = atan2(a, b); break;
            }
            st[(sp - 1) & (PGAS_PROG_STACK - 1)] = r;
        } else {
            const double a = st[(sp - 1) & (PGAS_PROG_STACK - 1)];
            double r;
            switch (op) {
                case PGAS_OP_NEG: r = -a; break;
                case PGAS_OP_SIN: r = sin(a); break;
                case PGAS_OP_COS: r = cos(a); break;
                case PGAS_OP_TAN: r = tan(a); break;
                case PGAS_OP_TANH: r = tanh(a); break;
                case PGAS_OP_ATAN: r = atan(a); break;
                case PGAS_OP_EXP: r = exp(a); break;
                case PGAS_OP_LOG: r = log(a); break;
                case PGAS_OP_SQRT: r = sqrt(a); break;
                default: r = fabs(a); break;
            }
            st[(sp - 1) & (PGAS_PROG_STACK - 1)] = r;
        }
    }
    for (int d = 0; d < n_out; ++d) out[d] = st[d];
}

// operands of a transition / output program: [state; interface variables]
__device__ __forceinline__ void mg_pack_operands(const MargDev& m, const double* x, const double* xi, double* v) {
#pragma unroll
    for (int k = 0; k < MG_NX + MG_GP; ++k) v[k] = 0.0;
#pragma unroll
    for (int k = 0; k < MG_NX; ++k) if (k < m.n_x) v[k] = x[k];
    if (xi) {
#pragma unroll
        for (int g = 0; g < MG_GP; ++g) if (g < m.G) v[m.n_x + g] = xi[g];
    }
}

// GP-input map z = p * link(a . x + b) + q   (pgas_b200.h, group B)
template <bool PROG>
__device__ __forceinline__ void gp_input(const MargDev& m, const MargGP& gp, int t, const double* x, double* z) {
    const int nx = m.n_x;
    if constexpr (PROG) {
        if (gp.prog.len) {
            double v[MG_NX + MG_GP], o[PGAS_PROG_STACK];
            mg_pack_operands(m, x, nullptr, v);
            mg_run_program(gp.prog.ops, gp.prog.consts, gp.prog.len, gp.D, v, m.inputs + (size_t)t * m.n_u, o);
#pragma unroll
            for (int d = 0; d < MG_D; ++d) if (d < gp.D) z[d] = o[d];
            return;
        }
    }
#pragma unroll
    for (int d = 0; d < MG_D; ++d) {
        if (d < gp.D) {
            const double* c = gp.gp_in + ((size_t)t * gp.D + d) * (nx + 1);
            const double* p = gp.gp_post + ((size_t)t * gp.D + d) * 2;
            const double p0 = p[0], p1 = p[1];
            double s = c[nx];
#pragma unroll
            for (int k = 0; k < MG_NX; ++k) if (k < nx) s = fma(c[k], x[k], s);
            if (gp.link == PGAS_LINK_ATAN) s = atan(s);
            z[d] = fma(p0, s, p1);
        }
    }
}

// sin(x), straight-line: Cody-Waite reduction r = x - n pi/2 with a three-part pi/2 (the first product is exact, so the reduced
// argument carries half an ulp of ITS OWN size for |x| up to ~1e5), then the sine / cosine polynomials of fastmath.cuh on r / pi
// (|r / pi| <= 1/4: that division's rounding is relative to r, not to x).  libm's sin() has the same fast path behind a slow-path
// branch for huge arguments; the shipped vehicle configuration amplifies basis errors enough that a reduction of x / pi in one
// rounding (error |x| 2^-53) fails the parity bar, so the reduction is done properly.
__device__ __forceinline__ double sin_bf(double x) {
    const double n = rint(x * 0.6366197723675814);
    double r = fma(-n, 1.5707963267948966, x);
    r = fma(-n, 6.123233995736766e-17, r);
    r = fma(-n, -1.4973849048591698e-33, r);
    const double rp = r * 0.3183098861837907;
    const double t = rp * rp;
    double ps = FM_SIN[0], pc = FM_COS[0];
#pragma unroll
    for (int i = 1; i < 10; ++i) {
        ps = fma(ps, t, FM_SIN[i]);
        pc = fma(pc, t, FM_COS[i]);
    }
    const double sn = rp * ps, cs = fma(pc, t, 1.0);
    const int q = (int)__double2ll_rn(n) & 3;
    const double v = (q & 1) ? cs : sn;
    return (q & 2) ? -v : v;
}

// _eigen_fnc (src/BasisFunctions.py:77-80): phi_m = prod_d sqrt(1/L_d) sin(sqrt(eig_md) (z_d - c_d + L_d)),
// lanes stride over m
__device__ __forceinline__ void basis_eval(const MargGP& gp, const double* z, double* out, int lane) {
    double zz[MG_D];
#pragma unroll
    for (int d = 0; d < MG_D; ++d) zz[d] = (d < gp.D) ? (z[d] - gp.center[d]) + gp.L[d] : 0.0;
    for (int mI = lane; mI < gp.M; mI += 32) {
        double p = 1.0;
#pragma unroll
        for (int d = 0; d < MG_D; ++d)
            if (d < gp.D) p *= gp.sqrt_invL[d] * sin_bf(gp.sqrt_eig[(size_t)mI * gp.D + d] * zz[d]);
        out[mI] = p;
    }
}

template <bool PROG>
__device__ __forceinline__ void transition(const MargDev& m, int t, const double* x, const double* xi, double* xn) {
    const int nx = m.n_x, G = m.G, W = nx + G + 1;
    if constexpr (PROG) {
        if (m.tprog.len) {
            double v[MG_NX + MG_GP], o[PGAS_PROG_STACK];
            mg_pack_operands(m, x, xi, v);
            mg_run_program(m.tprog.ops, m.tprog.consts, m.tprog.len, nx, v, m.inputs + (size_t)t * m.n_u, o);
#pragma unroll
            for (int r = 0; r < MG_NX; ++r) if (r < nx) xn[r] = o[r];
            return;
        }
    }
#pragma unroll
    for (int r = 0; r < MG_NX; ++r) {
        if (r < nx) {
            const double* c = m.trans + ((size_t)t * nx + r) * W;
            double s = c[nx + G];
#pragma unroll
            for (int g = 0; g < MG_GP; ++g) if (g < G) s = fma(c[nx + g], xi[g], s);
#pragma unroll
            for (int k = 0; k < MG_NX; ++k) if (k < nx) s = fma(c[k], x[k], s);
            xn[r] = s;
        }
    }
}

template <bool PROG>
__device__ __forceinline__ void output_mdl(const MargDev& m, int t, const double* x, const double* xi, double* y) {
    const int nx = m.n_x, G = m.G, W = nx + G + 1;
    if constexpr (PROG) {
        if (m.oprog.len) {
            double v[MG_NX + MG_GP], o[PGAS_PROG_STACK];
            mg_pack_operands(m, x, xi, v);
            mg_run_program(m.oprog.ops, m.oprog.consts, m.oprog.len, m.n_y, v, m.inputs + (size_t)t * m.n_u, o);
#pragma unroll
            for (int r = 0; r < MG_NY; ++r) if (r < m.n_y) y[r] = o[r];
            return;
        }
    }
#pragma unroll
    for (int r = 0; r < MG_NY; ++r) {
        if (r < m.n_y) {
            const double* c = m.outp + ((size_t)t * m.n_y + r) * W;
            double s = c[nx + G];
#pragma unroll
            for (int g = 0; g < MG_GP; ++g) if (g < G) s = fma(c[nx + g], xi[g], s);
#pragma unroll
            for (int k = 0; k < MG_NX; ++k) if (k < nx) s = fma(c[k], x[k], s);
            y[r] = (m.out_link == PGAS_LINK_TANH) ? tanh(s) : s;
        }
    }
}

// StateSpaceModel.log_likelihood (src/StateSpaceModel.py:75-87)
template <bool PROG>
__device__ __forceinline__ double log_likelihood(const MargDev& m, int t, const double* x, const double* xi) {
    double y[MG_NY], e[MG_NY];
    output_mdl<PROG>(m, t, x, xi, y);
#pragma unroll
    for (int r = 0; r < MG_NY; ++r) e[r] = (r < m.n_y) ? m.obs[(size_t)t * m.n_y + r] - y[r] : 0.0;
    double q = 0.0;
#pragma unroll
    for (int r = 0; r < MG_NY; ++r) {
        if (r < m.n_y) {
            double w = 0.0;
#pragma unroll
            for (int k = 0; k <= r; ++k) w = fma(m.Rw[r][k], e[k], w);
            q = fma(w, w, q);
        }
    }
    return -0.5 * q + m.R_logc;
}

// log N(target; mean, Q)  (h_x, src/Algorithm3.py:107-114)
__device__ __forceinline__ double log_trans_density(const MargDev& m, const double* target, const double* mean) {
    double e[MG_NX], q = 0.0;
#pragma unroll
    for (int r = 0; r < MG_NX; ++r) e[r] = (r < m.n_x) ? target[r] - mean[r] : 0.0;
#pragma unroll
    for (int r = 0; r < MG_NX; ++r) {
        if (r < m.n_x) {
            double w = 0.0;
#pragma unroll
            for (int k = 0; k <= r; ++k) w = fma(m.Qw[r][k], e[k], w);
            q = fma(w, w, q);
        }
    }
    return -0.5 * q + m.Q_logc;
}

// prior_mniw_log_base_measure (src/BayesianInferrence.py:111-124) for n = 1 from the factorisation:
// logdet = log det T1, psi = T2 - T0^T T1^-1 T0
__device__ __forceinline__ double log_base_measure(int M, double logdet, double psi, double nu) {
    const double t1 = -0.5 * (double)M * 1.8378770664093453;      // log(2 pi)
    const double t2 = 0.5 * logdet;
    const double t3 = -0.5 * nu * 0.6931471805599453;
    const double t4 = -lgamma(0.5 * nu);                          // multigammaln(nu/2, 1)
    const double t5 = log(psi) * nu * 0.5;
    return t1 + t2 + t3 + t4 + t5;
}

// Student-t variate of jax.random.t: z sqrt(a / g), a = df/2, g ~ Gamma(a) by Marsaglia-Tsang with
// Philox-indexed attempts.  Counter word 1 = t | slot << 20 (slot 0: z, 1: boost, 2+2k / 3+2k: attempt k).
__device__ double philox_student_t(unsigned long long seed, unsigned purpose, unsigned chain, unsigned iter, unsigned t, unsigned i,
                                   double df) {
    const double tiny = 1.0 / 9007199254740992.0;
    const double a0 = 0.5 * df;
    double a = a0, boost = 1.0, z, zb;
    philox_normal2(seed, purpose, chain, iter, t, i, z, zb);
    if (a < 1.0) {
        double u, ub;
        philox_uniform2(seed, purpose, chain, iter, t | (1u << 20), i, u, ub);
        boost = pow(u + tiny, 1.0 / a);
        a += 1.0;
    }
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    double g = d;
    for (unsigned attempt = 0;; ++attempt) {
        double x, xb, u, ub;
        philox_normal2(seed, purpose, chain, iter, t | ((2u + 2u * attempt) << 20), i, x, xb);
        philox_uniform2(seed, purpose, chain, iter, t | ((3u + 2u * attempt) << 20), i, u, ub);
        double v = 1.0 + c * x;
        if (v <= 0.0 && attempt <= 1000u) continue;
        v = v * v * v;
        if (log(u + tiny) < 0.5 * x * x + d - d * v + d * log(v) || attempt > 1000u) { g = d * v * boost; break; }
    }
    return z * sqrt(a0 / g);
}

// The Philox blocks one particle needs in a step, evaluated in parallel across lanes (the serial versions
// above define the stream; this only changes who computes which block):
//   lanes 3g, 3g+1, 3g+2 : GP g — Student-t normal (slot 0), first Marsaglia-Tsang attempt normal / uniform (slots 2, 3)
//   lanes 3G + p         : state normals of pair p
struct StepVariates {
    double zs[MG_NX];
    double tz[MG_GP], tx[MG_GP], tu[MG_GP];
};
__device__ __forceinline__ StepVariates philox_step_variates(unsigned long long seed, unsigned chain, unsigned iter, unsigned t, unsigned i,
                                                             int G, int nx, int lane) {
    const int ng = 3 * G, npair = (nx + 1) >> 1;
    const int g = lane / 3, slot = lane - 3 * g;
    const bool is_t = lane < ng;
    const unsigned purpose = is_t ? (unsigned)(PURPOSE_TVAR + g) : (unsigned)PURPOSE_STATE;
    const unsigned code = slot == 0 ? 0u : (slot == 1 ? 2u : 3u);
    const unsigned c1 = is_t ? (t | (code << 20)) : (t | ((unsigned)(lane - ng) << 28));
    uint32_t o[4];
    philox4x32_10(i, c1, iter, (purpose << 24) | (chain & 0xFFFFFFu), (uint32_t)seed, (uint32_t)(seed >> 32), o);
    const double ua = u53(o[0], o[1]), ub = u53(o[2], o[3]);
    const double r = sqrt_bf(fmax(-2.0 * log_unit_bf(ua + (1.0 / 9007199254740992.0)), 1e-30));
    double sn, cs;
    sincospi_bf(2.0 * ub, sn, cs);
    const double za = r * cs, zb = r * sn;
    StepVariates v;
    for (int gg = 0; gg < MG_GP; ++gg) {
        v.tz[gg] = __shfl_sync(FULL, za, min(3 * gg, 31));
        v.tx[gg] = __shfl_sync(FULL, za, min(3 * gg + 1, 31));
        v.tu[gg] = __shfl_sync(FULL, ua, min(3 * gg + 2, 31));
    }
    for (int p = 0; p < (MG_NX >> 1); ++p) {
        v.zs[2 * p] = __shfl_sync(FULL, za, min(ng + p, 31));
        v.zs[2 * p + 1] = __shfl_sync(FULL, zb, min(ng + p, 31));
    }
    (void)npair;
    return v;
}

// Student-t variate from pre-computed first-attempt pieces; falls back to the serial stream when the first
// attempt is rejected or the shape is below one (identical values by construction).
__device__ __forceinline__ double student_t_from(const StepVariates& v, int g, unsigned long long seed, unsigned purpose, unsigned chain,
                                                 unsigned iter, unsigned t, unsigned i, double df) {
    const double a = 0.5 * df;
    if (a >= 1.0) {
        const double tiny = 1.0 / 9007199254740992.0;
        const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
        const double x = v.tx[g];
        double w = 1.0 + c * x;
        if (w > 0.0) {
            w = w * w * w;
            if (log(v.tu[g] + tiny) < 0.5 * x * x + d - d * w + d * log(w)) return v.tz[g] * sqrt(a / (d * w));
        }
    }
    return philox_student_t(seed, purpose, chain, iter, t, i, df);
}

// ---------------------------------------------------------------------------------- warp linear algebra
// In-place left-looking Cholesky of a packed (row-major, lower) matrix in shared memory by one warp.
// The leading M x M block is factored; rows M .. R-1 ride along as right-hand sides (row M+e becomes
// (L^-1 b_e)^T in its first M entries; entries beyond column M-1 are left untouched).  Diagonal slots
// receive 1 / L_jj.  Returns log det = sum_j log(pivot_j) (identical on all lanes).
//
// Columns are processed in panels of four: a lane owns row J + lane (+32 per pass) and accumulates the
// four dot products of the panel at once — one load of its own row entry and four broadcast loads feed
// four independent DFMAs — then the 4 x 4 diagonal block is resolved with shuffles from the lanes that
// own the panel's rows.  Pivots are parked in registers (lane j % 32) so that the logarithms of the
// determinant run once, in parallel, after the sweep.
__device__ double warp_chol_packed(double* A, int M, int R, int lane, int& fail) {
    double pv0 = 1.0, pv1 = 1.0, pv2 = 1.0, pv3 = 1.0;        // pivot of column lane + 32 q
#define MG_SAVE_PIVOT(col, d) do { if (lane == ((col) & 31)) { const int q_ = (col) >> 5; \
    pv0 = q_ == 0 ? (d) : pv0; pv1 = q_ == 1 ? (d) : pv1; pv2 = q_ == 2 ? (d) : pv2; pv3 = q_ == 3 ? (d) : pv3; } } while (0)
    for (int J = 0; J < M; J += 4) {
        const int nc = min(4, M - J);
        const double* r0 = A + tri(J);
        const double* r1 = A + tri(J + (nc > 1 ? 1 : 0));
        const double* r2 = A + tri(J + (nc > 2 ? 2 : 0));
        const double* r3 = A + tri(J + (nc > 3 ? 3 : 0));
        double i0v = 0.0, i1v = 0.0, i2v = 0.0, i3v = 0.0;             // 1 / L_jj of the panel columns
        double b10 = 0.0, b20 = 0.0, b21 = 0.0, b30 = 0.0, b31 = 0.0, b32 = 0.0;   // L[J+c][J+c'] of the diagonal block
        for (int ib = J; ib < R; ib += 32) {
            const int i = ib + lane;
            const bool act = i < R;
            const double* ri = A + tri(act ? i : R - 1);
            double s0 = ri[J], s1 = ri[J + (nc > 1 ? 1 : 0)], s2 = ri[J + (nc > 2 ? 2 : 0)], s3 = ri[J + (nc > 3 ? 3 : 0)];
            int k = 0;
            for (; k + 1 < J; k += 2) {
                const double a0 = ri[k], a1 = ri[k + 1];
                s0 = fma(-a0, r0[k], s0); s1 = fma(-a0, r1[k], s1); s2 = fma(-a0, r2[k], s2); s3 = fma(-a0, r3[k], s3);
                s0 = fma(-a1, r0[k + 1], s0); s1 = fma(-a1, r1[k + 1], s1); s2 = fma(-a1, r2[k + 1], s2); s3 = fma(-a1, r3[k + 1], s3);
            }
            if (k < J) {
                const double a0 = ri[k];
                s0 = fma(-a0, r0[k], s0); s1 = fma(-a0, r1[k], s1); s2 = fma(-a0, r2[k], s2); s3 = fma(-a0, r3[k], s3);
            }
            const bool first = ib == J;
            // column J
            if (first) {
                const double d = __shfl_sync(FULL, s0, 0);
                if (!(d > 0.0)) fail = 1;
                i0v = rsqrt(d);
                MG_SAVE_PIVOT(J, d);
            }
            const double l0 = s0 * i0v;
            // column J + 1
            if (first) b10 = __shfl_sync(FULL, l0, 1);
            s1 = fma(-l0, b10, s1);
            if (first && nc > 1) {
                const double d = __shfl_sync(FULL, s1, 1);
                if (!(d > 0.0)) fail = 1;
                i1v = rsqrt(d);
                MG_SAVE_PIVOT(J + 1, d);
            }
            const double l1 = s1 * i1v;
            // column J + 2
            if (first) { b20 = __shfl_sync(FULL, l0, 2); b21 = __shfl_sync(FULL, l1, 2); }
            s2 = fma(-l1, b21, fma(-l0, b20, s2));
            if (first && nc > 2) {
                const double d = __shfl_sync(FULL, s2, 2);
                if (!(d > 0.0)) fail = 1;
                i2v = rsqrt(d);
                MG_SAVE_PIVOT(J + 2, d);
            }
            const double l2 = s2 * i2v;
            // column J + 3
            if (first) { b30 = __shfl_sync(FULL, l0, 3); b31 = __shfl_sync(FULL, l1, 3); b32 = __shfl_sync(FULL, l2, 3); }
            s3 = fma(-l2, b32, fma(-l1, b31, fma(-l0, b30, s3)));
            if (first && nc > 3) {
                const double d = __shfl_sync(FULL, s3, 3);
                if (!(d > 0.0)) fail = 1;
                i3v = rsqrt(d);
                MG_SAVE_PIVOT(J + 3, d);
            }
            const double l3 = s3 * i3v;
            if (act) {
                double* wi = A + tri(i) + J;
                if (i >= J) wi[0] = (i == J) ? i0v : l0;
                if (nc > 1 && i >= J + 1) wi[1] = (i == J + 1) ? i1v : l1;
                if (nc > 2 && i >= J + 2) wi[2] = (i == J + 2) ? i2v : l2;
                if (nc > 3 && i >= J + 3) wi[3] = (i == J + 3) ? i3v : l3;
            }
        }
        __syncwarp();
    }
#undef MG_SAVE_PIVOT
    double logdet = 0.0;
    if (lane < M) logdet += log(pv0);
    if (lane + 32 < M) logdet += log(pv1);
    if (lane + 64 < M) logdet += log(pv2);
    if (lane + 96 < M) logdet += log(pv3);
    return warp_sum(logdet);
}

// w = L^-1 b by columns: lane owns rows lane, lane + 32, ...; L packed with inverse diagonal.  The column loop is cut at the row
// segments (see warp_fused_update).
template <int ROWS>
__device__ __forceinline__ void warp_fwd_solve(const double* Lpk, const double* b, int M, int lane, double (&w)[ROWS]) {
#pragma unroll
    for (int r = 0; r < ROWS; ++r) { const int idx = lane + 32 * r; w[r] = idx < M ? b[idx] : 0.0; }
#pragma unroll
    for (int seg = 0; seg < ROWS; ++seg) {
        const int kend = min(M, 32 * (seg + 1));
        for (int k = 32 * seg; k < kend; ++k) {
            const int src = k - 32 * seg;
            const double wk = __shfl_sync(FULL, w[seg], src) * Lpk[tri(k) + k];
            if (lane == src) w[seg] = wk;
#pragma unroll
            for (int r = seg; r < ROWS; ++r) {
                const int idx = lane + 32 * r;
                if ((r > seg || idx > k) && idx < M) w[r] = fma(-Lpk[tri(idx) + k], wk, w[r]);
            }
        }
    }
}

// 1 / sqrt(a), a > 0: the bare MUFU.RSQ64H seed (2^-22) and two Newton steps in residual form — straight-line, ~2 ulp.  CUDA's
// rsqrt() wraps the same seed in special-case handling (17 instructions and a branch per call; three calls per column made it
// 12 % of the kernel's instruction stream, profiles/r02_marg_sweep_summary.md).  Non-positive arguments give NaN / inf, which the
// callers detect on the argument itself.
__device__ __forceinline__ double rsqrt_bf(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double h = 0.5 * a;
    y = fma(y, fma(-(h * y), y, 0.5), y);
    y = fma(y, fma(-(h * y), y, 0.5), y);
    return y;
}

// One sweep over the columns that does, in lockstep (three independent rotation chains and a substitution per column):
//   A <- chol(A A^T + z z^T)                       (Givens)          statistics gain [phi; xi][phi; xi]^T
//   B <- chol(B B^T + z z^T - r r^T)               (Givens, hyperbolic)  ... and the reference part loses [phi_ref; xi_ref](.)^T
//   w <- (new A, leading M x M block)^-1 rhs       (column-oriented forward substitution; 1 / L'_kk is the rsqrt of the rotation)
// A, B: packed augmented factors with true diagonal, n = M + 1 rows; z, r: n entries; rhs: M entries.
// The column loop is cut at the lanes' row segments (rows lane + 32 q), so that the register holding column k's entry and the set of
// rows still below the diagonal are compile-time facts: no selects, and rows of finished segments are not visited.  The squared
// diagonal of B after the Givens step IS rho_b, so the hyperbolic pivot rho_b - c_k^2 does not wait for the first square root: the three
// inverse square roots of a column are independent.
template <int ROWS>
__device__ __forceinline__ void warp_fused_update(double* A, double* B, const double* z, const double* r, const double* rhs, int M,
                                                  int lane, int& bad, double (&w)[ROWS]) {
    const int n = M + 1;
    double xa[ROWS], xb[ROWS], xc[ROWS];
#pragma unroll
    for (int q = 0; q < ROWS; ++q) {
        const int idx = lane + 32 * q;
        xa[q] = xb[q] = idx < n ? z[idx] : 0.0;
        xc[q] = idx < n ? r[idx] : 0.0;
        w[q] = idx < M ? rhs[idx] : 0.0;
    }
#pragma unroll
    for (int seg = 0; seg < ROWS; ++seg) {
        const int kend = min(n, 32 * (seg + 1));
        for (int k = 32 * seg; k < kend; ++k) {
            const int src = k - 32 * seg, dk = tri(k) + k;
            // the column's entries below the diagonal are read first: they do not depend on the rotation, and the __syncwarp that
            // guards the diagonal would otherwise pin the loads behind the whole coefficient chain
            double Aik[ROWS], Bik[ROWS];
#pragma unroll
            for (int q = seg; q < ROWS; ++q) {
                const int idx = lane + 32 * q;
                const bool act = (q > seg || idx > k) && idx < n;
                Aik[q] = act ? A[tri(idx) + k] : 0.0;
                Bik[q] = act ? B[tri(idx) + k] : 0.0;
            }
            const double Akk = A[dk], Bkk = B[dk];
            const double ak = __shfl_sync(FULL, xa[seg], src), bk = __shfl_sync(FULL, xb[seg], src), ck = __shfl_sync(FULL, xc[seg], src);
            const double wk0 = __shfl_sync(FULL, w[seg], src);
            const double rhoa = fma(ak, ak, Akk * Akk), rhob = fma(bk, bk, Bkk * Bkk);
            const double rhoc = fma(-ck, ck, rhob);
            if (!(rhoc > 0.0) || !(rhoa > 0.0)) bad = 1;
            const double ria = rsqrt_bf(rhoa), rib = rsqrt_bf(rhob), ric = rsqrt_bf(rhoc);
            const double Bk1 = rhob * rib;                                        // diagonal of B after the update
            const double aa = Akk * ria, ba = ak * ria, ab = Bkk * rib, bb = bk * rib, ac = Bk1 * ric, bc = ck * ric;
            const double wk = wk0 * ria;                                          // forward substitution on the NEW factor
            __syncwarp();
            if (lane == 0) { A[dk] = rhoa * ria; B[dk] = rhoc * ric; }
            if (lane == src && k < M) w[seg] = wk;
#pragma unroll
            for (int q = seg; q < ROWS; ++q) {
                const int idx = lane + 32 * q;
                if ((q > seg || idx > k) && idx < n) {
                    const double An = fma(ba, xa[q], aa * Aik[q]);
                    xa[q] = fma(aa, xa[q], -ba * Aik[q]);
                    const double B1 = fma(bb, xb[q], ab * Bik[q]);
                    xb[q] = fma(ab, xb[q], -bb * Bik[q]);
                    const double B2 = fma(-bc, xc[q], ac * B1);
                    xc[q] = fma(ac, xc[q], -bc * B1);
                    A[tri(idx) + k] = An;
                    B[tri(idx) + k] = B2;
                    if (k < M && idx < M) w[q] = fma(-An, wk, w[q]);
                }
            }
        }
    }
    __syncwarp();
}

// w = L^-1 b with the inverse diagonal given separately (inv[k] = 1 / L_kk); column loop cut at the row segments as above
template <int ROWS>
__device__ __forceinline__ void warp_fwd_solve_inv(const double* Lpk, const double* inv, const double* b, int M, int lane, double (&w)[ROWS]) {
#pragma unroll
    for (int r = 0; r < ROWS; ++r) { const int idx = lane + 32 * r; w[r] = idx < M ? b[idx] : 0.0; }
#pragma unroll
    for (int seg = 0; seg < ROWS; ++seg) {
        const int kend = min(M, 32 * (seg + 1));
        for (int k = 32 * seg; k < kend; ++k) {
            const int src = k - 32 * seg;
            const double wk = __shfl_sync(FULL, w[seg], src) * inv[k];
            if (lane == src) w[seg] = wk;
#pragma unroll
            for (int r = seg; r < ROWS; ++r) {
                const int idx = lane + 32 * r;
                if ((r > seg || idx > k) && idx < M) w[r] = fma(-Lpk[tri(idx) + k], wk, w[r]);
            }
        }
    }
}

// after warp_chol_packed on an augmented matrix with ONE extra row (row M = eta0^T, slot [M][M] = eta2): turn it into the
// true-diagonal augmented factor (diagonal L_kk, last diagonal sqrt(Psi)); returns log det eta1 and Psi
__device__ __forceinline__ void finish_aug_factor(double* A, int M, int lane, double& psi) {
    for (int k = lane; k < M; k += 32) A[tri(k) + k] = 1.0 / A[tri(k) + k];
    double yy = 0.0;
    for (int k = lane; k < M; k += 32) { const double y = A[tri(M) + k]; yy = fma(y, y, yy); }
    yy = warp_sum(yy);
    psi = A[tri(M) + M] - yy;
    __syncwarp();
    if (lane == 0) A[tri(M) + M] = sqrt(psi);
    __syncwarp();
}

// log det of the leading M x M block of a true-diagonal factor: 2 sum log L_kk
__device__ __forceinline__ double factor_logdet(const double* A, int M, int lane) {
    double s = 0.0;
    for (int k = lane; k < M; k += 32) s += log(A[tri(k) + k]);
    return 2.0 * warp_sum(s);
}

// ---------------------------------------------------------------------------------- CTA-wide softmax / CDF
// cdf[0..N) <- cumulative sums of softmax(lw) (src/Algorithm3.py:119-121), or — sisr — the table
// systematic_SISR searches: clip, renormalise (uniform when the sum is not > 0), cumsum, clip to [0,1]
// (src/Filtering.py:23-32).  Plain mode (cumulative = false) leaves the softmax weights themselves.
// Every CTA of a cluster runs this on the same data with the same thread count -> identical bits.
__device__ void cta_softmax_cdf(const double* __restrict__ lw, int N, bool sisr, bool cumulative, double* cdf, double* red, int tid,
                                int nthr) {
    const int lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
    double mx = -INFINITY;
    int isnan_ = 0;
    for (int i = tid; i < N; i += nthr) {
        const double v = ldcg(lw + i);
        cdf[i] = v;
        isnan_ |= (v != v);
        mx = fmax(mx, v);
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    isnan_ = __syncthreads_or(isnan_);
    mx = red[0];
    for (int w = 1; w < nw; ++w) mx = fmax(mx, red[w]);
    if (isnan_) mx = NAN;                                   // jnp.max propagates NaN -> all weights NaN
    __syncthreads();
    double sm = 0.0;
    for (int i = tid; i < N; i += nthr) {
        const double e = exp(cdf[i] - mx);
        cdf[i] = e;
        sm += e;
    }
    sm = warp_sum(sm);
    if (lane == 0) red[warp] = sm;
    __syncthreads();
    if (warp == 0) {
        double tot = 0.0;
        for (int w = 0; w < nw; ++w) tot += red[w];
        bool ok = true;
        double s2 = 1.0;
        if (sisr) {
            double s = 0.0;
            for (int i = lane; i < N; i += 32) s += fmax(__ddiv_rn(cdf[i], tot), 0.0);
            s2 = warp_sum(s);
            ok = s2 > 0.0;
            if (tot != tot) ok = false;
        }
        double carry = 0.0;
        for (int i0 = 0; i0 < N; i0 += 32) {
            const int i = i0 + lane;
            double w = 0.0;
            if (i < N) {
                w = __ddiv_rn(cdf[i], tot);
                if (sisr) w = ok ? __ddiv_rn(fmax(w, 0.0), s2) : __ddiv_rn(1.0, (double)N);
            }
            if (cumulative) {
                const double s = carry + warp_scan_incl(w, lane);
                carry = __shfl_sync(FULL, s, 31);
                if (i < N) cdf[i] = sisr ? fmin(fmax(s, 0.0), 1.0) : s;
            } else if (i < N) {
                cdf[i] = w;
            }
        }
    }
    __syncthreads();
}

// The same table built by ONE WARP into its own buffer (every warp of the chain runs this on the same global log-weights with the
// same code: identical bits everywhere, no CTA barrier, no idle warps).  Lane l owns the c = ceil(N / 32) consecutive entries
// [l c, (l + 1) c): sums and the cumulative sum are thread-serial over the lane's entries plus one butterfly / one scan.  Quotients by
// the two normalisers are Markstein-corrected products with the reciprocal.
__device__ __forceinline__ double div_rcp(double x, double d, double rd) {
    const double q = x * rd;
    return fma(fma(-q, d, x), rd, q);
}
// own entries of a lane, eight at a time with the eight bodies unrolled side by side (independent chains interleave)
#define MG_FOR_OWN(...)                                                        \
    for (int j0_ = 0; j0_ < c; j0_ += 8) {                                     \
        _Pragma("unroll") for (int j_ = 0; j_ < 8; ++j_) {                     \
            const int i = lo + j0_ + j_;                                       \
            if (j0_ + j_ < c && i < N) { __VA_ARGS__ }                         \
        }                                                                      \
    }
__device__ void warp_softmax_cdf(const double* __restrict__ lw, int N, bool sisr, double* cdf, int lane) {
    // global -> shared, eight loads in flight per lane before the first store (a load-store loop would pay one L2 round trip
    // per entry)
    for (int base = 0; base < N; base += 256) {
        double v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int i = base + lane + 32 * j; v[j] = i < N ? ldcg(lw + i) : 0.0; }
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int i = base + lane + 32 * j; if (i < N) cdf[i] = v[j]; }
    }
    __syncwarp();
    const int c = (N + 31) >> 5, lo = lane * c;
    double mx = -INFINITY;
    int isnan_ = 0;
    MG_FOR_OWN(const double v = cdf[i]; isnan_ |= (v != v); mx = fmax(mx, v);)
    mx = warp_max(mx);
    if (__any_sync(FULL, isnan_)) mx = NAN;                 // jnp.max propagates NaN -> all weights NaN
    double sm = 0.0;
    MG_FOR_OWN(const double e = exp_neg_bf(cdf[i] - mx); cdf[i] = e; sm += e;)
    const double tot = warp_sum(sm), rtot = 1.0 / tot;
    bool ok = true;
    double s2 = 1.0, rs2 = 1.0;
    if (sisr) {
        double s = 0.0;
        MG_FOR_OWN(s += fmax(div_rcp(cdf[i], tot, rtot), 0.0);)
        s2 = warp_sum(s);
        rs2 = 1.0 / s2;
        ok = (s2 > 0.0) && (tot == tot);
    }
    const double unif = 1.0 / (double)N;
    double run = 0.0;
    MG_FOR_OWN(double w = div_rcp(cdf[i], tot, rtot); if (sisr) w = ok ? div_rcp(fmax(w, 0.0), s2, rs2) : unif; run += w; cdf[i] = run;)
    const double off = warp_scan_incl(run, lane) - run;
    MG_FOR_OWN(const double v = off + cdf[i]; cdf[i] = sisr ? fmin(fmax(v, 0.0), 1.0) : v;)
    __syncwarp();
}
#undef MG_FOR_OWN

// #{j : cdf[j] < u} counted by the warp (the table is non-decreasing: this is searchsorted side = left)
__device__ __forceinline__ int warp_count_below(const double* cdf, int N, double u, int lane) {
    int n = 0;
    for (int i = lane; i < N; i += 32) n += (cdf[i] < u) ? 1 : 0;
    return __reduce_add_sync(FULL, n);
}

// #{j : cdf[j] < u}  (searchsorted side = left)
__device__ __forceinline__ int count_below(const double* cdf, int N, double u) {
    int lo = 0, hi = N;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] < u) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---------------------------------------------------------------------------------- the persistent kernel
constexpr int MG_REFRESH = 32;     // Algorithm3: steps between full re-factorisations of the rank-1-updated factors

struct WarpCtx {
    double* A[MG_GP];      // augmented packed matrices (M + 2 rows)
    double* B;             // staging: ancestor's factor / second factorisation (M + 1 rows)
    double* phi;           // basis vector of the new state
    double* Bg[MG_GP];     // Algorithm3: the ancestor's second augmented factor, prefetched per GP
    double* T1s[MG_GP];    // Algorithm3: the ancestor's packed T1, prefetched per GP
    double* inv;           // inverse diagonal of a factor (Algorithm3)
    double* zv;            // rank-1 vectors [phi; xi] and [phi_ref; xi_ref] (Algorithm3)
    double* rv;
    double* cdf;           // the warp's copy of the resampling table (N entries)
};

// WIDE: the launch has at most 128 threads per CTA and two CTAs per SM (the wide geometry of mg_geometry), so a thread may use 255
// registers instead of 128.
// PROG: the model carries at least one expression program (model plug-in); only this instantiation contains the interpreter.
template <int MODE, int ROWS, bool WIDE, bool PROG = false>
__global__ void __launch_bounds__(WIDE ? 128 : 512, WIDE ? 2 : 1) marg_sweep_kernel(const __grid_constant__ MargArgs a) {
    extern __shared__ double smem[];
    const MargDev& m = a.m;
    // the thread index goes through a shuffle once: ptxas otherwise re-reads the special register wherever `lane` is used (100
    // S2R per particle-step, 4 % of the stall samples of the round-1 kernel)
    const int tid = __shfl_sync(FULL, (int)threadIdx.x, (int)(threadIdx.x & 31u));
    const int lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
    const int CS = a.CS, NW = a.NW, N = a.N, T = m.T, G = m.G, nx = m.n_x;
    const int chain = blockIdx.x / CS, rank = blockIdx.x % CS;
    const int WT = CS * NW, wg = rank * NW + warp;
    const bool multi = CS > 1 && !a.sw_barrier;
    const double lam = a.lambda;

    // ---- shared memory carve-up: CTA part, then per-warp parts
    double* wsm = smem + ((N + 3) & ~3);                  // [N]   (mode 0 statistics trace)
    double* red = wsm + ((N + 3) & ~3);                   // [64]
    // All loops of the carve-up run over the compile-time maximum with a guard: the buffer pointers are then plain values derived
    // from `smem`, the compiler keeps their address space, and every access below is an LDS / STS.  (Filled in a run-time loop the
    // pointer arrays lived in local memory and every access to the factors was a GENERIC load / store.)
    unsigned* ijt[MG_GP];
    {
        unsigned* q = (unsigned*)(red + 64);
#pragma unroll
        for (int g = 0; g < MG_GP; ++g) { ijt[g] = q; q += (g < G) ? ((m.gp[g].npk + 1) & ~1) : 0; }
    }
    WarpCtx wc;
    {
        double* q = smem + a.cta_doubles + (size_t)warp * a.warp_doubles;
        int mmax = 0;
#pragma unroll
        for (int g = 0; g < MG_GP; ++g) {
            wc.A[g] = q;
            if (g < G) { q += (tri(m.gp[g].M + 2) + 3) & ~3; mmax = max(mmax, m.gp[g].M); }
        }
        wc.B = q; q += (MODE == 1) ? 0 : ((tri(mmax + 1) + 3) & ~3);       // Algorithm3 keeps per-GP buffers instead (Bg)
        wc.phi = q; q += (mmax + 3) & ~3;
        wc.inv = q; q += (mmax + 3) & ~3;
        wc.zv = q; q += (mmax + 4) & ~3;
        wc.rv = q; q += (mmax + 4) & ~3;
        wc.cdf = q; q += (N + 3) & ~3;
#pragma unroll
        for (int g = 0; g < MG_GP; ++g) {
            wc.Bg[g] = wc.B;
            wc.T1s[g] = q;
            if (MODE == 1) {
                wc.Bg[g] = q;
                if (g < G) q += (mg_naugp(m.gp[g].M) + 3) & ~3;
                wc.T1s[g] = q;
                if (g < G) q += (mg_npkp(m.gp[g].M) + 3) & ~3;
            }
        }
    }
#pragma unroll
    for (int g = 0; g < MG_GP; ++g) {
        if (g >= G) break;
        for (int i = tid; i < m.gp[g].M; i += nthr)
            for (int j = 0; j <= i; ++j) ijt[g][tri(i) + j] = (unsigned)i | ((unsigned)j << 16);
        if (tid == 0) ijt[g][m.gp[g].npk] = 0u;                        // pad entry of the pair-wise loops (table length is even)
    }
    __syncthreads();

    // ---- per-chain pointers
    const MargWs L = marg_ws_layout(m, N);
    double* wsc = a.ws + (size_t)chain * L.chain_stride;
    double* strace = a.state_trace + (size_t)chain * T * N * nx;
    double* xtrace = a.xi_trace + (size_t)chain * G * T * N;
    double* lwtrace = a.logw_trace + (size_t)chain * T * N;
    int* atrace = a.anc_trace + (size_t)chain * (T - 1) * N;
    const double* refx = (MODE == 1) ? a.ref_x + (size_t)chain * a.ref_x_stride : nullptr;
    const double* refxi = (MODE == 1) ? a.ref_xi + (size_t)chain * a.ref_xi_stride : nullptr;
    const double* Zc = a.Z ? a.Z + (size_t)chain * T * N * nx : nullptr;
    const double* ZXc = a.ZXI0 ? a.ZXI0 + (size_t)chain * G * N : nullptr;
    const double* Uc = a.U ? a.U + (size_t)chain * T * 2 : nullptr;
    const double* TSc = a.TS ? a.TS + (size_t)chain * G * T * N : nullptr;
    const unsigned pchain = a.chain_base + (unsigned)chain;
    int fail = 0;

    // ------------------------------------------------------------------ one particle, one pass
    // head: the part of step t after the resampling (src/Algorithm1.py:347-389, src/Algorithm3.py:127-188);
    // tail: the part of step t+1 before it (auxiliary quantities, :315-343 / :63-121).
    auto particle_pass = [&](int t, int i, int anc) {
        const int par = t & 1;
        double* wp = wsc + (size_t)par * L.parity_stride;            // written
        const double* wq = wsc + (size_t)(par ^ 1) * L.parity_stride; // gathered by ancestor
        const int ac = min(max(anc, 0), N - 1);                      // JAX gathers clamp
        const bool pinned = (MODE == 1) && (i == N - 1);
        if constexpr (MODE == 1) {
            if (t > 0) {
                // the three per-GP gathers by ancestor (two augmented factors, packed T1: ~21 KB at M = 41) start now and
                // land in shared memory while the new state, the basis and the variates are computed
                __syncwarp();
#pragma unroll
                for (int g = 0; g < MG_GP; ++g) {
                    if (g >= G) break;
                    const int Mg = m.gp[g].M, na = mg_naugp(Mg), np = mg_npkp(Mg);
                    const double* s0 = wq + L.Lp[g] + (size_t)ac * na;
                    const double* s1 = wq + L.LB[g] + (size_t)ac * na;
                    const double* s2 = wq + L.T1p[g] + (size_t)ac * np;
                    for (int e = 2 * lane; e < na; e += 64) { cp_async16(wc.A[g] + e, s0 + e); cp_async16(wc.Bg[g] + e, s1 + e); }
                    for (int e = 2 * lane; e < np; e += 64) cp_async16(wc.T1s[g] + e, s2 + e);
                }
                cp_async_commit();
            }
        }
        double x[MG_NX], xi[MG_GP], z[MG_D], T2v[MG_GP], T3v[MG_GP];
        double ldA[MG_GP], ldB[MG_GP], psA[MG_GP], psB[MG_GP];     // Algorithm3: log det eta1 / Psi of the two factors
        double T2av[MG_GP], T3av[MG_GP], ax1[MG_NX], axi1[MG_GP];
        StepVariates sv;
        // ---- new state
        if (t == 0) {
            double zz[MG_NX];
#pragma unroll
            for (int k = 0; k < MG_NX; k += 2) {
                double za = 0.0, zb = 0.0;
                if (k < nx) {
                    if (a.rng_mode == 1) { za = Zc[(size_t)i * nx + k]; zb = (k + 1 < nx) ? Zc[(size_t)i * nx + k + 1] : 0.0; }
                    else philox_normal2(a.seed, PURPOSE_STATE, pchain, a.iteration, (unsigned)(k >> 1) << 28, (unsigned)i, za, zb);
                }
                zz[k] = za;
                if (k + 1 < MG_NX) zz[k + 1] = zb;
            }
#pragma unroll
            for (int r = 0; r < MG_NX; ++r) {
                double s = 0.0;
                if (r < nx) {
                    s = m.m0[r];
#pragma unroll
                    for (int k = 0; k <= r; ++k) s = fma(m.P0c[r][k], zz[k], s);
                }
                x[r] = s;
            }
        } else {
            double zz[MG_NX];
            if (a.rng_mode != 1) sv = philox_step_variates(a.seed, pchain, a.iteration, (unsigned)t, (unsigned)i, G, nx, lane);
#pragma unroll
            for (int k = 0; k < MG_NX; ++k) {
                double za = 0.0;
                if (k < nx && !m.deterministic) za = (a.rng_mode == 1) ? Zc[((size_t)t * N + i) * nx + k] : sv.zs[k];
                zz[k] = za;
            }
#pragma unroll
            for (int r = 0; r < MG_NX; ++r) {
                double s = 0.0;
                if (r < nx) {
                    s = ldcg(wq + L.auxx + (size_t)ac * nx + r);        // f(x[a], u_{t-1}, xi[a]) = aux state of the ancestor
#pragma unroll
                    for (int k = 0; k <= r; ++k) s = fma(m.Qc[r][k], zz[k], s);
                }
                x[r] = s;
            }
        }
        if (pinned) {
#pragma unroll
            for (int r = 0; r < MG_NX; ++r) if (r < nx) x[r] = refx[(size_t)t * nx + r];
        }
        // ---- interface variables and statistics, GP by GP (unrolled over the compile-time maximum: everything indexed by g —
        // buffers, workspace offsets, the per-GP scalars — is then a register, not a local-memory array)
#pragma unroll
        for (int g = 0; g < MG_GP; ++g) {
            if (g >= G) break;
            const MargGP& gp = m.gp[g];
            const int M = gp.M, npk = gp.npk;
            gp_input<PROG>(m, gp, t, x, z);
            basis_eval(gp, z, wc.phi, lane);
            double xiv, T2a = 0.0, T3a = 0.0;
            if (t == 0) {
                double za, zb;
                if (a.rng_mode == 1) za = ZXc[(size_t)g * N + i];
                else philox_normal2(a.seed, PURPOSE_XI0, pchain, a.iteration, (unsigned)g, (unsigned)i, za, zb);
                xiv = fma(gp.xi_sd, za, gp.xi_mean);
                __syncwarp();
            } else {
                // predictive Student-t from the ancestor's factor (src/Algorithm1.py:249-272)
                double cs = 0.0, ms = 0.0, psia;
                if constexpr (MODE == 1) {
                    // Algorithm3: augmented true-diagonal factor of the ancestor -> wc.A[g] (updated in place below)
                    double* Ag = wc.A[g];
                    if (g == 0) { cp_async_wait_all(); __syncwarp(); }       // prefetched at the top of the pass
                    for (int k = lane; k < M; k += 32) wc.inv[k] = 1.0 / Ag[tri(k) + k];
                    __syncwarp();
                    double w[ROWS];
                    warp_fwd_solve_inv<ROWS>(Ag, wc.inv, wc.phi, M, lane, w);
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) {
                        const int idx = lane + 32 * r;
                        if (idx < M) { cs = fma(w[r], w[r], cs); ms = fma(Ag[tri(M) + idx], w[r], ms); }
                    }
                    const double sp = Ag[tri(M) + M];
                    psia = sp * sp;
                } else {
                    const double* La = wq + L.Lp[g] + (size_t)ac * mg_naugp(M);
#pragma unroll 8
                    for (int e = lane; e < npk; e += 32) wc.B[e] = ldcg(La + e);        // independent L2 gathers in flight
                    __syncwarp();
                    double w[ROWS];
                    warp_fwd_solve<ROWS>(wc.B, wc.phi, M, lane, w);
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) {
                        const int idx = lane + 32 * r;
                        if (idx < M) { cs = fma(w[r], w[r], cs); ms = fma(ldcg(wq + L.yv[g] + (size_t)ac * M + idx), w[r], ms); }
                    }
                    psia = ldcg(wq + L.psi[g] + ac);                      // eta2 - mean eta0
                }
                cs = warp_sum(cs) + 1.0;                                  // basis V basis^T + 1
                ms = warp_sum(ms);                                        // basis mean^T
                T2a = ldcg(wq + L.T2[g] + ac);
                T3a = ldcg(wq + L.T3[g] + ac);
                const double df = (gp.p3 + lam * T3a) + 1.0 - 1.0;        // df + 1 - n_xi (src/BayesianInferrence.py:78)
                double tv;
                if (a.rng_mode == 1) tv = TSc[((size_t)g * T + t) * N + i];
                else tv = student_t_from(sv, g, a.seed, PURPOSE_TVAR + g, pchain, a.iteration, (unsigned)t, (unsigned)i, df);
                xiv = ms + sqrt(psia / df) * tv * sqrt(cs);               // src/BayesianInferrence.py:98-108
            }
            if (pinned) xiv = refxi[(size_t)g * a.ref_xi_gstride + t];
            xi[g] = xiv;
            T2av[g] = T2a;
            T3av[g] = T3a;
            if constexpr (MODE == 0) {
            // statistics: S_t = lambda S_{t-1}[a] + T(xi, phi)  (src/Algorithm1.py:315-318, :356-375)
            double* T1w = wp + L.T1p[g] + (size_t)i * mg_npkp(M);
            const double* T1a = wq + L.T1p[g] + (size_t)ac * mg_npkp(M);
            double* Ag = wc.A[g];
            const int rowM = tri(M);
            const double T2n = (t > 0) ? fma(lam, T2a, xiv * xiv) : xiv * xiv;
            const double T3n = (t > 0) ? fma(lam, T3a, 1.0) : 1.0;
            T2v[g] = T2n;
            T3v[g] = T3n;
#pragma unroll 4
                for (int e = lane; e < npk; e += 32) {
                    const unsigned ij = ijt[g][e];
                    double v = wc.phi[ij & 0xffffu] * wc.phi[ij >> 16];
                    if (t > 0) v = fma(lam, ldcg(T1a + e), v);
                    T1w[e] = v;
                    Ag[e] = fma(lam, v, gp.p1[e]);                            // eta1 of the NEXT step: prior + lambda T1
                }
                for (int k = lane; k < M; k += 32) {
                    double v = wc.phi[k] * xiv;
                    if (t > 0) v = fma(lam, ldcg(wq + L.T0[g] + (size_t)ac * M + k), v);
                    wp[L.T0[g] + (size_t)i * M + k] = v;
                    Ag[rowM + k] = fma(lam, v, gp.p0[k]);                     // extra row 1: eta0
                }
                if (lane == 0) {
                    wp[L.T2[g] + i] = T2n;
                    wp[L.T3[g] + i] = T3n;
                    Ag[rowM + M] = fma(lam, T2n, gp.p2);                      // eta2
                    xtrace[((size_t)g * T + t) * N + i] = xiv;
                }
                __syncwarp();
            }
        }
        if constexpr (MODE == 1) {
            // Algorithm3 (lambda = 1), second pass over the GPs, once every xi is known.  The statistics change by the rank-one
            // term [phi; xi][phi; xi]^T and the remaining reference statistics lose [phi_ref; xi_ref][phi_ref; xi_ref]^T
            // (src/Algorithm3.py:153-174), so both augmented factors follow by Givens / hyperbolic rotation sweeps, O(M^2),
            // instead of two O(M^3) factorisations; the same sweep also solves L'^-1 phi(aux state) for the NEXT step's
            // auxiliary interface variable.  Every MG_REFRESH steps (and whenever a downdate loses definiteness) the
            // factors are rebuilt from the statistics.
            const bool more = t < T - 1;
            if (more) transition<PROG>(m, t, x, xi, ax1);
#pragma unroll
            for (int g = 0; g < MG_GP; ++g) {
                if (g >= G) break;
                const MargGP& gp = m.gp[g];
                const int M = gp.M, npk = gp.npk, rowM = tri(M), naug = npk + M + 1;
                const size_t trow = (size_t)chain * T + t;
                const double xiv = xi[g];
                double* T1w = wp + L.T1p[g] + (size_t)i * mg_npkp(M);
                double* Ag = wc.A[g];
                double* Bq = wc.Bg[g];
                const double T2n = (t > 0) ? T2av[g] + xiv * xiv : xiv * xiv;
                const double T3n = (t > 0) ? T3av[g] + 1.0 : 1.0;
                T2v[g] = T2n;
                T3v[g] = T3n;
                if (G > 1 || t == 0) { gp_input<PROG>(m, gp, t, x, z); basis_eval(gp, z, wc.phi, lane); }     // G == 1: still in wc.phi
                if (more) { gp_input<PROG>(m, gp, t + 1, ax1, z); basis_eval(gp, z, wc.inv, lane); }           // phi(aux state) -> wc.inv
                __syncwarp();
                // two packed entries per lane and trip (16-byte loads / stores; the tables are padded to even length)
#pragma unroll 4
                for (int e = 2 * lane; e < npk; e += 64) {
                    const uint2 ij = *reinterpret_cast<const uint2*>(ijt[g] + e);
                    double2 v;
                    v.x = wc.phi[ij.x & 0xffffu] * wc.phi[ij.x >> 16];
                    v.y = wc.phi[ij.y & 0xffffu] * wc.phi[ij.y >> 16];
                    if (t > 0) { const double2 o = *reinterpret_cast<const double2*>(wc.T1s[g] + e); v.x += o.x; v.y += o.y; }
                    *reinterpret_cast<double2*>(T1w + e) = v;
                }
                for (int k = lane; k < M; k += 32) {
                    double v = wc.phi[k] * xiv;
                    if (t > 0) v += ldcg(wq + L.T0[g] + (size_t)ac * M + k);
                    wp[L.T0[g] + (size_t)i * M + k] = v;
                    wc.zv[k] = wc.phi[k];
                    wc.rv[k] = a.tab.RPHI[g][trow * M + k];
                    if (!more) wc.inv[k] = 0.0;
                }
                if (lane == 0) {
                    wp[L.T2[g] + i] = T2n;
                    wp[L.T3[g] + i] = T3n;
                    wc.zv[M] = xiv;
                    wc.rv[M] = refxi[(size_t)g * a.ref_xi_gstride + t];
                    xtrace[((size_t)g * T + t) * N + i] = xiv;
                }
                __syncwarp();
                int bad = 0;
                const bool refresh = (t % MG_REFRESH) == 0;
                double w[ROWS];
                if (!refresh) {
                    warp_fused_update<ROWS>(Ag, Bq, wc.zv, wc.rv, wc.inv, M, lane, bad, w);
                    bad = __any_sync(FULL, bad);
                }
                if (refresh || bad) {
                    const double* PR1 = a.tab.PR1[g] + trow * npk;
                    const double* PR0 = a.tab.PR0[g] + trow * M;
#pragma unroll 8
                    for (int e = lane; e < npk; e += 32) { const double v = T1w[e]; Ag[e] = gp.p1[e] + v; Bq[e] = PR1[e] + v; }
                    for (int k = lane; k < M; k += 32) {
                        const double v = wp[L.T0[g] + (size_t)i * M + k];
                        Ag[rowM + k] = gp.p0[k] + v;
                        Bq[rowM + k] = PR0[k] + v;
                    }
                    if (lane == 0) { Ag[rowM + M] = gp.p2 + T2n; Bq[rowM + M] = a.tab.PR2[g][trow] + T2n; }
                    __syncwarp();
                    double psi_tmp;
                    warp_chol_packed(Ag, M, M + 1, lane, fail);
                    finish_aug_factor(Ag, M, lane, psi_tmp);
                    warp_chol_packed(Bq, M, M + 1, lane, fail);
                    finish_aug_factor(Bq, M, lane, psi_tmp);
                    for (int k = lane; k < M; k += 32) wc.zv[k] = 1.0 / Ag[tri(k) + k];      // zv is free now: inverse diagonal
                    __syncwarp();
                    warp_fwd_solve_inv<ROWS>(Ag, wc.zv, wc.inv, M, lane, w);
                }
                double yv = 0.0;
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const int idx = lane + 32 * r;
                    if (idx < M) yv = fma(Ag[rowM + idx], w[r], yv);
                }
                // one butterfly carries both sums: prior_mniw_mean . phi_aux of the next step, and log det A - log det B (g_t - g_T
                // needs only the difference, src/Algorithm3.py:92-106)
                double ld = 0.0;
                for (int k = lane; k < M; k += 32) ld += log(Ag[tri(k) + k]) - log(Bq[tri(k) + k]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { yv += __shfl_xor_sync(FULL, yv, o); ld += __shfl_xor_sync(FULL, ld, o); }
                axi1[g] = yv;
                ldA[g] = 2.0 * ld;
                ldB[g] = 0.0;
                { const double sa = Ag[rowM + M], sb = Bq[rowM + M]; psA[g] = sa * sa; psB[g] = sb * sb; }
                double* LBw = wp + L.LB[g] + (size_t)i * mg_naugp(M);
                double* LAw = wp + L.Lp[g] + (size_t)i * mg_naugp(M);
#pragma unroll 4
                for (int e = 2 * lane; e < naug; e += 64) {
                    *reinterpret_cast<double2*>(LBw + e) = *reinterpret_cast<const double2*>(Bq + e);
                    *reinterpret_cast<double2*>(LAw + e) = *reinterpret_cast<const double2*>(Ag + e);
                }
                __syncwarp();
            }
        }
        // ---- weights and traces
        double lw = 0.0;
        if (t > 0) lw = log_likelihood<PROG>(m, t, x, xi) - ldcg(wq + L.ellaux + ac);
        if (lane == 0) {
            lwtrace[(size_t)t * N + i] = lw;
            if (t > 0) atrace[(size_t)(t - 1) * N + i] = anc;
        }
        {
            double xl = x[0];
#pragma unroll
            for (int r = 1; r < MG_NX; ++r) xl = (lane == r) ? x[r] : xl;
            if (lane < nx) strace[((size_t)t * N + i) * nx + lane] = xl;
        }
        if (t == T - 1) return;

        // ---- tail: auxiliary quantities of step t + 1
        double ax[MG_NX], axi[MG_GP], gdiff = 0.0;
        if constexpr (MODE == 1) {
#pragma unroll
            for (int r = 0; r < MG_NX; ++r) ax[r] = ax1[r];               // computed before the factor updates
        } else {
            transition<PROG>(m, t, x, xi, ax);
        }
#pragma unroll
        for (int g = 0; g < MG_GP; ++g) {
            if (g >= G) break;
            const MargGP& gp = m.gp[g];
            const int M = gp.M, npk = gp.npk, rowM = tri(M), rowV = tri(M + 1);
            double* Ag = wc.A[g];
            if constexpr (MODE == 0) gp_input<PROG>(m, gp, t + 1, ax, z);
            if constexpr (MODE == 1) {
                // Algorithm3: the factors, the auxiliary interface variable and the log-determinants were produced in the second
                // GP pass above.  g_t - g_T (src/Algorithm3.py:92-106): only the particle-dependent terms of
                // prior_mniw_log_base_measure are kept — -n m/2 log(2 pi) cancels between g_t and g_T, and -nu n/2 log 2 -
                // multigammaln(nu/2, n) depend on T3 alone, which is the same for every particle, so they shift all ancestor
                // log-weights equally and vanish in the softmax (src/Algorithm3.py:115-118).
                axi[g] = axi1[g];
                const size_t trow = (size_t)chain * T + t;
                const double gt = 0.5 * ldA[g] + log(psA[g]) * (0.5 * (gp.p3 + T3v[g]));
                const double gT = 0.5 * ldB[g] + log(psB[g]) * (0.5 * (a.tab.PR3[g][trow] + T3v[g]));
                gdiff += gt - gT;
            } else {
                basis_eval(gp, z, Ag + rowV, lane);                       // extra row 2: phi(aux state)
                __syncwarp();
                const double logdet = warp_chol_packed(Ag, M, M + 2, lane, fail);
                double yy = 0.0, yv = 0.0;
                for (int k = lane; k < M; k += 32) {
                    const double y = Ag[rowM + k];
                    yy = fma(y, y, yy);
                    yv = fma(y, Ag[rowV + k], yv);
                    wp[L.yv[g] + (size_t)i * M + k] = y;
                }
                yy = warp_sum(yy);
                axi[g] = warp_sum(yv);                                    // prior_mniw_mean . phi_aux
                const double psi = Ag[rowM + M] - yy;
                double* Lw = wp + L.Lp[g] + (size_t)i * mg_naugp(M);
#pragma unroll 8
                for (int e = lane; e < npk; e += 32) Lw[e] = Ag[e];
                if (lane == 0) wp[L.psi[g] + i] = psi;
                (void)logdet;
                __syncwarp();
            }
        }
        const double ell = log_likelihood<PROG>(m, t + 1, ax, axi);
        const double lwa = ell + lw;
        if (lane == 0) {
            wp[L.ellaux + i] = ell;
            wp[L.lwaux + i] = lwa;
            if (MODE == 1) wp[L.lwanc + i] = (lwa + gdiff) + log_trans_density(m, refx + (size_t)(t + 1) * nx, ax);
        }
        {
            double xl = ax[0];
#pragma unroll
            for (int r = 1; r < MG_NX; ++r) xl = (lane == r) ? ax[r] : xl;
            if (lane < nx) wp[L.auxx + (size_t)i * nx + lane] = xl;
        }
    };

    // weighted means of the per-particle statistics (src/Algorithm1.py:165-169, :445-457) for step tp
    auto weighted_trace = [&](int tp) {
        const double* wq = wsc + (size_t)(tp & 1) * L.parity_stride;
        cta_softmax_cdf(lwtrace + (size_t)tp * N, N, false, false, wsm, red, tid, nthr);
        int base = 0;
        for (int g = 0; g < G; ++g) {
            const int M = m.gp[g].M, npk = m.gp[g].npk, E = npk + M + 2;
            for (int e = rank * nthr + tid; e < E; e += CS * nthr) {
                const double* src;
                size_t stride;
                if (e < npk) { src = wq + L.T1p[g] + e; stride = mg_npkp(M); }
                else if (e < npk + M) { src = wq + L.T0[g] + (e - npk); stride = M; }
                else if (e == npk + M) { src = wq + L.T2[g]; stride = 1; }
                else { src = wq + L.T3[g]; stride = 1; }
                double acc = 0.0;
#pragma unroll 8
                for (int i = 0; i < N; ++i) acc = fma(ldcg(src + (size_t)i * stride), wsm[i], acc);
                const size_t row = (size_t)chain * T + tp;
                if (e < npk) {
                    const unsigned ij = ijt[g][e];
                    const int r = ij & 0xffffu, c = ij >> 16;
                    a.sst[base + 1][(row * M + r) * M + c] = acc;
                    a.sst[base + 1][(row * M + c) * M + r] = acc;
                } else if (e < npk + M) a.sst[base + 0][row * M + (e - npk)] = acc;
                else if (e == npk + M) a.sst[base + 2][row] = acc;
                else a.sst[base + 3][row] = acc;
            }
            base += 4;
        }
        __syncthreads();
    };

    // ------------------------------------------------------------------ t = 0 .. T-1
    const bool want_sst = (MODE == 0) && a.sst[0] != nullptr;
    const int last_owner = (N - 1) % WT;                 // the warp that owns the conditioned particle N - 1
    int bar_gen = 0;
    const double dN = (double)N;
    for (int t = 0; t < T; ++t) {
        double u_res = 0.0, u_anc = 0.0;
        int refidx = 0;
        if (t > 0) {
            // the step's only chain-wide dependency: every particle's auxiliary log-weight.  Warps are the participants of the
            // software barrier, and every warp then builds the resampling table for itself (same data, same code, same bits).
            if (a.sw_barrier) mg_warp_barrier(a.bar_ctr + chain, (unsigned)WT * (unsigned)(++bar_gen), lane);
            else mg_cluster_barrier(multi);
            const double* wq = wsc + (size_t)((t - 1) & 1) * L.parity_stride;
            if (want_sst) weighted_trace(t - 1);
            if (a.rng_mode == 1) { u_res = Uc[(size_t)t * 2]; u_anc = Uc[(size_t)t * 2 + 1]; }
            else philox_uniform2(a.seed, PURPOSE_STEP_U, pchain, a.iteration, (unsigned)t, 0u, u_res, u_anc);
            if (MODE == 1 && wg == last_owner) {
                // ancestor of the conditioned path (src/Algorithm3.py:115-125); not clipped in the reference
                warp_softmax_cdf(wq + L.lwanc, N, false, wc.cdf, lane);
                refidx = warp_count_below(wc.cdf, N, u_anc, lane);
                __syncwarp();
            }
            warp_softmax_cdf(wq + L.lwaux, N, true, wc.cdf, lane);
        }
        for (int i = wg; i < N; i += WT) {
            int anc = 0;
            if (t > 0) {
                anc = min(warp_count_below(wc.cdf, N, __ddiv_rn(__dadd_rn(u_res, (double)i), dN), lane), N - 1);   // src/Filtering.py:28-35
                if (MODE == 1 && i == N - 1) anc = refidx;
            }
            particle_pass(t, i, anc);
        }
    }
    if (want_sst) {
        if (a.sw_barrier) mg_warp_barrier(a.bar_ctr + chain, (unsigned)WT * (unsigned)(++bar_gen), lane);
        else mg_cluster_barrier(multi);
        weighted_trace(T - 1);
    }
    if (fail && lane == 0) atomicMax(a.status + chain, 1);
}

// ---------------------------------------------------------------------------------- reference statistics
// Algorithm2's reference statistics (src/Algorithm2.py:83-96, :139-152): totals over all T steps, and the
// per-step tables Algorithm3 consumes: PR_j[t] = prior_j + (total_j - sum_{s<=t} T_j(s)), subtracted
// sequentially like the reference does (src/Algorithm3.py:235-246, :163-174).  One CTA per (chain, GP).
struct RefStatArgs {
    MargDev m;
    const double* x;  long long x_stride;
    const double* xi; long long xi_stride, xi_gstride;
    const double* tot_in[4 * MG_GP];   // given totals (n_chains, ...) or null -> computed here
    double* tot_out[4 * MG_GP];        // optional totals out (full matrices)
    long long tot_out_stride[4 * MG_GP];
    double* PR0[MG_GP]; double* PR1[MG_GP]; double* PR2[MG_GP]; double* PR3[MG_GP];   // optional tables
    double* RPHI[MG_GP];               // optional (n_chains, T, M): basis of the reference trajectory
};

constexpr int RS_THREADS = 512;
constexpr int RS_CHUNK = 16;

__global__ void __launch_bounds__(RS_THREADS) marg_refstats_kernel(const __grid_constant__ RefStatArgs a) {
    extern __shared__ double sm[];
    const MargDev& m = a.m;
    const int g = blockIdx.y, chain = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const MargGP& gp = m.gp[g];
    const int M = gp.M, npk = gp.npk, T = m.T, E = npk + M + 1;
    double* phis = sm;                      // [RS_CHUNK][M]
    double* xis = phis + RS_CHUNK * M;      // [RS_CHUNK]
    const double* xt = a.x + (size_t)chain * a.x_stride;
    const double* xit = a.xi + (size_t)chain * a.xi_stride + (size_t)g * a.xi_gstride;
    constexpr int EPT = 4;                  // elements per thread (E <= EPT * RS_THREADS checked by the host)
    double acc[EPT];
    int er[EPT], ec[EPT];
    for (int q = 0; q < EPT; ++q) {
        const int e = tid + q * RS_THREADS;
        acc[q] = 0.0;
        er[q] = ec[q] = 0;
        if (e < npk) {
            int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
            while (tri(i + 1) <= e) ++i;
            while (tri(i) > e) --i;
            er[q] = i; ec[q] = e - tri(i);
        }
    }
    const bool have_tot = a.tot_in[4 * g] != nullptr;
    for (int pass = have_tot ? 1 : 0; pass < 2; ++pass) {
        if (pass == 1) {
            if (!a.PR1[g]) break;
            if (have_tot) {
                for (int q = 0; q < EPT; ++q) {
                    const int e = tid + q * RS_THREADS;
                    if (e < npk) acc[q] = a.tot_in[4 * g + 1][((size_t)chain * M + er[q]) * M + ec[q]];
                    else if (e < npk + M) acc[q] = a.tot_in[4 * g][(size_t)chain * M + (e - npk)];
                    else if (e == npk + M) acc[q] = a.tot_in[4 * g + 2][chain];
                }
            }
        }
        for (int t0 = 0; t0 < T; t0 += RS_CHUNK) {
            const int nt = min(RS_CHUNK, T - t0);
            __syncthreads();
            for (int s = warp; s < nt; s += RS_THREADS / 32) {
                double x[MG_NX], z[MG_D];
                for (int k = 0; k < m.n_x; ++k) x[k] = xt[(size_t)(t0 + s) * m.n_x + k];
                gp_input<true>(m, gp, t0 + s, x, z);
                basis_eval(gp, z, phis + s * M, lane);
                if (lane == 0) xis[s] = xit[t0 + s];
            }
            __syncthreads();
            if (pass == 1 && a.RPHI[g])
                for (int e = tid; e < nt * M; e += RS_THREADS) a.RPHI[g][((size_t)chain * T + t0) * M + e] = phis[e];
            for (int s = 0; s < nt; ++s) {
                const double* ph = phis + s * M;
                const double xv = xis[s];
                for (int q = 0; q < EPT; ++q) {
                    const int e = tid + q * RS_THREADS;
                    if (e >= E) break;
                    double v;
                    if (e < npk) v = ph[er[q]] * ph[ec[q]];
                    else if (e < npk + M) v = ph[e - npk] * xv;
                    else v = xv * xv;
                    if (pass == 0) acc[q] += v;
                    else {
                        acc[q] -= v;
                        const size_t trow = (size_t)chain * T + t0 + s;
                        if (e < npk) a.PR1[g][trow * npk + e] = gp.p1[e] + acc[q];
                        else if (e < npk + M) a.PR0[g][trow * M + (e - npk)] = gp.p0[e - npk] + acc[q];
                        else a.PR2[g][trow] = gp.p2 + acc[q];
                    }
                }
            }
        }
        if (pass == 0) {
            for (int q = 0; q < EPT; ++q) {
                const int e = tid + q * RS_THREADS;
                if (e >= E) break;
                if (a.tot_out[4 * g]) {
                    if (e < npk) {
                        double* o = a.tot_out[4 * g + 1] + (size_t)chain * a.tot_out_stride[4 * g + 1];
                        o[(size_t)er[q] * M + ec[q]] = acc[q];
                        o[(size_t)ec[q] * M + er[q]] = acc[q];
                    } else if (e < npk + M) a.tot_out[4 * g][(size_t)chain * a.tot_out_stride[4 * g] + (e - npk)] = acc[q];
                    else a.tot_out[4 * g + 2][(size_t)chain * a.tot_out_stride[4 * g + 2]] = acc[q];
                }
            }
            if (tid == 0 && a.tot_out[4 * g + 3]) a.tot_out[4 * g + 3][(size_t)chain * a.tot_out_stride[4 * g + 3]] = (double)T;
        }
    }
    // T3: total T (sum of ones) or the given value; the table subtracts one per step
    if (a.PR3[g]) {
        const double tot3 = have_tot ? a.tot_in[4 * g + 3][chain] : (double)T;
        for (int t = tid; t < T; t += RS_THREADS) {
            a.PR3[g][(size_t)chain * T + t] = gp.p3 + (tot3 - (double)(t + 1));
        }
    }
}

// ---------------------------------------------------------------------------------- final pick + backward trace
// idx = searchsorted(cumsum(softmax(logw_T-1)), u) (src/Algorithm3.py:291-293), then reconstruct_trajectory
// for the state and every interface variable (:294-298).  One CTA per chain.
__global__ void __launch_bounds__(256) marg_pick_trace_kernel(const double* __restrict__ logw_trace, const double* __restrict__ state_trace,
                                                              const double* __restrict__ xi_trace, const int* __restrict__ anc_trace, int T,
                                                              int N, int nx, int G, int rng_mode, unsigned long long seed,
                                                              unsigned chain_base, unsigned iteration, const double* __restrict__ U,
                                                              int* __restrict__ final_idx, double* __restrict__ traj_x, long long x_stride,
                                                              double* __restrict__ traj_xi, long long xi_stride, long long xi_gstride) {
    extern __shared__ double sm[];
    __shared__ double red[8];
    __shared__ int s_idx;
    const int chain = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    double u, ub;
    if (rng_mode == 1) u = U[(size_t)chain * T * 2];
    else philox_uniform2(seed, PURPOSE_STEP_U, chain_base + chain, iteration, 0u, 0u, u, ub);
    cta_softmax_cdf(logw_trace + ((size_t)chain * T + (T - 1)) * N, N, false, true, sm, red, tid, blockDim.x);
    if (tid == 0) {
        s_idx = count_below(sm, N, u);
        if (final_idx) final_idx[chain] = s_idx;
    }
    __syncthreads();
    if (tid < 32) {
        const double* st = state_trace + (size_t)chain * T * N * nx;
        const double* xt = xi_trace + (size_t)chain * G * T * N;
        const int* an = anc_trace + (size_t)chain * (T - 1) * N;
        int aidx = min(max(s_idx, 0), N - 1);
        for (int t = T - 1; t >= 0; --t) {
            if (lane < nx) traj_x[(size_t)chain * x_stride + (size_t)t * nx + lane] = st[((size_t)t * N + aidx) * nx + lane];
            else if (lane - nx < G) {
                const int g = lane - nx;
                traj_xi[(size_t)chain * xi_stride + (size_t)g * xi_gstride + t] = xt[((size_t)g * T + t) * N + aidx];
            }
            if (t > 0) {
                int nxt = 0;
                if (lane == 0) nxt = an[(size_t)(t - 1) * N + aidx];
                nxt = __shfl_sync(FULL, nxt, 0);
                aidx = min(max(nxt, 0), N - 1);
            }
        }
    }
}

// ---------------------------------------------------------------------------------- small kernels
// per-particle statistics of the last step, unpacked to the reference's shapes (src/Algorithm1.py:488)
__global__ void marg_unpack_stats_kernel(const MargDev m, const double* __restrict__ ws, int N, int n_chains, int parity, int g,
                                         double* __restrict__ T0, double* __restrict__ T1, double* __restrict__ T2, double* __restrict__ T3) {
    const MargWs L = marg_ws_layout(m, N);
    const int M = m.gp[g].M, npk = m.gp[g].npk;
    const size_t per = (size_t)M * M + M + 2, total = per * N * n_chains;
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const size_t ci = q / per, r = q % per;
        const int c = (int)(ci / N), i = (int)(ci % N);
        const double* wq = ws + (size_t)c * L.chain_stride + (size_t)parity * L.parity_stride;
        if (r < (size_t)M * M) {
            const int rr = (int)(r / M), cc = (int)(r % M);
            const int hi = max(rr, cc), lo = min(rr, cc);
            T1[ci * M * M + r] = wq[L.T1p[g] + (size_t)i * mg_npkp(M) + tri(hi) + lo];
        } else if (r < (size_t)M * M + M) T0[ci * M + (r - (size_t)M * M)] = wq[L.T0[g] + (size_t)i * M + (r - (size_t)M * M)];
        else if (r == (size_t)M * M + M) T2[ci] = wq[L.T2[g] + i];
        else T3[ci] = wq[L.T3[g] + i];
    }
}

__global__ void marg_outputs_kernel(const MargDev m, const double* __restrict__ states, const double* __restrict__ xi, int n,
                                    double* __restrict__ obs_out, double* __restrict__ ll_out) {
    const size_t total = (size_t)m.T * n;
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(q / n), i = (int)(q % n);
        double x[MG_NX], xv[MG_GP], y[MG_NY];
        for (int k = 0; k < m.n_x; ++k) x[k] = states[q * m.n_x + k];
        for (int g = 0; g < m.G; ++g) xv[g] = xi[((size_t)g * m.T + t) * n + i];
        if (obs_out) {
            output_mdl<true>(m, t, x, xv, y);
            for (int r = 0; r < m.n_y; ++r) obs_out[q * m.n_y + r] = y[r];
        }
        if (ll_out) ll_out[q] = log_likelihood<true>(m, t, x, xv);
    }
}

// vmap(prior_mniw_log_base_measure), one warp per matrix
__global__ void __launch_bounds__(128) marg_lbm_kernel(const double* __restrict__ T0, const double* __restrict__ T1, const double* __restrict__ T2,
                                                       const double* __restrict__ T3, int n, int M, double* __restrict__ out) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double* A = sm + (size_t)warp * ((tri(M + 1) + 3) & ~3);
    for (int s = blockIdx.x * nw + warp; s < n; s += gridDim.x * nw) {
        for (int i = 0; i < M; ++i)
            for (int j = lane; j <= i; j += 32) A[tri(i) + j] = T1[((size_t)s * M + i) * M + j];
        for (int k = lane; k < M; k += 32) A[tri(M) + k] = T0[(size_t)s * M + k];
        __syncwarp();
        int fail = 0;
        const double logdet = warp_chol_packed(A, M, M + 1, lane, fail);
        double yy = 0.0;
        for (int k = lane; k < M; k += 32) yy = fma(A[tri(M) + k], A[tri(M) + k], yy);
        const double psi = T2[s] - warp_sum(yy);
        if (lane == 0) out[s] = log_base_measure(M, logdet, psi, T3[s]);
        __syncwarp();
    }
}

__global__ void marg_philox_variates_kernel(unsigned long long seed, unsigned chain_base, unsigned iteration, int n_chains, int G, int T,
                                            int N, int n_x, const double* __restrict__ df, double* __restrict__ Z, double* __restrict__ ZXI0,
                                            double* __restrict__ U, double* __restrict__ TS) {
    const size_t total = (size_t)n_chains * T * N;
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(q % N), t = (int)((q / N) % T), c = (int)(q / ((size_t)N * T));
        for (int k = 0; k < n_x; k += 2) {
            double za, zb;
            philox_normal2(seed, PURPOSE_STATE, chain_base + c, iteration, (unsigned)t | ((unsigned)(k >> 1) << 28), (unsigned)i, za, zb);
            Z[q * n_x + k] = za;
            if (k + 1 < n_x) Z[q * n_x + k + 1] = zb;
        }
        for (int g = 0; g < G; ++g) {
            if (t == 0) {
                double za, zb;
                philox_normal2(seed, PURPOSE_XI0, chain_base + c, iteration, (unsigned)g, (unsigned)i, za, zb);
                ZXI0[((size_t)c * G + g) * N + i] = za;
                TS[(((size_t)c * G + g) * T + t) * N + i] = 0.0;
            } else {
                TS[(((size_t)c * G + g) * T + t) * N + i] =
                    philox_student_t(seed, PURPOSE_TVAR + g, chain_base + c, iteration, (unsigned)t, (unsigned)i, df[(size_t)g * T + t]);
            }
        }
        if (i == 0) {
            double ua, ub;
            philox_uniform2(seed, PURPOSE_STEP_U, chain_base + c, iteration, (unsigned)t, 0u, ua, ub);
            U[((size_t)c * T + t) * 2] = ua;
            U[((size_t)c * T + t) * 2 + 1] = ub;
        }
    }
}

// ================================================================== host side
static bool mg_small_chol(const double* A, int n, int lda, double* Lo /* n x n */) {
    for (int i = 0; i < n * n; ++i) Lo[i] = 0.0;
    for (int j = 0; j < n; ++j) {
        double d = A[j * lda + j];
        for (int k = 0; k < j; ++k) d -= Lo[j * n + k] * Lo[j * n + k];
        if (!(d > 0.0)) return false;
        d = sqrt(d);
        Lo[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double v = A[i * lda + j];
            for (int k = 0; k < j; ++k) v -= Lo[i * n + k] * Lo[j * n + k];
            Lo[i * n + j] = v / d;
        }
    }
    return true;
}

// W = L^-1 (lower), logdet_half = sum log diag L
static void mg_tri_inv(const double* Lo, int n, double* W, double* logdet_half) {
    *logdet_half = 0.0;
    for (int i = 0; i < n * n; ++i) W[i] = 0.0;
    for (int j = 0; j < n; ++j) {
        *logdet_half += log(Lo[j * n + j]);
        W[j * n + j] = 1.0 / Lo[j * n + j];
        for (int i = j + 1; i < n; ++i) {
            double v = 0.0;
            for (int k = j; k < i; ++k) v -= Lo[i * n + k] * W[k * n + j];
            W[i * n + j] = v / Lo[i * n + i];
        }
    }
}

// a program of the model plug-in: known opcodes, operands in range, stack within bounds, exactly n_out results (len 0: no program)
static int mg_check_program(const pgas_marg_program& q, int n_var, int n_u, int n_out, const char* what) {
    if (q.len == 0) return 0;
    if (q.len < 0 || q.len > PGAS_MAX_PROG || q.n_const < 0 || q.n_const > PGAS_MAX_PROG || !q.ops || (q.n_const > 0 && !q.consts))
        PGAS_FAIL(-2, "%s: expression program of %d instructions / %d constants (limits %d)", what, q.len, q.n_const, PGAS_MAX_PROG);
    int sp = 0;
    for (int i = 0; i < q.len; ++i) {
        const int op = q.ops[i] & 0xff, arg = q.ops[i] >> 8;
        if (op == PGAS_OP_PUSH_X) { if (arg < 0 || arg >= n_var) PGAS_FAIL(-2, "%s: instruction %d reads variable %d of %d", what, i, arg, n_var); ++sp; }
        else if (op == PGAS_OP_PUSH_U) { if (arg < 0 || arg >= n_u) PGAS_FAIL(-2, "%s: instruction %d reads input component %d of %d", what, i, arg, n_u); ++sp; }
        else if (op == PGAS_OP_PUSH_C) { if (arg < 0 || arg >= q.n_const) PGAS_FAIL(-2, "%s: instruction %d reads constant %d of %d", what, i, arg, q.n_const); ++sp; }
        else if ((op >= PGAS_OP_ADD && op <= PGAS_OP_DIV) || op == PGAS_OP_POW || op == PGAS_OP_ATAN2) { if (sp < 2) PGAS_FAIL(-2, "%s: instruction %d: stack underflow", what, i); --sp; }
        else if (op >= PGAS_OP_NEG && op <= PGAS_OP_ABS) { if (sp < 1) PGAS_FAIL(-2, "%s: instruction %d: stack underflow", what, i); }
        else PGAS_FAIL(-2, "%s: instruction %d: unknown opcode %d", what, i, op);
        if (sp > PGAS_PROG_STACK) PGAS_FAIL(-2, "%s: instruction %d: more than %d operands on the stack", what, i, PGAS_PROG_STACK);
    }
    if (sp != n_out) PGAS_FAIL(-2, "%s: expression program leaves %d values, expected %d", what, sp, n_out);
    return 0;
}

extern "C" int pgas_marg_model_create(const pgas_marg_params* p, pgas_marg_model** out) {
    if (!p || !out) PGAS_FAIL(-1, "pgas_marg_model_create: null argument");
    if (p->n_x < 1 || p->n_x > MG_NX) PGAS_FAIL(-2, "n_x=%d outside [1,%d]", p->n_x, MG_NX);
    if (p->n_y < 1 || p->n_y > MG_NY) PGAS_FAIL(-2, "n_y=%d outside [1,%d]", p->n_y, MG_NY);
    if (p->n_gp < 1 || p->n_gp > MG_GP) PGAS_FAIL(-2, "n_gp=%d outside [1,%d]", p->n_gp, MG_GP);
    if (p->T < 2) PGAS_FAIL(-2, "need T >= 2 (T=%d)", p->T);
    if (!p->observations) PGAS_FAIL(-1, "observations must not be null");
    if ((!p->trans && p->trans_prog.len <= 0) || (!p->outp && p->outp_prog.len <= 0)) PGAS_FAIL(-1, "trans / outp: neither a table nor a program");
    if (p->n_u < 0 || p->n_u > 64 || (p->n_u > 0 && !p->inputs)) PGAS_FAIL(-2, "n_u=%d outside [0,64] or inputs null", p->n_u);
    if (int rc = mg_check_program(p->trans_prog, p->n_x + p->n_gp, p->n_u, p->n_x, "transition_model")) return rc;
    if (int rc = mg_check_program(p->outp_prog, p->n_x + p->n_gp, p->n_u, p->n_y, "output_model")) return rc;
    if (p->out_link != PGAS_LINK_IDENTITY && p->out_link != PGAS_LINK_TANH) PGAS_FAIL(-2, "unknown output link %d", p->out_link);
    MargDev dm;
    memset(&dm, 0, sizeof(dm));
    dm.n_x = p->n_x; dm.n_y = p->n_y; dm.G = p->n_gp; dm.T = p->T; dm.out_link = p->out_link;
    const int nx = p->n_x, ny = p->n_y, G = p->n_gp, T = p->T, W = nx + G + 1;
    {   // process noise (src/StateSpaceModel.py:30, :67-73)
        bool allzero = true;
        double Qm[MG_NX * MG_NX], Lq[MG_NX * MG_NX], Wq[MG_NX * MG_NX], ld = 0.0;
        for (int i = 0; i < nx; ++i) for (int j = 0; j < nx; ++j) { Qm[i * nx + j] = p->Q[i][j]; allzero = allzero && p->Q[i][j] == 0.0; }
        dm.deterministic = allzero ? 1 : 0;
        if (!allzero) {
            if (!mg_small_chol(Qm, nx, nx, Lq)) PGAS_FAIL(-3, "process noise Q is not positive definite");
            mg_tri_inv(Lq, nx, Wq, &ld);
            for (int i = 0; i < nx; ++i) for (int j = 0; j < nx; ++j) { dm.Qc[i][j] = Lq[i * nx + j]; dm.Qw[i][j] = Wq[i * nx + j]; }
            dm.Q_logc = -0.5 * nx * log(2.0 * M_PI) - ld;
        } else {
            dm.Q_logc = NAN;        // Algorithm3's h_x is undefined for a deterministic model (the reference yields NaN)
        }
        double Rm[MG_NY * MG_NY], Lr[MG_NY * MG_NY], Wr[MG_NY * MG_NY];
        for (int i = 0; i < ny; ++i) for (int j = 0; j < ny; ++j) Rm[i * ny + j] = p->R[i][j];
        if (!mg_small_chol(Rm, ny, ny, Lr)) PGAS_FAIL(-3, "output noise R is not positive definite");
        mg_tri_inv(Lr, ny, Wr, &ld);
        for (int i = 0; i < ny; ++i) for (int j = 0; j < ny; ++j) dm.Rw[i][j] = Wr[i * ny + j];
        dm.R_logc = -0.5 * ny * log(2.0 * M_PI) - ld;
        double Pm[MG_NX * MG_NX], Lp[MG_NX * MG_NX];
        for (int i = 0; i < nx; ++i) { dm.m0[i] = p->m0[i]; for (int j = 0; j < nx; ++j) Pm[i * nx + j] = p->P0[i][j]; }
        if (!mg_small_chol(Pm, nx, nx, Lp)) PGAS_FAIL(-3, "initial covariance P0 is not positive definite");
        for (int i = 0; i < nx; ++i) for (int j = 0; j < nx; ++j) dm.P0c[i][j] = Lp[i * nx + j];
    }
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    size_t total = al(sizeof(double) * (size_t)T * nx * W) + al(sizeof(double) * (size_t)T * ny * W) + al(sizeof(double) * (size_t)T * ny);
    std::vector<std::vector<double>> p1pk(G);
    for (int g = 0; g < G; ++g) {
        const pgas_marg_gp& q = p->gp[g];
        if (q.M < 1 || q.M > MG_MAX_M) PGAS_FAIL(-2, "GP %d: M=%d outside [1,%d]", g, q.M, MG_MAX_M);
        if (q.D < 1 || q.D > MG_D) PGAS_FAIL(-2, "GP %d: D=%d outside [1,%d]", g, q.D, MG_D);
        if (!q.sqrt_eig || !q.eta0 || !q.eta1) PGAS_FAIL(-1, "GP %d: null table", g);
        if ((!q.gp_in || !q.gp_post) && q.prog.len <= 0) PGAS_FAIL(-1, "GP %d: neither input tables nor a program", g);
        if (int rc = mg_check_program(q.prog, p->n_x, p->n_u, q.D, "basis_fcn")) return rc;
        if (q.link != PGAS_LINK_IDENTITY && q.link != PGAS_LINK_ATAN) PGAS_FAIL(-2, "GP %d: unknown input link %d", g, q.link);
        if (!(q.xi_var > 0.0)) PGAS_FAIL(-3, "GP %d: init_int_var_cov must be positive", g);
        MargGP& d = dm.gp[g];
        d.M = q.M; d.D = q.D; d.link = q.link; d.npk = q.M * (q.M + 1) / 2;
        for (int k = 0; k < q.D; ++k) {
            if (!(q.half_width[k] > 0.0)) PGAS_FAIL(-2, "GP %d: half_width[%d] must be positive", g, k);
            d.center[k] = q.center[k]; d.L[k] = q.half_width[k]; d.sqrt_invL[k] = sqrt(1.0 / q.half_width[k]);
        }
        d.p2 = q.eta2; d.p3 = q.eta3; d.xi_mean = q.xi_mean; d.xi_sd = sqrt(q.xi_var);
        p1pk[g].resize(d.npk);
        for (int i = 0; i < q.M; ++i)
            for (int j = 0; j <= i; ++j) p1pk[g][i * (i + 1) / 2 + j] = 0.5 * (q.eta1[(size_t)i * q.M + j] + q.eta1[(size_t)j * q.M + i]);
        total += al(sizeof(double) * q.M * q.D) + al(sizeof(double) * (size_t)T * q.D * (nx + 1)) + al(sizeof(double) * (size_t)T * q.D * 2) +
                 al(sizeof(double) * q.M) + al(sizeof(double) * d.npk);
    }
    total += al(sizeof(double) * (size_t)T * std::max(p->n_u, 1)) + (size_t)(2 + G) * (al(sizeof(int) * PGAS_MAX_PROG) + al(sizeof(double) * PGAS_MAX_PROG));
    char* arena = nullptr;
    PGAS_CUDA(cudaMalloc((void**)&arena, total));
    size_t o = 0;
    auto up = [&](const void* src, size_t bytes) -> const double* {
        const double* d = (const double*)(arena + o);
        cudaMemcpy(arena + o, src, bytes, cudaMemcpyHostToDevice);
        o += al(bytes);
        return d;
    };
    auto up_prog = [&](const pgas_marg_program& q, int n_out) {
        MargProg d;
        d.ops = nullptr; d.consts = nullptr; d.len = 0; d.n_out = n_out;
        if (q.len > 0) {
            d.ops = (const int*)up(q.ops, sizeof(int) * q.len);
            d.consts = q.n_const > 0 ? up(q.consts, sizeof(double) * q.n_const) : nullptr;
            d.len = q.len;
            dm.any_prog = 1;
        }
        return d;
    };
    dm.trans = p->trans ? up(p->trans, sizeof(double) * (size_t)T * nx * W) : nullptr;
    dm.outp = p->outp ? up(p->outp, sizeof(double) * (size_t)T * ny * W) : nullptr;
    dm.obs = up(p->observations, sizeof(double) * (size_t)T * ny);
    dm.n_u = p->n_u;
    dm.inputs = p->n_u > 0 ? up(p->inputs, sizeof(double) * (size_t)T * p->n_u) : nullptr;
    dm.tprog = up_prog(p->trans_prog, nx);
    dm.oprog = up_prog(p->outp_prog, ny);
    for (int g = 0; g < G; ++g) {
        const pgas_marg_gp& q = p->gp[g];
        MargGP& d = dm.gp[g];
        d.sqrt_eig = up(q.sqrt_eig, sizeof(double) * q.M * q.D);
        d.gp_in = q.gp_in ? up(q.gp_in, sizeof(double) * (size_t)T * q.D * (nx + 1)) : nullptr;
        d.gp_post = q.gp_post ? up(q.gp_post, sizeof(double) * (size_t)T * q.D * 2) : nullptr;
        d.prog = up_prog(q.prog, q.D);
        d.p0 = up(q.eta0, sizeof(double) * q.M);
        d.p1 = up(p1pk[g].data(), sizeof(double) * d.npk);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(arena); PGAS_FAIL((int)e, "marginal model upload failed: %s", cudaGetErrorString(e)); }
    pgas_marg_model* mdl = new pgas_marg_model;
    mdl->dev = dm; mdl->arena = arena; mdl->arena_bytes = total;
    *out = mdl;
    return 0;
}

extern "C" int pgas_marg_model_destroy(pgas_marg_model* model) {
    if (!model) return 0;
    cudaFree(model->arena);
    delete model;
    return 0;
}

// ---- workspace: [per-particle ping-pong | reference tables | run traces]
struct MargHostWs {
    unsigned* bar;                                 // n_chains software-barrier counters
    double* part;                                  // n_chains * chain_stride
    double *PR0[MG_GP], *PR1[MG_GP], *PR2[MG_GP], *PR3[MG_GP], *RPHI[MG_GP];
    size_t total;
};

static MargHostWs mg_carve(const MargDev& m, int N, int n_chains, char* base) {
    MargHostWs w;
    size_t o = 0;
    auto take = [&](size_t bytes) { char* p = base ? base + o : nullptr; o += (bytes + 255) & ~(size_t)255; return (double*)p; };
    const MargWs L = marg_ws_layout(m, N);
    w.bar = (unsigned*)take(sizeof(unsigned) * n_chains);
    w.part = take(sizeof(double) * L.chain_stride * n_chains);
    const size_t CT = (size_t)n_chains * m.T;
    for (int g = 0; g < MG_GP; ++g) {
        const bool on = g < m.G;
        w.PR0[g] = on ? take(sizeof(double) * CT * m.gp[g].M) : nullptr;
        w.PR1[g] = on ? take(sizeof(double) * CT * m.gp[g].npk) : nullptr;
        w.PR2[g] = on ? take(sizeof(double) * CT) : nullptr;
        w.PR3[g] = on ? take(sizeof(double) * CT) : nullptr;
        w.RPHI[g] = on ? take(sizeof(double) * CT * m.gp[g].M) : nullptr;
    }
    w.total = o + 256;
    return w;
}

extern "C" size_t pgas_marg_workspace_bytes(const pgas_marg_model* model, int32_t N, int32_t n_chains) {
    if (!model || N < 1 || n_chains < 1) return 0;
    return mg_carve(model->dev, N, n_chains, nullptr).total;
}

static size_t mg_warp_doubles(const MargDev& m, int mode, int N) {
    size_t d = (size_t)((N + 3) & ~3);
    int mmax = 0;
    for (int g = 0; g < m.G; ++g) { d += ((m.gp[g].M + 2) * (m.gp[g].M + 3) / 2 + 3) & ~3; mmax = std::max(mmax, m.gp[g].M); }
    if (mode != 1) d += ((mmax + 1) * (mmax + 2) / 2 + 3) & ~3;
    d += 2 * ((mmax + 3) & ~3) + 2 * ((mmax + 4) & ~3);
    if (mode == 1)
        for (int g = 0; g < m.G; ++g) d += ((mg_naugp(m.gp[g].M) + 3) & ~3) + ((mg_npkp(m.gp[g].M) + 3) & ~3);
    return d;
}
static size_t mg_cta_doubles(const MargDev& m, int N) {
    size_t d = 2 * (size_t)((N + 3) & ~3) + 64;
    for (int g = 0; g < m.G; ++g) d += ((m.gp[g].npk + 1) & ~1) / 2;
    return (d + 3) & ~(size_t)3;
}

static int mg_fill_rng(MargArgs& a, const pgas_marg_rng* rng) {
    if (!rng) PGAS_FAIL(-1, "rng must not be null");
    if (rng->mode != 0 && rng->mode != 1) PGAS_FAIL(-2, "unknown rng mode %d", rng->mode);
    if (rng->mode == 1 && (!rng->Z || !rng->ZXI0 || !rng->U || !rng->TS)) PGAS_FAIL(-1, "injected rng mode needs Z, ZXI0, U and TS");
    a.rng_mode = rng->mode; a.seed = rng->seed; a.chain_base = rng->chain_base; a.iteration = rng->iteration;
    a.Z = rng->Z; a.ZXI0 = rng->ZXI0; a.U = rng->U; a.TS = rng->TS;
    return 0;
}

template <int MODE, int ROWS, bool WIDE, bool PROG = false>
static int mg_launch_variant(const MargArgs& a, size_t smem, cudaStream_t st) {
    auto kern = marg_sweep_kernel<MODE, ROWS, WIDE, PROG>;
    PGAS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (a.CS > 8) PGAS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    if (a.sw_barrier) {
        int dev = 0, sms = 0, occ = 0;
        PGAS_CUDA(cudaGetDevice(&dev));
        PGAS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        PGAS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, a.NW * 32, smem));
        if ((long long)a.n_chains * a.CS > (long long)sms * occ) return -100;       // not co-resident: caller falls back to clusters
        MargArgs ac = a;
        void* params[1] = {(void*)&ac};
        PGAS_CUDA(cudaMemsetAsync(a.bar_ctr, 0, sizeof(unsigned) * a.n_chains, st));
        PGAS_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)(a.n_chains * a.CS)), dim3((unsigned)(a.NW * 32)), params, smem, st));
        PGAS_KERNEL_CHECK();
        return 0;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(a.n_chains * a.CS));
    cfg.blockDim = dim3((unsigned)(a.NW * 32));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PGAS_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    PGAS_KERNEL_CHECK();
    return 0;
}

// geometry: warps per CTA from the shared-memory budget, cluster size so that every particle has a warp
static int mg_geometry(MargArgs& a, int requested_cs, size_t* smem_out, bool allow_wide = true) {
    const MargDev& m = a.m;
    a.warp_doubles = mg_warp_doubles(m, a.mode, a.N);
    a.cta_doubles = mg_cta_doubles(m, a.N);
    const size_t budget = 225 * 1024;
    const size_t cta_b = sizeof(double) * a.cta_doubles, warp_b = sizeof(double) * a.warp_doubles;
    if (cta_b + warp_b > budget) PGAS_FAIL(-21, "marginalised filter: N=%d / M too large for the shared-memory carve-up (%zu + %zu bytes)", a.N, cta_b, warp_b);
    int nw = (int)std::min<size_t>(16, (budget - cta_b) / warp_b);
    a.sw_barrier = 0;
    if (allow_wide && requested_cs <= 0 && !getenv("PGAS_MARG_CLUSTER")) {
        // Wide geometry: the step is bound by per-SM instruction issue / latency, not by the barrier, so a chain is
        // spread over as many SMs as there are (4 warps = 4 particles per CTA, one warp per scheduler); the CTAs of a
        // chain then synchronise through a global counter, which needs every CTA of the launch to be co-resident.
        int dev = 0, sms = 0;
        PGAS_CUDA(cudaGetDevice(&dev));
        PGAS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int wnw = std::min(nw, 4);
        const size_t wsmem = cta_b + warp_b * wnw;
        const int per_sm = (int)std::min<size_t>(std::min<size_t>(65536 / (128 * 32 * (size_t)wnw), 16), (228 * 1024) / (wsmem + 1024));
        const int wcs = std::min(std::min(64, (a.N + wnw - 1) / wnw), per_sm >= 1 ? (sms * per_sm) / std::max(a.n_chains, 1) : 0);
        if (wcs > 16 && per_sm >= 1 && (long long)a.n_chains * wcs <= (long long)sms * per_sm) {
            a.sw_barrier = 1; a.NW = wnw; a.CS = wcs;
            *smem_out = wsmem;
            return 0;
        }
    }
    int cs = requested_cs;
    if (cs <= 0) {
        cs = 1;
        while (cs < 16 && cs * nw < a.N) cs *= 2;
    }
    if (cs != 1 && cs != 2 && cs != 4 && cs != 8 && cs != 16) PGAS_FAIL(-2, "cluster_size must be 0 (auto), 1, 2, 4, 8 or 16 (got %d)", cs);
    nw = std::min(nw, std::max(1, (a.N + cs - 1) / cs));     // no idle warps
    a.NW = nw; a.CS = cs;
    *smem_out = cta_b + warp_b * nw;
    return 0;
}

static int mg_launch_sweep(MargArgs& a, int requested_cs, cudaStream_t st) {
    int mmax = 0;
    for (int g = 0; g < a.m.G; ++g) mmax = std::max(mmax, a.m.gp[g].M);
    const bool small = mmax <= 62;
    for (int attempt = 0; attempt < 2; ++attempt) {
        size_t smem = 0;
        if (int rc = mg_geometry(a, requested_cs, &smem, attempt == 0)) return rc;
        int rc;
        // 255-register variant: at most 128 threads per CTA and every CTA of the launch resident at two CTAs per SM
        int dev = 0, sms = 0;
        PGAS_CUDA(cudaGetDevice(&dev));
        PGAS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const bool wide = a.NW <= 4 && (!a.sw_barrier || (long long)a.n_chains * a.CS <= 2ll * sms) && !getenv("PGAS_MARG_NARROW");
        if (a.m.any_prog) {
            // model plug-in: one general instantiation per mode (four register rows cover every M, 128 registers every geometry)
            rc = a.mode == 0 ? mg_launch_variant<0, 4, false, true>(a, smem, st) : mg_launch_variant<1, 4, false, true>(a, smem, st);
        } else if (a.mode == 0) {
            if (wide) rc = small ? mg_launch_variant<0, 2, true>(a, smem, st) : mg_launch_variant<0, 4, true>(a, smem, st);
            else rc = small ? mg_launch_variant<0, 2, false>(a, smem, st) : mg_launch_variant<0, 4, false>(a, smem, st);
        } else {
            if (wide) rc = small ? mg_launch_variant<1, 2, true>(a, smem, st) : mg_launch_variant<1, 4, true>(a, smem, st);
            else rc = small ? mg_launch_variant<1, 2, false>(a, smem, st) : mg_launch_variant<1, 4, false>(a, smem, st);
        }
        if (rc != -100) return rc;
    }
    PGAS_FAIL(-23, "marginalised filter: no launch geometry fits");
}

static int mg_launch_refstats(const MargDev& m, const double* x, long long x_stride, const double* xi, long long xi_stride,
                              long long xi_gstride, int n_chains, const double* const* tot_in, double* const* tot_out,
                              const long long* tot_out_stride, const MargHostWs* tab, cudaStream_t st) {
    RefStatArgs r;
    memset(&r, 0, sizeof(r));
    r.m = m; r.x = x; r.x_stride = x_stride; r.xi = xi; r.xi_stride = xi_stride; r.xi_gstride = xi_gstride;
    int mmax = 0;
    for (int g = 0; g < m.G; ++g) {
        mmax = std::max(mmax, m.gp[g].M);
        if (m.gp[g].npk + m.gp[g].M + 1 > 4 * RS_THREADS) PGAS_FAIL(-20, "GP %d: M=%d exceeds the reference-statistics kernel (M <= 62)", g, m.gp[g].M);
        for (int j = 0; j < 4; ++j) {
            r.tot_in[4 * g + j] = tot_in ? tot_in[4 * g + j] : nullptr;
            r.tot_out[4 * g + j] = tot_out ? tot_out[4 * g + j] : nullptr;
            r.tot_out_stride[4 * g + j] = tot_out_stride ? tot_out_stride[4 * g + j] : 0;
        }
        if (tab) { r.PR0[g] = tab->PR0[g]; r.PR1[g] = tab->PR1[g]; r.PR2[g] = tab->PR2[g]; r.PR3[g] = tab->PR3[g]; r.RPHI[g] = tab->RPHI[g]; }
    }
    const size_t smem = sizeof(double) * ((size_t)RS_CHUNK * mmax + RS_CHUNK);
    marg_refstats_kernel<<<dim3(n_chains, m.G), RS_THREADS, smem, st>>>(r);
    PGAS_KERNEL_CHECK();
    return 0;
}

static int mg_launch_pick(const MargDev& m, const MargArgs& a, int* final_idx, double* traj_x, long long x_stride, double* traj_xi,
                          long long xi_stride, long long xi_gstride, cudaStream_t st) {
    const size_t smem = sizeof(double) * (size_t)a.N;
    if (smem > 48 * 1024) PGAS_CUDA(cudaFuncSetAttribute(marg_pick_trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    marg_pick_trace_kernel<<<a.n_chains, 256, smem, st>>>(a.logw_trace, a.state_trace, a.xi_trace, a.anc_trace, m.T, a.N, m.n_x, m.G,
                                                         a.rng_mode, a.seed, a.chain_base, a.iteration, a.U, final_idx, traj_x, x_stride,
                                                         traj_xi, xi_stride, xi_gstride);
    PGAS_KERNEL_CHECK();
    return 0;
}

extern "C" int pgas_marg_filter_f64(const pgas_marg_model* model, int32_t N, int32_t n_chains, double forgetting_factor,
                                    const pgas_marg_rng* rng, double* state_trace, double* xi_trace, double* logw_trace,
                                    int32_t* anc_trace, double* const* sst_trace, double* const* final_stats, int32_t* status,
                                    int32_t cluster_size, void* workspace, size_t workspace_bytes, void* stream) {
    if (!model || !state_trace || !xi_trace || !logw_trace || !anc_trace || !status || !workspace) PGAS_FAIL(-1, "pgas_marg_filter_f64: null argument");
    if (N < 2 || n_chains < 1) PGAS_FAIL(-2, "bad sizes (N=%d n_chains=%d)", N, n_chains);
    const MargDev& m = model->dev;
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    MargHostWs w = mg_carve(m, N, n_chains, base);
    if (workspace_bytes < w.total) PGAS_FAIL(-5, "workspace too small: need %zu bytes, got %zu", w.total, workspace_bytes);
    MargArgs a;
    memset(&a, 0, sizeof(a));
    a.m = m; a.N = N; a.n_chains = n_chains; a.mode = 0; a.lambda = forgetting_factor;
    a.state_trace = state_trace; a.xi_trace = xi_trace; a.logw_trace = logw_trace; a.anc_trace = anc_trace;
    a.ws = w.part; a.status = status; a.bar_ctr = w.bar;
    if (sst_trace)
        for (int q = 0; q < 4 * m.G; ++q) {
            if (!sst_trace[q]) PGAS_FAIL(-1, "sst_trace[%d] is null", q);
            a.sst[q] = sst_trace[q];
        }
    if (int rc = mg_fill_rng(a, rng)) return rc;
    PGAS_CUDA(cudaMemsetAsync(status, 0, sizeof(int) * n_chains, st));
    if (int rc = mg_launch_sweep(a, cluster_size, st)) return rc;
    if (final_stats)
        for (int g = 0; g < m.G; ++g) {
            marg_unpack_stats_kernel<<<148, 256, 0, st>>>(m, w.part, N, n_chains, (m.T - 1) & 1, g, final_stats[4 * g], final_stats[4 * g + 1],
                                                        final_stats[4 * g + 2], final_stats[4 * g + 3]);
            PGAS_KERNEL_CHECK();
        }
    return 0;
}

extern "C" int pgas_marg_refstats_f64(const pgas_marg_model* model, const double* x_traj, int64_t x_stride, const double* xi_traj,
                                      int64_t xi_stride, int64_t xi_gstride, int32_t n_chains, double* const* stats_out, void* stream) {
    if (!model || !x_traj || !xi_traj || !stats_out) PGAS_FAIL(-1, "pgas_marg_refstats_f64: null argument");
    const MargDev& m = model->dev;
    long long strides[4 * MG_GP];
    for (int g = 0; g < m.G; ++g) {
        strides[4 * g] = m.gp[g].M; strides[4 * g + 1] = (long long)m.gp[g].M * m.gp[g].M; strides[4 * g + 2] = 1; strides[4 * g + 3] = 1;
        for (int j = 0; j < 4; ++j) if (!stats_out[4 * g + j]) PGAS_FAIL(-1, "stats_out[%d] is null", 4 * g + j);
    }
    return mg_launch_refstats(m, x_traj, x_stride, xi_traj, xi_stride, xi_gstride, n_chains, nullptr, stats_out, strides, nullptr,
                              (cudaStream_t)stream);
}

static int mg_csmc(const pgas_marg_model* model, int N, int n_chains, const double* ref_x, long long ref_x_stride, const double* ref_xi,
                   long long ref_xi_stride, long long ref_xi_gstride, const double* const* ref_stats, const pgas_marg_rng* rng,
                   double* state_trace, double* xi_trace, double* logw_trace, int* anc_trace, int* final_idx, double* traj_x,
                   long long tx_stride, double* traj_xi, long long txi_stride, long long txi_gstride, int* status, int cluster_size,
                   const MargHostWs& w, bool tables_ready, cudaStream_t st) {
    const MargDev& m = model->dev;
    if (m.deterministic) PGAS_FAIL(-3, "Algorithm3 needs a positive-definite process noise (the ancestor weights use log N(x_ref; f(x), Q))");
    if (!tables_ready)
        if (int rc = mg_launch_refstats(m, ref_x, ref_x_stride, ref_xi, ref_xi_stride, ref_xi_gstride, n_chains, ref_stats, nullptr, nullptr,
                                        &w, st))
            return rc;
    MargArgs a;
    memset(&a, 0, sizeof(a));
    a.m = m; a.N = N; a.n_chains = n_chains; a.mode = 1; a.lambda = 1.0;       // src/Algorithm3.py:34
    a.ref_x = ref_x; a.ref_x_stride = ref_x_stride;
    a.ref_xi = ref_xi; a.ref_xi_stride = ref_xi_stride; a.ref_xi_gstride = ref_xi_gstride;
    for (int g = 0; g < m.G; ++g) { a.tab.PR0[g] = w.PR0[g]; a.tab.PR1[g] = w.PR1[g]; a.tab.PR2[g] = w.PR2[g]; a.tab.PR3[g] = w.PR3[g]; a.tab.RPHI[g] = w.RPHI[g]; }
    a.state_trace = state_trace; a.xi_trace = xi_trace; a.logw_trace = logw_trace; a.anc_trace = anc_trace;
    a.ws = w.part; a.status = status; a.bar_ctr = w.bar;
    if (int rc = mg_fill_rng(a, rng)) return rc;
    if (int rc = mg_launch_sweep(a, cluster_size, st)) return rc;
    if (traj_x) return mg_launch_pick(m, a, final_idx, traj_x, tx_stride, traj_xi, txi_stride, txi_gstride, st);
    return 0;
}

extern "C" int pgas_marg_csmc_f64(const pgas_marg_model* model, int32_t N, int32_t n_chains, const double* ref_x, const double* ref_xi,
                                  const double* const* ref_stats, const pgas_marg_rng* rng, double* state_trace, double* xi_trace,
                                  double* logw_trace, int32_t* anc_trace, int32_t* final_idx, double* traj_x_out, double* traj_xi_out,
                                  int32_t* status, int32_t cluster_size, void* workspace, size_t workspace_bytes, void* stream) {
    if (!model || !ref_x || !ref_xi || !state_trace || !xi_trace || !logw_trace || !anc_trace || !status || !workspace)
        PGAS_FAIL(-1, "pgas_marg_csmc_f64: null argument");
    if ((traj_x_out == nullptr) != (traj_xi_out == nullptr)) PGAS_FAIL(-1, "traj_x_out and traj_xi_out must be given together");
    if (N < 2 || n_chains < 1) PGAS_FAIL(-2, "bad sizes (N=%d n_chains=%d)", N, n_chains);
    const MargDev& m = model->dev;
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    MargHostWs w = mg_carve(m, N, n_chains, base);
    if (workspace_bytes < w.total) PGAS_FAIL(-5, "workspace too small: need %zu bytes, got %zu", w.total, workspace_bytes);
    if (ref_stats)
        for (int q = 0; q < 4 * m.G; ++q) if (!ref_stats[q]) PGAS_FAIL(-1, "ref_stats[%d] is null", q);
    PGAS_CUDA(cudaMemsetAsync(status, 0, sizeof(int) * n_chains, st));
    return mg_csmc(model, N, n_chains, ref_x, (long long)m.T * m.n_x, ref_xi, (long long)m.G * m.T, m.T, ref_stats, rng, state_trace,
                   xi_trace, logw_trace, anc_trace, final_idx, traj_x_out, (long long)m.T * m.n_x, traj_xi_out, (long long)m.G * m.T, m.T,
                   status, cluster_size, w, false, st);
}

// ---- Algorithm2: K iterations, stream-ordered
struct MargRunWs {
    MargHostWs base;
    double *state, *xi, *logw;
    int* anc;
    size_t total;
};

static MargRunWs mg_carve_run(const MargDev& m, int N, int n_chains, char* base) {
    MargRunWs r;
    r.base = mg_carve(m, N, n_chains, base);
    size_t o = r.base.total;
    auto take = [&](size_t bytes) { char* p = base ? base + o : nullptr; o += (bytes + 255) & ~(size_t)255; return p; };
    const size_t C = n_chains, T = m.T;
    r.state = (double*)take(sizeof(double) * C * T * N * m.n_x);
    r.xi = (double*)take(sizeof(double) * C * m.G * T * N);
    r.logw = (double*)take(sizeof(double) * C * T * N);
    r.anc = (int*)take(sizeof(int) * C * (T - 1) * N);
    r.total = o + 256;
    return r;
}

extern "C" size_t pgas_marg_run_workspace_bytes(const pgas_marg_model* model, int32_t N, int32_t n_chains) {
    if (!model || N < 1 || n_chains < 1) return 0;
    return mg_carve_run(model->dev, N, n_chains, nullptr).total;
}

extern "C" int pgas_marg_run_f64(const pgas_marg_model* model, int32_t N, int32_t K, int32_t n_chains, const double* init_x,
                                 const double* init_xi, const pgas_marg_rng* rng, double* x_trace_out, double* xi_trace_out,
                                 double* const* sst_out, int32_t* status, int32_t cluster_size, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    if (!model || !init_x || !init_xi || !rng || !x_trace_out || !xi_trace_out || !status || !workspace)
        PGAS_FAIL(-1, "pgas_marg_run_f64: null argument");
    if (N < 2 || n_chains < 1 || K < 1) PGAS_FAIL(-2, "bad sizes (N=%d n_chains=%d K=%d)", N, n_chains, K);
    const MargDev& m = model->dev;
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    MargRunWs w = mg_carve_run(m, N, n_chains, base);
    if (workspace_bytes < w.total) PGAS_FAIL(-5, "workspace too small: need %zu bytes, got %zu", w.total, workspace_bytes);
    const size_t T = m.T, nx = m.n_x, G = m.G;
    const long long xs = (long long)K * T * nx;                 // chain stride of x_trace_out (n_chains, K, T, n_x)
    const long long xis = (long long)G * K * T, xig = (long long)K * T;   // xi_trace_out (n_chains, G, K, T)
    PGAS_CUDA(cudaMemsetAsync(status, 0, sizeof(int) * n_chains, st));
    PGAS_CUDA(cudaMemcpy2DAsync(x_trace_out, sizeof(double) * xs, init_x, sizeof(double) * T * nx, sizeof(double) * T * nx, n_chains,
                                cudaMemcpyDeviceToDevice, st));
    for (size_t g = 0; g < G; ++g)      // init_xi (n_chains, G, T) -> xi_trace_out[:, g, 0, :]
        PGAS_CUDA(cudaMemcpy2DAsync(xi_trace_out + g * xig, sizeof(double) * xis, init_xi + g * T, sizeof(double) * G * T, sizeof(double) * T,
                                    n_chains, cudaMemcpyDeviceToDevice, st));
    long long sst_stride[4 * MG_GP];
    for (size_t g = 0; g < G; ++g) {
        const long long M = m.gp[g].M;
        sst_stride[4 * g] = K * M; sst_stride[4 * g + 1] = K * M * M; sst_stride[4 * g + 2] = K; sst_stride[4 * g + 3] = K;
    }
    for (int k = 0; k < K; ++k) {
        // statistics of trajectory k (src/Algorithm2.py:83-96, :139-152) and the tables sweep k+1 consumes
        double* tot[4 * MG_GP];
        for (size_t g = 0; g < G; ++g) {
            const size_t M = m.gp[g].M;
            tot[4 * g] = sst_out ? sst_out[4 * g] + (size_t)k * M : nullptr;
            tot[4 * g + 1] = sst_out ? sst_out[4 * g + 1] + (size_t)k * M * M : nullptr;
            tot[4 * g + 2] = sst_out ? sst_out[4 * g + 2] + k : nullptr;
            tot[4 * g + 3] = sst_out ? sst_out[4 * g + 3] + k : nullptr;
        }
        const double* xk = x_trace_out + (size_t)k * T * nx;
        const double* xik = xi_trace_out + (size_t)k * T;
        const bool more = k + 1 < K;
        if (sst_out || more)
            if (int rc = mg_launch_refstats(m, xk, xs, xik, xis, xig, n_chains, nullptr, sst_out ? tot : nullptr, sst_stride,
                                            more ? &w.base : nullptr, st))
                return rc;
        if (!more) break;
        pgas_marg_rng r = *rng;
        r.iteration = rng->iteration + (unsigned)(k + 1);
        if (rng->mode == 1) {
            const size_t kk = (size_t)(k + 1) * n_chains;
            r.Z = rng->Z + kk * T * N * nx;
            r.ZXI0 = rng->ZXI0 + kk * G * N;
            r.U = rng->U + kk * T * 2;
            r.TS = rng->TS + kk * G * T * N;
        }
        if (int rc = mg_csmc(model, N, n_chains, xk, xs, xik, xis, xig, nullptr, &r, w.state, w.xi, w.logw, w.anc, nullptr,
                             x_trace_out + (size_t)(k + 1) * T * nx, xs, xi_trace_out + (size_t)(k + 1) * T, xis, xig, status, cluster_size,
                             w.base, true, st))
            return rc;
    }
    return 0;
}

extern "C" int pgas_marg_outputs_f64(const pgas_marg_model* model, const double* states, const double* xi, int32_t n, double* obs_out,
                                     double* loglik_out, void* stream) {
    if (!model || !states || !xi) PGAS_FAIL(-1, "pgas_marg_outputs_f64: null argument");
    if (n < 1) return 0;
    marg_outputs_kernel<<<148 * 2, 256, 0, (cudaStream_t)stream>>>(model->dev, states, xi, n, obs_out, loglik_out);
    PGAS_KERNEL_CHECK();
    return 0;
}

extern "C" int pgas_mniw_log_base_measure_f64(const double* T0, const double* T1, const double* T2, const double* T3, int32_t n, int32_t M,
                                              double* out, void* stream) {
    if (!T0 || !T1 || !T2 || !T3 || !out) PGAS_FAIL(-1, "pgas_mniw_log_base_measure_f64: null argument");
    if (M < 1 || M > MG_MAX_M) PGAS_FAIL(-2, "M=%d outside [1,%d]", M, MG_MAX_M);
    if (n < 1) return 0;
    const int nw = 4;
    const size_t smem = sizeof(double) * nw * (size_t)(((M + 1) * (M + 2) / 2 + 3) & ~3);
    PGAS_CUDA(cudaFuncSetAttribute(marg_lbm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    marg_lbm_kernel<<<std::min((n + nw - 1) / nw, 148 * 4), nw * 32, smem, (cudaStream_t)stream>>>(T0, T1, T2, T3, n, M, out);
    PGAS_KERNEL_CHECK();
    return 0;
}

extern "C" int pgas_philox_marg_variates_f64(const pgas_marg_rng* rng, int32_t n_chains, int32_t n_gp, int32_t T, int32_t N, int32_t n_x,
                                             const double* df_host, double* Z_out, double* ZXI0_out, double* U_out, double* TS_out,
                                             void* stream) {
    if (!rng || !df_host || !Z_out || !ZXI0_out || !U_out || !TS_out) PGAS_FAIL(-1, "pgas_philox_marg_variates_f64: null argument");
    double* df_dev = nullptr;
    PGAS_CUDA(cudaMalloc((void**)&df_dev, sizeof(double) * (size_t)n_gp * T));
    cudaMemcpyAsync(df_dev, df_host, sizeof(double) * (size_t)n_gp * T, cudaMemcpyHostToDevice, (cudaStream_t)stream);
    marg_philox_variates_kernel<<<148 * 2, 256, 0, (cudaStream_t)stream>>>(rng->seed, rng->chain_base, rng->iteration, n_chains, n_gp, T, N, n_x,
                                                                         df_dev, Z_out, ZXI0_out, U_out, TS_out);
    cudaError_t e = cudaGetLastError();
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(df_dev);
    if (e != cudaSuccess) PGAS_FAIL((int)e, "variates kernel: %s", cudaGetErrorString(e));
    __atomic_add_fetch(&g_pgas_launches, 1, __ATOMIC_RELAXED);
    return 0;
}
