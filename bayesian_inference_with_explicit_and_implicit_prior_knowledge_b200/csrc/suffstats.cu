// suffstats.cu — MNIW sufficient statistics of a trajectory (first half of PGAS.sample_params,
// reference src/PGAS.py:294-303; prior_mniw_calcStatistics, src/BayesianInferrence.py:53-61):
//   T0 = Phi^T Y (M x n_x),  T1 = Phi^T Phi (M x M),  T2 = Y^T Y (n_x x n_x),  T3 = T-1
// with Phi[t] = basis(x_t, u_t), Y[t] = x_{t+1}, t = 0..T-2.  The reference materialises the
// (T-1, M, M) outer products and sums them; here Phi is recomputed chunk by chunk from the
// trajectory (sine tables in shared memory) and never stored, and T1 is a SYRK over time on the
// FP64 tensor pipe: lower-triangular 64x64 tiles, one CTA per (tile, chain).
//
// Warp-specialised (round 2): one PRODUCER warp builds the chunk's sine tables (one sincospi and a three-term recurrence per time
// step and dimension) into a double-buffered stage; four CONSUMER warps own a 32x32 quadrant each = 4x4 accumulator fragments of
// mma.sync.m8n8k4.f64 (SASS: DMMA.8x8x4) and form their fragments ON THE FLY from the table (D shared loads and D - 1 multiplies per
// fragment element; the lattice positions of a thread's eight fragment columns are loop-invariant registers), one step of four time
// rows ahead of the 16 DMMAs that use them: Phi itself is never staged.  Hand-over through named barriers (full / empty per stage,
// bar.arrive on the giving side, bar.sync on the taking side).  Before, all four warps alternated between table, basis-value and
// tensor phases behind __syncthreads: the scalar FP64 operations of the first two queued behind other CTAs' DMMAs on the one FP64
// pipe and the tensor sub-pipe was 52 % busy (profiles/r02_tail_kernels_summary.md).  On diagonal tiles the upper quadrant is the
// mirror of the lower one and is not computed.  Deterministic: every output element is accumulated by one warp in time order
// (no atomics, no split over time).
#include "basis_eval.cuh"
#include "sweep_args.cuh"

constexpr int ST = 64;         // tile edge
constexpr int S_CONS = 128;    // consumer threads: 4 warps, a 32x32 quadrant each
constexpr int S_PROD = 32;     // producer threads (one warp: with 160 threads ptxas grants 128 registers at three CTAs per SM, with 192 only 96)
constexpr int SNT = S_CONS + S_PROD;
constexpr int MAXPOS = 48;     // lattice positions per dimension held in the shared sine table
constexpr int BAR_FULL = 1, BAR_EMPTY = 3;      // named barriers: full[2], empty[2]

struct SuffArgs {
    DevModel m;
    int n_chains, ntile, npos, TK;   // TK = time steps per chunk (multiple of 4)
    const double* traj;      // (n_chains, T, n_x), chain stride traj_stride elements
    long long traj_stride;
    double* T0;              // (n_chains, M, n_x)
    double* T1;              // (n_chains, M, M)
    double* T2;              // (n_chains, n_x, n_x)
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__host__ __device__ constexpr int suff_row(int npos) { return (npos + 1) | 1; }   // table row: >= npos + 1 (a zero slot), odd
__host__ __device__ inline size_t suff_stage_doubles(int D, int npos, int TK) { return (size_t)D * suff_row(npos) * TK + (size_t)TK * PGAS_MAX_NX; }

template <int D>
__global__ void __launch_bounds__(SNT, 3) suffstats_kernel(const __grid_constant__ SuffArgs a) {
    const DevModel& m = a.m;
    // tile pair (I >= J) from the linear block index
    int I = 0, rem = blockIdx.x;
    while (rem > I) { rem -= I + 1; ++I; }
    const int J = rem;
    const int chain = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nx = m.n_x, M = m.M, npos = a.npos, TK = a.TK;
    const double* traj = a.traj + (size_t)chain * a.traj_stride;
    const int nsteps = m.T - 1;
    const int nchunks = (nsteps + TK - 1) / TK;
    const bool diag = (I == J);

    extern __shared__ __align__(16) double sm[];
    // stage: sine table [D][TK][nposp] (position fastest, odd row length, slot npos of every row is zero: columns past M point
    // there), then the Y rows [TK][PGAS_MAX_NX]
    const int nposp = suff_row(npos), dstride = TK * nposp;
    const size_t stage_sz = suff_stage_doubles(D, npos, TK);

    if (warp >= S_CONS / 32) {
        // ------------------------------------------------------------------------------------------------ producers
        const int pt = tid - S_CONS;
        for (int c = 0; c < nchunks; ++c) {
            const int t0 = c * TK;
            double* tab = sm + (size_t)(c & 1) * stage_sz;
            double* ych = tab + (size_t)D * dstride;
            if (c >= 2) bar_sync(BAR_EMPTY + (c & 1), SNT);   // the consumers have released this stage (chunk c - 2)
            // sine tables of the chunk: item (tt, d); the normalisation rides on dimension 0
            for (int it = pt; it < TK * D; it += S_PROD) {
                const int tt = it % TK, d = it / TK, t = t0 + tt;
                double* tb = tab + ((size_t)d * TK + tt) * nposp;
                if (t < nsteps) {
                    double x[PGAS_MAX_NX], u[PGAS_MAX_NU], z = 0.0;
#pragma unroll
                    for (int k = 0; k < PGAS_MAX_NX; ++k) x[k] = (k < nx) ? traj[(size_t)t * nx + k] : 0.0;
#pragma unroll
                    for (int k = 0; k < PGAS_MAX_NU; ++k) u[k] = (k < m.n_u) ? m.inputs[(size_t)t * m.n_u + k] : 0.0;   // x_t pairs with u_t (src/PGAS.py:294-296)
                    if (m.map_kind == PGAS_MAP_AFFINE) {
                        z = m.bz[d];
#pragma unroll
                        for (int k = 0; k < PGAS_MAX_NX; ++k) z = (k < nx) ? fma(m.Az[d][k], x[k], z) : z;
#pragma unroll
                        for (int k = 0; k < PGAS_MAX_NU; ++k) z = (k < m.n_u) ? fma(m.Az[d][nx + k], u[k], z) : z;
                    } else {                                   // slip angles / expression program: all components, this item keeps one
                        double zz[PGAS_MAX_D];
                        gp_map_any(m, x, u, zz);
                        z = zz[d];
                    }
                    const double tn = (z - m.center[d] + m.L[d]) * m.inv2L[d];
                    double cur, prev, twoc;
                    sine_seed(tn, m.f_start, m.f_step, cur, prev, twoc);
                    if (d == 0) { cur *= m.norm; prev *= m.norm; }                 // the recurrence is linear
                    for (int p = 0; p < npos; ++p) {
                        tb[p] = cur;
                        const double n = fma(twoc, cur, -prev);
                        prev = cur; cur = n;
                    }
                } else {
                    for (int p = 0; p < npos; ++p) tb[p] = 0.0;
                }
                tb[npos] = 0.0;
            }
            for (int tt = pt; tt < TK; tt += S_PROD) {    // Y rows of the chunk
                const int t = t0 + tt;
                for (int k = 0; k < nx; ++k) ych[tt * PGAS_MAX_NX + k] = (t < nsteps) ? traj[(size_t)(t + 1) * nx + k] : 0.0;
            }
            bar_arrive(BAR_FULL + (c & 1), SNT);
        }
        return;
    }

    // ---------------------------------------------------------------------------------------------------- consumers
    // accumulator fragments: quadrant rows wi*32 + 8 fi + lane/4, columns wj*32 + 8 fj + 2 (lane%4) + {0,1}
    const int wi = warp >> 1, wj = warp & 1;
    const bool work = !(diag && wj > wi);               // diagonal tile: the upper quadrant is the mirror of the lower one
    // table offsets of this thread's fragment elements: time row lane%4, basis columns wi*32 + 8 f + lane/4 (A) / wj*32 + ... (B)
    int offA[4][D], offB[4][D], off0[D];
    {
        const int r = lane & 3, q = lane >> 2;
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            const int gi = I * ST + wi * 32 + 8 * f + q, gj = J * ST + wj * 32 + 8 * f + q;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const int pi = (gi < M) ? (m.freq[(size_t)gi * D + d] - m.f_start) / m.f_step : (d == 0 ? npos : 0);
                const int pj = (gj < M) ? (m.freq[(size_t)gj * D + d] - m.f_start) / m.f_step : (d == 0 ? npos : 0);
                offA[f][d] = d * dstride + r * nposp + pi;
                offB[f][d] = d * dstride + r * nposp + pj;
            }
        }
        const int g0 = I * ST + (tid % ST);               // T0 column of this thread (diagonal tiles)
#pragma unroll
        for (int d = 0; d < D; ++d) off0[d] = d * dstride + ((g0 < M) ? (m.freq[(size_t)g0 * D + d] - m.f_start) / m.f_step : (d == 0 ? npos : 0));
    }
    double c0[4][4], c1[4][4];
#pragma unroll
    for (int fi = 0; fi < 4; ++fi)
#pragma unroll
        for (int fj = 0; fj < 4; ++fj) { c0[fi][fj] = 0.0; c1[fi][fj] = 0.0; }
    double acc0 = 0.0;             // T0 element e = tid of the diagonal tile (mi = e % ST, k = e / ST)
    double acc2 = 0.0;             // T2 element on tile (0,0)

    for (int c = 0; c < nchunks; ++c) {
        const double* tab = sm + (size_t)(c & 1) * stage_sz;
        const double* ych = tab + (size_t)D * dstride;
        bar_sync(BAR_FULL + (c & 1), SNT);
        // rank-TK update on the tensor pipe: per 4 time steps 4 A + 4 B fragment elements feed 16 DMMAs; the elements of the next
        // four time steps are formed while the DMMAs of the current ones run
        if (work) {
            double af[4], bf[4];
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                double va = tab[offA[f][0]], vb = tab[offB[f][0]];
#pragma unroll
                for (int d = 1; d < D; ++d) { va *= tab[offA[f][d]]; vb *= tab[offB[f][d]]; }
                af[f] = va; bf[f] = vb;
            }
#pragma unroll 2
            for (int kk = 0; kk < TK; kk += 4) {
                double an[4], bn[4];
                const double* tn = tab + ((kk + 4 < TK) ? kk + 4 : kk) * nposp;
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    double va = tn[offA[f][0]], vb = tn[offB[f][0]];
#pragma unroll
                    for (int d = 1; d < D; ++d) { va *= tn[offA[f][d]]; vb *= tn[offB[f][d]]; }
                    an[f] = va; bn[f] = vb;
                }
#pragma unroll
                for (int fi = 0; fi < 4; ++fi)
#pragma unroll
                    for (int fj = 0; fj < 4; ++fj) dmma884(c0[fi][fj], c1[fi][fj], af[fi], bf[fj]);
#pragma unroll
                for (int f = 0; f < 4; ++f) { af[f] = an[f]; bf[f] = bn[f]; }
            }
        }
        if (diag) {
            if (tid < ST * nx) {
                const int k = tid / ST;
                const double* tb = tab;
                for (int tt = 0; tt < TK; ++tt, tb += nposp) {
                    double v = tb[off0[0]];
#pragma unroll
                    for (int d = 1; d < D; ++d) v *= tb[off0[d]];
                    acc0 = fma(v, ych[tt * PGAS_MAX_NX + k], acc0);
                }
            }
            if (I == 0 && tid >= S_CONS - nx * nx) {
                const int e = tid - (S_CONS - nx * nx), r = e / nx, cc = e % nx;
                for (int tt = 0; tt < TK; ++tt) acc2 = fma(ych[tt * PGAS_MAX_NX + r], ych[tt * PGAS_MAX_NX + cc], acc2);
            }
        }
        if (c + 2 < nchunks) bar_arrive(BAR_EMPTY + (c & 1), SNT);
    }
    // write back: lower tile and its mirror
    double* T1 = a.T1 + (size_t)chain * M * M;
    if (work) {
        const bool mirror = !diag || wi != wj;
#pragma unroll
        for (int fi = 0; fi < 4; ++fi)
#pragma unroll
            for (int fj = 0; fj < 4; ++fj) {
                const int gi = I * ST + wi * 32 + fi * 8 + (lane >> 2), gj = J * ST + wj * 32 + fj * 8 + 2 * (lane & 3);
                if (gi < M && gj < M) {
                    T1[(size_t)gi * M + gj] = c0[fi][fj];
                    if (mirror) T1[(size_t)gj * M + gi] = c0[fi][fj];
                }
                if (gi < M && gj + 1 < M) {
                    T1[(size_t)gi * M + gj + 1] = c1[fi][fj];
                    if (mirror) T1[(size_t)(gj + 1) * M + gi] = c1[fi][fj];
                }
            }
    }
    if (diag) {
        if (tid < ST * nx) {
            const int mi = tid % ST, k = tid / ST, gi = I * ST + mi;
            if (gi < M) a.T0[((size_t)chain * M + gi) * nx + k] = acc0;
        }
        if (I == 0 && tid >= S_CONS - nx * nx) {
            const int e = tid - (S_CONS - nx * nx);
            a.T2[(size_t)chain * nx * nx + e] = acc2;
        }
    }
}

static size_t suff_smem(int D, int npos, int TK) {
    return sizeof(double) * 2 * suff_stage_doubles(D, npos, TK);
}

int pgas_launch_suffstats(const DevModel& m, const double* traj, long long traj_stride, int n_chains, double* T0, double* T1,
                          double* T2, cudaStream_t st) {
    const int npos = m.npos;
    SuffArgs a;
    a.m = m;
    a.n_chains = n_chains;
    a.ntile = (m.M + ST - 1) / ST;
    a.npos = npos;
    a.traj = traj; a.traj_stride = traj_stride; a.T0 = T0; a.T1 = T1; a.T2 = T2;
    if (npos > MAXPOS) PGAS_FAIL(-20, "basis uses %d lattice positions per dimension; this build supports <= %d", npos, MAXPOS);
    if (ST * m.n_x > S_CONS) PGAS_FAIL(-20, "n_x too large for the statistics kernel");
    // time steps per chunk: the largest multiple of 4 (<= 32) that leaves room for three CTAs per SM (74 KB each, two stages)
    int TK = 32;
    while (TK > 8 && suff_smem(m.D, npos, TK) > 74 * 1024) TK -= 4;
    a.TK = TK;
    const size_t smem = suff_smem(m.D, npos, TK);
    auto kern = (m.D == 1) ? suffstats_kernel<1> : (m.D == 2) ? suffstats_kernel<2> : suffstats_kernel<3>;
    PGAS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(a.ntile * (a.ntile + 1) / 2), (unsigned)n_chains, 1);
    kern<<<grid, SNT, smem, st>>>(a);
    PGAS_KERNEL_CHECK();
    return 0;
}


extern "C" int pgas_suffstats_f64(const pgas_model* model, const double* traj, int32_t n_chains, double* T0_out, double* T1_out,
                                  double* T2_out, void* stream) {
    if (!model || !traj || !T0_out || !T1_out || !T2_out) PGAS_FAIL(-1, "pgas_suffstats_f64: null argument");
    if (n_chains < 1) PGAS_FAIL(-2, "n_chains must be >= 1");
    return pgas_launch_suffstats(model->dev, traj, (long long)model->dev.T * model->dev.n_x, n_chains, T0_out, T1_out, T2_out,
                                 (cudaStream_t)stream);
}
