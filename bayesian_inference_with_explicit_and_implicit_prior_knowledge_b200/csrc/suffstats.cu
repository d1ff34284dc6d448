// suffstats.cu — MNIW sufficient statistics of a trajectory (first half of PGAS.sample_params,
// reference src/PGAS.py:294-303; prior_mniw_calcStatistics, src/BayesianInferrence.py:53-61):
//   T0 = Phi^T Y (M x n_x),  T1 = Phi^T Phi (M x M),  T2 = Y^T Y (n_x x n_x),  T3 = T-1
// with Phi[t] = basis(x_t, u_t), Y[t] = x_{t+1}, t = 0..T-2.  The reference materialises the
// (T-1, M, M) outer products and sums them; here Phi is recomputed chunk by chunk from the
// trajectory (sine tables in shared memory) and never stored, and T1 is a SYRK over time on the
// FP64 tensor pipe: lower-triangular 64x64 tiles, one CTA of 4 warps per (tile, chain), every warp
// owns a 32x32 quadrant = 4x4 accumulator fragments of mma.sync.m8n8k4.f64 (SASS: DMMA.8x8x4) fed
// by 8 shared-memory loads per 16 DMMAs.  Deterministic: every output element is accumulated by one
// warp in time order (no atomics, no split over time).
#include "basis_eval.cuh"
#include "sweep_args.cuh"

constexpr int ST = 64;       // tile edge
constexpr int SNT = 128;     // threads: 4 warps, a 32x32 quadrant each
constexpr int LDP = ST + 8;  // row length of the basis chunk [time][basis]: == 8 (mod 16) doubles, so the 4 time rows x 8 basis
                             // columns of one fragment load hit every bank twice (the minimum for 256 bytes)
constexpr int MAXPOS = 48;   // lattice positions per dimension held in the shared sine table

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

template <int D>
__global__ void __launch_bounds__(SNT, 5) suffstats_kernel(const __grid_constant__ SuffArgs a) {
    const DevModel& m = a.m;
    // tile pair (I >= J) from the linear block index
    int I = 0, rem = blockIdx.x;
    while (rem > I) { rem -= I + 1; ++I; }
    const int J = rem;
    const int chain = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nx = m.n_x, M = m.M, npos = a.npos, TK = a.TK;
    const double* traj = a.traj + (size_t)chain * a.traj_stride;

    extern __shared__ __align__(16) double sm[];
    // sine table [D][TK][npos | 1]: position fastest, odd row length (the basis products read it with lanes that share the
    // time step and differ in the position)
    const int nposp = npos | 1;
    double* tab = sm;
    double* phiI = tab + (size_t)D * nposp * TK;        // [TK][LDP]
    double* phiJ = phiI + TK * LDP;                     // [TK][LDP] (aliases phiI on diagonal tiles)
    double* ych = phiJ + TK * LDP;                      // [TK][PGAS_MAX_NX]
    const bool diag = (I == J);
    if (diag) phiJ = phiI;

    // every thread evaluates ONE basis column of each tile block (mi = tid % 64) at every second time step of a chunk: its lattice
    // positions are loop-invariant registers, a basis value costs D shared loads and D multiplies
    const int mi = tid & (ST - 1), tt0 = tid >> 6;
    int pI[D], pJ[D];
    bool vI, vJ;
    {
        const int gi = I * ST + mi, gj = J * ST + mi;
        vI = gi < M; vJ = gj < M;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            pI[d] = vI ? (m.freq[(size_t)gi * D + d] - m.f_start) / m.f_step : 0;
            pJ[d] = vJ ? (m.freq[(size_t)gj * D + d] - m.f_start) / m.f_step : 0;
        }
    }
    // accumulator fragments: quadrant rows wi*32 + 8 fi + lane/4, columns wj*32 + 8 fj + 2 (lane%4) + {0,1}
    const int wi = warp >> 1, wj = warp & 1;
    double c0[4][4], c1[4][4];
#pragma unroll
    for (int fi = 0; fi < 4; ++fi)
#pragma unroll
        for (int fj = 0; fj < 4; ++fj) { c0[fi][fj] = 0.0; c1[fi][fj] = 0.0; }
    double acc0[2] = {0.0, 0.0};   // T0 elements e = tid, tid + SNT of the diagonal tile (mi = e % ST, k = e / ST)
    double acc2 = 0.0;             // T2 element on tile (0,0)
    const int nsteps = m.T - 1;
    const int aoff = (lane & 3) * LDP + wi * 32 + (lane >> 2);      // A fragment element: time row lane%4, basis column lane/4
    const int boff = (lane & 3) * LDP + wj * 32 + (lane >> 2);

    for (int t0 = 0; t0 < nsteps; t0 += TK) {
        __syncthreads();
        // 1. sine tables of the chunk: item (tt, d)
        for (int it = tid; it < TK * D; it += SNT) {
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
                for (int p = 0; p < npos; ++p) {
                    tb[p] = cur;
                    const double n = fma(twoc, cur, -prev);
                    prev = cur; cur = n;
                }
            } else {
                for (int p = 0; p < npos; ++p) tb[p] = 0.0;
            }
        }
        for (int tt = tid; tt < TK; tt += SNT) {     // Y rows of the chunk
            const int t = t0 + tt;
            for (int k = 0; k < nx; ++k) ych[tt * PGAS_MAX_NX + k] = (t < nsteps) ? traj[(size_t)(t + 1) * nx + k] : 0.0;
        }
        __syncthreads();
        // 2. basis values of the two tile blocks (one on diagonal tiles)
        {
            const double* tb = tab + (size_t)tt0 * nposp;
            const int dstride = TK * nposp;
            for (int tt = tt0; tt < TK; tt += 2, tb += 2 * nposp) {
                double vi = vI ? m.norm : 0.0;
#pragma unroll
                for (int d = 0; d < D; ++d) vi *= tb[d * dstride + pI[d]];
                phiI[tt * LDP + mi] = vi;
                if (!diag) {
                    double vj = vJ ? m.norm : 0.0;
#pragma unroll
                    for (int d = 0; d < D; ++d) vj *= tb[d * dstride + pJ[d]];
                    phiJ[tt * LDP + mi] = vj;
                }
            }
        }
        __syncthreads();
        // 3. rank-TK update on the tensor pipe: per 4 time steps 4 A + 4 B fragment loads feed 16 DMMAs
#pragma unroll 1
        for (int kk = 0; kk < TK; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                af[f] = phiI[kk * LDP + aoff + 8 * f];
                bf[f] = phiJ[kk * LDP + boff + 8 * f];
            }
#pragma unroll
            for (int fi = 0; fi < 4; ++fi)
#pragma unroll
                for (int fj = 0; fj < 4; ++fj) dmma884(c0[fi][fj], c1[fi][fj], af[fi], bf[fj]);
        }
        if (diag) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int e = tid + h * SNT;
                if (e < ST * nx) {
                    const int mi = e % ST, k = e / ST;
                    for (int tt = 0; tt < TK; ++tt) acc0[h] = fma(phiI[tt * LDP + mi], ych[tt * PGAS_MAX_NX + k], acc0[h]);
                }
            }
            if (I == 0 && tid >= SNT - nx * nx) {
                const int e = tid - (SNT - nx * nx), r = e / nx, c = e % nx;
                for (int tt = 0; tt < TK; ++tt) acc2 = fma(ych[tt * PGAS_MAX_NX + r], ych[tt * PGAS_MAX_NX + c], acc2);
            }
        }
    }
    // write back: lower tile and its mirror
    double* T1 = a.T1 + (size_t)chain * M * M;
#pragma unroll
    for (int fi = 0; fi < 4; ++fi)
#pragma unroll
        for (int fj = 0; fj < 4; ++fj) {
            const int gi = I * ST + wi * 32 + fi * 8 + (lane >> 2), gj = J * ST + wj * 32 + fj * 8 + 2 * (lane & 3);
            if (gi < M && gj < M) {
                T1[(size_t)gi * M + gj] = c0[fi][fj];
                if (!diag) T1[(size_t)gj * M + gi] = c0[fi][fj];
            }
            if (gi < M && gj + 1 < M) {
                T1[(size_t)gi * M + gj + 1] = c1[fi][fj];
                if (!diag) T1[(size_t)(gj + 1) * M + gi] = c1[fi][fj];
            }
        }
    if (diag) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int e = tid + h * SNT;
            if (e < ST * nx) {
                const int mi = e % ST, k = e / ST, gi = I * ST + mi;
                if (gi < M) a.T0[((size_t)chain * M + gi) * nx + k] = acc0[h];
            }
        }
        if (I == 0 && tid >= SNT - nx * nx) {
            const int e = tid - (SNT - nx * nx);
            a.T2[(size_t)chain * nx * nx + e] = acc2;
        }
    }
}

static size_t suff_smem(int D, int npos, int TK) {
    return sizeof(double) * ((size_t)D * (npos | 1) * TK + 2 * (size_t)TK * LDP + (size_t)TK * PGAS_MAX_NX);
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
    if (ST * m.n_x > 2 * SNT) PGAS_FAIL(-20, "n_x too large for the statistics kernel");
    // time steps per chunk: the largest multiple of 4 (<= 32) that leaves room for five CTAs per SM (45 KB each)
    int TK = 32;
    while (TK > 8 && suff_smem(m.D, npos, TK) > 45 * 1024) TK -= 4;
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
