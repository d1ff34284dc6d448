// suffstats.cu — MNIW sufficient statistics of a trajectory (first half of PGAS.sample_params,
// reference src/PGAS.py:294-303; prior_mniw_calcStatistics, src/BayesianInferrence.py:53-61):
//   T0 = Phi^T Y (M x n_x),  T1 = Phi^T Phi (M x M),  T2 = Y^T Y (n_x x n_x),  T3 = T-1
// with Phi[t] = basis(x_t, u_t), Y[t] = x_{t+1}, t = 0..T-2.  The reference materialises the
// (T-1, M, M) outer products and sums them; here Phi is recomputed chunk by chunk from the
// trajectory (sine tables in shared memory) and never stored, and only the lower-triangular
// 64x64 tiles of T1 are computed (the mirror is written at the end).  Deterministic: every
// output element is accumulated by one thread in time order.
#include "basis_eval.cuh"
#include "sweep_args.cuh"

constexpr int ST = 64;       // tile edge
constexpr int TK = 32;       // time steps per chunk
constexpr int SNT = 256;     // threads
constexpr int MAXPOS = 48;   // lattice positions per dimension held in the shared sine table

struct SuffArgs {
    DevModel m;
    int n_chains, ntile, npos;
    const double* traj;      // (n_chains, T, n_x), chain stride traj_stride elements
    long long traj_stride;
    double* T0;              // (n_chains, M, n_x)
    double* T1;              // (n_chains, M, M)
    double* T2;              // (n_chains, n_x, n_x)
};

__global__ void __launch_bounds__(SNT) suffstats_kernel(const __grid_constant__ SuffArgs a) {
    const DevModel& m = a.m;
    // tile pair (I >= J) from the linear block index
    int I = 0, rem = blockIdx.x;
    while (rem > I) { rem -= I + 1; ++I; }
    const int J = rem;
    const int chain = blockIdx.y;
    const int tid = threadIdx.x;
    const int nx = m.n_x, D = m.D, M = m.M, npos = a.npos;
    const double* traj = a.traj + (size_t)chain * a.traj_stride;

    extern __shared__ __align__(16) double sm[];
    // sine table [D][TK][npos | 1]: position fastest, odd row length — the basis products below read it with one bank per
    // lattice position (lanes of a warp share the time step and differ in the position), the table writes stride an odd length
    const int nposp = npos | 1;
    double* tab = sm;
    double* phiI = tab + (size_t)D * nposp * TK;        // [TK][ST]
    double* phiJ = phiI + TK * ST;                      // [TK][ST]
    double* ych = phiJ + TK * ST;                       // [TK][PGAS_MAX_NX]
    int* posI = reinterpret_cast<int*>(ych + TK * PGAS_MAX_NX);   // [ST][D]
    int* posJ = posI + ST * PGAS_MAX_D;

    for (int e = tid; e < ST * D; e += SNT) {
        const int mi = e / D, d = e % D;
        const int gi = I * ST + mi, gj = J * ST + mi;
        posI[mi * PGAS_MAX_D + d] = (gi < M) ? (m.freq[(size_t)gi * D + d] - m.f_start) / m.f_step : -1;
        posJ[mi * PGAS_MAX_D + d] = (gj < M) ? (m.freq[(size_t)gj * D + d] - m.f_start) / m.f_step : -1;
    }
    const int ty = tid / 16, tx = tid % 16;
    double acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
    double acc0 = 0.0;           // T0 element (mi = tid % ST, k = tid / ST) on diagonal tiles
    double acc2 = 0.0;           // T2 element on tile (0,0)
    const int nsteps = m.T - 1;

    for (int t0 = 0; t0 < nsteps; t0 += TK) {
        __syncthreads();
        // 1. sine tables of the chunk: thread (tt, d)
        if (tid < TK * D) {
            const int tt = tid % TK, d = tid / TK, t = t0 + tt;
            double* tb = tab + ((size_t)d * TK + tt) * nposp;
            if (t < nsteps) {
                double x[PGAS_MAX_NX], u[PGAS_MAX_NU], z = 0.0;
                for (int k = 0; k < nx; ++k) x[k] = traj[(size_t)t * nx + k];
                for (int k = 0; k < m.n_u; ++k) u[k] = m.inputs[(size_t)t * m.n_u + k];     // x_t pairs with u_t (src/PGAS.py:294-296)
                if (m.map_kind == PGAS_MAP_VEHICLE_SLIP) {
                    z = (d == 0) ? u[0] - atan((x[1] + x[0] * m.slip_lf) / u[1]) : -atan((x[1] - x[0] * m.slip_lr) / u[1]);
                } else {
                    z = m.bz[d];
                    for (int k = 0; k < nx; ++k) z = fma(m.Az[d][k], x[k], z);
                    for (int k = 0; k < m.n_u; ++k) z = fma(m.Az[d][nx + k], u[k], z);
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
        if (tid >= SNT - TK) {                      // Y rows of the chunk
            const int tt = tid - (SNT - TK), t = t0 + tt;
            for (int k = 0; k < nx; ++k) ych[tt * PGAS_MAX_NX + k] = (t < nsteps) ? traj[(size_t)(t + 1) * nx + k] : 0.0;
        }
        __syncthreads();
        // 2. basis values of the two tile blocks
        for (int e = tid; e < TK * ST; e += SNT) {
            const int mi = e % ST, tt = e / ST;
            double vi = m.norm, vj = m.norm;
            for (int d = 0; d < D; ++d) {
                const int pi = posI[mi * PGAS_MAX_D + d], pj = posJ[mi * PGAS_MAX_D + d];
                vi = (pi >= 0) ? vi * tab[((size_t)d * TK + tt) * nposp + pi] : 0.0;
                vj = (pj >= 0) ? vj * tab[((size_t)d * TK + tt) * nposp + pj] : 0.0;
            }
            phiI[tt * ST + mi] = vi;
            phiJ[tt * ST + mi] = vj;
        }
        __syncthreads();
        // 3. rank-TK update of the 4x4 register tile
#pragma unroll 4
        for (int tt = 0; tt < TK; ++tt) {
            double av[4], bv[4];
            const double2 a0 = *reinterpret_cast<const double2*>(&phiI[tt * ST + ty * 4]);
            const double2 a1 = *reinterpret_cast<const double2*>(&phiI[tt * ST + ty * 4 + 2]);
            const double2 b0 = *reinterpret_cast<const double2*>(&phiJ[tt * ST + tx * 4]);
            const double2 b1 = *reinterpret_cast<const double2*>(&phiJ[tt * ST + tx * 4 + 2]);
            av[0] = a0.x; av[1] = a0.y; av[2] = a1.x; av[3] = a1.y;
            bv[0] = b0.x; bv[1] = b0.y; bv[2] = b1.x; bv[3] = b1.y;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fma(av[r], bv[c], acc[r][c]);
        }
        if (I == J) {
            if (tid < ST * nx) {
                const int mi = tid % ST, k = tid / ST;
                for (int tt = 0; tt < TK; ++tt) acc0 = fma(phiI[tt * ST + mi], ych[tt * PGAS_MAX_NX + k], acc0);
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
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int gi = I * ST + ty * 4 + r, gj = J * ST + tx * 4 + c;
            if (gi < M && gj < M) {
                T1[(size_t)gi * M + gj] = acc[r][c];
                if (I != J) T1[(size_t)gj * M + gi] = acc[r][c];
            }
        }
    if (I == J) {
        if (tid < ST * nx) {
            const int mi = tid % ST, k = tid / ST, gi = I * ST + mi;
            if (gi < M) a.T0[((size_t)chain * M + gi) * nx + k] = acc0;
        }
        if (I == 0 && tid >= SNT - nx * nx) {
            const int e = tid - (SNT - nx * nx);
            a.T2[(size_t)chain * nx * nx + e] = acc2;
        }
    }
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
    if (ST * m.n_x > SNT - 16) PGAS_FAIL(-20, "n_x too large for the statistics kernel");
    const size_t smem = sizeof(double) * ((size_t)m.D * (npos | 1) * TK + 2 * TK * ST + TK * PGAS_MAX_NX) + sizeof(int) * 2 * ST * PGAS_MAX_D;
    auto kern = suffstats_kernel;
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
