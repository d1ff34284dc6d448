// suffstats.cu — MNIW sufficient statistics of a trajectory (first half of PGAS.sample_params,
// reference src/PGAS.py:294-303; prior_mniw_calcStatistics, src/BayesianInferrence.py:53-61):
//   T0 = Phi^T Y (M x n_x),  T1 = Phi^T Phi (M x M),  T2 = Y^T Y (n_x x n_x),  T3 = T-1
// with Phi[t] = basis(x_t, u_t), Y[t] = x_{t+1}, t = 0..T-2.  The reference materialises the
// (T-1, M, M) outer products and sums them; here Phi is recomputed chunk by chunk from the
// trajectory (sine tables in shared memory) and never stored, and T1 is a SYRK over time on the
// FP64 tensor pipe: lower-triangular 64x64 tiles, one CTA per (tile, chain).
//
// Two kernels (round 2).  suff_table_kernel evaluates, once per chain and time step, the sines of every lattice position in every
// dimension (one sincospi per dimension, then three interleaved three-term recurrences of stride 3) into a global table
// [chain][D][time][position] that stays in L2.  suffstats_kernel is warp-specialised: one PRODUCER warp copies the chunk's slice of
// that table into a double-buffered shared-memory stage with cp.async.bulk (the bulk-copy engine; completion on an mbarrier's
// transaction count — no arithmetic, no registers), four CONSUMER warps own a 32x32 quadrant each = 4x4 accumulator fragments of
// mma.sync.m8n8k4.f64 (SASS: DMMA.8x8x4) and form their fragments ON THE FLY from the stage (D shared loads and D - 1 multiplies per
// fragment element; the stage offsets of a thread's eight fragment columns are loop-invariant registers): Phi itself is never
// stored.  Tiles are walked in LATTICE order of the basis functions (DevModel::lat_perm), so that the eight columns of a fragment
// share their leading positions (broadcast loads) and are neighbours in the last dimension; the write-back permutes to the
// reference's order.  Why this shape: any scalar FP64 operation issued next to the DMMAs waits its turn behind whole batches of
// them on the one FP64 pipe (one slot in four, 16 cycles each) — with the tables built inside the kernel, first by all warps
// between __syncthreads, then by a producer warp, the tensor sub-pipe stayed 52-54 % busy and the consumers waited for the tables
// (profiles/r02_tail_kernels_summary.md).  On diagonal tiles the upper quadrant is the mirror of the lower one and is not computed.
// Deterministic: every output element is accumulated by one warp in time order (no atomics, no split over time).
#include "basis_eval.cuh"
#include "sweep_args.cuh"

constexpr int ST = 64;         // tile edge
constexpr int S_CONS = 128;    // consumer threads: 4 warps, a 32x32 quadrant each
constexpr int S_PROD = 32;     // producer threads (one warp: with 160 threads ptxas grants 128 registers at three CTAs per SM, with 192 only 96)
constexpr int SNT = S_CONS + S_PROD;
constexpr int MAXPOS = 48;     // lattice positions per dimension held in the shared sine table
constexpr int BAR_EMPTY = 3;   // named barriers empty[2] (consumers arrive, the producer waits); full[2] are mbarriers

struct SuffArgs {
    DevModel m;
    int n_chains, ntile, npos, TK;   // TK = time steps per chunk (multiple of 4)
    int TP;                          // time rows of the sine table: chunks * TK (rows past the trajectory are zero)
    double* table;                   // (n_chains, D, TP, suff_row(npos))
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

__device__ __forceinline__ uint32_t sf_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sf_mbar_init(uint32_t mbar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count)); }
__device__ __forceinline__ void sf_mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sf_mbar_arrive(uint32_t mbar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory"); }
__device__ __forceinline__ void sf_mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void sf_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

// One thread per (chain, time row): the sines of all lattice positions in all D dimensions, three interleaved recurrences of stride 3
// per dimension (s[p+3] = 2cos(3a) s[p] - s[p-3]); the normalisation rides on dimension 0 (the recurrence is linear); slot npos of
// every row is zero (tile columns past M point there); rows past the trajectory are zero.
template <int D>
__global__ void __launch_bounds__(128) suff_table_kernel(const __grid_constant__ SuffArgs a) {
    const DevModel& m = a.m;
    const int t = blockIdx.x * 128 + threadIdx.x, chain = blockIdx.y;
    if (t >= a.TP) return;
    const int nx = m.n_x, npos = a.npos, nposp = suff_row(npos), nsteps = m.T - 1;
    const bool live = t < nsteps;
    const double* traj = a.traj + (size_t)chain * a.traj_stride;
    double x[PGAS_MAX_NX], u[PGAS_MAX_NU], z[PGAS_MAX_D];
#pragma unroll
    for (int k = 0; k < PGAS_MAX_NX; ++k) x[k] = (live && k < nx) ? traj[(size_t)t * nx + k] : 0.0;
#pragma unroll
    for (int k = 0; k < PGAS_MAX_NU; ++k) u[k] = (live && k < m.n_u) ? m.inputs[(size_t)t * m.n_u + k] : 0.0;   // x_t pairs with u_t (src/PGAS.py:294-296)
    if (m.map_kind == PGAS_MAP_AFFINE) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            double v = m.bz[d];
#pragma unroll
            for (int k = 0; k < PGAS_MAX_NX; ++k) v = (k < nx) ? fma(m.Az[d][k], x[k], v) : v;
#pragma unroll
            for (int k = 0; k < PGAS_MAX_NU; ++k) v = (k < m.n_u) ? fma(m.Az[d][nx + k], u[k], v) : v;
            z[d] = v;
        }
    } else {
        gp_map_any(m, x, u, z);                     // slip angles / expression program
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const double tn = (z[d] - m.center[d] + m.L[d]) * m.inv2L[d];
        double s0, sm1, tc;
        sine_seed(tn, m.f_start, m.f_step, s0, sm1, tc);
        const double sc = (d == 0) ? m.norm : 1.0;
        s0 = live ? s0 * sc : 0.0;                  // rows past the trajectory are zero (their map may not even be finite)
        sm1 = live ? sm1 * sc : 0.0;
        tc = live ? tc : 0.0;
        const double s1 = fma(tc, s0, -sm1), sm2 = fma(tc, sm1, -s0);
        const double s2 = fma(tc, s1, -s0), sm3 = fma(tc, sm2, -sm1);
        const double c3 = tc * fma(tc, tc, -3.0);   // 2 cos(3a) from 2 cos(a)
        double ca = s0, cb = s1, cc = s2, pa = sm3, pb = sm2, pc = sm1;
        double* tb = a.table + (((size_t)chain * D + d) * a.TP + t) * nposp;
        for (int p = 0; p < npos; p += 3) {
            tb[p] = ca;
            if (p + 1 < npos) tb[p + 1] = cb;
            if (p + 2 < npos) tb[p + 2] = cc;
            const double na = fma(c3, ca, -pa), nb = fma(c3, cb, -pb), nc = fma(c3, cc, -pc);
            pa = ca; pb = cb; pc = cc;
            ca = na; cb = nb; cc = nc;
        }
        tb[npos] = 0.0;
    }
}

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

    __shared__ __align__(8) unsigned long long full_bar[2];
    if (tid == 0) {
        sf_mbar_init(sf_smem(&full_bar[0]), 32);
        sf_mbar_init(sf_smem(&full_bar[1]), 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= S_CONS / 32) {
        // ------------------------------------------------------------------------------------------------ producer
        // lane 0 starts the D bulk copies of the chunk's table slices (TK * nposp doubles each, 16-byte multiples at 16-byte aligned
        // addresses since TK is even); every lane stores one Y row and arrives: the stage is full when 32 arrivals and the copies'
        // bytes are in
        const int pt = tid - S_CONS;
        const uint32_t slice_bytes = (uint32_t)(dstride * sizeof(double));
        const double* gtab = a.table + (size_t)chain * D * a.TP * nposp;
        for (int c = 0; c < nchunks; ++c) {
            const int t = c * TK + pt;
            double* tab = sm + (size_t)(c & 1) * stage_sz;
            double* ych = tab + (size_t)D * dstride;
            const uint32_t mb = sf_smem(&full_bar[c & 1]);
            double y[PGAS_MAX_NX];
#pragma unroll
            for (int k = 0; k < PGAS_MAX_NX; ++k) y[k] = (pt < TK && t < nsteps && k < nx) ? traj[(size_t)(t + 1) * nx + k] : 0.0;
            if (c >= 2) bar_sync(BAR_EMPTY + (c & 1), SNT);   // the consumers have released this stage (chunk c - 2)
            if (pt == 0) {
                sf_mbar_expect_tx(mb, slice_bytes * D);
#pragma unroll
                for (int d = 0; d < D; ++d) sf_bulk_g2s(sf_smem(tab + (size_t)d * dstride), gtab + ((size_t)d * a.TP + (size_t)c * TK) * nposp, slice_bytes, mb);
            }
            if (pt < TK)
                for (int k = 0; k < nx; ++k) ych[pt * PGAS_MAX_NX + k] = y[k];
            if (pt != 0) sf_mbar_arrive(mb);
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
            const int ci = I * ST + wi * 32 + 8 * f + q, cj = J * ST + wj * 32 + 8 * f + q;     // tile columns, lattice order
            const int gi = (ci < M) ? m.lat_perm[ci] : -1, gj = (cj < M) ? m.lat_perm[cj] : -1;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const int pi = (gi >= 0) ? (m.freq[(size_t)gi * D + d] - m.f_start) / m.f_step : (d == 0 ? npos : 0);
                const int pj = (gj >= 0) ? (m.freq[(size_t)gj * D + d] - m.f_start) / m.f_step : (d == 0 ? npos : 0);
                offA[f][d] = d * dstride + r * nposp + pi;
                offB[f][d] = d * dstride + r * nposp + pj;
            }
        }
        const int cc0 = I * ST + (tid % ST);              // T0 column of this thread (diagonal tiles)
        const int g0 = (cc0 < M) ? m.lat_perm[cc0] : -1;
#pragma unroll
        for (int d = 0; d < D; ++d) off0[d] = d * dstride + ((g0 >= 0) ? (m.freq[(size_t)g0 * D + d] - m.f_start) / m.f_step : (d == 0 ? npos : 0));
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
        sf_mbar_wait(sf_smem(&full_bar[c & 1]), (uint32_t)((c >> 1) & 1));
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
    // write back: lower tile and its mirror, from lattice order to the reference's order of the basis functions
    double* T1 = a.T1 + (size_t)chain * M * M;
    if (work) {
        const bool mirror = !diag || wi != wj;
        int ri[4], rj[4][2];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            const int ci = I * ST + wi * 32 + f * 8 + (lane >> 2), cj = J * ST + wj * 32 + f * 8 + 2 * (lane & 3);
            ri[f] = (ci < M) ? m.lat_perm[ci] : -1;
            rj[f][0] = (cj < M) ? m.lat_perm[cj] : -1;
            rj[f][1] = (cj + 1 < M) ? m.lat_perm[cj + 1] : -1;
        }
#pragma unroll
        for (int fi = 0; fi < 4; ++fi)
#pragma unroll
            for (int fj = 0; fj < 4; ++fj) {
                const int gi = ri[fi];
                if (gi >= 0 && rj[fj][0] >= 0) {
                    T1[(size_t)gi * M + rj[fj][0]] = c0[fi][fj];
                    if (mirror) T1[(size_t)rj[fj][0] * M + gi] = c0[fi][fj];
                }
                if (gi >= 0 && rj[fj][1] >= 0) {
                    T1[(size_t)gi * M + rj[fj][1]] = c1[fi][fj];
                    if (mirror) T1[(size_t)rj[fj][1] * M + gi] = c1[fi][fj];
                }
            }
    }
    if (diag) {
        if (tid < ST * nx) {
            const int k = tid / ST, cc0 = I * ST + (tid % ST);
            if (cc0 < M) a.T0[((size_t)chain * M + m.lat_perm[cc0]) * nx + k] = acc0;
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
    a.TP = (m.T - 1 + TK - 1) / TK * TK;
    if (a.TP <= 0) PGAS_FAIL(-2, "the statistics need at least two time steps");
    // the sine table lives for the two launches only: stream-ordered allocation
    const size_t tbytes = sizeof(double) * (size_t)n_chains * m.D * a.TP * suff_row(npos);
    {   // keep the pool's memory across synchronisations (default: released to the OS at the next sync, ~0.3 ms per call to get it back)
        static thread_local int pool_dev = -1;
        int dev = 0;
        PGAS_CUDA(cudaGetDevice(&dev));
        if (pool_dev != dev) {
            cudaMemPool_t pool;
            PGAS_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
            unsigned long long keep = ~0ull;
            PGAS_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
            pool_dev = dev;
        }
    }
    PGAS_CUDA(cudaMallocAsync((void**)&a.table, tbytes, st));
    {
        dim3 g((unsigned)((a.TP + 127) / 128), (unsigned)n_chains, 1);
        if (m.D == 1) suff_table_kernel<1><<<g, 128, 0, st>>>(a);
        else if (m.D == 2) suff_table_kernel<2><<<g, 128, 0, st>>>(a);
        else suff_table_kernel<3><<<g, 128, 0, st>>>(a);
    }
    auto kern = (m.D == 1) ? suffstats_kernel<1> : (m.D == 2) ? suffstats_kernel<2> : suffstats_kernel<3>;
    PGAS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(a.ntile * (a.ntile + 1) / 2), (unsigned)n_chains, 1);
    kern<<<grid, SNT, smem, st>>>(a);
    __atomic_add_fetch(&g_pgas_launches, 2, __ATOMIC_RELAXED);
    const cudaError_t le = cudaGetLastError();
    PGAS_CUDA(cudaFreeAsync(a.table, st));
    if (le != cudaSuccess) PGAS_FAIL((int)le, "statistics kernels: %s", cudaGetErrorString(le));
    return 0;
}


extern "C" int pgas_suffstats_f64(const pgas_model* model, const double* traj, int32_t n_chains, double* T0_out, double* T1_out,
                                  double* T2_out, void* stream) {
    if (!model || !traj || !T0_out || !T1_out || !T2_out) PGAS_FAIL(-1, "pgas_suffstats_f64: null argument");
    if (n_chains < 1) PGAS_FAIL(-2, "n_chains must be >= 1");
    return pgas_launch_suffstats(model->dev, traj, (long long)model->dev.T * model->dev.n_x, n_chains, T0_out, T1_out, T2_out,
                                 (cudaStream_t)stream);
}
