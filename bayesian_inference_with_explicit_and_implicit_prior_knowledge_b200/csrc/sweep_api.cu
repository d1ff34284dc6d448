// sweep_api.cu — C-ABI entry points around the persistent sweep: full sweep, single step,
// final categorical pick + reconstruct_trajectory, standalone systematic resampling, basis
// evaluation, and the Philox variates export used by the parity tests.
#include <algorithm>
#include <vector>
#include "basis_eval.cuh"
#include "sweep_args.cuh"

constexpr int BT = 256;

// ------------------------------------------------------------------ pick + backward trace
// idx ~ Cat(softmax(logw_T-1)) via searchsorted(cumsum(w), u) (src/PGAS.py:224-225), then
// reconstruct_trajectory (src/Filtering.py:40-55).  One CTA per chain / set.
__global__ void __launch_bounds__(BT) pick_and_trace_kernel(const double* __restrict__ logw_last, const double* __restrict__ state_trace,
                                                            const int* __restrict__ anc_trace, const int* __restrict__ idx_in,
                                                            int T, int N, int n, int rng_mode, unsigned long long seed,
                                                            unsigned chain_base, unsigned iteration, const double* __restrict__ U,
                                                            int var_rows, int* __restrict__ final_idx, double* __restrict__ traj_out,
                                                            long long traj_stride) {
    const int chain = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ double red[BT / 32];
    __shared__ double wt[2][BT / 32];
    __shared__ int s_idx;
    __shared__ int s_cnt;
    if (tid == 0) s_cnt = 0;
    int idx;
    if (idx_in) {
        idx = idx_in[chain];
    } else {
        const double* lw = logw_last + (size_t)chain * N;
        double u;
        if (rng_mode == 1) u = U[(size_t)chain * var_rows * 2];
        else { double ub; philox_uniform2(seed, PURPOSE_STEP_U, chain_base + chain, iteration, 0u, 0u, u, ub); }
        double mx = -INFINITY;
        for (int i = tid; i < N; i += BT) mx = fmax(mx, lw[i]);
        mx = warp_max(mx);
        if (lane == 0) red[warp] = mx;
        __syncthreads();
        mx = red[0];
        for (int w = 1; w < BT / 32; ++w) mx = fmax(mx, red[w]);
        __syncthreads();
        double sm = 0.0;
        for (int i = tid; i < N; i += BT) sm += exp(lw[i] - mx);
        sm = warp_sum(sm);
        if (lane == 0) red[warp] = sm;
        __syncthreads();
        sm = 0.0;
        for (int w = 0; w < BT / 32; ++w) sm += red[w];
        const double rs = 1.0 / sm;
        double carry = 0.0;
        int cnt = 0, q = 0;
        for (int i0 = 0; i0 < N; i0 += BT, ++q) {
            const int i = i0 + tid;
            const double e = (i < N) ? exp(lw[i] - mx) : 0.0;
            const double s = warp_scan_incl(e, lane);
            if (lane == 31) wt[q & 1][warp] = s;
            __syncthreads();
            double run = carry, my = 0.0;
            for (int w = 0; w < BT / 32; ++w) { if (w == warp) my = run; run += wt[q & 1][w]; }
            carry = run;
            if (i < N && (my + s) * rs < u) ++cnt;
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0 && cnt) atomicAdd(&s_cnt, cnt);
        __syncthreads();
        idx = s_cnt;
    }
    if (tid == 0) {
        if (final_idx) final_idx[chain] = idx;
        s_idx = idx;
    }
    __syncthreads();
    // backward walk: T - 1 DEPENDENT loads a_{t-1} = anc[t-1][a_t], each an HBM miss (~1 us) when walked alone.  Warp 0
    // walks; the other warps run ahead of it and pull whole ancestor rows (N ints) towards the walker — rows t-L2W..
    // into L2, rows t-L1W.. into this SM's L1 — so that the dependent load is an L1 / L2 hit.  The trajectory rows
    // themselves are not on the dependent chain: the walker only records the path a_t; all threads gather the rows afterwards
    // (a load-then-store inside the walk would stall every hop on an HBM miss).
    constexpr int L1W = 8, L2W = 96;                       // look-ahead in rows (L1: 8 x 16 KB at N = 4096)
    __shared__ volatile int s_t;
    if (tid == 0) s_t = T - 1;
    __syncthreads();
    const int* an = anc_trace + (size_t)chain * (T - 1) * N;
    extern __shared__ int s_path[];                        // a_t for all t: the state rows are gathered afterwards by all threads
    if (warp == 0) {
        int a = min(max(s_idx, 0), N - 1);
        for (int t = T - 1; t >= 0; --t) {
            if (lane == 0) s_path[t] = a;
            if (t > 0) {
                int nx = 0;
                if (lane == 0) { nx = an[(size_t)(t - 1) * N + a]; s_t = t - 1; }
                nx = __shfl_sync(0xffffffffu, nx, 0);
                a = min(max(nx, 0), N - 1);
            }
        }
    } else {
        const int lines = (N + 31) / 32;                   // 128-byte lines per ancestor row
        const bool near = warp < 4;                        // warps 1-3: L1 window, warps 4-7: L2 window
        const int nth = near ? 96 : BT - 128, me = near ? tid - 32 : tid - 128, win = near ? L1W : L2W;
        for (int r = T - 2; r >= 0; --r) {
            while (r < s_t - win) __nanosleep(64);         // stay `win` rows ahead of the walker, not more
            if (r >= s_t) continue;                        // the walker is already past this row
            const int* row = an + (size_t)r * N;
            for (int l = me; l < lines; l += nth) {
                if (near) asm volatile("prefetch.global.L1 [%0];" ::"l"(row + l * 32));
                else asm volatile("prefetch.global.L2 [%0];" ::"l"(row + l * 32));
            }
        }
    }
    __syncthreads();
    // the trajectory rows are independent loads once the path is known (src/Filtering.py:46-53)
    const double* st = state_trace + (size_t)chain * T * N * n;
    double* tr = traj_out + (size_t)chain * traj_stride;
    for (int e = tid; e < T * n; e += BT) {
        const int t = e / n, k = e - t * n;
        tr[e] = st[((size_t)t * N + s_path[t]) * n + k];
    }
}

// ------------------------------------------------------------------ standalone resampling
// systematic_SISR (src/Filtering.py:6-37): clip, normalise (uniform if the sum is not > 0),
// cumsum, clip, searchsorted(side=left), clip.  One CTA per weight set, CDF in shared memory.
__global__ void __launch_bounds__(BT) resample_kernel(const double* __restrict__ w_in, int N, const double* __restrict__ u_in,
                                                      int* __restrict__ idx_out) {
    extern __shared__ double cdf[];
    __shared__ double red[BT / 32];
    __shared__ double wt[2][BT / 32];
    const int set = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* w = w_in + (size_t)set * N;
    const double u = u_in[set];
    double sm = 0.0;
    for (int i = tid; i < N; i += BT) sm += (w[i] != w[i]) ? w[i] : fmax(w[i], 0.0);      // clip keeps NaN (jnp.clip)
    sm = warp_sum(sm);
    if (lane == 0) red[warp] = sm;
    __syncthreads();
    sm = 0.0;
    for (int k = 0; k < BT / 32; ++k) sm += red[k];
    const bool ok = sm > 0.0;                               // NaN -> false -> uniform weights (:25)
    const double dN = (double)N;
    double carry = 0.0;
    int q = 0;
    for (int i0 = 0; i0 < N; i0 += BT, ++q) {
        const int i = i0 + tid;
        double e = 0.0;
        if (i < N) e = ok ? __ddiv_rn(fmax(w[i], 0.0), sm) : __ddiv_rn(1.0, dN);
        const double s = warp_scan_incl(e, lane);
        if (lane == 31) wt[q & 1][warp] = s;
        __syncthreads();
        double run = carry, my = 0.0;
        for (int k = 0; k < BT / 32; ++k) { if (k == warp) my = run; run += wt[q & 1][k]; }
        carry = run;
        if (i < N) cdf[i] = fmin(fmax(my + s, 0.0), 1.0);
    }
    __syncthreads();
    for (int j = tid; j < N; j += BT) {
        const double uj = __ddiv_rn(__dadd_rn(u, (double)j), dN);
        int lo = 0, hi = N;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cdf[mid] < uj) lo = mid + 1; else hi = mid;
        }
        idx_out[(size_t)set * N + j] = min(lo, N - 1);
    }
}

// ------------------------------------------------------------------ basis evaluation
// phi (n, M) in the reference's basis order, using the same sine recurrences as the sweep.
__global__ void __launch_bounds__(128) hgp_eval_kernel(const __grid_constant__ DevModel m, const double* __restrict__ states, const double* __restrict__ inputs,
                                                       int input_stride, int n, int npos, double* __restrict__ phi) {
    // one warp per sample: lanes build the per-dimension sine tables in shared memory, then
    // stride over the M basis functions.
    extern __shared__ double tab[];                         // [warps][D][npos]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    double* mytab = tab + (size_t)warp * m.D * npos;
    for (int smp = blockIdx.x * nwarp + warp; smp < n; smp += gridDim.x * nwarp) {
        double x[PGAS_MAX_NX], u[PGAS_MAX_NU], z[PGAS_MAX_D];
        for (int k = 0; k < m.n_x; ++k) x[k] = states[(size_t)smp * m.n_x + k];
        for (int k = 0; k < m.n_u; ++k) u[k] = inputs[(size_t)smp * input_stride + k];
        for (int k = m.n_x; k < PGAS_MAX_NX; ++k) x[k] = 0.0;
        for (int k = m.n_u; k < PGAS_MAX_NU; ++k) u[k] = 0.0;
        gp_map_any(m, x, u, z);
        __syncwarp();
        if (lane < m.D) {
            const int d = lane;
            const double t = (z[d] - m.center[d] + m.L[d]) * m.inv2L[d];
            double cur, prev, twoc;
            sine_seed(t, m.f_start, m.f_step, cur, prev, twoc);
            double* tb = mytab + (size_t)d * npos;
            for (int p = 0; p < npos; ++p) {
                tb[p] = cur;
                const double nx = fma(twoc, cur, -prev);
                prev = cur; cur = nx;
            }
        }
        __syncwarp();
        for (int mm = lane; mm < m.M; mm += 32) {
            double v = m.norm;
            for (int d = 0; d < m.D; ++d) v *= mytab[(size_t)d * npos + (m.freq[(size_t)mm * m.D + d] - m.f_start) / m.f_step];
            phi[(size_t)smp * m.M + mm] = v;
        }
    }
}

// ------------------------------------------------------------------ Philox variates export
__global__ void philox_variates_kernel(unsigned long long seed, unsigned chain_base, unsigned iteration, int n_chains, int T, int N,
                                       int n_x, double* __restrict__ Z, double* __restrict__ U) {
    const size_t total = (size_t)n_chains * T * N;
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(g % N);
        const int t = (int)((g / N) % T);
        const int c = (int)(g / ((size_t)N * T));
        for (int k = 0; k < n_x; k += 2) {
            double za, zb;
            philox_normal2(seed, PURPOSE_STATE, chain_base + c, iteration, (unsigned)t | ((unsigned)(k >> 1) << 28), (unsigned)i, za, zb);
            Z[g * n_x + k] = za;
            if (k + 1 < n_x) Z[g * n_x + k + 1] = zb;
        }
        if (i == 0) {
            double ua, ub;
            philox_uniform2(seed, PURPOSE_STEP_U, chain_base + c, iteration, (unsigned)t, 0u, ua, ub);
            U[((size_t)c * T + t) * 2] = ua;
            U[((size_t)c * T + t) * 2 + 1] = ub;
        }
    }
}

int pgas_launch_pick_and_trace(const double* logw_last, const double* state_trace, const int* anc_trace, const int* idx_in,
                               int n_sets, int T, int N, int n, const pgas_rng* rng, int var_rows, int* final_idx, double* traj_out,
                               long long traj_stride, cudaStream_t st) {
    if ((size_t)T * sizeof(int) > 200 * 1024) PGAS_FAIL(-20, "trajectory of %d steps is too long for the backward-trace kernel", T);
    PGAS_CUDA(cudaFuncSetAttribute(pick_and_trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(T * sizeof(int))));
    pick_and_trace_kernel<<<n_sets, BT, (size_t)T * sizeof(int), st>>>(logw_last, state_trace, anc_trace, idx_in, T, N, n, rng ? rng->mode : 1,
                                                 rng ? rng->seed : 0ull, rng ? rng->chain_base : 0u, rng ? rng->iteration : 0u,
                                                 rng ? rng->U : nullptr, var_rows, final_idx, traj_out, traj_stride);
    PGAS_KERNEL_CHECK();
    return 0;
}

// ================================================================== C ABI
static int fill_rng(SweepArgs& a, const pgas_rng* rng) {
    if (!rng) PGAS_FAIL(-1, "rng must not be null");
    a.rng_mode = rng->mode;
    a.seed = rng->seed;
    a.chain_base = rng->chain_base;
    a.iteration = rng->iteration;
    a.Z = rng->Z;
    a.U = rng->U;
    if (rng->mode == 1 && (!rng->Z || !rng->U)) PGAS_FAIL(-1, "injected rng mode needs Z and U");
    if (rng->mode != 0 && rng->mode != 1) PGAS_FAIL(-2, "unknown rng mode %d", rng->mode);
    return 0;
}

static int check_cluster(int C, int N) {
    if (C < 1 || C > 16) PGAS_FAIL(-2, "cluster_size must be 0 (auto) or 1..16 (got %d)", C);
    if (N < 2) PGAS_FAIL(-2, "need at least 2 particles (N=%d)", N);
    return 0;
}

// [logw_last fallback (N * n_chains doubles) | split-form buffers (sweep.cu: pgas_sweep_split_workspace)]
extern "C" size_t pgas_csmc_sweep_workspace_bytes(const pgas_model* model, int32_t N, int32_t n_chains) {
    if (!model || N < 1 || n_chains < 1) return 0;
    return (((size_t)N * n_chains * sizeof(double) + 255) & ~(size_t)255) + pgas_sweep_split_workspace(model->dev, N, n_chains) + 256;
}

extern "C" int pgas_csmc_sweep_f64(const pgas_model* model, int32_t N, int32_t n_chains, const double* ref_traj, const double* Theta,
                                   const double* Sigma, const pgas_rng* rng, double* state_trace, int32_t* anc_trace,
                                   double* logw_last, int32_t* final_idx, double* traj_out, int32_t cluster_size, void* workspace,
                                   size_t workspace_bytes, void* stream) {
    if (!model || !ref_traj || !Theta || !Sigma || !state_trace || !anc_trace) PGAS_FAIL(-1, "pgas_csmc_sweep_f64: null argument");
    if (n_chains < 1) PGAS_FAIL(-2, "n_chains must be >= 1");
    const DevModel& m = model->dev;
    cudaStream_t st = (cudaStream_t)stream;
    SweepArgs a;
    memset(&a, 0, sizeof(a));
    a.m = m;
    a.N = N; a.n_chains = n_chains;
    a.C = pgas_choose_cluster(m, N, n_chains, cluster_size);
    if (int rc = check_cluster(a.C, N)) return rc;
    a.P = (N + a.C - 1) / a.C;
    a.t_begin = 1; a.t_end = m.T;
    a.row_off = 0; a.anc_shift = 0;
    a.ref_rows = m.T; a.trace_rows = m.T; a.anc_rows = m.T - 1; a.var_rows = m.T;
    a.ref = ref_traj; a.ref_stride = (long long)m.T * m.n_x; a.Theta = Theta; a.Sigma = Sigma;
    a.state_trace = state_trace; a.anc_trace = anc_trace;
    double* lw = logw_last;
    const size_t lw_bytes = ((size_t)N * n_chains * sizeof(double) + 255) & ~(size_t)255;
    if (!lw) {
        if (!workspace || workspace_bytes < sizeof(double) * (size_t)N * n_chains)
            PGAS_FAIL(-5, "logw_last is NULL and the workspace is smaller than N*n_chains doubles");
        lw = (double*)workspace;
    }
    a.logw_last = lw;
    if (workspace && workspace_bytes > lw_bytes) {          // the rest feeds the split form (state kernel ahead of the resampling kernel)
        a.ws = (char*)workspace + lw_bytes;
        a.ws_bytes = workspace_bytes - lw_bytes;
    }
    if (int rc = fill_rng(a, rng)) return rc;
    if (int rc = pgas_launch_sweep(a, st)) return rc;
    if (traj_out || final_idx) {
        if (!traj_out) PGAS_FAIL(-1, "final_idx requested without traj_out");
        if (int rc = pgas_launch_pick_and_trace(lw, state_trace, anc_trace, nullptr, n_chains, m.T, N, m.n_x, rng, a.var_rows, final_idx,
                                                traj_out, (long long)m.T * m.n_x, st))
            return rc;
    }
    return 0;
}

extern "C" int pgas_csmc_step_f64(const pgas_model* model, int32_t N, int32_t t, const double* logw, const double* state,
                                  const double* Theta, const double* Sigma, const double* ref_t, const double* u2, const double* z,
                                  double* logw_out, double* state_out, int32_t* anc_out, int32_t cluster_size, void* stream) {
    if (!model || !logw || !state || !Theta || !Sigma || !ref_t || !u2 || !z || !logw_out || !state_out || !anc_out)
        PGAS_FAIL(-1, "pgas_csmc_step_f64: null argument");
    const DevModel& m = model->dev;
    if (t < 1 || t >= m.T) PGAS_FAIL(-2, "t=%d outside [1,%d)", t, m.T);
    SweepArgs a;
    memset(&a, 0, sizeof(a));
    a.m = m;
    a.N = N; a.n_chains = 1;
    a.C = pgas_choose_cluster(m, N, 1, cluster_size);
    if (int rc = check_cluster(a.C, N)) return rc;
    a.P = (N + a.C - 1) / a.C;
    a.t_begin = t; a.t_end = t + 1;
    a.row_off = t; a.anc_shift = 1;
    a.ref_rows = 1; a.trace_rows = 1; a.anc_rows = 1; a.var_rows = 1;
    a.ref = ref_t; a.ref_stride = m.n_x; a.Theta = Theta; a.Sigma = Sigma;
    a.init_state = state; a.init_logw = logw;
    a.state_trace = state_out; a.anc_trace = anc_out; a.logw_last = logw_out;
    a.rng_mode = 1; a.Z = z; a.U = u2;
    return pgas_launch_sweep(a, (cudaStream_t)stream);
}

extern "C" int pgas_reconstruct_trajectory_f64(const double* particles, const int32_t* ancestry, const int32_t* idx, int32_t n_sets,
                                               int32_t T, int32_t N, int32_t n, double* traj_out, void* stream) {
    if (!particles || !ancestry || !idx || !traj_out) PGAS_FAIL(-1, "pgas_reconstruct_trajectory_f64: null argument");
    if (n < 1 || n > 32 || T < 1 || N < 1 || n_sets < 1) PGAS_FAIL(-2, "bad shape (n_sets=%d T=%d N=%d n=%d)", n_sets, T, N, n);
    return pgas_launch_pick_and_trace(nullptr, particles, ancestry, idx, n_sets, T, N, n, nullptr, 0, nullptr, traj_out, (long long)T * n,
                                      (cudaStream_t)stream);
}

extern "C" int pgas_resample_f64(const double* w, int32_t N, int32_t n_sets, const double* u, int32_t* idx_out, void* stream) {
    if (!w || !u || !idx_out) PGAS_FAIL(-1, "pgas_resample_f64: null argument");
    if (N < 1 || n_sets < 1) PGAS_FAIL(-2, "bad shape (N=%d n_sets=%d)", N, n_sets);
    const size_t smem = sizeof(double) * (size_t)N;
    if (smem > 200 * 1024) PGAS_FAIL(-2, "N=%d exceeds the single-CTA resampler (use the sweep)", N);
    PGAS_CUDA(cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    resample_kernel<<<n_sets, BT, smem, (cudaStream_t)stream>>>(w, N, u, idx_out);
    PGAS_KERNEL_CHECK();
    return 0;
}

extern "C" int pgas_hgp_eval_f64(const pgas_model* model, const double* states, const double* inputs, int32_t input_stride, int32_t n,
                                 double* phi_out, void* stream) {
    if (!model || !states || !phi_out) PGAS_FAIL(-1, "pgas_hgp_eval_f64: null argument");
    const DevModel& m = model->dev;
    if (m.n_u > 0 && !inputs) PGAS_FAIL(-1, "inputs required (n_u=%d)", m.n_u);
    if (n < 1) return 0;
    const int warps = 4;
    const int maxpos = m.npos;
    const size_t bytes = sizeof(double) * warps * m.D * (size_t)maxpos;
    const int blocks = std::min((n + warps - 1) / warps, 148 * 8);
    hgp_eval_kernel<<<blocks, warps * 32, bytes, (cudaStream_t)stream>>>(m, states, inputs, input_stride, n, maxpos, phi_out);
    PGAS_KERNEL_CHECK();
    return 0;
}

extern "C" int pgas_philox_sweep_variates_f64(const pgas_rng* rng, int32_t n_chains, int32_t T, int32_t N, int32_t n_x, double* Z_out,
                                              double* U_out, void* stream) {
    if (!rng || !Z_out || !U_out) PGAS_FAIL(-1, "pgas_philox_sweep_variates_f64: null argument");
    philox_variates_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(rng->seed, rng->chain_base, rng->iteration, n_chains, T, N, n_x,
                                                                     Z_out, U_out);
    PGAS_KERNEL_CHECK();
    return 0;
}
