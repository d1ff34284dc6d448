// chains.cu — PGAS.__call__ (reference src/PGAS.py:345-397) for n_chains independent chains,
// stream-ordered on the device: K iterations of
//   sweep (sweep.cu) -> final pick + backward trace -> sufficient statistics (suffstats.cu)
//   -> eta = prior + statistics -> MNIW draw (mniw_draw.cu)
// with no host synchronisation inside the loop.  Chains never communicate; the multi-GPU layer
// (Python, torch.distributed) shards chain ids across ranks and gathers the traces at the end.
#include "sweep_args.cuh"
#include <algorithm>

__global__ void add_prior_kernel(const double* __restrict__ p0, const double* __restrict__ p1, const double* __restrict__ p2,
                                 const double* __restrict__ T0, const double* __restrict__ T1, const double* __restrict__ T2, int M,
                                 int nx, int n_chains, double* __restrict__ e0, double* __restrict__ e1, double* __restrict__ e2) {
    // eta_j = prior_j + T_j (src/PGAS.py:298-303), prior shared by all chains
    const size_t n0 = (size_t)M * nx, n1 = (size_t)M * M, n2 = (size_t)nx * nx;
    const size_t per = n0 + n1 + n2, total = per * n_chains;
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const size_t c = g / per, r = g % per;
        if (r < n0) e0[c * n0 + r] = p0[r] + T0[c * n0 + r];
        else if (r < n0 + n1) e1[c * n1 + (r - n0)] = p1[r - n0] + T1[c * n1 + (r - n0)];
        else e2[c * n2 + (r - n0 - n1)] = p2[r - n0 - n1] + T2[c * n2 + (r - n0 - n1)];
    }
}

struct ChainWs {
    double *state, *logw, *T0, *T1, *T2, *e0, *e1, *e2, *A, *S;
    int *anc, *status;
    void* draw;
    void* sweep;
    size_t draw_bytes, sweep_bytes, total;
};

static ChainWs carve(const DevModel& m, int N, int n_chains, char* base) {
    ChainWs w;
    size_t o = 0;
    auto take = [&](size_t bytes) { char* p = base ? base + o : nullptr; o += (bytes + 255) & ~(size_t)255; return p; };
    const size_t C = n_chains, T = m.T, nx = m.n_x, M = m.M;
    w.state = (double*)take(sizeof(double) * C * T * N * nx);
    w.anc = (int*)take(sizeof(int) * C * (T - 1) * N);
    w.logw = (double*)take(sizeof(double) * C * N);
    w.T0 = (double*)take(sizeof(double) * C * M * nx);
    w.T1 = (double*)take(sizeof(double) * C * M * M);
    w.T2 = (double*)take(sizeof(double) * C * nx * nx);
    w.e0 = (double*)take(sizeof(double) * C * M * nx);
    w.e1 = (double*)take(sizeof(double) * C * M * M);
    w.e2 = (double*)take(sizeof(double) * C * nx * nx);
    w.A = (double*)take(sizeof(double) * C * nx * M);
    w.S = (double*)take(sizeof(double) * C * nx * nx);
    w.status = (int*)take(sizeof(int) * C);
    w.draw_bytes = pgas_mniw_draw_workspace_bytes((int)M, (int)nx, n_chains);
    w.draw = take(w.draw_bytes);
    w.sweep_bytes = pgas_sweep_split_workspace(m, N, n_chains);
    w.sweep = take(w.sweep_bytes);
    w.total = o + 256;
    return w;
}

extern "C" size_t pgas_run_chains_workspace_bytes(const pgas_model* model, int32_t N, int32_t n_chains) {
    if (!model) return 0;
    return carve(model->dev, N, n_chains, nullptr).total;
}

// dst[c * dstride + e] = src[c * sstride + e], e < len, c < rows
__global__ void __launch_bounds__(256) strided_rows_copy_kernel(double* __restrict__ dst, long long dstride, const double* __restrict__ src,
                                                                long long sstride, size_t len, int rows) {
    const size_t total = len * (size_t)rows;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
        const size_t c = e / len, i = e % len;
        dst[c * dstride + i] = src[c * sstride + i];
    }
}
static int strided_rows_copy(double* dst, long long dstride, const double* src, long long sstride, size_t len, int rows, cudaStream_t st) {
    const size_t total = len * (size_t)rows;
    if (!total) return 0;
    const unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 8);
    strided_rows_copy_kernel<<<grid, 256, 0, st>>>(dst, dstride, src, sstride, len, rows);
    PGAS_KERNEL_CHECK();
    return 0;
}

extern "C" int pgas_run_chains_f64(const pgas_model* model, int32_t N, int32_t K, int32_t n_chains, const double* eta0,
                                   const double* eta1, const double* eta2, double eta3, const double* init_ref, const pgas_rng* rng,
                                   double* state_trace_out, double* A_trace_out, double* S_trace_out, int32_t cluster_size,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    if (!model || !eta0 || !eta1 || !eta2 || !init_ref || !rng || !state_trace_out || !workspace)
        PGAS_FAIL(-1, "pgas_run_chains_f64: null argument");
    if (K < 1 || n_chains < 1 || N < 2) PGAS_FAIL(-2, "bad sizes (K=%d n_chains=%d N=%d)", K, n_chains, N);
    const DevModel& m = model->dev;
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    ChainWs w = carve(m, N, n_chains, base);
    if (workspace_bytes < w.total) PGAS_FAIL(-5, "workspace too small: need %zu bytes, got %zu", w.total, workspace_bytes);
    const size_t T = m.T, nx = m.n_x, M = m.M;
    const long long tstride = (long long)K * T * nx;           // chain stride of state_trace_out (n_chains, K, T, nx)
    if (rng->mode == 1 && (!rng->chi2 || !rng->G || !rng->Nrm || (K > 1 && (!rng->Z || !rng->U))))
        PGAS_FAIL(-1, "injected rng mode needs Z, U, chi2, G and Nrm");

    // per-chain blocks into the (n_chains, K, ...) traces: one strided copy kernel (a pitched cudaMemcpy2D rejects destination
    // pitches above cudaDevAttrMaxPitch, i.e. long runs with K T n_x > ~2.7e8)
    if (int rc = strided_rows_copy(state_trace_out, tstride, init_ref, (long long)(T * nx), T * nx, n_chains, st)) return rc;
    const int C = pgas_choose_cluster(m, N, n_chains, cluster_size);
    if (C < 1 || C > 16) PGAS_FAIL(-2, "cluster_size must be 0 (auto) or 1..16 (got %d -> %d)", cluster_size, C);
    for (int k = 0; k < K; ++k) {
        pgas_rng r = *rng;
        r.iteration = rng->iteration + (unsigned)k;
        if (rng->mode == 1) {
            r.Z = rng->Z ? rng->Z + (size_t)k * n_chains * T * N * nx : nullptr;
            r.U = rng->U ? rng->U + (size_t)k * n_chains * T * 2 : nullptr;
            r.chi2 = rng->chi2 + (size_t)k * n_chains * nx;
            r.G = rng->G + (size_t)k * n_chains * nx * nx;
            r.Nrm = rng->Nrm + (size_t)k * n_chains * nx * M;
        }
        double* traj_k = state_trace_out + (size_t)k * T * nx;
        if (k > 0) {
            SweepArgs a;
            memset(&a, 0, sizeof(a));
            a.m = m;
            a.N = N; a.n_chains = n_chains; a.C = C; a.P = (N + C - 1) / C;
            a.t_begin = 1; a.t_end = m.T;
            a.ref_rows = m.T; a.trace_rows = m.T; a.anc_rows = m.T - 1; a.var_rows = m.T;
            a.ref = traj_k - T * nx; a.ref_stride = tstride;
            a.Theta = w.A; a.Sigma = w.S;
            a.state_trace = w.state; a.anc_trace = w.anc; a.logw_last = w.logw;
            a.rng_mode = r.mode; a.seed = r.seed; a.chain_base = r.chain_base; a.iteration = r.iteration;
            a.Z = r.Z; a.U = r.U;
            a.ws = w.sweep; a.ws_bytes = w.sweep_bytes;
            if (int rc = pgas_launch_sweep(a, st)) return rc;
            if (int rc = pgas_launch_pick_and_trace(w.logw, w.state, w.anc, nullptr, n_chains, m.T, N, m.n_x, &r, m.T, nullptr, traj_k,
                                                    tstride, st))
                return rc;
        }
        if (int rc = pgas_launch_suffstats(m, traj_k, tstride, n_chains, w.T0, w.T1, w.T2, st)) return rc;
        add_prior_kernel<<<148 * 2, 256, 0, st>>>(eta0, eta1, eta2, w.T0, w.T1, w.T2, (int)M, (int)nx, n_chains, w.e0, w.e1, w.e2);
        PGAS_KERNEL_CHECK();
        if (int rc = pgas_launch_mniw_draw(w.e0, w.e1, w.e2, eta3 + (double)(T - 1), false, (int)M, (int)nx, n_chains, &r, m.flags, w.A,
                                           w.S, w.status, w.draw, w.draw_bytes, st))
            return rc;
        if (A_trace_out)
            if (int rc = strided_rows_copy(A_trace_out + (size_t)k * nx * M, (long long)(K * nx * M), w.A, (long long)(nx * M), nx * M, n_chains, st)) return rc;
        if (S_trace_out)
            if (int rc = strided_rows_copy(S_trace_out + (size_t)k * nx * nx, (long long)(K * nx * nx), w.S, (long long)(nx * nx), nx * nx, n_chains, st)) return rc;
    }
    return 0;
}
