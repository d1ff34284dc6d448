// sweep_args.cuh — argument block of the persistent sweep kernel (sweep.cu) and its launcher.
#pragma once
#include "common.cuh"

struct SweepArgs {
    DevModel m;
    int N, n_chains, C, P;       // P = particles per CTA = ceil(N / C)
    int t_begin, t_end;          // steps t = t_begin .. t_end-1
    int row_off, anc_shift;      // time-indexed caller arrays use row t - row_off (anc: t-1-row_off+anc_shift)
    int ref_rows, trace_rows, anc_rows, var_rows;   // leading (time) extents of ref / state_trace / anc_trace / Z,U per chain
    const double* ref;           // (n_chains, ref_rows, NX), chain stride ref_stride elements
    long long ref_stride;
    const double* Theta;         // (n_chains, NX, M)
    const double* Sigma;         // (n_chains, NX, NX)
    const double* init_state;    // (n_chains, N, NX) or null -> sample x_0
    const double* init_logw;     // (n_chains, N) or null
    double* state_trace;         // (n_chains, trace_rows, N, NX)
    int* anc_trace;              // (n_chains, anc_rows, N)
    double* logw_last;           // (n_chains, N) or null
    int rng_mode;                // 0 Philox, 1 injected
    unsigned long long seed;
    unsigned chain_base, iteration;
    const double* Z;             // (n_chains, var_rows, N, NX)
    const double* U;             // (n_chains, var_rows, 2)
    long long* dbg;              // optional phase clocks of CTA 0 (8 per step), developer aid
    // split form (sweep_split.cu): the state recursion does not depend on the resampling in the reference's
    // semantics (src/PGAS.py:131-133 propagates particle i from particle i), so a separate kernel runs it ahead
    // and leaves, per step, l_aux = log p(y_t | mu), h = log N(x_ref,t; mu, Sigma) and log p(y_t | x_t) here:
    const double* pre_la;        // (n_chains, pre_rows, N), row t - pre_off
    const double* pre_lr;
    const double* pre_ll;
    int pre_rows, pre_off;
    // caller workspace (pgas_csmc_sweep_workspace_bytes); null / too small -> fused kernel only
    void* ws;
    size_t ws_bytes;
};

int pgas_launch_sweep(const SweepArgs& a, cudaStream_t stream);
int pgas_launch_sweep_fused(const SweepArgs& a, cudaStream_t stream);     // one kernel, all phases (sweep.cu)
int pgas_launch_sweep_pre(const SweepArgs& a, cudaStream_t stream);       // resampling recursion on precomputed log-densities
int pgas_launch_weights(const SweepArgs& a, cudaStream_t stream);         // dedicated resampling kernel of the split form (weights.cu)
int pgas_weights_cluster(int N);                                           // its cluster size for N particles; 0 = not applicable
int pgas_launch_weights_lat(const SweepArgs& a, cudaStream_t stream);     // latency form: cluster per chain, st.async + mbarrier hand-offs (weights_lat.cu)
int pgas_weights_lat_cluster(int N);                                       // its cluster size for N particles; 0 = not applicable
bool pgas_weights_lat_fits_beside_big_state(int N);                        // its CTAs leave half of an SM's registers to a 256-thread state CTA
bool pgas_sweep_split_eligible(const SweepArgs& a);
size_t pgas_sweep_split_workspace(const DevModel& m, int N, int n_chains);
int pgas_choose_cluster(const DevModel& m, int N, int n_chains, int requested);
size_t pgas_sweep_smem_for(const DevModel& m, int P);

// final categorical pick (src/PGAS.py:224-225) + reconstruct_trajectory (src/Filtering.py:40-55)
int pgas_launch_pick_and_trace(const double* logw_last, const double* state_trace, const int* anc_trace, const int* idx_in,
                               int n_sets, int T, int N, int n, const pgas_rng* rng, int var_rows, int* final_idx, double* traj_out,
                               long long traj_stride, cudaStream_t st);
int pgas_launch_suffstats(const DevModel& m, const double* traj, long long traj_stride, int n_chains, double* T0, double* T1,
                          double* T2, cudaStream_t st);
int pgas_launch_mniw_draw(const double* eta0, const double* eta1, const double* eta2, double eta3, bool shared_eta, int M, int nx,
                          int n_chains, const pgas_rng* rng, int flags, double* A, double* S, int* status, void* ws, size_t ws_bytes,
                          cudaStream_t st);
