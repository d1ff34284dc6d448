/*
 * pgas_b200.h — C ABI of libpgas_b200.so: the B200-native (sm_100a) implementation of the
 * particle-Gibbs-with-ancestor-sampling hot path of
 * VolkmannB/bayesian-inference-with-explicit-and-implicit-prior-knowledge.
 *
 * The reference has no FFI: its seam is the Python API of src/*.py.  Each entry point below
 * names the reference function (file:line) it replaces; INTEGRATION.md shows the ctypes stub a
 * maintainer would put behind that function.  Conventions:
 *   - plain pointers and sizes only; every array pointer is a DEVICE pointer to float64 /
 *     int32 data in C (row-major) order unless the comment says "host";
 *   - the library never allocates result memory and never frees caller memory; scratch comes
 *     from a caller workspace sized by the matching *_workspace_bytes query;
 *   - all work is enqueued on the cudaStream_t passed as `void* stream` (0 = legacy default);
 *   - every function returns 0 on success, <0 for an argument/shape error, >0 for a CUDA error
 *     code; pgas_last_error() returns a thread-local message.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef PGAS_B200_H
#define PGAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGAS_MAX_NX 4      /* state dimension n_x                      */
#define PGAS_MAX_NY 2      /* observation dimension n_y                */
#define PGAS_MAX_NU 4      /* input dimension n_u                      */
#define PGAS_MAX_D 3       /* GP input dimension D                     */
#define PGAS_MAX_GP 2      /* GPs of the marginalised path (group B)   */

/* GP-input map g: (state, input) -> z in R^D fed to the Hilbert basis. */
enum { PGAS_MAP_AFFINE = 0,      /* z = Az [state; input] + bz  (src/EMPS.py:110-113, src/Toy_Example.py:146) */
       PGAS_MAP_VEHICLE_SLIP = 1,/* z = (alpha_f, alpha_r) of src/Vehicle.py:50-57, input = [delta, v_x]     */
       PGAS_MAP_PROGRAM = 2      /* z = an expression program over (state, input): the model plug-in */ };

/* Expression programs: the reference hands an arbitrary Python callable basis_fcn(state, input) to the sampler
 * (src/PGAS.py:24-43, src/StateSpaceModel.py:19-30) and lets JAX trace it.  Here the host traces the callable's GP-input map
 * symbolically (models.py: Sym) into a postfix program that every kernel evaluating the map interprets per particle — uniform
 * control flow, no run-time compiler in the loop.  prog_op[i] = opcode | (argument << 8); after the last instruction the
 * stack holds z_0 .. z_{D-1}. */
#define PGAS_MAX_PROG 64
#define PGAS_PROG_STACK 8
enum { PGAS_OP_PUSH_X = 1,   /* argument: state component          */
       PGAS_OP_PUSH_U = 2,   /* argument: input component          */
       PGAS_OP_PUSH_C = 3,   /* argument: index into prog_const    */
       PGAS_OP_ADD = 4, PGAS_OP_SUB = 5, PGAS_OP_MUL = 6, PGAS_OP_DIV = 7, PGAS_OP_NEG = 8,
       PGAS_OP_SIN = 9, PGAS_OP_COS = 10, PGAS_OP_TAN = 11, PGAS_OP_TANH = 12, PGAS_OP_ATAN = 13, PGAS_OP_EXP = 14,
       PGAS_OP_LOG = 15, PGAS_OP_SQRT = 16, PGAS_OP_ABS = 17, PGAS_OP_POW = 18, PGAS_OP_ATAN2 = 19,
       PGAS_OP_PUSH_Y = 20   /* argument: observation component (likelihood programs only) */ };

/* reference quirks; the default 0 reproduces the reference bit for bit */
enum { PGAS_FLAG_ANCESTOR_GATHER = 1, /* propagate x_t^i from x_{t-1}^{a_i} instead of x_{t-1}^i (src/PGAS.py:131-133) */
       PGAS_FLAG_INPUT_PREV      = 2, /* sweep pairs x_{t-1} with inputs[t-1] instead of inputs[t] (src/PGAS.py:52-54)  */
       PGAS_FLAG_VCHOL_TRANSPOSE = 4  /* A = mean + S_chol Nrm V_chol^T instead of V_chol (src/PGAS.py:341)             */ };

/* Host-side description of one Theta-conditioned model: the data and the two user callables of
 * condSequentialMonteCarlo.__init__ (src/PGAS.py:24-43) restricted to the shipped families:
 * basis_fcn = Hilbert-space GP basis of a map of (state,input); likelihood_fcn = Gaussian
 * log-density of the observation around H state + h0, or an expression program (lik_prog_*).
 * All pointers are HOST pointers, copied. */
typedef struct pgas_model_params {
    int32_t n_x, n_y, n_u, D, M, T;
    /* Hilbert basis (src/BasisFunctions.py:8-80): integer frequencies S (M x D), phi_m(z) =
     * prod_d L_d^-1/2 sin(pi S[m,d] (z_d - center_d + L_d) / (2 L_d)); frequencies lie on the
     * lattice idx_start + k*idx_step (src/BasisFunctions.py:24-25). */
    const int32_t* freq;          /* (M, D) host */
    int32_t idx_start, idx_step;
    double center[PGAS_MAX_D];
    double half_width[PGAS_MAX_D]; /* L_d = domain_size_d / 2 */
    /* GP-input map */
    int32_t map_kind;
    double Az[PGAS_MAX_D][PGAS_MAX_NX + PGAS_MAX_NU];
    double bz[PGAS_MAX_D];
    double slip_lf, slip_lr;
    int32_t prog_len;             /* PGAS_MAP_PROGRAM: number of instructions (<= PGAS_MAX_PROG) */
    int32_t prog_op[PGAS_MAX_PROG];
    double prog_const[PGAS_MAX_PROG];
    /* Gaussian likelihood N(y; H x + h0, R) */
    double H[PGAS_MAX_NY][PGAS_MAX_NX];
    double h0[PGAS_MAX_NY];
    double R[PGAS_MAX_NY][PGAS_MAX_NY];
    /* data and initial distribution */
    const double* observations;   /* (T, n_y) host */
    const double* inputs;         /* (T, n_u) host; may be NULL when n_u == 0 */
    double m0[PGAS_MAX_NX];
    double P0[PGAS_MAX_NX][PGAS_MAX_NX];
    int32_t flags;                /* PGAS_FLAG_* */
    /* Model plug-in for likelihood_fcn(obs, state, input) (src/PGAS.py:93-100, :137-147): lik_prog_len > 0 replaces the Gaussian
     * family above by an expression program over (state, inputs[t], observations[t]) that leaves ONE value, the log-density.
     * Together with the GP-input program at most PGAS_MAX_PROG instructions and constants.  Such a model runs the fused sweep
     * kernel (the split form's state kernel is specialised for the Gaussian family). */
    int32_t lik_prog_len;
    int32_t lik_prog_op[PGAS_MAX_PROG];
    double lik_prog_const[PGAS_MAX_PROG];
} pgas_model_params;

typedef struct pgas_model pgas_model;   /* opaque; owns device copies of the tables/data */

/* Random-number source of a sweep / draw.  Injected mode (testing): device
 * arrays of variates.  Philox mode (production): Philox-4x32-10, key = seed, counter =
 * (particle, time, iteration, purpose<<24 | chain). */
typedef struct pgas_rng {
    int32_t mode;                 /* 0 = Philox, 1 = injected */
    uint64_t seed;
    uint32_t chain_base;          /* chain c of this call uses chain id chain_base + c */
    uint32_t iteration;
    /* injected sweep variates (device): Z (n_chains,T,N,n_x) normals, U (n_chains,T,2) uniforms:
     * U[t,0]=u_res(t), U[t,1]=u_anc(t) for t>=1; U[0,0]=u_idx */
    const double* Z;
    const double* U;
    /* injected draw variates (device): chi2 (n_chains,n_x), G (n_chains,n_x,n_x), Nrm (n_chains,n_x,M) */
    const double* chi2;
    const double* G;
    const double* Nrm;
} pgas_rng;

const char* pgas_last_error(void);
int pgas_version(void);
/* PGAS_ABI_VERSION of the header the library was compiled against: a binding compares it with its own copy of
 * the header before the first call (struct layouts and argument orders are only meaningful when they agree). */
#define PGAS_ABI_VERSION 204
int pgas_abi_version(void);
int pgas_device_count(void);
/* number of CUDA kernels this library has launched in this process (bench.py: gpu_launches) */
long long pgas_launch_count(void);

/* condSequentialMonteCarlo.__init__ (src/PGAS.py:24-43) + generate_Hilbert_BasisFunction's closure
 * (src/BasisFunctions.py:63-66): uploads tables/data, builds the packed tensor-product row layout. */
int pgas_model_create(const pgas_model_params* params, pgas_model** out);
int pgas_model_destroy(pgas_model* model);
/* largest last-dimension lattice position + 1 (register-table length the sweep needs) */
int pgas_model_jmax(const pgas_model* model);

/* systematic_SISR (src/Filtering.py:6-37).  w (n_sets,N) unnormalised weights, u (n_sets) uniforms
 * -> idx_out (n_sets,N) int32.  One CTA per set. */
int pgas_resample_f64(const double* w, int32_t N, int32_t n_sets, const double* u,
                      int32_t* idx_out, void* stream);

/* vmap(basis_fcn) (src/BasisFunctions.py:77-80 via src/PGAS.py:52,67,294): states (n,n_x), inputs
 * (n,n_u) [input_stride 0 = one shared input row] -> phi_out (n,M) in the reference's basis order. */
int pgas_hgp_eval_f64(const pgas_model* model, const double* states, const double* inputs,
                      int32_t input_stride, int32_t n, double* phi_out, void* stream);

/* condSequentialMonteCarlo.step (src/PGAS.py:79-153), one step for testing/teacher forcing:
 * logw (N), state (N,n_x), Theta (n_x,M), Sigma (n_x,n_x), ref_t (n_x), u2 = {u_res,u_anc} (2),
 * z (N,n_x) -> logw_out (N), state_out (N,n_x), anc_out (N) int32.  Runs the fused sweep kernel for a
 * single time step `t`.  anc_out[N-1] is the reference particle's ancestor exactly as src/PGAS.py:122-127
 * stores it: unclipped, i.e. N when rounding leaves cumsum(w)[-1] below u_anc (JAX's gather at :146 clamps;
 * NumPy indexing would not) - clamp to N-1 before indexing with it on the host. */
int pgas_csmc_step_f64(const pgas_model* model, int32_t N, int32_t t, const double* logw,
                       const double* state, const double* Theta, const double* Sigma,
                       const double* ref_t, const double* u2, const double* z,
                       double* logw_out, double* state_out, int32_t* anc_out,
                       int32_t cluster_size, void* stream);

/* condSequentialMonteCarlo.__call__ (src/PGAS.py:176-228) for n_chains independent chains: the sweep, the final categorical
 * pick (:224-225) and reconstruct_trajectory (src/Filtering.py:40-55).  The sweep runs in one of two schedules of the same
 * arithmetic.  Split form (two-dimensional bases, reference semantics, workspace given): csmc_state_kernel in launches of <= 16
 * steps runs ahead and writes traces plus three log-densities per particle-step to HBM; a resampling kernel (cluster per chain
 * with st.async / mbarrier hand-offs up to 32 chains, one CTA per chain beyond) runs the weight recursion chunk-wise behind it —
 * about 290 launches per sweep at T = 2000, nothing resident across the whole sweep.  Fused form (otherwise): one kernel per
 * launch range with the particle set in shared memory / DSMEM.
 *   ref_traj (n_chains,T,n_x), Theta (n_chains,n_x,M), Sigma (n_chains,n_x,n_x)
 *   -> traj_out (n_chains,T,n_x); optional (may be NULL... see below) traces:
 *      state_trace (n_chains,T,N,n_x), anc_trace (n_chains,T-1,N) int32, logw_last (n_chains,N),
 *      final_idx (n_chains) int32.
 * state_trace and anc_trace are required (the backward pass reads them); logw_last/final_idx may
 * be NULL.  anc_trace[.., N-1] (reference particle) is unclipped as in the reference and may equal N
 * (see pgas_csmc_step_f64); every device consumer clamps.  cluster_size: 0 = choose automatically, else 1,2,4,8,16.
 * workspace: pgas_csmc_sweep_workspace_bytes bytes enable the split form for two-dimensional bases (a state
 * kernel running ahead of the resampling kernel on a library-owned low-priority stream, fenced against
 * `stream` with events); with a NULL / smaller workspace the fused single-kernel form runs.  Both forms
 * produce the same ancestors and traces. */
size_t pgas_csmc_sweep_workspace_bytes(const pgas_model* model, int32_t N, int32_t n_chains);
int pgas_csmc_sweep_f64(const pgas_model* model, int32_t N, int32_t n_chains,
                        const double* ref_traj, const double* Theta, const double* Sigma,
                        const pgas_rng* rng, double* state_trace, int32_t* anc_trace,
                        double* logw_last, int32_t* final_idx, double* traj_out,
                        int32_t cluster_size, void* workspace, size_t workspace_bytes,
                        void* stream);

/* reconstruct_trajectory (src/Filtering.py:40-55): particles (n_sets,T,N,n), ancestry
 * (n_sets,T-1,N) int32, idx (n_sets) int32 -> traj (n_sets,T,n). */
int pgas_reconstruct_trajectory_f64(const double* particles, const int32_t* ancestry,
                                    const int32_t* idx, int32_t n_sets, int32_t T, int32_t N,
                                    int32_t n, double* traj_out, void* stream);

/* First half of PGAS.sample_params (src/PGAS.py:294-303; prior_mniw_calcStatistics,
 * src/BayesianInferrence.py:53-61): traj (n_chains,T,n_x) -> T0 (n_chains,M,n_x) = Phi^T Y,
 * T1 (n_chains,M,M) = Phi^T Phi, T2 (n_chains,n_x,n_x) = Y^T Y; T3 = T-1 is implicit.
 * Phi is recomputed on the fly from the trajectory and never stored. */
int pgas_suffstats_f64(const pgas_model* model, const double* traj, int32_t n_chains,
                       double* T0_out, double* T1_out, double* T2_out, void* stream);

/* Second half of PGAS.sample_params (src/PGAS.py:306-343; prior_mniw_2naturalPara_inv,
 * src/BayesianInferrence.py:35-45): eta0 (n_chains,M,n_x), eta1 (n_chains,M,M), eta2
 * (n_chains,n_x,n_x), eta3 (scalar, host) -> A_out (n_chains,n_x,M), S_out (n_chains,n_x,n_x).
 * eta1 is not modified.  status_out (n_chains) int32 device: 0 ok, k>0 = eta1 lost positive
 * definiteness at pivot k (the reference would return NaN). */
size_t pgas_mniw_draw_workspace_bytes(int32_t M, int32_t n_x, int32_t n_chains);
int pgas_mniw_draw_f64(const double* eta0, const double* eta1, const double* eta2, double eta3,
                       int32_t M, int32_t n_x, int32_t n_chains, const pgas_rng* rng,
                       int32_t flags, double* A_out, double* S_out, int32_t* status_out,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Post-processing of the traced statistics: what the reference's figure scripts do with
 * jax.vmap(prior_mniw_2naturalPara_inv) over K iterations (SingleMassOscillator_Figures.py:58-89;
 * src/BayesianInferrence.py:35-45) followed by a loop of prior_mniw_Predictive on the plot grid (:131-140;
 * src/BayesianInferrence.py:64-89), batched on the device.
 *   eta0 (K,M,n), eta1 (K,M,M), eta2 (K,n,n), eta3 (K) device: natural parameters INCLUDING the prior;
 *   basis (G,M) device: basis functions on the grid (pgas_hgp_eval_f64); G may be 0
 *   -> mean_out (K,n,M), row_scale_out (K,n,n) = eta2 - mean eta0, df_out (K) = eta3      [2naturalPara_inv]
 *      pred_mean_out (K,G,n) = basis mean^T, pred_colscale_out (K,G) = diag(basis V basis^T + I)   [Predictive]
 *   (the predictive row scale / df are row_scale_out / (df_out + 1 - n) and df_out + 1 - n; the full col_cov (M,M) and
 *   col_scale (G,G) of the reference are not formed).  status_out (K): 0 ok, j>0 = eta1 not positive definite at pivot j. */
size_t pgas_mniw_posterior_batch_workspace_bytes(int32_t M, int32_t n, int32_t K);
int pgas_mniw_posterior_batch_f64(int32_t M, int32_t n, int32_t K, const double* eta0, const double* eta1,
                                  const double* eta2, const double* eta3, const double* basis, int32_t G,
                                  double* mean_out, double* row_scale_out, double* df_out,
                                  double* pred_mean_out, double* pred_colscale_out, int32_t* status_out,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* PGAS.__call__ (src/PGAS.py:345-397) for n_chains independent chains, entirely on the device:
 * K iterations of (sweep -> pick -> backward trace -> sufficient statistics -> MNIW draw).
 *   prior eta0 (M,n_x), eta1 (M,M), eta2 (n_x,n_x) device, eta3 host scalar (shared by chains);
 *   init_ref (n_chains,T,n_x)
 *   -> state_trace_out (n_chains,K,T,n_x)  [row k = trajectory of iteration k; row 0 = init_ref],
 *      optional A_trace_out (n_chains,K,n_x,M), S_trace_out (n_chains,K,n_x,n_x) (may be NULL).
 * In injected mode rng->Z/U hold (K,n_chains,...) blocks (block k used by sweep k; block 0 unused)
 * and rng->chi2/G/Nrm hold (K,n_chains,...) blocks (block k = draw after trajectory k). */
size_t pgas_run_chains_workspace_bytes(const pgas_model* model, int32_t N, int32_t n_chains);
int pgas_run_chains_f64(const pgas_model* model, int32_t N, int32_t K, int32_t n_chains,
                        const double* eta0, const double* eta1, const double* eta2, double eta3,
                        const double* init_ref, const pgas_rng* rng, double* state_trace_out,
                        double* A_trace_out, double* S_trace_out, int32_t cluster_size,
                        void* workspace, size_t workspace_bytes, void* stream);

/* The variates the Philox mode would use, written out so a CPU checker can be fed the same
 * numbers: Z (n_chains,T,N,n_x), U (n_chains,T,2) for sweep `rng->iteration`. */
int pgas_philox_sweep_variates_f64(const pgas_rng* rng, int32_t n_chains, int32_t T, int32_t N,
                                   int32_t n_x, double* Z_out, double* U_out, void* stream);

/* ====================================================================================
 * Group B — the marginalised particle filters: Algorithm1 (online
 * marginalised auxiliary particle filter, src/Algorithm1.py), Algorithm3 (marginalised conditional
 * SMC with ancestor sampling, src/Algorithm3.py) and Algorithm2 (the PGAS outer loop,
 * src/Algorithm2.py).  Every particle carries the MNIW sufficient statistics of each GP.
 *
 * The reference's StateSpaceModel callables (src/StateSpaceModel.py:19-30) are arbitrary Python.
 * The host layer traces them once per time step (the input u_t is concrete there) and reduces them to
 * per-step coefficient tables of the compiled-in family
 *   transition   x' = A_t x + B_t xi + c_t                              (trans, row t moves t -> t+1)
 *   output       y  = link(C_t x + D_t xi + e_t),  link = identity | tanh (outp, row t)
 *   GP input g   z_d = p_t,d * link(a_t,d . x + b_t,d) + q_t,d, link = identity | atan
 * which covers every shipped example (RK4 of src/SingleMassOscillator.py:32-44 and src/EMPS.py:177-183
 * is affine in (x, xi); src/Vehicle.py:60-128 is affine in (x, xi) given u_t, with a tanh output and
 * the atan slip angles of :50-57).  Interface variables are scalar (n_xi = 1), as in every example.
 *
 * Model plug-in: a transition_model / output_model / GP-input map OUTSIDE those families is traced once,
 * with the input symbolic as well, into a postfix expression program (same instruction set as PGAS_MAP_PROGRAM above) that the
 * kernels interpret per particle.  Operands: PGAS_OP_PUSH_X k = state component k for k < n_x, interface variable k - n_x
 * otherwise (GP-input programs read the state only); PGAS_OP_PUSH_U k = component k of inputs[t]; PGAS_OP_PUSH_C k = consts[k].
 * After the last instruction the stack holds the n_x (transition), n_y (output) or D (GP input) results in order.  A program of
 * length 0 means "use the table"; with a program the corresponding table pointer may be null.
 * ==================================================================================== */
enum { PGAS_LINK_IDENTITY = 0, PGAS_LINK_ATAN = 1, PGAS_LINK_TANH = 2 };

typedef struct pgas_marg_program {
    int32_t len, n_const;          /* instructions (<= PGAS_MAX_PROG), constants (<= PGAS_MAX_PROG) */
    const int32_t* ops;            /* (len) host: opcode | (argument << 8) */
    const double* consts;          /* (n_const) host */
} pgas_marg_program;

typedef struct pgas_marg_gp {
    int32_t M, D;
    const double* sqrt_eig;        /* (M, D) host: sqrt(eigenvalue) = pi j / size  (src/BasisFunctions.py:60,79) */
    double center[PGAS_MAX_D];
    double half_width[PGAS_MAX_D];
    int32_t link;                  /* PGAS_LINK_IDENTITY or PGAS_LINK_ATAN */
    const double* gp_in;           /* (T, D, n_x + 1) host: a_t,d (n_x) then b_t,d */
    const double* gp_post;         /* (T, D, 2) host: p_t,d, q_t,d */
    const double* eta0;            /* (M) host      prior natural parameters (src/BayesianInferrence.py:18-32) */
    const double* eta1;            /* (M, M) host */
    double eta2, eta3;
    double xi_mean, xi_var;        /* init_int_var_mean / init_int_var_cov (src/Algorithm1.py:36-37) */
    pgas_marg_program prog;        /* len > 0: z = program(state, inputs[t]) instead of gp_in / gp_post / link */
} pgas_marg_gp;

typedef struct pgas_marg_params {
    int32_t n_x, n_y, n_gp, T;
    pgas_marg_gp gp[PGAS_MAX_GP];
    const double* trans;           /* (T, n_x, n_x + n_gp + 1) host */
    const double* outp;            /* (T, n_y, n_x + n_gp + 1) host */
    int32_t out_link;              /* PGAS_LINK_IDENTITY or PGAS_LINK_TANH */
    const double* observations;    /* (T, n_y) host */
    double Q[PGAS_MAX_NX][PGAS_MAX_NX];   /* process noise (src/StateSpaceModel.py:9) */
    double R[PGAS_MAX_NY][PGAS_MAX_NY];   /* output noise */
    double m0[PGAS_MAX_NX];
    double P0[PGAS_MAX_NX][PGAS_MAX_NX];
    int32_t n_u;                   /* components of inputs[t] the programs may read (0: none) */
    const double* inputs;          /* (T, n_u) host; may be null when n_u = 0 */
    pgas_marg_program trans_prog;  /* len > 0: x' = program(state, xi, inputs[t]) instead of trans */
    pgas_marg_program outp_prog;   /* len > 0: y  = program(state, xi, inputs[t]) instead of outp / out_link */
} pgas_marg_params;

typedef struct pgas_marg_model pgas_marg_model;

/* Variates of the marginalised filters.  Injected mode (device arrays):
 *   Z (n_chains,T,N,n_x) normals [row 0: initial states], ZXI0 (n_chains,n_gp,N) normals,
 *   U (n_chains,T,2) uniforms [U[t,0]=u_res(t), U[t,1]=u_anc(t) for t>=1; U[0,0]=u_idx],
 *   TS (n_chains,n_gp,T,N) Student-t variates with the predictive's degrees of freedom
 *   (jax.random.t, src/BayesianInferrence.py:103).
 * Philox mode generates the same slots in-kernel (t = z sqrt(a/g), g ~ Gamma(a), a = df/2). */
typedef struct pgas_marg_rng {
    int32_t mode;                  /* 0 = Philox, 1 = injected */
    uint64_t seed;
    uint32_t chain_base, iteration;
    const double* Z;
    const double* ZXI0;
    const double* U;
    const double* TS;
} pgas_marg_rng;

int pgas_marg_model_create(const pgas_marg_params* params, pgas_marg_model** out);
int pgas_marg_model_destroy(pgas_marg_model* model);
size_t pgas_marg_workspace_bytes(const pgas_marg_model* model, int32_t N, int32_t n_chains);

/* Statistic arrays are passed as arrays of 4*n_gp device pointers, entry 4g+j = statistic j of GP g:
 * j=0: T0 (.., M), j=1: T1 (.., M, M), j=2: T2 (..), j=3: T3 (..)   [n_xi = 1]. */

/* Algorithm1.__call__ (src/Algorithm1.py:399-492): one persistent kernel for all T steps.
 *   -> state_trace (n_chains,T,N,n_x), xi_trace (n_chains,n_gp,T,N), logw_trace (n_chains,T,N),
 *      anc_trace (n_chains,T-1,N) int32; optional sst_trace[4g+j] (n_chains,T,...) = weighted means
 *      of the per-particle statistics (:445-457); optional final_stats[4g+j] (n_chains,N,...) (:488).
 * status (n_chains) int32 device: 0 or the first time step at which a factorisation lost positive
 * definiteness (the reference would propagate NaN). */
int pgas_marg_filter_f64(const pgas_marg_model* model, int32_t N, int32_t n_chains, double forgetting_factor,
                         const pgas_marg_rng* rng, double* state_trace, double* xi_trace, double* logw_trace,
                         int32_t* anc_trace, double* const* sst_trace, double* const* final_stats,
                         int32_t* status, int32_t cluster_size, void* workspace, size_t workspace_bytes, void* stream);

/* Reference-trajectory statistics of Algorithm2 (src/Algorithm2.py:83-96, :139-152): sums over ALL T
 * steps.  x_traj (n_chains,T,n_x) with chain stride x_stride, xi_traj (n_chains,n_gp,T) with chain
 * stride xi_stride and GP stride xi_gstride (elements) -> stats_out[4g+j] (n_chains, ...). */
int pgas_marg_refstats_f64(const pgas_marg_model* model, const double* x_traj, int64_t x_stride, const double* xi_traj,
                           int64_t xi_stride, int64_t xi_gstride, int32_t n_chains, double* const* stats_out,
                           void* stream);

/* Algorithm3.__call__ (src/Algorithm3.py:199-303): ref_x (n_chains,T,n_x), ref_xi (n_chains,n_gp,T),
 * ref_stats[4g+j] (n_chains,...) -> traces as above, final_idx (n_chains) int32 (may be NULL),
 * traj_x_out (n_chains,T,n_x), traj_xi_out (n_chains,n_gp,T). */
int pgas_marg_csmc_f64(const pgas_marg_model* model, int32_t N, int32_t n_chains, const double* ref_x,
                       const double* ref_xi, const double* const* ref_stats, const pgas_marg_rng* rng,
                       double* state_trace, double* xi_trace, double* logw_trace, int32_t* anc_trace,
                       int32_t* final_idx, double* traj_x_out, double* traj_xi_out, int32_t* status,
                       int32_t cluster_size, void* workspace, size_t workspace_bytes, void* stream);

/* Algorithm2.__call__ (src/Algorithm2.py:106-187), K iterations stream-ordered on the device:
 * init_x (n_chains,T,n_x), init_xi (n_chains,n_gp,T) -> x_trace_out (n_chains,K,T,n_x),
 * xi_trace_out (n_chains,n_gp,K,T), sst_out[4g+j] (n_chains,K,...) reference statistics per iteration.
 * Injected mode: rng arrays carry a leading K axis (block k feeds sweep k; block 0 unused). */
size_t pgas_marg_run_workspace_bytes(const pgas_marg_model* model, int32_t N, int32_t n_chains);
int pgas_marg_run_f64(const pgas_marg_model* model, int32_t N, int32_t K, int32_t n_chains, const double* init_x,
                      const double* init_xi, const pgas_marg_rng* rng, double* x_trace_out, double* xi_trace_out,
                      double* const* sst_out, int32_t* status, int32_t cluster_size, void* workspace,
                      size_t workspace_bytes, void* stream);

/* vmap(vmap(output_mdl)) and vmap(vmap(log_likelihood)) over a (T, n) table (src/Algorithm1.py:463-480,
 * src/Algorithm2.py:160-178): states (T,n,n_x), xi (n_gp,T,n) -> obs_out (T,n,n_y), loglik_out (T,n). */
int pgas_marg_outputs_f64(const pgas_marg_model* model, const double* states, const double* xi, int32_t n,
                          double* obs_out, double* loglik_out, void* stream);

/* vmap(prior_mniw_log_base_measure) (src/BayesianInferrence.py:111-124) for n_xi = 1:
 * T0 (n,M), T1 (n,M,M), T2 (n), T3 (n) -> out (n). */
int pgas_mniw_log_base_measure_f64(const double* T0, const double* T1, const double* T2, const double* T3,
                                   int32_t n, int32_t M, double* out, void* stream);

/* The variates the Philox mode of the marginalised filters uses for sweep `rng->iteration`:
 * Z (n_chains,T,N,n_x), ZXI0 (n_chains,n_gp,N), U (n_chains,T,2), TS (n_chains,n_gp,T,N) with
 * df (n_gp,T) host = degrees of freedom of the predictive at each step. */
int pgas_philox_marg_variates_f64(const pgas_marg_rng* rng, int32_t n_chains, int32_t n_gp, int32_t T, int32_t N,
                                  int32_t n_x, const double* df_host, double* Z_out, double* ZXI0_out,
                                  double* U_out, double* TS_out, void* stream);

/* Machine denominators for the FP64 rooflines: a register-resident DFMA loop and
 * a register-resident DMMA (mma.sync m8n8k4 f64) loop on all SMs.  Results in TFLOP/s (host). */
int pgas_measure_fp64_peaks(double* dfma_tflops, double* dmma_tflops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PGAS_B200_H */
