/*
 * pnmol_b200.h -- C ABI of the B200-native EK1 filter loop (libpnmol_b200.so).
 *
 * The reference (schmidtjonathan/pnmol-experiments) is pure Python/JAX and has no FFI
 * layer; the drop-in boundary is its Python solver API.  This header is the C ABI the
 * Python host package (pnmol_b200) binds with ctypes, one entry point per reference
 * interface on the hot path.  Citations are file:line in the reference tree.
 *
 * Conventions
 *  - all floating point is IEEE double; matrices are row-major unless stated;
 *  - "dev" pointers are CUDA device pointers on the handle's device, "host" pointers are
 *    ordinary host memory; the caller owns every buffer it passes, the library owns only
 *    the workspace inside the handle;
 *  - every function returns 0 on success and a negative code on error, with a
 *    human-readable message available from pnmol_b200_last_error();
 *  - compute calls are asynchronous on the cudaStream_t passed as `stream`
 *    (NULL = legacy default stream); a handle is bound to one device and is not
 *    thread-safe; no exceptions or callbacks cross the boundary;
 *  - a per-member int32 status word is written by the kernels: 0 = finite results,
 *    1 = a non-finite value was produced (reference behaviour: NaNs propagate silently,
 *    tests/test_pdefilter.py:143-146 only checks for them afterwards).
 *
 * State layout (per ensemble member), identical to the reference's PDEFilterState
 * (src/pnmol/pdefilter.py:17-22, src/pnmol/base/rv.py:9-14):
 *    mean      (n, dd)  row i = i-th time derivative; dd = d (white) or 2d (latent, state
 *                       columns then latent-force columns: src/pnmol/latent.py:165-176);
 *    cov_sqrtm (D, D)   D = n*dd, lower triangular up to the sign of its columns.
 */
#ifndef PNMOL_B200_H
#define PNMOL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pnmol_b200_handle pnmol_b200_handle;

/* Solver kinds: the four PDEFilter subclasses of src/pnmol/white.py:169,189 and
 * src/pnmol/latent.py:237,266. */
enum {
    PNMOL_B200_WHITE_LINEAR = 0,
    PNMOL_B200_WHITE_SEMILINEAR = 1,
    PNMOL_B200_LATENT_LINEAR = 2,
    PNMOL_B200_LATENT_SEMILINEAR = 3
};

/* Point-wise reaction terms f / df shipped with the reference
 * (src/pnmol/pde/examples.py:151-165, 228-235, 311-315), evaluated on the device. */
enum {
    PNMOL_B200_REACTION_NONE = 0,
    PNMOL_B200_REACTION_SPRUCE = 1,        /* params: growth_rate               (1 component)  */
    PNMOL_B200_REACTION_SIR = 2,           /* params: beta, gamma               (3 components) */
    PNMOL_B200_REACTION_LOTKA_VOLTERRA = 3 /* params: a, b, c, d                (2 components) */
};

#define PNMOL_B200_MAX_REACTION_PARAMS 4

/* Flags for pnmol_b200_step / pnmol_b200_run. */
#define PNMOL_B200_FLAG_DENSE_FACTOR 1 /* input cov_sqrtm is not lower triangular */
#define PNMOL_B200_FLAG_NO_ERROR_ESTIMATE 2 /* skip white.py:153-162 (unused by Constant steps) */

const char* pnmol_b200_last_error(void);
int pnmol_b200_version(void);

/* Number of kernels this library has launched since it was loaded (bench.py's
 * gpu_launches claim). */
int64_t pnmol_b200_launch_count(void);

/* Diagnostics: with enable != 0 the step kernel accumulates clock64 cycles per phase (thread 0 of
 * every CTA) into 24 counters; a later call copies them to cycles_out host [24] (may be NULL)
 * and resets them.  Phases: 0 predict mean + evaluate_ode, 1 build predict stack, 2 predict QR,
 * 3 error estimate, 4 build update matrix, 5 update QR, 6 triangular solves + mean, 7 outputs.
 * The counters are only fed by a library compiled with -DPNMOL_PROFILE=1 (tools/build_variant.sh); the shipped
 * library carries no phase marks (they cost local-memory traffic on the critical path) and reports zeros. */
int pnmol_b200_profile(pnmol_b200_handle* h, int enable, uint64_t* cycles_out);

/* Which kernel family serves this handle (valid after pnmol_b200_set_operator):
 *   0 = one CTA per member (blocked QR with 16-column panels; row lists up to 512 rows; the default for ensembles),
 *   2 = one WARP per member, the whole QR workspace in the warp's slice of shared memory (small state dimension,
 *       D <~ 48: the meshes the reference tests and plots),
 *   1 = the whole grid per member (multi-CTA blocked QR with FP64 tensor-core trailing updates fed by TMA-staged
 *       reflector tiles; BASELINE configs C2-C4).
 * The environment variable PNMOL_B200_PATH=small|cta|large overrides the choice where the problem fits (used by
 * the parity tests to run every family on the same inputs).  <0 = error. */
int pnmol_b200_path(pnmol_b200_handle* h);

/* Diagnostics of the multi-CTA path: the thread-block cluster size the kernels are launched with (`requested`; the
 * panel factorisation runs on cluster 0) and the size the last pnmol_b200_run launch observed on the device
 * (cooperative_groups::this_cluster().num_blocks(); 0 before the first run or off the multi-CTA path). */
int pnmol_b200_cluster_size(pnmol_b200_handle* h, int* requested, int* observed);

/* Create a solver handle for `batch` independent members of one discretised problem.
 * Replaces PDEFilter.__init__ + the shape bookkeeping of initialize()
 * (src/pnmol/pdefilter.py:37-70, src/pnmol/white.py:17, src/pnmol/latent.py:52).
 *   d = pde.L.shape[0], nb = pde.B.shape[0], ncomp = number of PDE components,
 *   num_derivatives = nu.  device = CUDA ordinal. */
int pnmol_b200_create(pnmol_b200_handle** out, int kind, int d, int num_derivatives, int nb, int ncomp,
                      int batch, int reaction_id, int device);
int pnmol_b200_destroy(pnmol_b200_handle* h);

/* Discretised operator (host pointers, copied): the attributes the solvers read from the
 * problem object, src/pnmol/pde/mixins.py:37-54,98-117.
 *   L as ELL rows: L_col/L_val [d, wl] (col < 0 = padding); E_diag [d] = diag(pde.E_sqrtm);
 *   B as ELL rows: B_col/B_val [nb, wb]; R_sqrtm [nb, nb] dense. */
int pnmol_b200_set_operator(pnmol_b200_handle* h, const int32_t* L_col, const double* L_val, int wl,
                            const double* E_diag, const int32_t* B_col, const double* B_val, int wb,
                            const double* R_sqrtm);

/* IWP prior (host pointers, copied): A_1d, L_Q1d [n, n] of
 * IntegratedWienerTransition.preconditioned_discretize_1d (src/pnmol/base/iwp.py:13-30) and
 * Lk = chol(k(X, X)) [d, d] (src/pnmol/white.py:85, src/pnmol/latent.py:139). */
int pnmol_b200_set_prior(pnmol_b200_handle* h, const double* A1d, const double* LQ1d, const double* Lk);

/* Per-member ensemble axes (host pointers, copied; NULL = all ones / reaction defaults):
 *   diff_scale [batch, ncomp] multiplies L and E_sqrtm per component
 *   (src/pnmol/pde/mixins.py:37-38,87-89); prior_scale [batch] multiplies Lk;
 *   reaction_params [batch, nparams]. */
int pnmol_b200_set_members(pnmol_b200_handle* h, const double* diff_scale, const double* prior_scale,
                           const double* reaction_params, int nparams);

/* Householder row-support envelopes the kernels use for the two QR factorisations
 * (host-only arithmetic, no GPU needed; exported so that the structure model can be
 * tested on CPU).  te_p/be_p [D]: predict stack [(A Cl)^T ; Ql^T]; te_u/be_u [m + D]:
 * update block matrix of src/pnmol/base/sqrt.py:60-65.  Entry j = last (inclusive) row of
 * the top / bottom row segment that column j may touch. */
int pnmol_b200_structure(int kind, int d, int num_derivatives, int nb, const int32_t* L_col, int wl,
                         const int32_t* B_col, int wb, int ncomp, int dense_factor, int32_t* te_p,
                         int32_t* be_p, int32_t* te_u, int32_t* be_u);

/* initialize(pde) of src/pnmol/white.py:12-80 / src/pnmol/latent.py:20-134.
 *   y0 dev [batch, d]; mean_out dev [batch, n, dd]; chol_out dev [batch, D, D];
 *   status dev [batch] int32. */
int pnmol_b200_initialize(pnmol_b200_handle* h, const double* y0, double t0, double diffuse_prior_scale,
                          double* mean_out, double* chol_out, int32_t* status, void* stream);

/* attempt_step(state, dt, pde) of src/pnmol/white.py:96-146 / src/pnmol/latent.py:155-225.
 *   precond/precond_inv host [n]: nordsieck_preconditioner_1d_raw(dt)
 *   (src/pnmol/base/iwp.py:55-62), computed by the caller so that both sides use the same
 *   libm pow; t_new = state.t + dt is passed to the reaction term;
 *   err_out dev [batch, d] (error_estimate, white only), ref_out dev [batch, d]
 *   (reference_state = |m_new[0]|, white only), diff_out dev [batch]
 *   (diffusion_squared_local); err_out / ref_out may be NULL. */
int pnmol_b200_step(pnmol_b200_handle* h, double t_new, double dt, const double* precond,
                    const double* precond_inv, const double* mean_in, const double* chol_in,
                    double* mean_out, double* chol_out, double* err_out, double* ref_out, double* diff_out,
                    int32_t* status, int flags, void* stream);

/* The time loop of solution_generator / perform_full_step with step.Constant
 * (src/pnmol/pdefilter.py:118-227, src/pnmol/odetools/step.py:30-55) as ONE persistent
 * launch: nsteps steps with step sizes dts host [nsteps], preconditioners host
 * [nsteps, n] each.  mean / chol dev are updated in place (scratch = second state buffer
 * owned by the caller: mean_tmp, chol_tmp).  diff_sum dev [batch] receives the sum of the
 * local diffusions (for the mean of src/pnmol/pdefilter.py:95,113).  If mean_traj /
 * chol_traj are non-NULL they receive every step's state ([nsteps, batch, ...]).
 * Inside the launch only the last step writes the D x D factor to `chol` (unless a trajectory is requested): the
 * others hand it to their successor inside the workspace; flags bit 2 (value 4) makes every step go through the state
 * buffers (diagnostics; bitwise the same result).  The CTAs of the launch keep pace with each other after every step
 * (environment PNMOL_B200_PACE=0 switches that off; performance only). */
int pnmol_b200_run(pnmol_b200_handle* h, double t0, const double* dts, const double* precond,
                   const double* precond_inv, int nsteps, double* mean, double* chol, double* mean_tmp,
                   double* chol_tmp, double* err_out, double* ref_out, double* diff_last, double* diff_sum,
                   double* mean_traj, double* chol_traj, int32_t* status, int flags, void* stream);

/* Persistent multi-step run that returns MARGINALS instead of the factor trajectory:
 * mean_traj dev [nsteps, batch, n, dd] and std_traj dev [nsteps, batch, dd] with
 * std[j] = sqrt((E0 L L^T E0^T)[j][j]) = norm of row j n of the factor, the read-out of
 * experiments/figure1.py:76-89 and figure3.py:87-93 (read_mean_and_std*), fused into the step kernel so that the
 * T x D^2 factor trajectory of pnmol_b200_run is never written.  Other arguments as pnmol_b200_run. */
int pnmol_b200_run_marginals(pnmol_b200_handle* h, double t0, const double* dts, const double* precond,
                             const double* precond_inv, int nsteps, double* mean, double* chol, double* mean_tmp,
                             double* chol_tmp, double* diff_last, double* diff_sum, double* mean_traj, double* std_traj,
                             int32_t* status, int flags, void* stream);

/* Adaptive time loop on the device: solution_generator + perform_full_step with
 * step.Adaptive (src/pnmol/pdefilter.py:118-227, src/pnmol/odetools/step.py:58-119) for every member, each with its own
 * step size; accept/reject, the step-size proposal and the Nordsieck preconditioner are evaluated in the kernel.
 * White-noise solvers (the latent-force solvers have no error estimate); all three kernel families.
 *   dt0 dev [batch] first step (Adaptive.first_dt), mean/chol dev: state at t0 in, state at tmax out (unscaled factor;
 *   pnmol_b200_rescale applies the calibration with nsteps = num_steps), *_tmp scratch of the same shapes,
 *   t_out/dt_out/diff_sum/diff_last dev [batch], num_steps/num_attempts/status dev int32 [batch]
 *   (status bit 0: non-finite state, bit 1: max_attempts reached before tmax). */
int pnmol_b200_run_adaptive(pnmol_b200_handle* h, double t0, double tmax, const double* dt0, double abstol, double reltol,
                            double change_min, double change_max, double safety_scale, int max_attempts, double* mean,
                            double* chol, double* mean_tmp, double* chol_tmp, double* t_out, double* dt_out,
                            double* diff_sum, double* diff_last, int32_t* num_steps, int32_t* num_attempts,
                            int32_t* status, int flags, void* stream);

/* The same loop with the TRAJECTORY of the accepted states, i.e. solve() with step.Adaptive (src/pnmol/pdefilter.py:75-103,
 * 192-227) in one launch: slot s of mean_traj dev [max_traj, batch, n, dd] / chol_traj dev [max_traj, batch, D, D] /
 * t_traj dev [batch, max_traj] receives the state after accepted step s + 1 (unscaled factors).  A member that accepts
 * more than max_traj steps keeps going and sets status bit 2 (value 4: trajectory truncated; num_steps tells the
 * capacity a second call needs).  err_last / ref_last dev [batch, d] or NULL: error estimate and reference state of the
 * last attempted step (the PDEFilterState fields of the final state).  All trajectory pointers NULL = pnmol_b200_run_adaptive. */
int pnmol_b200_run_adaptive_trajectory(pnmol_b200_handle* h, double t0, double tmax, const double* dt0, double abstol,
                                       double reltol, double change_min, double change_max, double safety_scale,
                                       int max_attempts, double* mean, double* chol, double* mean_tmp, double* chol_tmp,
                                       double* t_out, double* dt_out, double* diff_sum, double* diff_last,
                                       int32_t* num_steps, int32_t* num_attempts, int32_t* status, double* err_last,
                                       double* ref_last, double* t_traj, double* mean_traj, double* chol_traj, int max_traj,
                                       int flags, void* stream);

/* Marginal standard deviations of `count` factors: chol dev [count, D, D] -> std_out dev [count, D / (nu + 1)]
 * (same read-out as above for states that already exist, e.g. the initial state or PDESolution.cov_sqrtm). */
int pnmol_b200_marginal_std(const double* chol, double* std_out, int D, int num_derivatives, int count, int device,
                            void* stream);

/* cov_sqrtm <- cov_sqrtm * sqrt(mean of the local diffusions), the last line of
 * simulate_final_state (src/pnmol/pdefilter.py:113-116).  chol dev [batch, D, D] in place,
 * diff_sum dev [batch] as returned by pnmol_b200_run, diff_cal_out dev [batch] or NULL
 * receives diffusion_squared_calibrated = diff_sum / nsteps. */
int pnmol_b200_rescale(pnmol_b200_handle* h, double* chol, const double* diff_sum, int nsteps,
                       double* diff_cal_out, void* stream);

/* simulate_final_state (src/pnmol/pdefilter.py:105-116) through HOST buffers: copies y0
 * host [batch, d] to the device, runs initialize + the constant-step loop, rescales the
 * factor by sqrt(mean local diffusion) and copies mean [batch, n, dd], chol [batch, D, D],
 * diffusion_calibrated [batch] and status [batch] back.  Blocks until done. */
int pnmol_b200_simulate_final_state_host(pnmol_b200_handle* h, const double* y0_host, double t0,
                                         double diffuse_prior_scale, const double* dts, const double* precond,
                                         const double* precond_inv, int nsteps, double* mean_host,
                                         double* chol_host, double* diff_cal_host, int32_t* status_host,
                                         int flags, void* stream);

/* propagate_cholesky_factor(S1, S2) of src/pnmol/base/sqrt.py:9-23 (batched like
 * batched_propagate_cholesky_factor, sqrt.py:27-29): S1 dev [batch, r, c1], S2 dev
 * [batch, r, c2] (c2 may be 0), out dev [batch, r, min(r, c1 + c2)]. */
int pnmol_b200_sqrt_propagate(const double* S1, const double* S2, double* out, int r, int c1, int c2,
                              int batch, int device, void* stream);

/* update_sqrt(H, C, meascov_sqrtm) / update_sqrt_no_meascov(H, C) of
 * src/pnmol/base/sqrt.py:34-95 for dense inputs: H dev [batch, m, D], C dev [batch, D, D],
 * meascov dev [batch, m, m] or NULL; outputs C_out [batch, D, D], K_out [batch, D, m],
 * S_out [batch, m, m]. */
int pnmol_b200_sqrt_update(const double* H, const double* C, const double* meascov, double* C_out,
                           double* K_out, double* S_out, int m, int D, int batch, int device, void* stream);

/* Square-root RTS smoother step (src/pnmol/base/kalman.py:49-66 smoother_step_sqrt):
 *   mean_out = m - sgain (mp - m_fut);  chol_out = (R[d:2d, d:])^T of the QR of [[x^T, sc^T], [sq^T, 0], [0, sc_fut^T sgain^T]].
 * All arrays dev float64 row-major with a leading batch dimension: m, m_fut, mp, mean_out [batch, d]; sc, sc_fut, sgain,
 * sq, x, chol_out [batch, d, d]. */
int pnmol_b200_smoother_step(const double* m, const double* sc, const double* m_fut, const double* sc_fut, const double* sgain,
                             const double* sq, const double* mp, const double* x, double* mean_out, double* chol_out, int d,
                             int batch, int device, void* stream);

/* Batched probabilistic finite-difference stencils (src/pnmol/discretize.py:177-201
 * fd_coefficients, vmapped over the mesh points as in fd_probabilistic :61-77 and, here, over `nbatch` kernel
 * hyper-parameter sets).  kernel_kind: 0 SquareExponential, 1 Matern52, 2 Polynomial (src/pnmol/kernels.py:107-144);
 * kernel_params dev [nbatch, 4] = (input_scale, output_scale, order, const); x dev [npoints] evaluation points,
 * neighbors dev [npoints, stencil] stencil points (1-D meshes; stencil <= 8); diffop: 0 gradient, 1 laplace.
 * Outputs: weights dev [nbatch, npoints, stencil] (rows of L), uncertainties dev [nbatch, npoints] (diagonal of
 * E_sqrtm).  Asynchronous on `stream`. */
int pnmol_b200_fd_coefficients(int kernel_kind, const double* kernel_params, int nbatch, const double* x,
                               const double* neighbors, int npoints, int stencil, int diffop, double nugget_gram_matrix,
                               double* weights, double* uncertainties, int device, void* stream);

/* Batched spatial Gram matrix and its Cholesky factor, chol(k(X, X) + diagonal_add I) (src/pnmol/white.py:82-94
 * initialize_iwp; diagonal_add = WhiteNoise output_scale^2 + nugget), and optionally the Gaussian log-likelihood of
 * `data` under each Gram matrix (src/pnmol/kernels.py:186-211 mle_input_scale / log_likelihood).  points dev [d]
 * (1-D mesh), data dev [d] or NULL, chol_out dev [nbatch, d, d] (lower, zeros above the diagonal), loglik_out dev
 * [nbatch] or NULL (needs d (d + 2) doubles of shared memory: d <= 168), status_out dev int32 [nbatch] or NULL (1 = not
 * positive definite).  Asynchronous on `stream`. */
int pnmol_b200_gram_cholesky(int kernel_kind, const double* kernel_params, int nbatch, const double* points, int d,
                             double diagonal_add, const double* data, double* chol_out, double* loglik_out,
                             int32_t* status_out, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PNMOL_B200_H */
