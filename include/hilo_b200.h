/* hilo_b200.h — C ABI of the B200 batched MPC / estimation solve engine.
 *
 * The reference (HILO-MPC) has no FFI seam on this path; its seams are Python callables created in setup() and
 * invoked once per control step:
 *   NMPC : solver = ca.nlpsol('solver','ipopt',{f,x,p,g},opts); sol = solver(x0=,lbx=,ubx=,lbg=,ubg=,p=)
 *          hilo_mpc/modules/controller/mpc.py:1778-1787, :722
 *   MHE  : same shape, hilo_mpc/modules/estimator/mhe.py:782-790, :375
 *   LMPC : ca.conic('solver','qpoases',{h,a}); solver(h=,g=,a=,lbx=,ubx=,lba=,uba=)   mpc.py:2268-2276, :2374
 *          (precedent for swapping the back-end: solver='muaompc', hilo_mpc/modules/embedded/muaompc.py:52-128)
 *   KF/EKF/UKF : self._function(x0=[x|P], y=, p=[u;p], Q=, R=)                          estimator/kf.py:261-265, :304-307
 *   PF   : self._function(X,y,p,w,v,R) + np.random.choice(n_s, n_s, p=q)                estimator/pf.py:324-328, :374, :406
 *   GP   : self._function(X=, x0=hyper, p=[X_train; y_train])                           machine_learning/gp/gp.py:623-629, :709
 * Each entry point below replaces one of those calls for a whole batch of independent instances.
 *
 * Conventions: plain pointers and sizes only; every data pointer is a CUDA DEVICE pointer owned by the caller
 * (e.g. a torch tensor) unless it says "host"; arrays are dense, row-major, fp64; functions return 0 on success,
 * <0 for API misuse, >0 = cudaError_t; nothing throws across the boundary; work is enqueued on `stream` and is
 * asynchronous with respect to the host unless documented otherwise.  Per-instance `status` uses the reference's
 * codes (hilo_mpc/modules/optimizer.py:1091-1104): 1 optimal, 2 acceptable, 5 max iterations, -1 failure.
 */
#ifndef HILO_B200_H
#define HILO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------ library */
#define HB_ABI_VERSION 8              /* bumps on every incompatible change of a struct or signature below */
int hb_version(void);                 /* HB_ABI_VERSION the library was built with; generated modules export the same
                                         number (hbm_abi_version) and hb_ocp_open / hb_est_open refuse a mismatch (-5) */
const char* hb_last_error(void);      /* host string describing the last failure on this thread */
int hb_device_info(int device, int* sm_count, size_t* total_mem, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------ NMPC / MHE
 * A transcribed optimal-control problem lives in a generated module (shared object, built by
 * hilo_mpc_b200/codegen.py from the Model/cost expressions at NMPC.setup()/MHE.setup() time; it contains the
 * model's residual/Jacobian device functions).  hb_ocp_open loads it; hb_ocp_solve = one NMPC.optimize /
 * MHE.estimate for B instances (replaces mpc.py:722 / mhe.py:375).                                            */
typedef struct hb_ocp_dims {
  int nx, nu, N, nc;        /* states, inputs, horizon, collocation states per interval (deg*nx or 0)           */
  int nsp, ntp, nuold, nap; /* per-stage parameters, terminal parameters, previous inputs, arrival parameters   */
  int stage_doubles;        /* doubles of one stage QP block [A B b Q S R q r] = 2nx^2+2nx*nu+nu^2+2nx+nu        */
  int elim_doubles;         /* doubles of one elimination block                                                  */
  int method;               /* 0 collocation, 1 explicit RK, 2 discrete                                         */
  int fixed_x0;             /* 1: x_0 pinned to the measured state (NMPC), 0: free (MHE)                        */
  int n_state;              /* doubles of per-instance solver statistics                                        */
  int block_a, block_h;     /* doubles of one stage block of qa = [A B b] / qh = [Q S R q r qbar rbar] as the sweep
                               reads them (hb_ocp_riccati)                                                      */
  int frag;                 /* 1: blocks in tensor-core operand order (nx <= 4, nu <= 2, x0 fixed): qa = the 4 x 8 matrix
                               [A B b 0] column-major, qh = the 6 x 8 tile [Q S^T q qbar; S R r rbar] row-major, zero
                               padded (ones on the diagonal of the R pad); 0: compact row-major blocks
                               [A | B | b] and [Q | S | R | q | r | qbar | rbar], padded to an even number of doubles */
} hb_ocp_dims;

typedef struct hb_ocp_call {
  int B;                    /* batch size                                                                        */
  int n_alpha;              /* line-search candidates evaluated in parallel (1 = full step to the boundary rule) */
  int max_iter;             /* interior-point iteration cap (IPOPT max_iter)                                     */
  int check_every;          /* host polls the active-instance count every this many iterations                  */
  double tol, acc_tol;      /* IPOPT tol / acceptable_tol                                                        */
  double mu_init, mu_min;   /* initial / final barrier parameter (IPOPT mu_init, tol/10)                         */
  double time0, dt;         /* time of node 0 and sampling interval                                              */
  const double* x0;         /* [B][nx] measured states (scaled); ignored when fixed_x0 == 0                      */
  const double* sp;         /* [B|1][N][nsp] per-stage parameters (tvp, cp, references, measurements)            */
  long long sp_stride;      /* doubles between instances, 0 = shared                                             */
  const double* tp;         /* [B|1][ntp] terminal parameters                                                    */
  long long tp_stride;
  const double* uold;       /* [B|1][nuold] previous inputs for the input-change cost                            */
  long long uold_stride;
  const double* ap;         /* [B|1][nap] arrival-cost parameters (MHE)                                          */
  long long ap_stride;
  const double *xlb, *xub, *ulb, *uub; /* [nx],[nx],[nu],[nu] bounds (|v| >= 1e19 = none)                        */
  double* x;                /* in/out [B][N+1][nx] node states: starting point in, solution out                  */
  double* u;                /* in/out [B][N][nu]                                                                 */
  double* xc;               /* in/out [B][N][nc] collocation states                                              */
  double* lam;              /* out    [B][N+1][nx] costates of the continuity constraints                        */
  int* status;              /* out    [B]                                                                         */
  int* iters;               /* out    [B+1] iterations per instance; slot B is scratch                           */
  double* stats;            /* out    [B][n_state] or NULL                                                       */
  double* work;             /* workspace of hb_ocp_work_doubles(B, n_alpha) doubles                              */
  size_t work_doubles;
  void* stream;             /* cudaStream_t                                                                      */
  int* n_iterations_launched; /* host out or NULL                                                                */
  long long* n_kernel_launches; /* host out or NULL: kernels launched by this call                                 */
  double* kernel_ms;        /* host in/out [HB_OCP_KERNELS] or NULL: accumulates device ms per kernel (CUDA events) */
  long long* kernel_count;  /* host in/out [HB_OCP_KERNELS] or NULL: accumulates launches per kernel              */
  int riccati_variant;      /* 0: one-thread-per-instance sweep (any size; the reference version of the sweep),
                               >= 2: fastest sweep the problem admits (default): warp-per-instance sweep on the fp64
                               tensor pipe for nx <= 4, nu <= 2, x0 fixed; cp.async ring with 8 lanes per instance
                               otherwise; 6 / 8 / 9: register-ring depth 4 / 5 / 3 of the former (tuning runs)      */
  int warm_start;           /* 0: IPOPT-default start (push 1e-2 into the bounds, z = 1, mu = mu_init);
                               1: x/u/xc AND the bound multipliers kept in `work` by the previous call are reused,
                                  tiny push, mu = mu_init_warm (needs the same `work` buffer and batch as before)    */
  double mu_init_warm;
  int eval_variant;         /* 1: 4-lanes-per-node register-resident evaluation kernel where supported (default),
                               0: one-thread-per-node kernel                                                       */
  int first_check;          /* > 0: first host poll after this many iterations, then after every further one (pass
                               *n_iterations_needed of the previous solve of the same closed loop); 0: every check_every */
  int* n_iterations_needed; /* host out or NULL: largest per-instance iteration count + 1                          */
  int step_variant;         /* expand / line search / apply of one iteration: 2: one kernel, one CTA per instance,
                               4 lanes per node, iterate register-resident (default where supported); 1: one kernel,
                               one thread per node; 0: separate kernels                                            */
  int no_compaction;        /* 0 (default): once a poll finds at most half of the batch active, the node / sweep / step
                               kernels run over the list of those instances; 1: always over the whole batch        */
  int regularization;       /* treatment of negative curvature of the exact Lagrangian Hessian.  0 (default): IPOPT's
                               inertia correction - the sweep is repeated with delta_w * I on the primal variables
                               until every Huu is positive definite (Waechter & Biegler 2006, algorithm IC);
                               1: additionally every reduced stage Hessian [Q S^T; S R] is convexified before the sweep
                               by mirroring its negative eigenvalues (|lambda|, floor regularization_eps)            */
  double regularization_eps;
  double regularization_mu;         /* regularization == 1 acts only while the instance's barrier parameter is above this
                                       value (far from the solution); below it the exact Hessian is kept and the inertia
                                       correction alone guards the sweep (fast local convergence).  0: always           */
  int shift_warm_start;             /* with warm_start = 1: the kept solution (primal variables, bound multipliers,
                                       costates) is moved one stage towards the present before the solve (receding-
                                       horizon shift; the last stage keeps its values).  0 (default): the previous
                                       solution as it is, like the reference (mpc.py:725-726)                            */
  double* kernel_ms_by_iter;        /* host in/out [HB_OCP_TIMED_ITERS][HB_OCP_KERNELS] or NULL: kernel_ms split by the iteration
                                       index of the solve (last row = that index and beyond).  Iteration 0 of a solve is the
                                       one launch in which every instance of the batch is active: the roofline of a kernel
                                       is quoted on it                                                                   */
  long long* kernel_count_by_iter;  /* host in/out, same shape, or NULL                                                  */
} hb_ocp_call;

/* kernel indices of kernel_ms / kernel_count */
#define HB_OCP_KERNELS 12
#define HB_OCP_TIMED_ITERS 8
#define HB_OCP_KERNEL_NAMES "init,eval,riccati,expand,prep,trial,decide,apply,commit,count,step,ricfix"

/* The Riccati sweep of one interior-point iteration alone (unit tests, profiling): the block-banded KKT system
 *   min sum_k 1/2 [dx;du]^T [Q S^T; S R] [dx;du] + [q;r]^T [dx;du] + terminal,  dx_{k+1} = A dx_k + B du_k + b,  dx_0 = 0
 * of B instances, stage data in the module's block layout (hb_ocp_dims.block_a / block_h / frag).                      */
typedef struct hb_ocp_riccati_call {
  int B;
  int riccati_variant;      /* as in hb_ocp_call */
  const double* qa;         /* [B][N][block_a] */
  const double* qh;         /* [B][N][block_h] */
  const double* qpN;        /* [B][nx*nx + 2 nx] terminal block [Q_N | q_N | qbar_N] */
  double* dx;               /* out [B][N+1][nx] */
  double* du;               /* out [B][N][nu]   */
  double* lam;              /* out [B][N+1][nx] adjoint of the multiplier-form gradient (lam_N = qbar_N) */
  double* edual;            /* out [B] max_k |rbar_k + B_k^T lam_{k+1}| (dual infeasibility of the inputs) */
  int* status;              /* out [B] 0 ok, -1 factorisation failed */
  double* work;             /* hb_ocp_riccati_work_doubles(B) doubles */
  size_t work_doubles;
  void* stream;
} hb_ocp_riccati_call;

typedef struct hb_ocp* hb_ocp_t;
int hb_ocp_open(const char* module_path, hb_ocp_t* out);
int hb_ocp_get_dims(hb_ocp_t h, hb_ocp_dims* dims);
size_t hb_ocp_work_doubles(hb_ocp_t h, int B, int n_alpha);
int hb_ocp_solve(hb_ocp_t h, const hb_ocp_call* call);   /* synchronises `stream` only to poll convergence */
size_t hb_ocp_riccati_work_doubles(hb_ocp_t h, int B);
int hb_ocp_riccati(hb_ocp_t h, const hb_ocp_riccati_call* call);   /* asynchronous on `stream` */
int hb_ocp_close(hb_ocp_t h);

/* ------------------------------------------------------------------------------------------------ estimators
 * A generated estimator module holds the model's f / h device functions and their Jacobians.  One call = one
 * estimate() of the reference for B independent filters:
 *   hb_ekf_step  KF / EKF predict + update   (estimator/kf.py:71-186, 304-307)
 *   hb_ukf_step  UKF predict + update        (estimator/kf.py:486-604)
 *   hb_pf_step   PF propagate / weight / resample / moments (estimator/pf.py:103-166, 363-421)
 *   hb_sim_step  one sampling interval of Model.simulate   (dynamic_model/dynamic_model.py:3911-4212)           */
typedef struct hb_est_dims { int nx, nu, np, ny, nye, discrete; } hb_est_dims;

typedef struct hb_est_call {
  int B;                     /* number of filter instances                                                       */
  int substeps;              /* RK4 sub-steps per sampling interval for continuous models                        */
  int n_particles;           /* PF only                                                                          */
  int do_predict, do_update; /* KF.predict()/update() pass-throughs (kf.py:309-325); estimate() sets both        */
  double dt, t0;
  double alpha, beta, kappa; /* UKF tuning (kf.py:428)                                                           */
  double* x;                 /* in/out [B][nx]                                                                    */
  double* P;                 /* in/out [B][nx][nx]                                                                */
  double* X;                 /* UKF: in/out [B][2nx+1][nx] sigma points; PF: in/out [B][n_particles][nx]          */
  const double* y;           /* [B][nye] measurements                                                             */
  const double* u;           /* [B|1][nu]                                                                         */
  long long u_stride;
  const double* p;           /* [B|1][np]                                                                         */
  long long p_stride;
  const double* Q;           /* [nx][nx] process noise covariance                                                 */
  const double* R;           /* [nye][nye] measurement noise covariance                                           */
  double* y_pred;            /* out [B][nye]                                                                      */
  const double* w;           /* PF: [B][n_particles][nx] process-noise samples                                    */
  const double* v;           /* PF: [B][n_particles][nye] measurement-noise samples                               */
  const double* unif;        /* PF: [B][n_particles] uniforms consumed by the resampling                          */
  double* q;                 /* PF out: [B][n_particles] normalised weights                                       */
  long long* idx;            /* PF out: [B][n_particles] resampled indices (== np.random.choice on the same RNG)  */
  double* work;              /* PF workspace of hb_est_work_doubles(B, n_particles) doubles                       */
  size_t work_doubles;
  int* status;               /* out [B]: 0 ok, -1 numerical failure                                               */
  void* stream;
  int variant;               /* UKF: 1 (default of the Python shell) one lane per sigma point, reductions by warp shuffles;
                                PF: 1 one warp per filter, particles on the lanes (bit-identical results);
                                0 one thread per filter                                                           */
} hb_est_call;

typedef struct hb_est* hb_est_t;
int hb_est_open(const char* module_path, hb_est_t* out);
int hb_est_get_dims(hb_est_t h, hb_est_dims* dims);
size_t hb_est_work_doubles(hb_est_t h, int B, int n_particles);
int hb_sim_step(hb_est_t h, const hb_est_call* call);
int hb_ekf_step(hb_est_t h, const hb_est_call* call);
int hb_ukf_step(hb_est_t h, const hb_est_call* call);
int hb_pf_step(hb_est_t h, const hb_est_call* call);
int hb_est_close(hb_est_t h);

/* Model-independent resampling (replaces np.random.choice(n, n, p=q), estimator/pf.py:406): q [B][n], unif [B][n],
 * cdf_work [B][n] scratch, idx [B][n] out; indices are bit-identical to NumPy's for the same q and uniforms.      */
int hb_pf_resample(int B, int n, const double* q, const double* unif, double* cdf_work, long long* idx, void* stream);

/* ------------------------------------------------------------------------------------------------ linear MPC
 * LMPC.optimize for B instances (replaces ca.conic('solver','qpoases',{h,a}) + solver(h=,g=,a=,lbx=,ubx=,lba=,uba=),
 * mpc.py:2268-2276, :2374): the horizon is condensed once on the host (LMPC.setup),
 *     x_pred = Phi x0 + Gam u,   H = Gam^T Qbar Gam + Rbar [n][n],   F = Gam^T Qbar Phi [n][nx],
 * n = N nu, mx = N nx, and every instance solves the dense QP
 *     min_u 1/2 u^T H u + (F x0)^T u   s.t.  ulb <= u <= uub [n],  xlb <= Phi x0 + Gam u <= xub [mx]
 * with a primal-dual interior-point method (relative KKT tolerance tol; |bound| >= 1e19 = none).  The equality block of the
 * reference's QP enters only through Phi / Gam, so both the intended input block I_N (x) B and the reference's kron(B, I_N)
 * (mpc.py:2243, LMPC(bug_compat=True)) are served by the same entry point.
 * u [B][n] out, xpred [B][mx] out or NULL, status [B]: 1 optimal, 5 iteration cap, -1 numerical failure.            */
size_t hb_lmpc_work_doubles(int B, int n, int mx);
int hb_lmpc_solve(int B, int n, int mx, int nx, const double* H, const double* F, const double* Gam, const double* Phi,
                  const double* ulb, const double* uub, const double* xlb, const double* xub, const double* x0,
                  double* u, double* xpred, int* status, int* iters, double tol, int max_iter, double* work,
                  size_t work_doubles, void* stream);

/* ------------------------------------------------------------------------------------------------ Gaussian process
 * Exact inference with the squared-exponential kernel (machine_learning/gp/inference.py:197-217, kernel.py:650-734).
 * X is (d, n) row-major like the reference's training data (gp.py:347-354); hyper-parameters are passed in natural
 * units (signal variance sigma_f^2, length scales l_k, noise variance sigma_n^2).
 *   hb_gp_fit     : L = chol(sigma_n^2 I + K) [n][n], Linv = L^-1 [n][n], alpha [n], lml [1]; info [1] != 0 when the
 *                   matrix is not positive definite (index of the failing pivot + 1)
 *   hb_gp_predict : mean [M] and (if var != NULL) variance [M] at the M query points Xq (d, M); replaces
 *                   GaussianProcess.predict -> self._function(X=, x0=hyper, p=[X_train; y_train])  (gp.py:623-629, :709) */
size_t hb_gp_fit_work_doubles(int n);
int hb_gp_fit(int n, int d, const double* X, const double* y, const double* length_scales, double signal_variance,
              double noise_variance, double prior_mean, double* L, double* Linv, double* alpha, double* lml,
              double* work, int* info, void* stream);
size_t hb_gp_predict_work_doubles(int n, int M);
int hb_gp_predict(int n, int d, int M, const double* X, const double* Xq, const double* length_scales,
                  double signal_variance, double noise_variance, double prior_mean, const double* Linv,
                  const double* alpha, double* mean, double* var, double* work, void* stream);
/* The same with another stationary covariance function of r^2 = sum_k ((x_k - x'_k) / l_k)^2: kind 0 = squared
 * exponential, 1 = exponential, 2 = Matern 3/2, 3 = Matern 5/2 (MaternKernel.get_covariance_function with p = 0, 1, 2,
 * kernel.py:783-826; classes :829-917).                                                                            */
int hb_gp_fit_kernel(int kind, int n, int d, const double* X, const double* y, const double* length_scales,
                     double signal_variance, double noise_variance, double prior_mean, double* L, double* Linv,
                     double* alpha, double* lml, double* work, int* info, void* stream);
int hb_gp_predict_kernel(int kind, int n, int d, int M, const double* X, const double* Xq, const double* length_scales,
                         double signal_variance, double noise_variance, double prior_mean, const double* Linv,
                         const double* alpha, double* mean, double* var, double* work, void* stream);
/* Gradient of the NEGATIVE log marginal likelihood with respect to the log hyper-parameters, from the factorisation
 * hb_gp_fit_kernel left behind: grad [d + 2] = [d/dlog sigma_n^2 | d/dlog l_1 .. l_d | d/dlog sigma_f^2]
 * = 1/2 tr((K^-1 - alpha alpha^T) dK/dtheta_j).  Replaces the automatic differentiation of the likelihood graph inside
 * GaussianProcess.fit_model (gp.py:660-697, inference.py:197-217).  work: hb_gp_lml_grad_work_doubles(n) doubles.   */
size_t hb_gp_lml_grad_work_doubles(int n);
int hb_gp_lml_grad(int kind, int n, int d, const double* X, const double* length_scales, double signal_variance,
                   double noise_variance, const double* Linv, const double* alpha, double* work, double* grad,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HILO_B200_H */
