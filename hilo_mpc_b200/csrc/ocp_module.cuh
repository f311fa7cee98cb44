// C entry points of one generated OCP module (one .so per transcribed problem, built by hilo_mpc_b200/codegen.py).
// The core library (hilo_b200.cu) dlopens the module and forwards hb_ocp_* calls to these hbm_* symbols.
#pragma once
#include "ocp_kernels.cuh"
#include "riccati2.cuh"
#include "riccati6.cuh"
#include "eval_quad.cuh"
#include "step_fused.cuh"
#include "step_quad.cuh"
#include "../../include/hilo_b200.h"

namespace hb {

struct OcpCount {  // (b): poll[0] = instances still active (their indices appended to alist, in any order), poll[1] =
                   // largest iteration count of the batch
  static HB_HD void run(const OcpArgs& a, int b) {
    const int it = a.iters[b];
#if HB_DEVICE_BUILD && defined(__CUDA_ARCH__)
    if (a.st[(long long)b * ST_COUNT + ST_ACTIVE] != 0.0) a.alist[atomicAdd(a.poll, 1)] = b;
    if (it > a.poll[1]) atomicMax(a.poll + 1, it);
#else
    if (a.st[(long long)b * ST_COUNT + ST_ACTIVE] != 0.0) { a.alist[a.poll[0]] = b; a.poll[0] += 1; }
    if (it > a.poll[1]) a.poll[1] = it;
#endif
  }
};

// Per-kernel device timing with CUDA events recorded on the launching stream (hb_ocp_call.kernel_ms != NULL).
// Events are only read back at the points where the driver synchronises anyway (convergence polls).
enum { K_INIT = 0, K_EVAL, K_RICCATI, K_EXPAND, K_PREP, K_TRIAL, K_DECIDE, K_APPLY, K_COMMIT, K_COUNT_ACTIVE, K_STEP, K_RICFIX,
       K_NUM };

struct KernelTimer {
  bool on = false;
  cudaStream_t s = nullptr;
#if HB_DEVICE_BUILD
  static constexpr int kPool = 512;
  cudaEvent_t ev0[kPool], ev1[kPool];
  int kind[kPool], iter_of[kPool];
  int n = 0;
  bool created = false;
#endif
  double* ms = nullptr;        // [K_NUM] accumulated milliseconds
  long long* count = nullptr;  // [K_NUM] launches
  double* ms_it = nullptr;     // [HB_OCP_TIMED_ITERS][K_NUM] the same split by iteration index of the solve, or NULL
  long long* count_it = nullptr;
  int it = 0;                  // iteration index of the launches being recorded
  long long launches = 0;

  void begin(int k) {
    ++launches;
    if (count) count[k] += 1;
    if (count_it) count_it[(it < HB_OCP_TIMED_ITERS ? it : HB_OCP_TIMED_ITERS - 1) * K_NUM + k] += 1;
#if HB_DEVICE_BUILD
    if (!on) return;
    if (!created) { for (int i = 0; i < kPool; ++i) { cudaEventCreate(&ev0[i]); cudaEventCreate(&ev1[i]); } created = true; }
    if (n == kPool) flush(true);
    kind[n] = k;
    iter_of[n] = it < HB_OCP_TIMED_ITERS ? it : HB_OCP_TIMED_ITERS - 1;
    cudaEventRecord(ev0[n], s);
#endif
  }
  void end() {
#if HB_DEVICE_BUILD
    if (!on) return;
    cudaEventRecord(ev1[n], s);
    ++n;
#endif
  }
  void flush(bool sync) {
#if HB_DEVICE_BUILD
    if (!on || n == 0) return;
    if (sync) cudaStreamSynchronize(s);
    for (int i = 0; i < n; ++i) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, ev0[i], ev1[i]) == cudaSuccess && ms) {
        ms[kind[i]] += (double)t;
        if (ms_it) ms_it[iter_of[i] * K_NUM + kind[i]] += (double)t;
      }
    }
    n = 0;
#endif
  }
  ~KernelTimer() {
#if HB_DEVICE_BUILD
    if (created) for (int i = 0; i < kPool; ++i) { cudaEventDestroy(ev0[i]); cudaEventDestroy(ev1[i]); }
#endif
  }
};

template <class P>
struct OcpModule {
  using T = OcpT<P>;

  static void dims(hb_ocp_dims* d) {
    d->nx = T::NX; d->nu = T::NU; d->N = T::N; d->nc = T::NC; d->nsp = T::NSP; d->ntp = T::NTP;
    d->nuold = T::NUOLD; d->nap = T::NAP; d->stage_doubles = T::NA + T::NH; d->elim_doubles = T::TD;
    d->method = P::METHOD; d->fixed_x0 = P::FIXED_X0 ? 1 : 0; d->n_state = ST_COUNT;
    d->block_a = T::NAp; d->block_h = T::NHp; d->frag = T::FRAG ? 1 : 0;
  }

  // the sweep of one iteration, fastest version the problem admits (riccati_variant: see hb_ocp_call)
  // fix_in_step: the step kernel that follows redoes flagged (non-convex) instances itself (step_quad.cuh), so the
  // separate fix-up launch is skipped
  static int launch_riccati(const OcpArgs& a, int variant, cudaStream_t s, KernelTimer* tm = nullptr,
                            bool fix_in_step = false) {
    bool ric_done = false;
    if (variant >= 2) {
      int rc6 = -100;
      auto try6 = [&](auto tag) {
        using R = decltype(tag);
        if constexpr (R::SUPPORTED) rc6 = launch_riccati6<R>(a, s);
      };
      switch (variant) {
        case 6: try6(Riccati6<P, 4>{}); break;
        case 8: try6(Riccati6<P, 5>{}); break;
        case 9: try6(Riccati6<P, 3>{}); break;
        default: try6(Riccati6<P, 2>{}); break;
      }
      if (rc6 > 0) return rc6;
      ric_done = rc6 == 0;
      if (ric_done && !fix_in_step) {   // instances flagged as non-convex (rare): timed as its own kernel
        if (tm) { tm->end(); tm->begin(K_RICFIX); }
        launch_1d<OcpRiccatiFix<P>>(a, a.n_instances(), s);
      }
    }
#if HB_DEVICE_BUILD
    if constexpr (RiccatiTma2<P>::SUPPORTED) {
      if (variant >= 2 && !ric_done) {
        cudaError_t e = launch_riccati_tma2<P>(a, s);
        if (e != cudaSuccess) return (int)e;
        ric_done = true;
      }
    }
#endif
    if (!ric_done) launch_1d<OcpRiccati<P>>(a, a.B, s);
    return 0;
  }

  struct RiccatiOnlyInit {  // (b): solver state of a stand-alone sweep
    static HB_HD void run(const OcpArgs& a, int b) {
      double* st = a.st + (long long)b * ST_COUNT;
      for (int i = 0; i < ST_COUNT; ++i) st[i] = 0.0;
      st[ST_ACTIVE] = 1.0;
      a.status[b] = 0;
    }
  };
  struct RiccatiOnlyFinish {  // (b): edual out
    static HB_HD void run(const OcpArgs& a, int b) { a.s_cinf[b] = a.st[(long long)b * ST_COUNT + ST_EDUAL]; }
  };
  static size_t riccati_work_doubles(int B) {
    return (size_t)B * ST_COUNT + (size_t)B * T::N * T::KD + 64;
  }
  static int riccati(const hb_ocp_riccati_call* c) {
    if (c == nullptr || c->B <= 0) return -1;
    if (c->work == nullptr || c->work_doubles < riccati_work_doubles(c->B)) return -2;
    if (!c->qa || !c->qh || !c->qpN || !c->dx || !c->du || !c->lam || !c->edual || !c->status) return -1;
    OcpArgs a;
    memset(&a, 0, sizeof(a));
    a.B = c->B;
    a.qa = const_cast<double*>(c->qa); a.qh = const_cast<double*>(c->qh); a.qpN = const_cast<double*>(c->qpN);
    a.dx = c->dx; a.du = c->du; a.lam = c->lam; a.status = c->status;
    double* w = c->work;
    w += ((128 - (reinterpret_cast<uintptr_t>(w) & 127)) & 127) / 8;
    a.st = w; w += ((size_t)c->B * ST_COUNT + 15) & ~(size_t)15;
    a.fb = w;
    a.s_cinf = c->edual;   // carrier of the per-instance result
    cudaStream_t s = (cudaStream_t)c->stream;
    launch_1d<RiccatiOnlyInit>(a, a.B, s);
    const int rc = launch_riccati(a, c->riccati_variant, s);
    if (rc != 0) return rc;
    launch_1d<RiccatiOnlyFinish>(a, a.B, s);
#if HB_DEVICE_BUILD
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
#else
    return 0;
#endif
  }

  // doubles of workspace needed for a batch of B instances with n_alpha line-search candidates
  static size_t work_doubles(int B, int n_alpha) {
    const size_t b = (size_t)B, n1 = T::N + 1, n = T::N;
    size_t w = 0;
    w += 2 * b * n1 * T::NX + 2 * b * n * T::NU + 2 * b * n * T::NC;       // z multipliers
    w += b * n * T::SD + b * T::SDN + b * n * T::TD + b * n * T::KD;       // qa+qh, qpN, el, fb
    w += b * n1 * T::NX + b * n * T::NU + b * n * T::NC;                   // dx, du, dxc
    w += 7 * b * n1;                                                       // per-node scalars
    w += 2 * b * (size_t)n_alpha * n1;                                     // line-search table
    w += b * ST_COUNT + (b * n + 1) / 2 + 2 + 2 + (b + 1) / 2 + 1;         // state, node flags, poll counters, active list
    return w + 32 * 16 + 16;  // slack: every array starts on a 128-byte line (30 arrays), so does the first one
  }

  static int solve(const hb_ocp_call* c) {
    if (c == nullptr || c->B <= 0) return -1;
    if (c->work == nullptr || c->work_doubles < work_doubles(c->B, c->n_alpha)) return -2;
    if (c->n_alpha < 1 || c->n_alpha > 32) return -3;
    OcpArgs a;
    memset(&a, 0, sizeof(a));
    a.B = c->B; a.n_alpha = c->n_alpha; a.max_iter = c->max_iter;
    a.tol = c->tol; a.acc_tol = c->acc_tol; a.mu_init = c->mu_init; a.mu_min = c->mu_min;
    a.time0 = c->time0; a.dt = c->dt; a.warm = c->warm_start;
    a.convexify = c->regularization == 1 ? 1 : 0;
    a.convexify_eps = c->regularization_eps > 0.0 ? c->regularization_eps : 1e-4;
    a.convexify_mu = c->regularization_mu;
    if (a.warm) a.mu_init = c->mu_init_warm;
    a.x0 = c->x0; a.sp = c->sp; a.sp_stride = c->sp_stride; a.tp = c->tp; a.tp_stride = c->tp_stride;
    a.uold = c->uold; a.uold_stride = c->uold_stride; a.ap = c->ap; a.ap_stride = c->ap_stride;
    a.xlb = c->xlb; a.xub = c->xub; a.ulb = c->ulb; a.uub = c->uub;
    a.x = c->x; a.u = c->u; a.xc = c->xc; a.lam = c->lam;
    a.status = c->status; a.iters = c->iters;
    const size_t b = (size_t)c->B, n1 = T::N + 1, n = T::N;
    double* w = c->work;
    w += ((128 - (reinterpret_cast<uintptr_t>(w) & 127)) & 127) / 8;   // 128-byte lines: vector loads, L2 line prefetches
    auto take = [&](size_t cnt) { double* p = w; w += (cnt + 15) & ~(size_t)15; return p; };
    a.zlx = take(b * n1 * T::NX); a.zux = take(b * n1 * T::NX);
    a.zlu = take(b * n * T::NU); a.zuu = take(b * n * T::NU);
    a.zlc = take(b * n * T::NC); a.zuc = take(b * n * T::NC);
    a.qa = take(b * n * T::NAp); a.qh = take(b * n * T::NHp); a.qpN = take(b * T::SDN); a.el = take(b * n * T::TD); a.fb = take(b * n * T::KD);
    a.dx = take(b * n1 * T::NX); a.du = take(b * n * T::NU); a.dxc = take(b * n * T::NC);
    a.s_cinf = take(b * n1); a.s_c1 = take(b * n1); a.s_phi = take(b * n1); a.s_comp = take(b * n1);
    a.s_ap = take(b * n1); a.s_ad = take(b * n1); a.s_dphi = take(b * n1);
    a.ls_phi = take(b * (size_t)c->n_alpha * n1); a.ls_c1 = take(b * (size_t)c->n_alpha * n1);
    a.st = take(b * ST_COUNT);
    a.nflag = reinterpret_cast<int*>(take((b * n + 1) / 2 + 1));
    a.poll = reinterpret_cast<int*>(take(2));
    a.alist = reinterpret_cast<int*>(take((b + 1) / 2 + 1));
    a.n_list = 0;
    cudaStream_t s = (cudaStream_t)c->stream;

    static thread_local KernelTimer tm;
    tm.on = c->kernel_ms != nullptr; tm.s = s; tm.ms = c->kernel_ms; tm.count = c->kernel_count; tm.launches = 0;
    tm.ms_it = c->kernel_ms_by_iter; tm.count_it = c->kernel_count_by_iter; tm.it = 0;
    tm.begin(K_INIT);
    if (c->warm_start && c->shift_warm_start) launch_2d<OcpShift<P>>(a, a.B, OcpShift<P>::PER, s);
    launch_2d<OcpInit<P>>(a, a.B, OcpInit<P>::ETOT, s);
    tm.end();
    // Host polling schedule: every check_every iterations, or - with a hint from the previous solve of the same closed
    // loop (first_check = iterations that one needed) - once after that many and then after every further iteration.
    // With compaction the first poll comes one iteration earlier: the last iterations of a solve belong to a few
    // stragglers (8 % of the instances in the BASELINE loop), and a launch over the whole batch costs half of a full one
    // however few are active, whereas a launch over the list costs about the latency of one CTA.
    const int check_every = c->check_every > 0 ? c->check_every : 4;
    const int hint = (c->first_check > 0 && c->first_check <= 64) ? c->first_check : 0;
    int next_check = hint > 0 ? ((c->no_compaction == 0 && hint > 2) ? hint - 1 : hint) : check_every;
    int launched = 0, needed = 0, single_polls = 0;
    for (int it = 0; it <= c->max_iter; ++it) {
      tm.it = it;
      tm.begin(K_EVAL);
#if HB_DEVICE_BUILD
      if constexpr (EvalQuad<P>::SUPPORTED) {
        if (c->eval_variant != 0) {
          cudaError_t e = launch_eval_quad<P>(a, s);
          if (e != cudaSuccess) return (int)e;
          if (a.convexify) launch_2d<OcpConvexify<P>>(a, a.n_instances(), T::N, s);
        } else {
          launch_2d<OcpEval<P>>(a, a.B, T::N + 1, s);
        }
      } else {
        launch_2d<OcpEval<P>>(a, a.B, T::N + 1, s);
      }
#else
      launch_2d<OcpEval<P>>(a, a.B, T::N + 1, s);
#endif
      tm.end();
      tm.begin(K_RICCATI);
      {
        // (folding the fix-up of flagged instances into the step kernel was measured: its cold call costs step_quad
        //  45 us per launch in spills - the 3 us launch stays)
        const int rcr = launch_riccati(a, c->riccati_variant, s, &tm, false);
        if (rcr != 0) return rcr;
      }
      tm.end();
      bool fused = false;
#if HB_DEVICE_BUILD
      if constexpr (StepQuad<P>::SUPPORTED) {
        if (c->step_variant >= 2) {
          tm.begin(K_STEP);
          cudaError_t e = launch_step_quad<P>(a, s);
          tm.end();
          if (e != cudaSuccess) return (int)e;
          fused = true;
        }
      }
      if constexpr (StepFused<P>::SUPPORTED) {
        if (c->step_variant != 0 && !fused) {
          tm.begin(K_STEP);
          cudaError_t e = launch_step_fused<P>(a, s);
          tm.end();
          if (e != cudaSuccess) return (int)e;
          fused = true;
        }
      }
#endif
      if (!fused) {
        tm.begin(K_EXPAND); launch_2d<OcpExpand<P>>(a, a.B, T::N + 1, s); tm.end();
        tm.begin(K_PREP); launch_1d<OcpPrep<P>>(a, a.B, s); tm.end();
        a.ls_phase = 0;
        tm.begin(K_TRIAL); launch_3d<OcpTrial<P>>(a, a.B, T::N + 1, 1, s); tm.end();
        tm.begin(K_DECIDE); launch_1d<OcpDecide<P>>(a, a.B, s); tm.end();
        if (a.n_alpha > 1) {
          a.ls_phase = 1;
          tm.begin(K_TRIAL); launch_3d<OcpTrial<P>>(a, a.B, T::N + 1, a.n_alpha - 1, s); tm.end();
          tm.begin(K_DECIDE); launch_1d<OcpDecide<P>>(a, a.B, s); tm.end();
        }
        tm.begin(K_APPLY); launch_2d<OcpApply<P>>(a, a.B, T::N + 1, s); tm.end();
        tm.begin(K_COMMIT); launch_1d<OcpCommit<P>>(a, a.B, s); tm.end();
      }
      ++launched;
      if (launched == next_check || it == c->max_iter) {
        int poll[2] = {0, 0};
#if HB_DEVICE_BUILD
        cudaMemsetAsync(a.poll, 0, 2 * sizeof(int), s);
        tm.begin(K_COUNT_ACTIVE); launch_1d<OcpCount>(a, a.B, s); tm.end();
        cudaMemcpyAsync(poll, a.poll, 2 * sizeof(int), cudaMemcpyDeviceToHost, s);
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return (int)e;
        tm.flush(false);
#else
        a.poll[0] = 0; a.poll[1] = 0;
        launch_1d<OcpCount>(a, a.B, s);
        poll[0] = a.poll[0]; poll[1] = a.poll[1];
#endif
        needed = poll[1] + 1;  // the iteration after the last step is the one that detects convergence
        if (poll[0] == 0) break;
        // stragglers: run the following iterations over the list of active instances the count kernel just wrote
        a.n_list = (c->no_compaction == 0 && 2 * poll[0] <= a.B) ? poll[0] : 0;
        if (hint > 0 && single_polls < 3) { next_check = launched + 1; ++single_polls; }
        else next_check = launched + check_every;
      }
    }
    if (c->stats != nullptr) {
#if HB_DEVICE_BUILD
      cudaMemcpyAsync(c->stats, a.st, b * ST_COUNT * sizeof(double), cudaMemcpyDeviceToDevice, s);
#else
      memcpy(c->stats, a.st, b * ST_COUNT * sizeof(double));
#endif
    }
    if (c->n_iterations_launched != nullptr) *c->n_iterations_launched = launched;
    if (c->n_iterations_needed != nullptr) *c->n_iterations_needed = needed;
    if (c->n_kernel_launches != nullptr) *c->n_kernel_launches = tm.launches;
#if HB_DEVICE_BUILD
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
#else
    return 0;
#endif
  }
};

}  // namespace hb

#define HB_DEFINE_OCP_MODULE(PROB)                                                                        \
  extern "C" int hbm_abi_version() { return HB_ABI_VERSION; }                                             \
  extern "C" void hbm_ocp_dims(hb_ocp_dims* d) { hb::OcpModule<PROB>::dims(d); }                          \
  extern "C" size_t hbm_ocp_work_doubles(int B, int n_alpha) {                                            \
    return hb::OcpModule<PROB>::work_doubles(B, n_alpha);                                                 \
  }                                                                                                       \
  extern "C" int hbm_ocp_solve(const hb_ocp_call* c) { return hb::OcpModule<PROB>::solve(c); }                   \
  extern "C" size_t hbm_ocp_riccati_work_doubles(int B) { return hb::OcpModule<PROB>::riccati_work_doubles(B); } \
  extern "C" int hbm_ocp_riccati(const hb_ocp_riccati_call* c) { return hb::OcpModule<PROB>::riccati(c); }
