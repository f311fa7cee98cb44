// C entry points of one generated estimator / simulation module (see est_kernels.cuh).
#pragma once
#include "est_kernels.cuh"
#include "est_lanes.cuh"
#include "../../include/hilo_b200.h"

namespace hb {

template <class E>
struct EstModule {
  static void dims(hb_est_dims* d) {
    d->nx = E::NX; d->nu = E::NU; d->np = E::NP; d->ny = E::NY; d->nye = E::NYE; d->discrete = E::DISCRETE ? 1 : 0;
  }
  static size_t work_doubles(int B, int n_particles) {
    return (size_t)B * (size_t)n_particles * (E::NX + 2) + 16;
  }
  static bool fill(const hb_est_call* c, EstArgs& a) {
    if (c == nullptr || c->B <= 0) return false;
    memset(&a, 0, sizeof(a));
    a.B = c->B; a.substeps = c->substeps > 0 ? c->substeps : 1; a.n_particles = c->n_particles;
    a.do_predict = c->do_predict; a.do_update = c->do_update; a.dt = c->dt; a.t0 = c->t0;
    // UKF parameters (kf.py:486-503)
    const double n = E::NX;
    const double lam = c->alpha * c->alpha * (n + c->kappa) - n;
    a.gamma = sqrt(n + lam);
    a.w_m0 = lam / (n + lam);
    a.w_c0 = lam / (n + lam) + 1.0 - c->alpha * c->alpha + c->beta;
    a.w_i = 1.0 / (2.0 * (n + lam));
    a.x = c->x; a.P = c->P; a.X = c->X; a.y = c->y; a.u = c->u; a.u_stride = c->u_stride; a.p = c->p;
    a.p_stride = c->p_stride; a.Q = c->Q; a.R = c->R; a.y_pred = c->y_pred; a.w = c->w; a.v = c->v; a.unif = c->unif;
    a.q = c->q; a.idx = c->idx; a.status = c->status;
    if (c->work != nullptr) {
      const size_t bn = (size_t)c->B * (size_t)(c->n_particles > 0 ? c->n_particles : 0);
      a.Xs = c->work; a.cdf = c->work + bn * E::NX; a.Ys = a.cdf + bn;
    }
    return true;
  }
  static int finish() {
#if HB_DEVICE_BUILD
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
#else
    return 0;
#endif
  }
  static int sim(const hb_est_call* c) {
    EstArgs a; if (!fill(c, a)) return -1;
    launch_1d<EstSim<E>>(a, a.B, (cudaStream_t)c->stream); return finish();
  }
  static int ekf(const hb_est_call* c) {
    EstArgs a; if (!fill(c, a)) return -1;
    if (a.status == nullptr || a.P == nullptr) return -4;
    launch_1d<EkfStep<E>>(a, a.B, (cudaStream_t)c->stream); return finish();
  }
  static int ukf(const hb_est_call* c) {
    EstArgs a; if (!fill(c, a)) return -1;
    if (a.status == nullptr || a.P == nullptr || a.X == nullptr) return -4;
#if HB_DEVICE_BUILD
    if constexpr (UkfLanes<E>::SUPPORTED) {   // one lane per sigma point (est_lanes.cuh); variant 0 keeps one thread per filter
      if (c->variant != 0) {
        cudaError_t e = launch_ukf_lanes<E>(a, (cudaStream_t)c->stream);
        return e == cudaSuccess ? finish() : (int)e;
      }
    }
#endif
    launch_1d<UkfStep<E>>(a, a.B, (cudaStream_t)c->stream); return finish();
  }
  static int pf(const hb_est_call* c) {
    EstArgs a; if (!fill(c, a)) return -1;
    if (a.n_particles < 2 || c->work == nullptr || c->work_doubles < work_doubles(a.B, a.n_particles)) return -2;
    if (a.status == nullptr || a.X == nullptr || a.q == nullptr || a.idx == nullptr) return -4;
#if HB_DEVICE_BUILD
    if (c->variant != 0) {   // one warp per filter, particles on the lanes (est_lanes.cuh); bit-identical to PfStep
      cudaError_t e = launch_pf_lanes<E>(a, (cudaStream_t)c->stream);
      return e == cudaSuccess ? finish() : (int)e;
    }
#endif
    launch_1d<PfStep<E>>(a, a.B, (cudaStream_t)c->stream); return finish();
  }
};

}  // namespace hb

#define HB_DEFINE_EST_MODULE(EPROB)                                                                    \
  extern "C" int hbm_abi_version() { return HB_ABI_VERSION; }                                          \
  extern "C" void hbm_est_dims(hb_est_dims* d) { hb::EstModule<EPROB>::dims(d); }                      \
  extern "C" size_t hbm_est_work_doubles(int B, int n) { return hb::EstModule<EPROB>::work_doubles(B, n); } \
  extern "C" int hbm_sim_step(const hb_est_call* c) { return hb::EstModule<EPROB>::sim(c); }           \
  extern "C" int hbm_ekf_step(const hb_est_call* c) { return hb::EstModule<EPROB>::ekf(c); }           \
  extern "C" int hbm_ukf_step(const hb_est_call* c) { return hb::EstModule<EPROB>::ukf(c); }           \
  extern "C" int hbm_pf_step(const hb_est_call* c) { return hb::EstModule<EPROB>::pf(c); }
