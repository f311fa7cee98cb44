// Unscented Kalman filter, warp-level version for sm_100a: one LANE per sigma point.
//
// The one-thread-per-filter kernel (UkfStep, est_kernels.cuh) propagates the 2 nx + 1 sigma points of a filter one after
// the other: 9 x substeps x 4 right-hand-side evaluations in sequence per thread for the CSTR, and a batch of 16384 filters
// is only 512 warps - 3.5 per SM - so the kernel runs at the latency of that chain.  Here the sigma points of a filter sit
// on the lanes of one group of LPF = 4 / 8 / 16 / 32 lanes (G = 32 / LPF filters per warp): the propagations run side by side
// (B x (2 nx + 1) independent chains), the weighted mean and the covariances Pxx, Pxy, Pyy are butterfly reductions over the
// group (__shfl_xor), the tiny factorisations (Cholesky of P for the sigma points, the gain) are done redundantly by every
// lane so that no broadcast is needed.  Loads of x / P are the same addresses for all lanes of a group (one transaction),
// the sigma-point matrix X [B][2nx+1][nx] is written / read with consecutive lanes on consecutive rows.
// Formulas and conventions are those of UkfStep (kf.py:486-604: columns of the UPPER Cholesky factor, update re-uses the
// propagated points); the summation order of the reductions differs (tree instead of sequence), i.e. results agree to
// rounding, not bit for bit.
#pragma once
#include "est_kernels.cuh"

namespace hb {
#if HB_DEVICE_BUILD

template <class E>
struct UkfLanes {
  static constexpr int NX = E::NX, NY = E::NYE, NS = 2 * E::NX + 1;
  static constexpr int LPF = NS <= 4 ? 4 : NS <= 8 ? 8 : NS <= 16 ? 16 : 32;
  static constexpr bool SUPPORTED = NS <= 32;
  static constexpr int G = 32 / LPF;
  static constexpr int NPXX = NX * (NX + 1) / 2, NPYY = NY * (NY + 1) / 2;

  static __device__ __forceinline__ double gsum(double v) {
#pragma unroll
    for (int m = LPF / 2; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m, LPF);
    return v;
  }

  static __device__ void run(const EstArgs& a) {
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int fb = gt / LPF, s = gt % LPF;
    const bool valid = fb < a.B;
    const int b = valid ? fb : a.B - 1;          // surplus groups shadow the last filter and write nothing
    const bool lane_on = s < NS;
    const double* u = a.u + (long long)b * a.u_stride;
    const double* p = a.p + (long long)b * a.p_stride;
    double x[NX], P[NX * NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = a.x[(long long)b * NX + i];
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) P[i] = a.P[(long long)b * NX * NX + i];
    double* Xg = a.X + ((long long)b * NS + (lane_on ? s : 0)) * NX;
    double Xs[NX];                                // this lane's propagated sigma point
    bool ok = true, dead = false;
    const double wm = !lane_on ? 0.0 : (s == 0 ? a.w_m0 : a.w_i), wc = !lane_on ? 0.0 : (s == 0 ? a.w_c0 : a.w_i);
    if (a.do_predict) {
      double L[NX * NX];
#pragma unroll
      for (int i = 0; i < NX * NX; ++i) L[i] = P[i];
      ok = chol_lower<NX>(L);                     // P = L L^T; the reference's upper factor is L^T, its column k is row k of L
      // P not positive definite (uniform over the group): x, P and the sigma points stay as they were.  The lanes keep
      // running - the other groups of the warp need them in the shuffles - and only the stores are suppressed.
      dead = !ok;
      double xs[NX];
      const int k = (s >= 1 && s <= NX) ? s - 1 : (s > NX && lane_on ? s - 1 - NX : 0);
      const double sgn = (s >= 1 && s <= NX) ? a.gamma : ((s > NX && lane_on) ? -a.gamma : 0.0);
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double lki = 0.0;
#pragma unroll
        for (int kk = 0; kk < NX; ++kk)
          if (kk == k && i <= kk) lki = L[kk * NX + i];
        xs[i] = x[i] + sgn * lki;
      }
      est_propagate<E>(xs, u, p, a.t0, a.dt, a.substeps, Xs);
      if (valid && lane_on && !dead) {
#pragma unroll
        for (int i = 0; i < NX; ++i) Xg[i] = Xs[i];
      }
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = gsum(wm * Xs[i]);
      int e = 0;
#pragma unroll
      for (int i = 0; i < NX; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j, ++e) {
          const double v = a.Q[i * NX + j] + gsum(wc * (Xs[i] - x[i]) * (Xs[j] - x[j]));
          P[i * NX + j] = v; P[j * NX + i] = v;
        }
    } else {
#pragma unroll
      for (int i = 0; i < NX; ++i) Xs[i] = lane_on ? Xg[i] : 0.0;
    }
    if (a.do_update) {
      double Y[NY], ym[NY], Pxy[NX * NY], Pyy[NY * NY], innov[NY];
      E::hE(Xs, u, p, Y);
#pragma unroll
      for (int i = 0; i < NY; ++i) ym[i] = gsum(wm * Y[i]);
#pragma unroll
      for (int i = 0; i < NX; ++i)
#pragma unroll
        for (int j = 0; j < NY; ++j) Pxy[i * NY + j] = gsum(wc * (Xs[i] - x[i]) * (Y[j] - ym[j]));
#pragma unroll
      for (int i = 0; i < NY; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
          const double v = a.R[i * NY + j] + gsum(wc * (Y[i] - ym[i]) * (Y[j] - ym[j]));
          Pyy[i * NY + j] = v; Pyy[j * NY + i] = v;
        }
#pragma unroll
      for (int i = 0; i < NY; ++i) {
        innov[i] = a.y[(long long)b * NY + i] - ym[i];
        if (valid && s == 0 && !dead) a.y_pred[(long long)b * NY + i] = ym[i];
      }
      ok = kalman_correct<NX, NY>(x, P, Pxy, Pyy, innov) && ok && !dead;
    }
    if (valid && s == 0) {
      if (ok && !dead) {  // a failed factorisation leaves the estimate untouched (status -1 tells the caller)
#pragma unroll
        for (int i = 0; i < NX; ++i) a.x[(long long)b * NX + i] = x[i];
#pragma unroll
        for (int i = 0; i < NX * NX; ++i) a.P[(long long)b * NX * NX + i] = P[i];
      }
      a.status[b] = (ok && !dead) ? 0 : -1;
    }
  }
};

template <class E>
__global__ void __launch_bounds__(128) ukf_lanes_kernel(EstArgs a) { UkfLanes<E>::run(a); }

template <class E>
inline cudaError_t launch_ukf_lanes(const EstArgs& a, cudaStream_t s) {
  const long long threads = (long long)a.B * UkfLanes<E>::LPF;
  ukf_lanes_kernel<E><<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(a);
  return cudaGetLastError();
}

#endif  // HB_DEVICE_BUILD
}  // namespace hb
