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

// ---------------------------------------------------------------------------------------------------------------
// Particle filter, warp-level version: one WARP per filter, its particles on the lanes (particle k on lane k % 32).
//
// The one-thread-per-filter kernel (PfStep) walks the n_s particles of a filter in sequence - n_s x substeps x 4
// right-hand-side evaluations per thread, and 16384 filters are 3.5 warps per SM.  Here the propagations, likelihoods,
// the binary searches of the resampling and the gather run side by side; what the reference computes as a SEQUENCE
// (normalisation sum, cumulative sum of the weights, sample mean and covariance: the order fixes the rounding, and the
// resampling indices must equal numpy's bit for bit) stays a sequence, walked by one lane per output - weights: lane 0,
// mean: lane i < nx, covariance: lane (i, j) - over data the warp has just written (L1-resident).  The results are
// therefore identical to PfStep's, bit for bit.
template <class E>
struct PfLanes {
  static constexpr int NX = E::NX, NY = E::NYE;

  static __device__ void run(const EstArgs& a) {
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = gt >> 5, lane = gt & 31;
    if (b >= a.B) return;                       // whole warps only: no shuffle / barrier partner is lost
    const double* u = a.u + (long long)b * a.u_stride;
    const double* p = a.p + (long long)b * a.p_stride;
    const int ns = a.n_particles;
    double* Xg = a.X + (long long)b * ns * NX;
    double* Xs = a.Xs + (long long)b * ns * NX;
    double* q = a.q + (long long)b * ns;
    double* Ys = a.Ys + (long long)b * ns;
    double* cdf = a.cdf + (long long)b * ns;
    long long* idx = a.idx + (long long)b * ns;
    const double* unif = a.unif + (long long)b * ns;
    const double yv = a.y[(long long)b * NY];
    const double sigma = sqrt(a.R[0]);          // univariate likelihood (pf.py:158)
    // ---- 1. propagate, add the noise, likelihood (parallel over the particles)
    for (int k = lane; k < ns; k += 32) {
      double xn[NX], yk[NY];
      est_propagate<E>(Xg + k * NX, u, p, a.t0, a.dt, a.substeps, xn);
      E::hE(xn, u, p, yk);
#pragma unroll
      for (int i = 0; i < NX; ++i) Xs[k * NX + i] = xn[i] + a.w[((long long)b * ns + k) * NX + i];
      const double Y = yk[0] + a.v[((long long)b * ns + k) * NY];
      const double r = (Y - yv) / sigma;
      q[k] = exp(-0.5 * r * r) / (sqrt(2.0 * M_PI) * sigma);
      Ys[k] = Y;
    }
    __syncwarp();
    if (!a.do_update) {                          // prediction pass only (see PfStep)
      if (lane == 0) a.status[b] = 0;
      return;
    }
    // ---- 2. weights: q / sum(q), cdf = cumsum(q) / cdf[-1]  (the two running sums in particle order)
    double qsum = 0.0;
    if (lane == 0)
      for (int k = 0; k < ns; ++k) qsum += q[k];
    qsum = __shfl_sync(0xffffffffu, qsum, 0);
    for (int k = lane; k < ns; k += 32) q[k] /= qsum;
    __syncwarp();
    double last = 0.0;
    if (lane == 0) {
      double sacc = 0.0;
      for (int k = 0; k < ns; ++k) { sacc += q[k]; cdf[k] = sacc; }
      last = sacc;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    __syncwarp();
    for (int k = lane; k < ns; k += 32) cdf[k] /= last;
    __syncwarp();
    // ---- 3. resampling indices: searchsorted(cdf, u, side='right'), one search per lane
    for (int k = lane; k < ns; k += 32) {
      const double uu = unif[k];
      int lo = 0, hi = ns;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] <= uu) lo = mid + 1; else hi = mid;
      }
      idx[k] = lo;
    }
    __syncwarp();
    // ---- 4. gather the survivors
    for (int k = lane; k < ns; k += 32) {
      const long long src = idx[k];
#pragma unroll
      for (int i = 0; i < NX; ++i) Xg[k * NX + i] = Xs[src * NX + i];
    }
    __syncwarp();
    // ---- 5. moments, each output a sequence over the particles like the reference's sum2 / np.cov
    for (int i = lane; i < NX; i += 32) {
      double sacc = 0.0;
      for (int k = 0; k < ns; ++k) sacc += Xg[k * NX + i];
      a.x[(long long)b * NX + i] = sacc / ns;
    }
    __syncwarp();
    for (int e = lane; e < NX * NX; e += 32) {
      const int i = e / NX, j = e % NX;
      const double mi = a.x[(long long)b * NX + i], mj = a.x[(long long)b * NX + j];
      double sacc = 0.0;
      for (int k = 0; k < ns; ++k) sacc += (Xg[k * NX + i] - mi) * (Xg[k * NX + j] - mj);
      a.P[(long long)b * NX * NX + e] = sacc / (ns - 1);
    }
    if (lane == 0) {
      if (a.y_pred != nullptr) {                // pf.py:419-420: mean of the resampled noisy measurements
        double ym = 0.0;
        for (int k = 0; k < ns; ++k) ym += Ys[idx[k]];
        a.y_pred[(long long)b * NY] = ym / ns;
      }
      a.status[b] = 0;
    }
  }
};

template <class E>
__global__ void __launch_bounds__(128) pf_lanes_kernel(EstArgs a) { PfLanes<E>::run(a); }

template <class E>
inline cudaError_t launch_pf_lanes(const EstArgs& a, cudaStream_t s) {
  const long long threads = (long long)a.B * 32;
  pf_lanes_kernel<E><<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(a);
  return cudaGetLastError();
}

#endif  // HB_DEVICE_BUILD
}  // namespace hb
