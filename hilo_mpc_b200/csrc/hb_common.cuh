// Common helpers for the hilo_mpc_b200 kernels (sm_100a).
//
// Every kernel body is written as a __host__ __device__ function of (instance, stage) indices and launched through
// the HB_LAUNCH_* macros below.  Under nvcc they become <<<>>> launches on the caller's stream; when the translation
// unit is compiled with -DHB_HOSTSIM (tests only, see tests/hostsim/) the same bodies run in plain host loops so the
// kernel *logic* can be checked against the oracle on a box without a GPU.  The product library is never built
// with HB_HOSTSIM and has no CPU fallback.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#if defined(__CUDACC__) && !defined(HB_HOSTSIM)
#include <cuda_runtime.h>
#define HB_HD __host__ __device__ __forceinline__
#define HB_DEVICE_BUILD 1
#else
#define HB_HD inline
#define HB_DEVICE_BUILD 0
#ifndef __CUDACC__
typedef void* cudaStream_t;
#endif
#endif

namespace hb {

constexpr double kInfBound = 1e19;  // |bound| above this means "no bound" (IPOPT nlp_lower_bound_inf)

HB_HD double dmin(double a, double b) { return a < b ? a : b; }
HB_HD double dmax(double a, double b) { return a > b ? a : b; }
HB_HD double dabs(double a) { return a < 0 ? -a : a; }

// ---------------------------------------------------------------------------------------------------------------
// small dense helpers on row-major arrays (compile-time sizes, fully unrollable)
template <int M, int N>
HB_HD void mat_zero(double* A) {
#pragma unroll
  for (int i = 0; i < M * N; ++i) A[i] = 0.0;
}

// C (MxN) += A (MxK) * B (KxN)
template <int M, int K, int N>
HB_HD void mat_mul_acc(const double* A, const double* B, double* C) {
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      double s = C[i * N + j];
      for (int l = 0; l < K; ++l) s += A[i * K + l] * B[l * N + j];
      C[i * N + j] = s;
    }
}

// C (MxN) += A^T (A is KxM) * B (KxN)
template <int M, int K, int N>
HB_HD void mat_tmul_acc(const double* A, const double* B, double* C) {
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      double s = C[i * N + j];
      for (int l = 0; l < K; ++l) s += A[l * M + i] * B[l * N + j];
      C[i * N + j] = s;
    }
}

// In-place Cholesky of a symmetric positive definite NxN matrix (lower factor in the lower triangle).
// Returns false when a pivot is not positive.
template <int N>
HB_HD bool chol_lower(double* A) {
  for (int j = 0; j < N; ++j) {
    double d = A[j * N + j];
    for (int l = 0; l < j; ++l) d -= A[j * N + l] * A[j * N + l];
    if (!(d > 0.0)) return false;
    d = sqrt(d);
    A[j * N + j] = d;
    for (int i = j + 1; i < N; ++i) {
      double s = A[i * N + j];
      for (int l = 0; l < j; ++l) s -= A[i * N + l] * A[j * N + l];
      A[i * N + j] = s / d;
    }
  }
  return true;
}

// Solve L L^T X = Bm for NRHS right-hand sides stored row-major (N x NRHS), in place.
template <int N, int NRHS>
HB_HD void chol_solve(const double* L, double* Bm) {
  for (int c = 0; c < NRHS; ++c) {
    for (int i = 0; i < N; ++i) {
      double s = Bm[i * NRHS + c];
      for (int l = 0; l < i; ++l) s -= L[i * N + l] * Bm[l * NRHS + c];
      Bm[i * NRHS + c] = s / L[i * N + i];
    }
    for (int i = N - 1; i >= 0; --i) {
      double s = Bm[i * NRHS + c];
      for (int l = i + 1; l < N; ++l) s -= L[l * N + i] * Bm[l * NRHS + c];
      Bm[i * NRHS + c] = s / L[i * N + i];
    }
  }
}

// Gaussian elimination with partial pivoting on an augmented row-major matrix M (N x (N + NRHS)); on return the
// right NRHS columns hold the solution of  M[:, :N] * Y = M[:, N:].  Returns false on a zero pivot.
template <int N, int NRHS>
HB_HD bool lu_solve_aug(double* M) {
  constexpr int W = N + NRHS;
  for (int c = 0; c < N; ++c) {
    int piv = c;
    double best = dabs(M[c * W + c]);
    for (int r = c + 1; r < N; ++r) {
      double v = dabs(M[r * W + c]);
      if (v > best) { best = v; piv = r; }
    }
    if (!(best > 0.0)) return false;
    if (piv != c)
      for (int j = c; j < W; ++j) { double tmp = M[c * W + j]; M[c * W + j] = M[piv * W + j]; M[piv * W + j] = tmp; }
    const double inv = 1.0 / M[c * W + c];
    for (int r = c + 1; r < N; ++r) {
      const double f = M[r * W + c] * inv;
      if (f != 0.0)
        for (int j = c + 1; j < W; ++j) M[r * W + j] -= f * M[c * W + j];
    }
  }
  for (int c = N - 1; c >= 0; --c) {
    const double inv = 1.0 / M[c * W + c];
    for (int j = N; j < W; ++j) {
      double s = M[c * W + j];
      for (int l = c + 1; l < N; ++l) s -= M[c * W + l] * M[l * W + j];
      M[c * W + j] = s * inv;
    }
  }
  return true;
}

// Convexification of a symmetric NxN matrix by eigenvalue mirroring: H = V diag(l) V^T -> V diag(max(|l|, eps)) V^T
// (cyclic Jacobi; N <= 16).  Negative curvature of magnitude |l| becomes positive curvature of the same magnitude, which
// keeps the length scale of the Newton step in that direction instead of flattening it (Verschueren et al. 2017).
// Returns true when the matrix was changed.
template <int N>
HB_HD bool sym_mirror(double* H, double eps) {
  {  // cheap test first: Cholesky with all pivots above eps -> nothing to do
    double L[N * N];
    for (int i = 0; i < N * N; ++i) L[i] = H[i];
    bool pd = true;
    for (int j = 0; j < N && pd; ++j) {
      double d = L[j * N + j];
      for (int l = 0; l < j; ++l) d -= L[j * N + l] * L[j * N + l];
      if (!(d > eps)) { pd = false; break; }
      d = sqrt(d);
      L[j * N + j] = d;
      for (int i = j + 1; i < N; ++i) {
        double s2 = L[i * N + j];
        for (int l = 0; l < j; ++l) s2 -= L[i * N + l] * L[j * N + l];
        L[i * N + j] = s2 / d;
      }
    }
    if (pd) return false;
  }
  double A[N * N], V[N * N];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) { A[i * N + j] = 0.5 * (H[i * N + j] + H[j * N + i]); V[i * N + j] = (i == j) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) { if (i != j) off += A[i * N + j] * A[i * N + j]; else diag += A[i * N + i] * A[i * N + i]; }
    if (off <= 1e-30 * (diag + 1e-300)) break;
    for (int p = 0; p < N - 1; ++p)
      for (int q = p + 1; q < N; ++q) {
        const double apq = A[p * N + q];
        if (apq == 0.0) continue;
        const double th = (A[q * N + q] - A[p * N + p]) / (2.0 * apq);
        const double t = (th >= 0.0 ? 1.0 : -1.0) / (dabs(th) + sqrt(th * th + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < N; ++k) {
          const double akp = A[k * N + p], akq = A[k * N + q];
          A[k * N + p] = c * akp - sn * akq; A[k * N + q] = sn * akp + c * akq;
        }
        for (int k = 0; k < N; ++k) {
          const double apk = A[p * N + k], aqk = A[q * N + k];
          A[p * N + k] = c * apk - sn * aqk; A[q * N + k] = sn * apk + c * aqk;
        }
        for (int k = 0; k < N; ++k) {
          const double vkp = V[k * N + p], vkq = V[k * N + q];
          V[k * N + p] = c * vkp - sn * vkq; V[k * N + q] = sn * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < N; ++i)
    for (int j = 0; j <= i; ++j) {
      double s2 = 0.0;
      for (int k = 0; k < N; ++k) s2 += V[i * N + k] * dmax(dabs(A[k * N + k]), eps) * V[j * N + k];
      H[i * N + j] = s2; H[j * N + i] = s2;
    }
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// Log-barrier pieces for one scalar variable with bounds [lb, ub] (either side may be infinite).
struct BarrierTerm {
  double grad;   // d/dv of  -mu*log(v-lb) - mu*log(ub-v)
  double sigma;  // primal-dual Hessian  zl/(v-lb) + zu/(ub-v)
  double dz;     // (-zl + zu) - grad : correction that turns the barrier gradient into the multiplier form
};

HB_HD bool has_lb(double lb) { return lb > -kInfBound; }
HB_HD bool has_ub(double ub) { return ub < kInfBound; }

// 1/x for the slack variables: hardware approximation + two Newton steps on the device (a full-precision IEEE
// division is ~25 instructions and a slow-path call), plain division on the host build.
HB_HD double rcp_nr(double x) {
#if HB_DEVICE_BUILD && defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
#else
  return 1.0 / x;
#endif
}

HB_HD BarrierTerm barrier_term(double v, double lb, double ub, double zl, double zu, double mu) {
  BarrierTerm t{0.0, 0.0, 0.0};
  if (has_lb(lb)) { const double is = rcp_nr(v - lb), ms = mu * is; t.grad -= ms; t.sigma += zl * is; t.dz += ms - zl; }
  if (has_ub(ub)) { const double is = rcp_nr(ub - v), ms = mu * is; t.grad += ms; t.sigma += zu * is; t.dz -= ms - zu; }
  return t;
}

HB_HD double barrier_value(double v, double lb, double ub, double mu) {
  double phi = 0.0;
  if (has_lb(lb)) { const double s = v - lb; phi += (s > 0.0) ? -mu * log(s) : HUGE_VAL; }
  if (has_ub(ub)) { const double s = ub - v; phi += (s > 0.0) ? -mu * log(s) : HUGE_VAL; }
  return phi;
}

// Sum of -mu*log(slack) over many bounds with one logarithm per four variables: the slacks are multiplied up (a missing
// bound contributes the factor 1) and log() is taken of the product.  A double-precision log costs ~45 instructions and
// the barrier objective needs two per bounded variable and node, which made it a third of the node evaluation.  Eight
// slacks in [1e-19, 2e19] keep the product inside the double range.  Any non-positive slack makes the value +inf.
struct BarrierSum {
  double prod = 1.0, logsum = 0.0;
  int n = 0;
  bool bad = false;
  HB_HD void add(double v, double lb, double ub) {
    const double sl = has_lb(lb) ? v - lb : 1.0, su = has_ub(ub) ? ub - v : 1.0;
    bad = bad || !(sl > 0.0) || !(su > 0.0);
    prod *= sl * su;
    if (++n == 4) { logsum += log(prod); prod = 1.0; n = 0; }
  }
  HB_HD double value(double mu) const {
    if (bad) return HUGE_VAL;
    return -mu * ((n > 0) ? logsum + log(prod) : logsum);
  }
};

// IPOPT's initial push of a variable into the interior of its bounds (bound_push = bound_frac = 1e-2).
HB_HD double push_interior(double v, double lb, double ub) {
  const double k1 = 1e-2, k2 = 1e-2;
  const bool l = has_lb(lb), u = has_ub(ub);
  if (l) {
    double p = k1 * dmax(1.0, dabs(lb));
    if (u) p = dmin(p, k2 * (ub - lb));
    v = dmax(v, lb + p);
  }
  if (u) {
    double p = k1 * dmax(1.0, dabs(ub));
    if (l) p = dmin(p, k2 * (ub - lb));
    v = dmin(v, ub - p);
  }
  return v;
}

// Warm start: keep the point where it is unless it sits (numerically) on a bound; multipliers are kept but bounded
// away from zero / infinity relative to the starting barrier parameter.
HB_HD double push_interior_warm(double v, double lb, double ub) {
  const double k = 1e-9;
  if (has_lb(lb)) v = dmax(v, lb + k * dmax(1.0, dabs(lb)));
  if (has_ub(ub)) v = dmin(v, ub - k * dmax(1.0, dabs(ub)));
  return v;
}
HB_HD double warm_z(double z, double mu) {
  if (!(z == z) || z <= 0.0) return 1.0;
  return dmin(dmax(z, 1e-3 * mu), 1e12);
}

// Largest alpha in (0, 1] with  s + alpha*ds >= (1 - tau) * s  for a positive quantity s.
HB_HD double frac_to_boundary(double s, double ds, double tau, double amax) {
  if (ds < 0.0) { const double a = -tau * s * rcp_nr(ds); if (a < amax) amax = a; }
  return amax;
}

// Newton step of a bound multiplier z whose slack s changes by ds:  dz = mu/s - z - (z/s) ds
HB_HD double dual_step(double z, double s, double ds, double mu) {
  const double is = rcp_nr(s);
  return mu * is - z - z * is * ds;
}

}  // namespace hb

// ---------------------------------------------------------------------------------------------------------------
// launch macros: body(args, i0, i1[, i2]) over a dense index space
#if HB_DEVICE_BUILD
namespace hb {
template <class Body, class Args>
__global__ void __launch_bounds__(128) kernel_2d(Args args, int n0, int n1) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n0 * n1) return;
  Body::run(args, (int)(idx / n1), (int)(idx % n1));
}
template <class Body, class Args>
__global__ void __launch_bounds__(128) kernel_3d(Args args, int n0, int n1, int n2) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n0 * n1 * n2) return;
  const int i2 = (int)(idx % n2);
  const long long r = idx / n2;
  Body::run(args, (int)(r / n1), (int)(r % n1), i2);
}
template <class Body, class Args>
__global__ void __launch_bounds__(128) kernel_1d(Args args, int n0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n0) return;
  Body::run(args, idx);
}
template <class Body, class Args>
inline void launch_1d(const Args& a, int n0, cudaStream_t s) {
  if (n0 <= 0) return;
  kernel_1d<Body, Args><<<(n0 + 127) / 128, 128, 0, s>>>(a, n0);
}
template <class Body, class Args>
inline void launch_2d(const Args& a, int n0, int n1, cudaStream_t s) {
  const long long n = (long long)n0 * n1;
  if (n <= 0) return;
  kernel_2d<Body, Args><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(a, n0, n1);
}
template <class Body, class Args>
inline void launch_3d(const Args& a, int n0, int n1, int n2, cudaStream_t s) {
  const long long n = (long long)n0 * n1 * n2;
  if (n <= 0) return;
  kernel_3d<Body, Args><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(a, n0, n1, n2);
}
}  // namespace hb
#else
namespace hb {
template <class Body, class Args>
inline void launch_1d(const Args& a, int n0, cudaStream_t) {
  for (int i = 0; i < n0; ++i) Body::run(a, i);
}
template <class Body, class Args>
inline void launch_2d(const Args& a, int n0, int n1, cudaStream_t) {
  for (int i = 0; i < n0; ++i)
    for (int j = 0; j < n1; ++j) Body::run(a, i, j);
}
template <class Body, class Args>
inline void launch_3d(const Args& a, int n0, int n1, int n2, cudaStream_t) {
  for (int i = 0; i < n0; ++i)
    for (int j = 0; j < n1; ++j)
      for (int l = 0; l < n2; ++l) Body::run(a, i, j, l);
}
}  // namespace hb
#endif
