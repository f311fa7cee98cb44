// Posterior mean of a Gaussian process with the squared-exponential (ARD) kernel as a DEVICE FUNCTION for the node
// kernels: value, gradient and Hessian with respect to the query point in one pass over the training set.
//
// The reference embeds a learned component by substituting the GP's symbolic prediction into the model equations
// (hilo_mpc/modules/dynamic_model/dynamic_model.py:3040-3125, gp.py:623-629, 699-718): an expression with one term per
// training point, evaluated by the CasADi VM inside IPOPT.  Here the training set lives in a table of the generated
// module ([n][D + 1] doubles: the inputs divided by their length scales, then alpha_i * sigma_f^2), 512 x 3 doubles =
// 12 KB for BASELINE config 3 - L1/L2 resident - and the generated model function calls this loop instead of carrying
// 512 unrolled terms:
//   m(x)      = sum_i c_i k_i,            k_i = exp(-1/2 |xs - Xs_i|^2),  xs_j = x_j / l_j
//   dm/dx_j   = sum_i c_i k_i (-(xs_j - Xs_ij)) / l_j
//   d2m/dxjdxl= sum_i c_i k_i ((xs_j - Xs_ij)(xs_l - Xs_il) - delta_jl) / (l_j l_l)
// out = [m | dm/dx (D) | lower triangle of the Hessian, row by row (D (D + 1) / 2)], as far as ORDER asks for.
#pragma once
#include "hb_common.cuh"

namespace hb {

template <int D, int ORDER>
HB_HD void gp_se_eval(const double* __restrict__ tab, int n, const double* __restrict__ inv_ls, const double* x,
                      double* out) {
  constexpr int NH = D * (D + 1) / 2;
  double xs[D], m = 0.0, g[D], H[NH];
#if HB_DEVICE_BUILD
#pragma unroll
#endif
  for (int j = 0; j < D; ++j) { xs[j] = x[j] * inv_ls[j]; g[j] = 0.0; }
#if HB_DEVICE_BUILD
#pragma unroll
#endif
  for (int e = 0; e < NH; ++e) H[e] = 0.0;
  for (int i = 0; i < n; ++i) {
    const double* row = tab + (long long)i * (D + 1);
    double r[D], d2 = 0.0;
#if HB_DEVICE_BUILD
#pragma unroll
#endif
    for (int j = 0; j < D; ++j) { r[j] = xs[j] - row[j]; d2 += r[j] * r[j]; }
    const double ck = row[D] * exp(-0.5 * d2);
    m += ck;
    if (ORDER >= 1) {
#if HB_DEVICE_BUILD
#pragma unroll
#endif
      for (int j = 0; j < D; ++j) g[j] -= ck * r[j];
    }
    if (ORDER >= 2) {
#if HB_DEVICE_BUILD
#pragma unroll
#endif
      for (int j = 0; j < D; ++j)
#if HB_DEVICE_BUILD
#pragma unroll
#endif
        for (int l = 0; l <= j; ++l) H[j * (j + 1) / 2 + l] += ck * (r[j] * r[l] - (j == l ? 1.0 : 0.0));
    }
  }
  out[0] = m;
  if (ORDER >= 1)
    for (int j = 0; j < D; ++j) out[1 + j] = g[j] * inv_ls[j];
  if (ORDER >= 2)
    for (int j = 0; j < D; ++j)
      for (int l = 0; l <= j; ++l) out[1 + D + j * (j + 1) / 2 + l] = H[j * (j + 1) / 2 + l] * inv_ls[j] * inv_ls[l];
}

}  // namespace hb
