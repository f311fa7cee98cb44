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

// ---------------------------------------------------------------------------------------------------------------
// One-hidden-layer perceptron  y = W2 act(W1 z + b1) + b2  (the graph of hilo_mpc/util/machine_learning.py:521-578 with the
// z-score scaling of inputs and outputs folded into the weights) as a device function: value, Jacobian and the second
// derivatives of every output in ONE loop over the hidden units.  The fully unrolled expression of a 1 x 32 network and
// its derivatives is ~20 000 SASS instructions per node kernel (333 KB of straight-line code: 37 % of the warp samples
// stalled on instruction fetch, profiles/ncu_evq_c5_r2_summary.csv); the loop is a few hundred.
// tab = [NH][NZ + 1] rows (W1[j][:], b1[j]) | [NO][NH] W2 | [NO] b2, in constant memory of the generated module.
// out = [y (NO) | J row-major (NO x NZ) | for every output the lower triangle of its Hessian (NO x NZ (NZ + 1) / 2)].
// ACT: 0 tanh, 1 logistic sigmoid.
template <int NZ, int NH, int NO, int ORDER, int ACT>
HB_HD void mlp1_eval(const double* __restrict__ tab, const double* z, double* out) {
  constexpr int NT = NZ * (NZ + 1) / 2;
  constexpr int NOUT = NO + (ORDER >= 1 ? NO * NZ : 0) + (ORDER >= 2 ? NO * NT : 0);
  const double* W2 = tab + NH * (NZ + 1);
  const double* b2 = W2 + NO * NH;
  for (int e = 0; e < NOUT; ++e) out[e] = 0.0;
  for (int o = 0; o < NO; ++o) out[o] = b2[o];
  for (int j = 0; j < NH; ++j) {
    const double* w1 = tab + j * (NZ + 1);
    double act = w1[NZ];
#if HB_DEVICE_BUILD
#pragma unroll
#endif
    for (int i = 0; i < NZ; ++i) act += w1[i] * z[i];
    double h, d1, d2;
    if (ACT == 0) { h = tanh(act); d1 = 1.0 - h * h; d2 = -2.0 * h * d1; }
    else { h = 1.0 / (1.0 + exp(-act)); d1 = h * (1.0 - h); d2 = d1 * (1.0 - 2.0 * h); }
    double ww[ORDER >= 2 ? NT : 1];
    if (ORDER >= 2) {
#if HB_DEVICE_BUILD
#pragma unroll
#endif
      for (int i = 0; i < NZ; ++i)
#if HB_DEVICE_BUILD
#pragma unroll
#endif
        for (int l = 0; l <= i; ++l) ww[i * (i + 1) / 2 + l] = w1[i] * w1[l];
    }
#if HB_DEVICE_BUILD
#pragma unroll
#endif
    for (int o = 0; o < NO; ++o) {
      const double w = W2[o * NH + j];
      out[o] += w * h;
      if (ORDER >= 1) {
        const double wd = w * d1;
#if HB_DEVICE_BUILD
#pragma unroll
#endif
        for (int i = 0; i < NZ; ++i) out[NO + o * NZ + i] += wd * w1[i];
      }
      if (ORDER >= 2) {
        const double wd2 = w * d2;
#if HB_DEVICE_BUILD
#pragma unroll
#endif
        for (int e = 0; e < NT; ++e) out[NO + NO * NZ + o * NT + e] += wd2 * ww[e];
      }
    }
  }
}

}  // namespace hb
