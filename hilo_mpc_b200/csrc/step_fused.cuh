// Step computation of one interior-point iteration as ONE kernel: one CTA per instance, one thread per node.
//
// The sequence  Expand -> Prep -> Trial(alpha_max) -> Decide -> [Trial(backtracking) -> Decide] -> Apply -> Commit
// of the reference pipeline (ocp_module.cuh) alternates node-parallel work with per-instance reductions/decisions.
// As separate launches every boundary costs a launch and a round trip of the iterate (x, u, X, dx, du, dX, z) through
// HBM.  All the dependencies are inside one instance, so a CTA that owns the instance can run the whole sequence with
// __syncthreads() between the phases: the node-parallel bodies are the same device functions as in the separate kernels
// (bit-identical arithmetic and summation order), thread 0 runs the per-instance bodies, and the second and third pass
// over the iterate hit L1/L2 instead of HBM.  The per-node scalar tables (s_*, ls_*) stay in global memory (2.8 KB per
// instance, L1-resident); __syncthreads() orders the global accesses of the CTA.
#pragma once
#include "ocp_kernels.cuh"

namespace hb {
#if HB_DEVICE_BUILD

template <class P>
struct StepFused {
  static constexpr int N = OcpT<P>::N;
  static constexpr int THREADS = ((N + 1 + 31) / 32) * 32;
  static constexpr bool SUPPORTED = THREADS <= 1024;
};

// a0 / a1: the same arguments with ls_phase = 0 / 1 (kernel parameters are read-only)
template <class P>
__global__ void __launch_bounds__(StepFused<P>::THREADS) step_fused_kernel(const OcpArgs a0, const OcpArgs a1) {
  constexpr int N = OcpT<P>::N;
  const int b = blockIdx.x, k = threadIdx.x;
  const double* st = a0.st + (long long)b * ST_COUNT;
  if (st[ST_ACTIVE] == 0.0) return;   // uniform over the CTA
  const bool node = k <= N;
  if (node) OcpExpand<P>::run(a0, b, k);
  __syncthreads();
  if (k == 0) OcpPrep<P>::run(a0, b);
  __syncthreads();
  if (node) OcpTrial<P>::run(a0, b, k, 0);
  __syncthreads();
  if (k == 0) OcpDecide<P>::run(a0, b);
  __syncthreads();
  if (st[ST_ACTIVE] == 0.0) return;   // converged / iteration cap: nothing to apply
  if (a0.n_alpha > 1 && st[ST_LS_PENDING] != 0.0) {   // full step rejected by the filter: backtracking candidates
    if (node)
      for (int ai = 0; ai < a0.n_alpha - 1; ++ai) OcpTrial<P>::run(a1, b, k, ai);
    __syncthreads();
    if (k == 0) OcpDecide<P>::run(a1, b);
    __syncthreads();
    if (st[ST_ACTIVE] == 0.0) return;
  }
  if (node) OcpApply<P>::run(a0, b, k);
  __syncthreads();
  if (k == 0) OcpCommit<P>::run(a0, b);
}

template <class P>
inline cudaError_t launch_step_fused(const OcpArgs& a, cudaStream_t s) {
  OcpArgs a0 = a, a1 = a;
  a0.ls_phase = 0; a1.ls_phase = 1;
  step_fused_kernel<P><<<a.B, StepFused<P>::THREADS, 0, s>>>(a0, a1);
  return cudaGetLastError();
}

#endif  // HB_DEVICE_BUILD
}  // namespace hb
