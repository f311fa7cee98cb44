// fp64 tensor-core (mma.sync.m8n8k4.f64) latency and throughput per SM sub-partition on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void k_dmma(double* out, long long* cyc, int n) {
  double c0[ILP], c1[ILP];
  double a = out[threadIdx.x] + 1.0, b = out[threadIdx.x] + 0.5;
  for (int j = 0; j < ILP; ++j) { c0[j] = j; c1[j] = -j; }
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < ILP; ++j) dmma(c0[j], c1[j], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int j = 0; j < ILP; ++j) s += c0[j] + c1[j];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
// dependent chain through the A operand: d -> a of the next mma (what the Riccati recursion does)
__global__ void k_dmma_dep(double* out, long long* cyc, int n) {
  double a = out[threadIdx.x] + 1.0, b = 1e-3;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) { double c0 = 0.0, c1 = 0.0; dmma(c0, c1, a, b); a = c0 + c1; }
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 4096 * 8); cudaMemset(out, 0, 4096 * 8);
  cudaMallocManaged(&cyc, 64 * 8);
  const int n = 1000;
  auto rep = [&](const char* name, double per) { cudaDeviceSynchronize(); printf("%-64s %8.2f cycles\n", name, (double)cyc[0] / per); };
  k_dmma_dep<<<1, 32>>>(out, cyc, n); rep("DMMA dependent (d->a, +DADD) latency per link", 8.0 * n);
  k_dmma<1><<<1, 32>>>(out, cyc, n); rep("DMMA accumulate chain (c->c) per op, 1 warp", 8.0 * n);
  k_dmma<4><<<1, 32>>>(out, cyc, n); rep("DMMA ILP4 per op, 1 warp", 32.0 * n);
  k_dmma<8><<<1, 32>>>(out, cyc, n); rep("DMMA ILP8 per op, 1 warp", 64.0 * n);
  k_dmma<8><<<148, 128>>>(out, cyc, n); rep("DMMA ILP8 per op per warp, 4 warps/SM (1 per SMSP)", 64.0 * n);
  k_dmma<8><<<148, 256>>>(out, cyc, n); rep("DMMA ILP8 per op per warp, 8 warps/SM", 64.0 * n);
  k_dmma<8><<<148, 512>>>(out, cyc, n); rep("DMMA ILP8 per op per warp, 16 warps/SM", 64.0 * n);
  k_dmma<4><<<148, 896>>>(out, cyc, n); rep("DMMA ILP4 per op per warp, 28 warps/SM", 32.0 * n);
  k_dmma<1><<<148, 896>>>(out, cyc, n); rep("DMMA ILP1 per op per warp, 28 warps/SM", 8.0 * n);
  return 0;
}
