// Latency / issue-rate probes for the fp64 kernels (dependent DFMA chain, independent DFMAs, LDS, SHFL, rcp) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma_dep(double* out, long long* cyc, double a, double b, int n) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x = fma(x, a, b);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void k_dfma_ilp(double* out, long long* cyc, double a, double b, int n) {
  double x[ILP];
  for (int j = 0; j < ILP; ++j) x[j] = out[threadIdx.x] + j;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int j = 0; j < ILP; ++j) s += x[j];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_lds_dep(double* out, long long* cyc, int n) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i + 32) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 32; ++j) p = idx[p];
  }
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_shfl_dep(double* out, long long* cyc, int n) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 3, 4);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_rcp_dep(double* out, long long* cyc, int n) {
  double x = out[threadIdx.x] + 1.5;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double r;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
      double e = fma(-x, r, 1.0);
      r = fma(r, e, r);
      e = fma(-x, r, 1.0);
      x = fma(r, e, r) + 1.0;
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_sts_lds(double* out, long long* cyc, int n) {   // store -> syncwarp -> load round trip
  __shared__ double buf[256];
  double x = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      buf[threadIdx.x] = x;
      __syncwarp();
      x = buf[threadIdx.x ^ 1] + 1.0;
      __syncwarp();
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMemset(out, 0, 1024 * 8);
  cudaMallocManaged(&cyc, 64 * 8);
  const int n = 2000;
  auto report = [&](const char* name, double per) { cudaDeviceSynchronize(); printf("%-40s %8.2f cycles\n", name, (double)cyc[0] / per); };
  for (int warps = 1; warps <= 4; warps *= 2) {
    printf("--- %d warp(s) per SM sub-partition (block of %d threads, 1 block)\n", warps, warps * 128);
    k_dfma_dep<<<1, 128 * warps>>>(out, cyc, 1.0000001, 1e-9, n); report("DFMA dependent, per op", 32.0 * n);
    k_dfma_ilp<2><<<1, 128 * warps>>>(out, cyc, 1.0000001, 1e-9, n); report("DFMA ILP2, per op", 16.0 * n);
    k_dfma_ilp<4><<<1, 128 * warps>>>(out, cyc, 1.0000001, 1e-9, n); report("DFMA ILP4, per op", 32.0 * n);
    k_dfma_ilp<8><<<1, 128 * warps>>>(out, cyc, 1.0000001, 1e-9, n); report("DFMA ILP8, per op", 64.0 * n);
    k_lds_dep<<<1, 128 * warps>>>(out, cyc, n); report("LDS.32 dependent, per load", 32.0 * n);
    k_shfl_dep<<<1, 128 * warps>>>(out, cyc, n); report("SHFL f64 (2x32) dependent, per shfl", 32.0 * n);
    k_rcp_dep<<<1, 128 * warps>>>(out, cyc, n); report("rcp.approx+2 Newton+DADD, per chain", 8.0 * n);
    k_sts_lds<<<1, 128 * warps>>>(out, cyc, n); report("STS+syncwarp+LDS+DADD+syncwarp", 8.0 * n);
  }
  k_dfma_dep<<<1, 32>>>(out, cyc, 1.0000001, 1e-9, n); report("single warp: DFMA dependent", 32.0 * n);
  k_dfma_ilp<8><<<1, 32>>>(out, cyc, 1.0000001, 1e-9, n); report("single warp: DFMA ILP8 per op", 64.0 * n);
  return 0;
}
