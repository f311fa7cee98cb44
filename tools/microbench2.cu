// Why does a lone warp issue fp64 so slowly?  Independent DFMA streams under different launch shapes.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int LANES>
__global__ void k_dfma(double* out, long long* cyc, double a, double b, int n) {
  double x[ILP];
  for (int j = 0; j < ILP; ++j) x[j] = out[threadIdx.x] + j;
  long long t0 = clock64();
  if ((threadIdx.x & 31) < LANES) {
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int j = 0; j < ILP; ++j) s += x[j];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP>
__global__ void k_imad(int* out, long long* cyc, int a, int b, int n) {
  int x[ILP];
  for (int j = 0; j < ILP; ++j) x[j] = out[threadIdx.x] + j;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < ILP; ++j) x[j] = x[j] * a + b;
  }
  long long t1 = clock64();
  int s = 0;
  for (int j = 0; j < ILP; ++j) s += x[j];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP>
__global__ void k_ffma(float* out, long long* cyc, float a, float b, int n) {
  float x[ILP];
  for (int j = 0; j < ILP; ++j) x[j] = out[threadIdx.x] + j;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < ILP; ++j) x[j] = fmaf(x[j], a, b);
  }
  long long t1 = clock64();
  float s = 0;
  for (int j = 0; j < ILP; ++j) s += x[j];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 4096 * 8); cudaMemset(out, 0, 4096 * 8);
  cudaMallocManaged(&cyc, 64 * 8);
  const int n = 2000;
  auto rep = [&](const char* name, double per) { cudaDeviceSynchronize(); printf("%-60s %8.2f cycles/op/warp\n", name, (double)cyc[0] / per); };
  for (int rpt = 0; rpt < 2; ++rpt) {
    k_dfma<8, 32><<<1, 32>>>(out, cyc, 1.0000001, 1e-9, n); rep("DFMA ILP8  grid 1   block 32", 64.0 * n);
    k_dfma<8, 32><<<1, 64>>>(out, cyc, 1.0000001, 1e-9, n); rep("DFMA ILP8  grid 1   block 64", 64.0 * n);
    k_dfma<8, 32><<<1, 96>>>(out, cyc, 1.0000001, 1e-9, n); rep("DFMA ILP8  grid 1   block 96", 64.0 * n);
    k_dfma<8, 32><<<1, 128>>>(out, cyc, 1.0000001, 1e-9, n); rep("DFMA ILP8  grid 1   block 128", 64.0 * n);
    k_dfma<8, 32><<<1, 256>>>(out, cyc, 1.0000001, 1e-9, n); rep("DFMA ILP8  grid 1   block 256", 64.0 * n);
    k_dfma<8, 32><<<148, 32>>>(out, cyc, 1.0000001, 1e-9, n); rep("DFMA ILP8  grid 148 block 32", 64.0 * n);
    k_dfma<8, 32><<<296, 32>>>(out, cyc, 1.0000001, 1e-9, n); rep("DFMA ILP8  grid 296 block 32", 64.0 * n);
    k_dfma<8, 32><<<592, 32>>>(out, cyc, 1.0000001, 1e-9, n); rep("DFMA ILP8  grid 592 block 32", 64.0 * n);
    k_dfma<8, 32><<<148, 128>>>(out, cyc, 1.0000001, 1e-9, n); rep("DFMA ILP8  grid 148 block 128", 64.0 * n);
    k_dfma<8, 8><<<148, 128>>>(out, cyc, 1.0000001, 1e-9, n); rep("DFMA ILP8  grid 148 block 128, 8 lanes active", 64.0 * n);
    k_dfma<8, 16><<<148, 64>>>(out, cyc, 1.0000001, 1e-9, n); rep("DFMA ILP8  grid 148 block 64, 16 lanes active", 64.0 * n);
    k_dfma<8, 1><<<1, 32>>>(out, cyc, 1.0000001, 1e-9, n); rep("DFMA ILP8  grid 1 block 32, 1 lane active", 64.0 * n);
    k_dfma<4, 32><<<1, 32>>>(out, cyc, 1.0000001, 1e-9, n); rep("DFMA ILP4  grid 1   block 32", 32.0 * n);
    k_dfma<16, 32><<<1, 32>>>(out, cyc, 1.0000001, 1e-9, n); rep("DFMA ILP16 grid 1   block 32", 128.0 * n);
    k_dfma<1, 32><<<1, 32>>>(out, cyc, 1.0000001, 1e-9, n); rep("DFMA ILP1  grid 1   block 32", 8.0 * n);
    k_dfma<1, 32><<<1, 128>>>(out, cyc, 1.0000001, 1e-9, n); rep("DFMA ILP1  grid 1   block 128", 8.0 * n);
    k_imad<8><<<1, 32>>>((int*)out, cyc, 3, 7, n); rep("IMAD ILP8  grid 1   block 32", 64.0 * n);
    k_imad<8><<<1, 128>>>((int*)out, cyc, 3, 7, n); rep("IMAD ILP8  grid 1   block 128", 64.0 * n);
    k_ffma<8><<<1, 32>>>((float*)out, cyc, 1.0001f, 1e-5f, n); rep("FFMA ILP8  grid 1   block 32", 64.0 * n);
    k_ffma<8><<<1, 128>>>((float*)out, cyc, 1.0001f, 1e-5f, n); rep("FFMA ILP8  grid 1   block 128", 64.0 * n);
    printf("\n");
  }
  return 0;
}
