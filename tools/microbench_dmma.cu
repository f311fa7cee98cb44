#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// m16n8k8: A 16x8 (4 regs), B 8x8 (2 regs), C 16x8 (4 regs)
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
// m16n8k16: A 16x16 (8 regs), B 16x8 (4 regs)
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
template <int MODE>
__global__ void k(double* out, int n) {
  double acc[8][4] = {};
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = threadIdx.x * 2e-3 + i;
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) { dmma884(acc[j][0], acc[j][1], a[0], b[0]); }
      else if (MODE == 1) { double a4[4] = {a[0], a[1], a[2], a[3]}; double b2[2] = {b[0], b[1]}; dmma1688(acc[j], a4, b2); }
      else dmma16816(acc[j], a, b);
    }
  }
  double s = 0; for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) s += acc[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 8 * 256 * 8);
  const int n = 20000;
  const double fl[3] = {8. * 8 * 4 * 2, 16. * 8 * 8 * 2, 16. * 8 * 16 * 2};
  for (int warps = 4; warps <= 16; warps *= 2)
  for (int mode = 0; mode < 3; ++mode) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148, warps * 32>>>(out, n); else if (mode == 1) k<1><<<148, warps * 32>>>(out, n); else k<2><<<148, warps * 32>>>(out, n);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("mode %d (%s) warps/SM %d: %.3f ms  %.2f TFLOP/s  err %d\n", mode, mode == 0 ? "m8n8k4" : mode == 1 ? "m16n8k8" : "m16n8k16", warps, ms,
           148.0 * warps * 8 * n * fl[mode] / ms / 1e9, (int)cudaGetLastError());
  }
  return 0;
}
