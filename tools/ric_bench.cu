// Stand-alone timing / correctness harness for the Riccati sweep kernels on synthetic stage data (B200 tuning runs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo \
//        -I hilo_mpc_b200/csrc [-DHB_R6_MODE=1|2] [-DPFD=2] tools/ric_bench.cu -o ric_bench && ./ric_bench [B]
// Compares riccati6 (warp per instance, tensor pipe) with the one-thread-per-instance reference sweep on the same data
// and prints the average launch time of both (CUDA events, 20 launches after 3 warm-ups, L2 flushed in between by the
// working set itself: 180 MB per launch at B = 4096).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "riccati6.cuh"

#ifndef PFD
#define PFD 2
#endif
#ifndef NXD
#define NXD 4
#endif
#ifndef NUD
#define NUD 2
#endif

struct Prob {
  static constexpr int NX = NXD, NU = NUD, N = 50, NC = 12, NSP = 0, NTP = 0, NUOLD = 0, NAP = 0;
  static constexpr bool FIXED_X0 = true;
};
using T = hb::OcpT<Prob>;
using R6 = hb::Riccati6<Prob, PFD>;

__global__ void ref_kernel(hb::OcpArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < a.B) hb::OcpRiccati<Prob>::run(a, b);
}

static double frand() { return rand() / (double)RAND_MAX - 0.5; }

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 4096;
  const int N = T::N, NX = T::NX, NU = T::NU;
  std::vector<double> qa((size_t)B * N * T::NAp, 0.0), qh((size_t)B * N * T::NHp, 0.0), qpN((size_t)B * T::SDN), st((size_t)B * hb::ST_COUNT, 0.0);
  srand(1);
  for (int b = 0; b < B; ++b) {
    for (int k = 0; k < N; ++k) {
      double* a = &qa[((size_t)b * N + k) * T::NAp];
      double* h = &qh[((size_t)b * N + k) * T::NHp];
      for (int e = 0; e < T::NHp; ++e) h[e] = 0.0;
      for (int i = 0; i < NX; ++i) {
        for (int j = 0; j < NX; ++j) a[T::iA(i, j)] = (i == j ? 0.9 : 0.0) + 0.1 * frand();
        for (int c = 0; c < NU; ++c) a[T::iB(i, c)] = 0.3 * frand();
        a[T::ib(i)] = 0.1 * frand();
      }
      // cost: G^T G + I on the joint (x, u) space, so that [Q S^T; S R] is positive definite (one stage in 64 is made
      // indefinite in R when argv[2] is given: exercises the regularising slow path)
      double G[6][6];
      for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) G[i][j] = frand();
      auto hij = [&](int i, int j) { double s = (i == j) ? 1.0 : 0.0; for (int l = 0; l < 6; ++l) s += G[l][i] * G[l][j]; return s; };
      for (int i = 0; i < NX; ++i) {
        for (int j = 0; j < NX; ++j) h[T::iQ(i, j)] = hij(i, j);
        h[T::iq(i)] = frand(); h[T::iqb(i)] = frand();
      }
      for (int c = 0; c < NU; ++c) {
        for (int j = 0; j < NX; ++j) { h[T::iS(c, j)] = hij(4 + c, j); if (T::FRAG) h[T::iSt(c, j)] = hij(4 + c, j); }
        for (int d = 0; d < NU; ++d) h[T::iR(c, d)] = hij(4 + c, 4 + d);
        h[T::ir(c)] = frand(); h[T::irb(c)] = frand();
      }
      if (argc > 2 && ((b * N + k) % 64) == 7) h[T::iR(0, 0)] = -5.0;
      T::fill_pads(a, h, 0, 1);
    }
    double* t = &qpN[(size_t)b * T::SDN];
    for (int i = 0; i < NX; ++i) {
      for (int j = 0; j < NX; ++j) t[i * NX + j] = (i == j) ? 2.0 : 0.1;
      t[NX * NX + i] = frand(); t[NX * NX + NX + i] = frand();
    }
    st[(size_t)b * hb::ST_COUNT + hb::ST_ACTIVE] = 1.0;
  }
  hb::OcpArgs a{};
  a.B = B;
  auto up = [&](const std::vector<double>& v) { double* d; cudaMalloc(&d, v.size() * 8); cudaMemcpy(d, v.data(), v.size() * 8, cudaMemcpyHostToDevice); return d; };
  auto al = [&](size_t n) { double* d; cudaMalloc(&d, n * 8); cudaMemset(d, 0, n * 8); return d; };
  a.qa = up(qa); a.qh = up(qh); a.qpN = up(qpN); a.st = up(st);
  a.fb = al((size_t)B * N * T::KD);
  int* status; cudaMalloc(&status, B * 4); cudaMemset(status, 0, B * 4); a.status = status;
  const size_t nx = (size_t)B * (N + 1) * NX, nu = (size_t)B * N * NU;
  double *dx[2], *du[2], *lam[2];
  for (int v = 0; v < 2; ++v) { dx[v] = al(nx); du[v] = al(nu); lam[v] = al(nx); }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms[2] = {0, 0};
  for (int v = 0; v < 2; ++v) {
    a.dx = dx[v]; a.du = du[v]; a.lam = lam[v];
    const int reps = 20;
    for (int it = 0; it < 3 + reps; ++it) {
      if (it == 3) cudaEventRecord(e0);
      if (v == 0) ref_kernel<<<(B + 127) / 128, 128>>>(a);
      else { int rc = hb::launch_riccati6<R6>(a, 0); if (rc) { printf("launch rc %d\n", rc); return 1; } }
    }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
    cudaEventElapsedTime(&ms[v], e0, e1); ms[v] /= reps;
  }
  std::vector<double> h0(nx), h1(nx), u0(nu), u1(nu), l0(nx), l1(nx), s1((size_t)B * hb::ST_COUNT);
  cudaMemcpy(h0.data(), dx[0], nx * 8, cudaMemcpyDeviceToHost); cudaMemcpy(h1.data(), dx[1], nx * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(u0.data(), du[0], nu * 8, cudaMemcpyDeviceToHost); cudaMemcpy(u1.data(), du[1], nu * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(l0.data(), lam[0], nx * 8, cudaMemcpyDeviceToHost); cudaMemcpy(l1.data(), lam[1], nx * 8, cudaMemcpyDeviceToHost);
  auto maxdiff = [](const std::vector<double>& x, const std::vector<double>& y) { double m = 0, s = 0; for (size_t i = 0; i < x.size(); ++i) { double d = fabs(x[i] - y[i]); if (!(d <= m)) m = d; if (fabs(x[i]) > s) s = fabs(x[i]); } printf("  (scale %.3g)", s); return m; };
  printf("B=%d PF=%d mode=%d smem/CTA=%d\n", B, PFD, HB_R6_MODE, R6::SMEM_BYTES);
  printf("max |dx - ref| = %.3e\n", maxdiff(h0, h1));
  printf("max |du - ref| = %.3e\n", maxdiff(u0, u1));
  printf("max |lam - ref| = %.3e\n", maxdiff(l0, l1));
  const double bytes = (double)B * ((double)N * (28 + 40) + 24 + (N + 1) * 2.0 * NX + N * NU) * 8;
  printf("reference sweep (thread per instance): %.1f us\n", ms[0] * 1e3);
  printf("riccati6: %.1f us  -> %.0f GB/s algorithmic (%.1f MB)\n", ms[1] * 1e3, bytes / (ms[1] * 1e-3) * 1e-9, bytes * 1e-6);
  return 0;
}
