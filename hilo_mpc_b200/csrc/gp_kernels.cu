// Gaussian-process regression kernels (squared-exponential covariance, exact inference) — model independent.
//
// Replace the reference's symbolic graph (hilo_mpc/modules/machine_learning/gp/inference.py:197-217: K = kernel(X, X),
// L = chol(sigma_n^2 I + K), alpha = L^-T L^-1 (y - m), LML, mu* = m + K*^T alpha, var* = k** - v^T v with v = L^-1 K*;
// SE covariance hilo_mpc/modules/machine_learning/gp/kernel.py:650-734: k = sigma_f^2 exp(-1/2 sum_k ((x_k - x'_k)/l_k)^2)).
// Data layout as in the reference: X is (n_features, n_observations), row-major.
//
// The dense contractions (trailing update of the blocked Cholesky, L^-1 K* for the predictive variance) run on the
// fp64 tensor pipe (mma.sync m8n8k4 f64 — tcgen05 has no fp64 path); the Gram/cross-covariance entries need one exp
// each and are elementwise.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../include/hilo_b200.h"

namespace {

// ------------------------------------------------------------------------------------------------ covariance
// C[i][j] = sf2 * g(r) with r^2 = sum_k ((A[k][i] - B[k][j]) / ls[k])^2 (+ noise on the diagonal when A == B);
// kind 0: squared exponential g = exp(-r^2/2) (kernel.py:650-734); Matern family (kernel.py:783-826, d = sqrt(2 nu) r):
// 1: exponential (nu = 1/2) g = exp(-r), 2: Matern 3/2 g = (1 + d) exp(-d), 3: Matern 5/2 g = (1 + d + d^2/3) exp(-d)
__global__ void se_cov_kernel(int kind, int d, int na, int nb, const double* __restrict__ A,
                              const double* __restrict__ Bm, const double* __restrict__ ls, double sf2, double noise,
                              double* __restrict__ C, int ldc) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= nb) return;
  double d2 = 0.0;
  for (int k = 0; k < d; ++k) {
    const double r = (A[(size_t)k * na + i] - Bm[(size_t)k * nb + j]) / ls[k];
    d2 += r * r;
  }
  double v;
  if (kind == 0) {
    v = sf2 * exp(-0.5 * d2);
  } else {
    const double r = sqrt(d2);
    if (kind == 1) v = sf2 * exp(-r);
    else if (kind == 2) { const double t = 1.7320508075688772 * r; v = sf2 * (1.0 + t) * exp(-t); }
    else { const double t = 2.23606797749979 * r; v = sf2 * (1.0 + t + t * t * (1.0 / 3.0)) * exp(-t); }
  }
  if (noise != 0.0 && i == j) v += noise;
  C[(size_t)i * ldc + j] = v;
}

// ------------------------------------------------------------------------------------------------ DMMA GEMM
// C (M x N, row-major, ldc) = beta * C + alpha * A (M x K, row-major, lda) * op(B)
//   transB == 0: B is K x N row-major (ldb);  transB == 1: B is N x K row-major (ldb), i.e. C += A * B^T
// lower_only: skip tiles strictly above the diagonal (symmetric rank-k update); tri_k: A is lower triangular, so the
// K loop of a row tile stops at its last column.
constexpr int TM = 32, TN = 32, TK = 16;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128) dgemm_dmma_kernel(int M, int N, int K, double alpha, const double* __restrict__ A,
                                                         int lda, const double* __restrict__ B, int ldb, int transB,
                                                         double beta, double* __restrict__ C, int ldc, int lower_only,
                                                         int tri_k) {
  const int tm = blockIdx.y * TM, tn = blockIdx.x * TN;
  if (lower_only && tn > tm + TM - 1) return;
  __shared__ double As[TM][TK + 1];
  __shared__ double Bs[TK][TN + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = (warp >> 1) * 16, wn = (warp & 1) * 16;  // each warp: 16 x 16 = 2 x 2 mma tiles
  double acc[2][2][2] = {};
  const int kmax = tri_k ? min(K, tm + TM) : K;
  for (int k0 = 0; k0 < kmax; k0 += TK) {
    for (int e = threadIdx.x; e < TM * TK; e += 128) {
      const int r = e / TK, c = e % TK;
      const int gr = tm + r, gc = k0 + c;
      As[r][c] = (gr < M && gc < K) ? A[(size_t)gr * lda + gc] : 0.0;
    }
    for (int e = threadIdx.x; e < TK * TN; e += 128) {
      const int r = e / TN, c = e % TN;  // r: k index, c: n index
      const int gk = k0 + r, gn = tn + c;
      double v = 0.0;
      if (gk < K && gn < N) v = transB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn];
      Bs[r][c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; kk += 4) {
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const double a = As[wm + mi * 8 + (lane >> 2)][kk + (lane & 3)];
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
          const double b = Bs[kk + (lane & 3)][wn + ni * 8 + (lane >> 2)];
          dmma(acc[mi][ni][0], acc[mi][ni][1], a, b);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r = tm + wm + mi * 8 + (lane >> 2), c = tn + wn + ni * 8 + 2 * (lane & 3) + e;
        if (r < M && c < N) {
          const double old = (beta == 0.0) ? 0.0 : beta * C[(size_t)r * ldc + c];
          C[(size_t)r * ldc + c] = old + alpha * acc[mi][ni][e];
        }
      }
}

void dgemm_dmma(int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, int transB,
                double beta, double* C, int ldc, int lower_only, int tri_k, cudaStream_t s) {
  if (M <= 0 || N <= 0) return;
  dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM);
  dgemm_dmma_kernel<<<grid, 128, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, transB, beta, C, ldc, lower_only, tri_k);
}

// ------------------------------------------------------------------------------------------------ blocked Cholesky
constexpr int NB = 32;

// factor the diagonal block (nb x nb, nb <= 32) in place, one CTA of 32 threads; flag != 0 on a non-positive pivot
__global__ void potrf_block_kernel(double* __restrict__ A, int lda, int j0, int nb, int* __restrict__ flag) {
  __shared__ double T[NB][NB + 1];
  const int t = threadIdx.x;
  for (int c = 0; c < nb; ++c)
    if (t < nb) T[t][c] = A[(size_t)(j0 + t) * lda + j0 + c];
  __syncwarp();
  for (int j = 0; j < nb; ++j) {
    double d = T[j][j];
    if (t == 0) {
      if (!(d > 0.0)) { *flag = j0 + j + 1; d = 1.0; }
      T[j][j] = sqrt(d);
    }
    __syncwarp();
    const double piv = T[j][j];
    if (t > j && t < nb) T[t][j] /= piv;
    __syncwarp();
    if (t > j && t < nb)
      for (int c = j + 1; c <= t; ++c) T[t][c] -= T[t][j] * T[c][j];
    __syncwarp();
  }
  for (int c = 0; c < nb; ++c)
    if (t < nb) A[(size_t)(j0 + t) * lda + j0 + c] = (c <= t) ? T[t][c] : 0.0;
}

// rows below the diagonal block: A[i, j0:j0+nb] <- A[i, j0:j0+nb] * L_jj^-T  (one thread per row)
__global__ void trsm_panel_kernel(double* __restrict__ A, int lda, int n, int j0, int nb) {
  __shared__ double Ld[NB][NB + 1];
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) Ld[e / nb][e % nb] = A[(size_t)(j0 + e / nb) * lda + j0 + e % nb];
  __syncthreads();
  const int i = j0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double row[NB];
  for (int c = 0; c < nb; ++c) row[c] = A[(size_t)i * lda + j0 + c];
  for (int c = 0; c < nb; ++c) {
    double s = row[c];
    for (int l = 0; l < c; ++l) s -= row[l] * Ld[c][l];
    row[c] = s / Ld[c][c];
  }
  for (int c = 0; c < nb; ++c) A[(size_t)i * lda + j0 + c] = row[c];
}

__global__ void zero_upper_kernel(double* __restrict__ A, int lda, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j < n && j > i) A[(size_t)i * lda + j] = 0.0;
}

// ------------------------------------------------------------------------------------------------ triangular inverse
// Linv = L^-1 (lower): column j by forward substitution, one CTA per column; rows are processed in order with the
// inner products reduced across the CTA.
__global__ void __launch_bounds__(128) tri_inverse_kernel(int n, const double* __restrict__ L, int ldl,
                                                          double* __restrict__ X, int ldx) {
  extern __shared__ double xs[];  // column j of the inverse
  __shared__ double red[4];
  const int j = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += blockDim.x) xs[i] = 0.0;
  __syncthreads();
  for (int i = j; i < n; ++i) {
    double s = 0.0;
    for (int k = j + threadIdx.x; k < i; k += blockDim.x) s += L[(size_t)i * ldl + k] * xs[k];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      const double tot = red[0] + red[1] + red[2] + red[3];
      xs[i] = (((i == j) ? 1.0 : 0.0) - tot) / L[(size_t)i * ldl + i];
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) X[(size_t)i * ldx + j] = xs[i];
}

// out[i] = sum_k op(M)[i][k] v[k]   (one warp per row; trans: use column i of M)
__global__ void gemv_kernel(int n, const double* __restrict__ Mx, int ld, int trans, const double* __restrict__ v,
                            double* __restrict__ out) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  double s = 0.0;
  for (int k = lane; k < n; k += 32) s += (trans ? Mx[(size_t)k * ld + i] : Mx[(size_t)i * ld + k]) * v[k];
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) out[i] = s;
}

__global__ void residual_kernel(int n, const double* __restrict__ y, double mean, double* __restrict__ r) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) r[i] = y[i] - mean;
}

// lml = -1/2 r^T alpha - sum log L_ii - n/2 log(2 pi)   (single CTA)
__global__ void lml_kernel(int n, const double* __restrict__ r, const double* __restrict__ alpha,
                           const double* __restrict__ L, int ldl, double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += -0.5 * r[i] * alpha[i] - log(L[(size_t)i * ldl + i]);
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) tot += red[w];
    out[0] = tot - 0.5 * n * log(2.0 * M_PI);
  }
}

// mean[m] = prior + sum_i Ks[i][m] alpha[i];   var[m] = sf2 (+ noise) - sum_i V[i][m]^2
__global__ void predict_reduce_kernel(int n, int M, const double* __restrict__ Ks, const double* __restrict__ V,
                                      const double* __restrict__ alpha, double prior, double sf2, double noise,
                                      double* __restrict__ mean, double* __restrict__ var) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double mu = prior, vv = 0.0;
  for (int i = 0; i < n; ++i) {
    mu += Ks[(size_t)i * M + m] * alpha[i];
    if (V != nullptr) { const double t = V[(size_t)i * M + m]; vv += t * t; }
  }
  mean[m] = mu;
  if (var != nullptr) var[m] = sf2 + noise - vv;
}

// ------------------------------------------------------------------------------------------------ fused prediction
// mean and variance of QT = 32 query points per CTA without materialising K* or V = L^-1 K* in HBM (the unfused path
// writes and re-reads 2 x n x M doubles: 537 MB for n = 512, M = 65536):
//   1. the cross-covariance tile Ks [n][QT] is built in shared memory (one exp per entry) together with the mean
//      sum_i alpha_i Ks[i][q];
//   2. V = Linv Ks row tile by row tile (32 rows) on the fp64 tensor pipe: A = Linv tile [32][16] per k-step, staged
//      through a double-buffered cp.async ring (Linv is L2 resident: n^2 / 2 doubles), B = the resident Ks; the K loop of
//      a row tile stops at its diagonal block (Linv is lower triangular);
//   3. the epilogue of a row tile squares its accumulators and adds them to the per-query sum ||V[:, q]||^2 - V itself is
//      never stored.
// 8 warps (one CTA per SM: the Ks tile of n = 512 takes 135 KB), warp w owns the 8 x 16 part (rows 8 (w >> 1), columns
// 16 (w & 1)) of the 32 x 32 output tile = 2 independent DMMA accumulators; n a multiple of 32 (else the unfused path).
constexpr int QT = 32, FK = 16, FT = 256, TMR = 64, NSTAGE = 4;   // queries per CTA, k-step, threads, rows of L^-1 per row block, tiles of the cp.async ring
// leading dimensions = 4 (mod 16) doubles: the 16 lanes of a half-warp read fragment element (k = lane & 3, row or
// column = lane >> 2) at k * LD + (lane >> 2), i.e. 4 k + (lane >> 2) (mod 16) - sixteen different 8-byte banks
constexpr int KS_LD = QT + 4, ALD = FK + 4;

__device__ __forceinline__ void cp_async16_gp(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

__global__ void __launch_bounds__(FT) gp_predict_fused_kernel(int kind, int n, int d, int M, const double* __restrict__ X,
                                                              const double* __restrict__ Xq, const double* __restrict__ ls,
                                                              double sf2, double noise, double prior,
                                                              const double* __restrict__ Linv, const double* __restrict__ alpha,
                                                              double* __restrict__ mean, double* __restrict__ var) {
  extern __shared__ __align__(16) double sm[];
  double* Ks = sm;                                   // [n][KS_LD]
  double* As = Ks + (size_t)n * KS_LD;               // [NSTAGE][TMR][ALD]
  double* colsum = As + NSTAGE * TMR * ALD;          // [QT] sum of squares, [QT] mean
  double* musum = colsum + QT;
  const int q0 = blockIdx.x * QT, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < QT) { colsum[tid] = 0.0; musum[tid] = 0.0; }
  __syncthreads();
  // ---- 1. cross covariance tile + mean: thread (q = tid % 32, rows tid / 32, +8, ...)
  {
    const int q = tid & 31, gq = q0 + q;
    double xq[16], il[16];   // reciprocal length scales once per thread: an fp64 division per (training point, feature) was
                             // half of this kernel's time
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      il[k] = (k < d) ? 1.0 / ls[k] : 0.0;
      xq[k] = (k < d && gq < M) ? Xq[(size_t)k * M + gq] * il[k] : 0.0;
    }
    double mu = 0.0;
    for (int i = tid >> 5; i < n; i += FT / 32) {
      double d2 = 0.0;
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (k < d) { const double r = __ldg(X + (size_t)k * n + i) * il[k] - xq[k]; d2 += r * r; }
      double v;
      if (kind == 0) v = sf2 * exp(-0.5 * d2);
      else {
        const double r = sqrt(d2);
        if (kind == 1) v = sf2 * exp(-r);
        else if (kind == 2) { const double t = 1.7320508075688772 * r; v = sf2 * (1.0 + t) * exp(-t); }
        else { const double t = 2.23606797749979 * r; v = sf2 * (1.0 + t + t * t * (1.0 / 3.0)) * exp(-t); }
      }
      if (gq >= M) v = 0.0;
      Ks[(size_t)i * KS_LD + q] = v;
      mu += v * alpha[i];
    }
    atomicAdd(&musum[q], mu);
  }
  __syncthreads();
  if (var != nullptr) {
    // warp tile: 16 rows of L^-1 (two fragments) x 16 queries (two fragments): 4 shared-memory loads per 4 DMMA
    const int wm = (warp >> 1) * 16, wn = (warp & 1) * 16;
    auto load_stage = [&](int buf, int tm, int k0) {   // TMR rows x 16 doubles = 512 chunks of 16 bytes, two per thread
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = tid + h * FT, r = c >> 3, cc = (c & 7) * 2;
        cp_async16_gp(As + (size_t)buf * TMR * ALD + r * ALD + cc, Linv + (size_t)(tm + r) * n + k0 + cc);
      }
    };
    // ONE pipeline over all (row block tm, k-step ks) pairs of the lower triangle, NSTAGE - 1 tiles in flight: the CTA
    // is alone on its SM (the K* tile takes the shared memory), so only depth hides the L2 latency of the L^-1 tiles,
    // and a pipeline that drained at every row block would pay that latency n / TMR times
    double csq[2][2] = {{0.0, 0.0}, {0.0, 0.0}};   // partial sums of squares of this thread's columns (ni, e)
    double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
    int ltm = 0, lks = 0, lstep = 0;                // load cursor: row block, k-step, tile number
    auto issue_next = [&]() {
      if (ltm < n) {
        load_stage(lstep % NSTAGE, ltm, lks * FK);
        ++lstep;
        if (++lks == (ltm + TMR) / FK) { lks = 0; ltm += TMR; }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");   // (possibly empty) group: the count stays uniform
    };
    int step = 0;
    for (int tm = 0; tm < n; tm += TMR) {
      const int nk = (tm + TMR) / FK;      // k-steps up to the diagonal block
      for (int ks = 0; ks < nk; ++ks, ++step) {
        if (step == 0) {
#pragma unroll
          for (int st = 0; st < NSTAGE - 1; ++st) issue_next();
        }
        asm volatile("cp.async.wait_group %0;" ::"n"(NSTAGE - 2) : "memory");
        __syncthreads();       // tile `step` has landed for everybody; everybody is done with tile step - 1
        issue_next();          // ... whose buffer takes tile step + NSTAGE - 1
        const double* A0 = As + (size_t)(step % NSTAGE) * TMR * ALD;
        const int k0 = ks * FK;
        // L^-1 is lower triangular: k-steps right of the warp's last row hold zeros only
        if (k0 < tm + wm + 16) {
#pragma unroll
          for (int kk = 0; kk < FK; kk += 4) {
            double av[2], bv[2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) av[mi] = A0[(wm + mi * 8 + (lane >> 2)) * ALD + kk + (lane & 3)];
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) bv[ni] = Ks[(size_t)(k0 + kk + (lane & 3)) * KS_LD + wn + ni * 8 + (lane >> 2)];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
              for (int ni = 0; ni < 2; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], av[mi], bv[ni]);
          }
        }
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
          csq[ni][0] += acc[mi][ni][0] * acc[mi][ni][0];
          csq[ni][1] += acc[mi][ni][1] * acc[mi][ni][1];
          acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // column c = wn + ni * 8 + 2 * (lane & 3) + e is shared by the 8 lanes with equal (lane & 3): reduce over lane >> 2
#pragma unroll
    for (int ni = 0; ni < 2; ++ni)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double v = csq[ni][e];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if ((lane >> 2) == 0) atomicAdd(&colsum[wn + ni * 8 + 2 * (lane & 3) + e], v);
      }
    __syncthreads();
  }
  if (tid < QT && q0 + tid < M) {
    mean[q0 + tid] = prior + musum[tid];
    if (var != nullptr) var[q0 + tid] = sf2 + noise - colsum[tid];
  }
}

// mean only: one thread per query point walks the training set (every thread of a warp reads the same training row: one
// broadcast per load); nothing but Xq and the means touches HBM
__global__ void __launch_bounds__(128) gp_mean_only_kernel(int kind, int n, int d, int M, const double* __restrict__ X,
                                                           const double* __restrict__ Xq, const double* __restrict__ ls,
                                                           double sf2, double prior, const double* __restrict__ alpha,
                                                           double* __restrict__ mean) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double xq[16], il[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) { il[k] = k < d ? 1.0 / ls[k] : 0.0; xq[k] = k < d ? Xq[(size_t)k * M + m] * il[k] : 0.0; }
  double mu = 0.0;
  for (int i = 0; i < n; ++i) {
    double d2 = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < d) { const double r = __ldg(X + (size_t)k * n + i) * il[k] - xq[k]; d2 += r * r; }
    double v;
    if (kind == 0) v = exp(-0.5 * d2);
    else {
      const double r = sqrt(d2);
      if (kind == 1) v = exp(-r);
      else if (kind == 2) { const double t = 1.7320508075688772 * r; v = (1.0 + t) * exp(-t); }
      else { const double t = 2.23606797749979 * r; v = (1.0 + t + t * t * (1.0 / 3.0)) * exp(-t); }
    }
    mu += v * __ldg(alpha + i);
  }
  mean[m] = prior + sf2 * mu;
}

// ------------------------------------------------------------------------------ gradient of the log marginal likelihood
// dNLML/dtheta_j = 1/2 tr(W dK/dtheta_j) with W = K^-1 - alpha alpha^T (Rasmussen & Williams (5.9); gp.py:660-697 obtains
// it by automatic differentiation of the CasADi graph).  Kinv comes from the own DMMA GEMM (L^-T L^-1); this kernel then
// makes ONE pass over the (i, j) pairs, recomputes r_ij and the covariance factors on the fly (nothing of size d x n x n is
// ever stored) and reduces  sum W,  sum W sf2 g(r),  sum W sf2 h(r) D2_k  (dK/dlog l_k = sf2 h(r) D2_k, D2_k = ((x_k -
// x'_k) / l_k)^2;  SE: h = g;  exponential: h = g / r;  Matern 3/2: h = 3 e^-t;  5/2: h = 5/3 (1 + t) e^-t).
// out = [1/2 sigma_n^2 tr W | 1/2 sum_k-terms (d) | 1/2 sum W K]  - derivatives with respect to the LOG parameters.
__global__ void transpose_kernel(int n, const double* __restrict__ A, double* __restrict__ At) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = by + r, j = bx + threadIdx.x;
    tile[r][threadIdx.x] = (i < n && j < n) ? A[(size_t)i * n + j] : 0.0;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = bx + r, j = by + threadIdx.x;
    if (i < n && j < n) At[(size_t)i * n + j] = tile[threadIdx.x][r];
  }
}

__global__ void __launch_bounds__(256) gp_lml_grad_kernel(int kind, int n, int d, const double* __restrict__ X,
                                                          const double* __restrict__ ls, double sf2, double noise,
                                                          const double* __restrict__ Kinv,
                                                          const double* __restrict__ alpha, double* __restrict__ out) {
  const int i = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[18];   // [0] trace part, [1 .. d] length scales, [d + 1] signal variance
#pragma unroll
  for (int k = 0; k < 18; ++k) acc[k] = 0.0;
  if (j < n) {
    const double w = Kinv[(size_t)i * n + j] - alpha[i] * alpha[j];
    double D2[16], r2 = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      D2[k] = 0.0;
      if (k < d) { const double t = (X[(size_t)k * n + i] - X[(size_t)k * n + j]) / ls[k]; D2[k] = t * t; r2 += D2[k]; }
    }
    double g, h;
    if (kind == 0) { g = exp(-0.5 * r2); h = g; }
    else {
      const double r = sqrt(r2);
      if (kind == 1) { g = exp(-r); h = (r > 0.0) ? g / r : 0.0; }
      else if (kind == 2) { const double t = 1.7320508075688772 * r, e = exp(-t); g = (1.0 + t) * e; h = 3.0 * e; }
      else { const double t = 2.23606797749979 * r, e = exp(-t); g = (1.0 + t + t * t * (1.0 / 3.0)) * e; h = (5.0 / 3.0) * (1.0 + t) * e; }
    }
    if (i == j) acc[0] = 0.5 * noise * w;
    const double wh = 0.5 * sf2 * w * h;
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < d) acc[1 + k] = wh * D2[k];
    acc[1 + d] = 0.5 * sf2 * w * g;
  }
  __shared__ double red[8][18];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < d + 2; ++k) {
    double v = acc[k];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < d + 2) {
    double v = 0.0;
    for (int wq = 0; wq < (int)(blockDim.x >> 5); ++wq) v += red[wq][threadIdx.x];
    atomicAdd(out + threadIdx.x, v);
  }
}

}  // namespace

extern "C" {

size_t hb_gp_fit_work_doubles(int n) { return (size_t)n * 2 + 8; }

int hb_gp_fit(int n, int d, const double* X, const double* y, const double* length_scales, double signal_variance,
              double noise_variance, double prior_mean, double* L, double* Linv, double* alpha, double* lml,
              double* work, int* info, void* stream) {
  return hb_gp_fit_kernel(0, n, d, X, y, length_scales, signal_variance, noise_variance, prior_mean, L, Linv, alpha,
                          lml, work, info, stream);
}

int hb_gp_fit_kernel(int kind, int n, int d, const double* X, const double* y, const double* length_scales,
                     double signal_variance, double noise_variance, double prior_mean, double* L, double* Linv,
                     double* alpha, double* lml, double* work, int* info, void* stream) {
  if (kind < 0 || kind > 3) return -1;
  if (n <= 0 || d <= 0 || !X || !y || !length_scales || !L || !Linv || !alpha || !lml || !work || !info) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(info, 0, sizeof(int), s);
  {
    dim3 grid((n + 127) / 128, n);
    se_cov_kernel<<<grid, 128, 0, s>>>(kind, d, n, n, X, X, length_scales, signal_variance, noise_variance, L, n);
  }
  for (int j0 = 0; j0 < n; j0 += NB) {
    const int nb = (n - j0 < NB) ? (n - j0) : NB;
    potrf_block_kernel<<<1, 32, 0, s>>>(L, n, j0, nb, info);
    const int rest = n - j0 - nb;
    if (rest > 0) {
      trsm_panel_kernel<<<(rest + 63) / 64, 64, 0, s>>>(L, n, n, j0, nb);
      // trailing update A22 -= L21 L21^T on the tensor pipe (lower triangle only)
      const double* L21 = L + (size_t)(j0 + nb) * n + j0;
      double* A22 = L + (size_t)(j0 + nb) * n + (j0 + nb);
      dgemm_dmma(rest, rest, nb, -1.0, L21, n, L21, n, 1, 1.0, A22, n, 1, 0, s);
    }
  }
  {
    dim3 grid((n + 127) / 128, n);
    zero_upper_kernel<<<grid, 128, 0, s>>>(L, n, n);
  }
  tri_inverse_kernel<<<n, 128, (size_t)n * sizeof(double), s>>>(n, L, n, Linv, n);
  double* r = work;
  double* tmp = work + n;
  residual_kernel<<<(n + 127) / 128, 128, 0, s>>>(n, y, prior_mean, r);
  gemv_kernel<<<(n + 3) / 4, 128, 0, s>>>(n, Linv, n, 0, r, tmp);      // tmp = L^-1 r
  gemv_kernel<<<(n + 3) / 4, 128, 0, s>>>(n, Linv, n, 1, tmp, alpha);  // alpha = L^-T tmp
  lml_kernel<<<1, 256, 0, s>>>(n, r, alpha, L, n, lml);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

size_t hb_gp_predict_work_doubles(int n, int M) { return (size_t)2 * n * M + 8; }

int hb_gp_predict(int n, int d, int M, const double* X, const double* Xq, const double* length_scales,
                  double signal_variance, double noise_variance, double prior_mean, const double* Linv,
                  const double* alpha, double* mean, double* var, double* work, void* stream) {
  return hb_gp_predict_kernel(0, n, d, M, X, Xq, length_scales, signal_variance, noise_variance, prior_mean, Linv,
                              alpha, mean, var, work, stream);
}

int hb_gp_predict_kernel(int kind, int n, int d, int M, const double* X, const double* Xq, const double* length_scales,
                         double signal_variance, double noise_variance, double prior_mean, const double* Linv,
                         const double* alpha, double* mean, double* var, double* work, void* stream) {
  if (kind < 0 || kind > 3) return -1;
  if (n <= 0 || d <= 0 || M <= 0 || !X || !Xq || !length_scales || !alpha || !mean || !work) return -1;
  if (var != nullptr && Linv == nullptr) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  {  // fused path: no K* / V in HBM (n a multiple of 64, d <= 16, the tile fits shared memory)
    const size_t smem = ((size_t)n * KS_LD + NSTAGE * TMR * ALD + 2 * QT) * sizeof(double);
    static bool fused_off = getenv("HB_GP_UNFUSED") != nullptr;
    if (!fused_off && var == nullptr && d <= 16) {
      gp_mean_only_kernel<<<(M + 127) / 128, 128, 0, s>>>(kind, n, d, M, X, Xq, length_scales, signal_variance, prior_mean,
                                                          alpha, mean);
      cudaError_t e = cudaGetLastError();
      return e == cudaSuccess ? 0 : (int)e;
    }
    if (!fused_off && n % TMR == 0 && d <= 16 && smem <= 220 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(gp_predict_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      gp_predict_fused_kernel<<<(M + QT - 1) / QT, FT, smem, s>>>(kind, n, d, M, X, Xq, length_scales, signal_variance,
                                                                   noise_variance, prior_mean, Linv, alpha, mean, var);
      e = cudaGetLastError();
      return e == cudaSuccess ? 0 : (int)e;
    }
  }
  double* Ks = work;                  // [n][M] cross covariance
  double* V = work + (size_t)n * M;   // [n][M] L^-1 K*
  dim3 grid((M + 127) / 128, n);
  se_cov_kernel<<<grid, 128, 0, s>>>(kind, d, n, M, X, Xq, length_scales, signal_variance, 0.0, Ks, M);
  if (var != nullptr) dgemm_dmma(n, M, n, 1.0, Linv, n, Ks, M, 0, 0.0, V, M, 0, 1, s);
  predict_reduce_kernel<<<(M + 127) / 128, 128, 0, s>>>(n, M, Ks, var ? V : nullptr, alpha, prior_mean,
                                                        signal_variance, noise_variance, mean, var);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

size_t hb_gp_lml_grad_work_doubles(int n) { return (size_t)2 * n * n; }

int hb_gp_lml_grad(int kind, int n, int d, const double* X, const double* length_scales, double signal_variance,
                   double noise_variance, const double* Linv, const double* alpha, double* work, double* grad,
                   void* stream) {
  if (kind < 0 || kind > 3 || n <= 0 || d <= 0 || d > 16) return -1;
  if (!X || !length_scales || !Linv || !alpha || !work || !grad) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  double* LinvT = work;                     // [n][n]
  double* Kinv = work + (size_t)n * n;      // [n][n] = L^-T L^-1
  {
    dim3 grid((n + 31) / 32, (n + 31) / 32), block(32, 8);
    transpose_kernel<<<grid, block, 0, s>>>(n, Linv, LinvT);
  }
  dgemm_dmma(n, n, n, 1.0, LinvT, n, Linv, n, 0, 0.0, Kinv, n, 0, 0, s);
  cudaMemsetAsync(grad, 0, (size_t)(d + 2) * sizeof(double), s);
  {
    dim3 grid((n + 255) / 256, n);
    gp_lml_grad_kernel<<<grid, 256, 0, s>>>(kind, n, d, X, length_scales, signal_variance, noise_variance, Kinv, alpha,
                                            grad);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // extern "C"
