// Riccati sweep, sm_100a version: bulk-async (TMA) staged stage data, LPI lanes per instance.
//
// Work decomposition (NX=4, NU=2 -> LPI=8, 4 instances per warp, 16 per CTA, 2 CTAs per SM => 4736 instance slots
// on 148 SMs, i.e. the whole BASELINE batch of 4096 is resident in one wave):
//   * stage blocks stream HBM -> shared memory through a per-instance ring of NBUF slots filled by
//     cp.async.bulk (1-D TMA) and tracked by mbarriers; the same ring carries the backward sweep
//     ([A B b | Q S R q r qbar rbar], 544 B per stage) and then the forward sweep ([A B b] only, 224 B, L2 hits),
//   * lane g < NE of an instance owns column g of M = [A B] / of the stage Hessian, lane NE owns the gradient column,
//   * the feedback gains K_k, k_k (4 KB per instance) never leave shared memory,
//   * the adjoint recursion lam_k = qbar_k + A_k^T lam_{k+1} and the reduced gradient rbar_k + B_k^T lam_{k+1}
//     (dual infeasibility) ride along the backward sweep.
// Algorithmic HBM traffic per instance: N*(NAp+NHp)*8 read + SDN*8 read + ((N+1)*2*NX + N*NU)*8 written
// (29 KB for the CSTR N=50 problem; SURVEY.md §8d counts 29 024 B).
#pragma once
#include "ocp_kernels.cuh"

#if HB_DEVICE_BUILD
namespace hb {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 1/x for x > 0: hardware approximation refined by two Newton steps (full fp64 accuracy, no slow-path branch)
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <class P>
struct RiccatiTma {
  using T = OcpT<P>;
  static constexpr int NX = T::NX, NU = T::NU, NE = T::NE, N = T::N;
  static constexpr int LPI = (NE + 1 <= 4) ? 4 : (NE + 1 <= 8) ? 8 : (NE + 1 <= 16) ? 16 : 32;  // lanes per instance
  static constexpr int G = 32 / LPI;          // instances per warp
  static constexpr int WARPS = 4;             // warps per CTA
  static constexpr int NBUF = 4;              // ring slots per instance (one stage each)
  static constexpr int SLOT = T::NAp + T::NHp;  // doubles per slot
  static constexpr int XCH = (NU * NE > NX * NX) ? NU * NE : NX * NX;  // exchange scratch (doubles)
  static constexpr int XCHp = (XCH + 1) & ~1;
  // per-instance shared memory (doubles): ring | feedback gains | exchange | mbarriers; the stride is padded so that
  // the instances of a warp start 8 banks apart (conflict-free 128-bit broadcasts)
  static constexpr int FBD = (N * T::KD + 1) & ~1;  // feedback-gain store, padded to keep 16-byte alignment behind it
  static constexpr int RAW = NBUF * SLOT + FBD + XCHp + NBUF;
  static constexpr int STRIDE = ((RAW + 15) / 16) * 16 + 4;  // doubles; (STRIDE*2) % 32 == 8 words
  static constexpr int SMEM_BYTES = WARPS * G * STRIDE * 8;
  static constexpr bool SUPPORTED = (NE + 1 <= 32) && (NX <= 6) && (SMEM_BYTES <= 110 * 1024);

  static __device__ __forceinline__ double gshfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src, LPI); }
  // arr[idx] for a run-time idx without dynamic register indexing (which would spill the array to local memory)
  template <int M>
  static __device__ __forceinline__ double pick(const double (&arr)[M], int idx) {
    double v = arr[0];
#pragma unroll
    for (int i = 1; i < M; ++i) v = (idx == i) ? arr[i] : v;
    return v;
  }

  static __device__ void run(const OcpArgs& a) {
    extern __shared__ __align__(128) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / LPI, g = lane % LPI;
    const int b = (blockIdx.x * WARPS + warp) * G + grp;
    const bool valid = b < a.B;
    const int bb = valid ? b : (a.B - 1);
    double* st = a.st + (long long)bb * ST_COUNT;
    const bool active = valid && st[ST_ACTIVE] != 0.0;
    if (!__any_sync(0xffffffffu, active)) return;

    double* base = smem + (size_t)(warp * G + grp) * STRIDE;
    double* ring = base;
    double* fbs = ring + NBUF * SLOT;
    double* xch = fbs + FBD;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(xch + XCHp);

    const double* qa_g = a.qa + (long long)bb * N * T::NAp;
    const double* qh_g = a.qh + (long long)bb * N * T::NHp;
    const double* tn = a.qpN + (long long)bb * T::SDN;

    if (g == 0) {
      for (int s = 0; s < NBUF; ++s) mbar_init(&bars[s], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // sequence j = 0..2N-1: backward stage N-1-j (full block) for j < N, forward stage j-N (dynamics part) after
    auto issue = [&](int j) {
      const int s = j % NBUF;
      double* dst = ring + s * SLOT;
      if (j < N) {
        const int k = N - 1 - j;
        mbar_expect_tx(&bars[s], (unsigned)((T::NAp + T::NHp) * 8));
        tma_load_1d(dst, qa_g + (long long)k * T::NAp, T::NAp * 8, &bars[s]);
        tma_load_1d(dst + T::NAp, qh_g + (long long)k * T::NHp, T::NHp * 8, &bars[s]);
      } else {
        const int k = j - N;
        mbar_expect_tx(&bars[s], (unsigned)(T::NAp * 8));
        tma_load_1d(dst, qa_g + (long long)k * T::NAp, T::NAp * 8, &bars[s]);
      }
    };
    if (g == 0)
      for (int j = 0; j < NBUF && j < 2 * N; ++j) issue(j);

    // ---- terminal data: P = Q_N (replicated in every lane), p = q_N, lam = qbar_N
    double Pm[NX][NX], pv[NX], lam[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
      for (int j = 0; j < NX; ++j) Pm[i][j] = tn[i * NX + j];
      pv[i] = tn[NX * NX + i];
      lam[i] = tn[NX * NX + NX + i];
    }
    if (active && g < NX) a.lam[((long long)b * (N + 1) + N) * NX + g] = pick<NX>(lam, g);
    double edual = 0.0;
    bool ok = true;

    // ---- per-lane offsets (doubles, relative to a ring slot), fixed for the whole sweep
    int moff[NX];   // column g of [A B] / the affine column b
    int coff[NE];   // column g of the cost Hessian [Q S^T; S R] / the gradient column [q; r]
#pragma unroll
    for (int l = 0; l < NX; ++l)
      moff[l] = (g < NX) ? (T::oA + l * NX + g) : (g < NE) ? (T::oB + l * NU + (g - NX)) : (T::ob + l);
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      int o;
      if (g < NX) o = (i < NX) ? (T::oQ + i * NX + g) : (T::oS + (i - NX) * NX + g);
      else if (g < NE) o = (i < NX) ? (T::oS + (g - NX) * NX + i) : (T::oR + (i - NX) * NU + (g - NX));
      else o = (i < NX) ? (T::oq + i) : (T::orr + (i - NX));
      coff[i] = T::NAp + o;
    }
    const int goff = T::NAp + T::oqb + (g < NE ? g : 0);

    // =============================================================== backward sweep
    for (int j = 0; j < N; ++j) {
      const int k = N - 1 - j, s = j % NBUF;
      mbar_wait(&bars[s], (unsigned)((j / NBUF) & 1));
      const double* slot = ring + s * SLOT;
      // whole dynamics block [A B] (broadcast 128-bit loads) and the lane's own column
      double Mf[NX * NE + (NX * NE & 1)];
      {
        const double2* v = reinterpret_cast<const double2*>(slot);
#pragma unroll
        for (int i = 0; i < (NX * NE + 1) / 2; ++i) { const double2 t2 = v[i]; Mf[2 * i] = t2.x; if (2 * i + 1 < NX * NE + (NX * NE & 1)) Mf[2 * i + 1] = t2.y; }
      }
      double m[NX];
#pragma unroll
      for (int l = 0; l < NX; ++l) m[l] = slot[moff[l]];
      // adjoint / reduced gradient with lam_{k+1}
      {
        double acc = (g < NE) ? slot[goff] : 0.0;
#pragma unroll
        for (int l = 0; l < NX; ++l) acc += m[l] * lam[l];
        if (g >= NX && g < NE) edual = dmax(edual, dabs(acc));
#pragma unroll
        for (int i = 0; i < NX; ++i) lam[i] = gshfl(acc, i);
        if (active && g < NX) a.lam[((long long)b * (N + 1) + k) * NX + g] = acc;
      }
      // w = P m (+ p for the affine column)
      double w[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double acc = (g >= NE) ? pv[i] : 0.0;
#pragma unroll
        for (int l = 0; l < NX; ++l) acc += Pm[i][l] * m[l];
        w[i] = acc;
      }
      // column g of H = [Q S^T; S R] + M^T P M   (affine column: [q; r] + M^T (P b + p))
      double h[NE];
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        double acc = slot[coff[i]];
#pragma unroll
        for (int l = 0; l < NX; ++l) {
          const double mli = (i < NX) ? Mf[T::oA + l * NX + i] : Mf[T::oB + l * NU + (i - NX)];
          acc += mli * w[l];
        }
        h[i] = acc;
      }
      // exchange: columns NX..NE-1 (rows 0..NE-1) are needed by everybody
      if (g >= NX && g < NE) {
#pragma unroll
        for (int i = 0; i < NE; ++i) xch[(g - NX) * NE + i] = h[i];
      }
      __syncwarp();
      double hu[NU][NE];  // hu[c][i] = H[i][NX + c]
#pragma unroll
      for (int c = 0; c < NU; ++c)
#pragma unroll
        for (int i = 0; i < NE; ++i) hu[c][i] = xch[c * NE + i];
      // Huu = L D L^T (unit lower L, no square roots; reciprocals by Newton-refined rcp.approx), every lane redundantly
      double Lm[NU][NU], dd[NU], dinv[NU];
      {
        double reg = 0.0;
        for (int attempt = 0; attempt < 12; ++attempt) {
          bool pd = true;
#pragma unroll
          for (int jj = 0; jj < NU; ++jj) {
            double d = hu[jj][NX + jj] + reg;
#pragma unroll
            for (int l = 0; l < jj; ++l) d -= Lm[jj][l] * Lm[jj][l] * dd[l];
            pd = pd && (d > 0.0);
            dd[jj] = d > 0.0 ? d : 1.0;
            dinv[jj] = fast_rcp(dd[jj]);
#pragma unroll
            for (int i = jj + 1; i < NU; ++i) {
              double sacc = 0.5 * (hu[jj][NX + i] + hu[i][NX + jj]);
#pragma unroll
              for (int l = 0; l < jj; ++l) sacc -= Lm[i][l] * Lm[jj][l] * dd[l];
              Lm[i][jj] = sacc * dinv[jj];
            }
          }
          if (pd) break;
          reg = (reg == 0.0) ? 1e-8 : reg * 100.0;
          if (attempt == 11) ok = false;
        }
      }
      // own feedback column: kk = -Huu^{-1} h[NX..NE)   (L y = -h, z = D^{-1} y, L^T kk = z)
      double kk[NU];
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        double sacc = -h[NX + i];
#pragma unroll
        for (int l = 0; l < i; ++l) sacc -= Lm[i][l] * kk[l];
        kk[i] = sacc;
      }
#pragma unroll
      for (int i = 0; i < NU; ++i) kk[i] *= dinv[i];
#pragma unroll
      for (int i = NU - 1; i >= 0; --i) {
        double sacc = kk[i];
#pragma unroll
        for (int l = i + 1; l < NU; ++l) sacc -= Lm[l][i] * kk[l];
        kk[i] = sacc;
      }
      double* fb = fbs + k * T::KD;
      if (g < NX) {
#pragma unroll
        for (int i = 0; i < NU; ++i) fb[i * NX + g] = kk[i];
      } else if (g == NE) {
#pragma unroll
        for (int i = 0; i < NU; ++i) fb[NU * NX + i] = kk[i];
      }
      // new P column / p:  pn[i] = h[i] + sum_c H[i][NX+c] kk[c]
      double pn[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double acc = h[i];
#pragma unroll
        for (int c = 0; c < NU; ++c) acc += hu[c][i] * kk[c];
        pn[i] = acc;
      }
      __syncwarp();  // everybody has read hu from the exchange buffer
      if (g < NX) {
#pragma unroll
        for (int i = 0; i < NX; ++i) xch[g * NX + i] = pn[i];  // column g stored contiguously (P is symmetric)
      } else if (g == NE) {
#pragma unroll
        for (int i = 0; i < NX; ++i) pv[i] = pn[i];
      }
      __syncwarp();
      {
        const double2* v = reinterpret_cast<const double2*>(xch);
        double Pf[NX * NX + (NX * NX & 1)];
#pragma unroll
        for (int i = 0; i < (NX * NX + 1) / 2; ++i) { const double2 t2 = v[i]; Pf[2 * i] = t2.x; if (2 * i + 1 < NX * NX + (NX * NX & 1)) Pf[2 * i + 1] = t2.y; }
#pragma unroll
        for (int i = 0; i < NX; ++i)
#pragma unroll
          for (int jj = 0; jj <= i; ++jj) { const double mm = 0.5 * (Pf[jj * NX + i] + Pf[i * NX + jj]); Pm[i][jj] = mm; Pm[jj][i] = mm; }
      }
      __syncwarp();  // slot s and the exchange buffer are free again
      if (g == 0 && j + NBUF < 2 * N) { fence_proxy_async(); issue(j + NBUF); }
    }

    // =============================================================== initial step
    double dx[NX];
    if (P::FIXED_X0) {
#pragma unroll
      for (int i = 0; i < NX; ++i) dx[i] = 0.0;
    } else {
      // free x_0: lam_0 must vanish; dx_0 = -P_0^{-1} p_0 (p_0 lives in lane NE)
#pragma unroll
      for (int i = 0; i < NX; ++i) { edual = dmax(edual, dabs(lam[i])); pv[i] = gshfl(pv[i], NE); }
      double L[NX][NX];
      double reg = 0.0;
      for (int attempt = 0; attempt < 12; ++attempt) {
        bool pd = true;
#pragma unroll
        for (int jj = 0; jj < NX; ++jj) {
          double d = Pm[jj][jj] + reg;
#pragma unroll
          for (int l = 0; l < jj; ++l) d -= L[jj][l] * L[jj][l];
          pd = pd && (d > 0.0);
          d = sqrt(d > 0.0 ? d : 1.0);
          L[jj][jj] = d;
#pragma unroll
          for (int i = jj + 1; i < NX; ++i) {
            double sacc = Pm[i][jj];
#pragma unroll
            for (int l = 0; l < jj; ++l) sacc -= L[i][l] * L[jj][l];
            L[i][jj] = sacc / d;
          }
        }
        if (pd) break;
        reg = (reg == 0.0) ? 1e-8 : reg * 100.0;
        if (attempt == 11) ok = false;
      }
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double sacc = -pv[i];
#pragma unroll
        for (int l = 0; l < i; ++l) sacc -= L[i][l] * dx[l];
        dx[i] = sacc / L[i][i];
      }
#pragma unroll
      for (int i = NX - 1; i >= 0; --i) {
        double sacc = dx[i];
#pragma unroll
        for (int l = i + 1; l < NX; ++l) sacc -= L[l][i] * dx[l];
        dx[i] = sacc / L[i][i];
      }
    }

    // =============================================================== forward sweep
    for (int j = N; j < 2 * N; ++j) {
      const int k = j - N, s = j % NBUF;
      mbar_wait(&bars[s], (unsigned)((j / NBUF) & 1));
      const double* qa = ring + s * SLOT;
      const double* fb = fbs + k * T::KD;
      double du[NU];
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        double acc = fb[NU * NX + i];
#pragma unroll
        for (int l = 0; l < NX; ++l) acc += fb[i * NX + l] * dx[l];
        du[i] = acc;
      }
      if (active) {
        if (g < NX) a.dx[((long long)b * (N + 1) + k) * NX + g] = pick<NX>(dx, g);
        else if (g < NE) a.du[((long long)b * N + k) * NU + (g - NX)] = pick<NU>(du, g - NX);
      }
      const int r = g < NX ? g : 0;
      double acc = qa[T::ob + r];
#pragma unroll
      for (int l = 0; l < NX; ++l) acc += qa[T::oA + r * NX + l] * dx[l];
#pragma unroll
      for (int l = 0; l < NU; ++l) acc += qa[T::oB + r * NU + l] * du[l];
#pragma unroll
      for (int i = 0; i < NX; ++i) dx[i] = gshfl(acc, i);
      __syncwarp();
      if (g == 0 && j + NBUF < 2 * N) { fence_proxy_async(); issue(j + NBUF); }
    }
    if (active && g < NX) a.dx[((long long)b * (N + 1) + N) * NX + g] = pick<NX>(dx, g);
    // ---- per-instance results
    double em = edual;
#pragma unroll
    for (int off = LPI / 2; off > 0; off >>= 1) em = dmax(em, __shfl_xor_sync(0xffffffffu, em, off, LPI));
    if (active && g == 0) {
      st[ST_EDUAL] = em;
      if (!ok) { a.status[b] = -1; st[ST_ACTIVE] = 0.0; }
    }
  }
};

template <class P>
__global__ void __launch_bounds__(RiccatiTma<P>::WARPS * 32, 2) riccati_tma_kernel(OcpArgs a) {
  RiccatiTma<P>::run(a);
}

template <class P>
inline cudaError_t launch_riccati_tma(const OcpArgs& a, cudaStream_t s) {
  using R = RiccatiTma<P>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(riccati_tma_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         R::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int per_cta = R::WARPS * R::G;
  const int grid = (a.B + per_cta - 1) / per_cta;
  riccati_tma_kernel<P><<<grid, R::WARPS * 32, R::SMEM_BYTES, s>>>(a);
  return cudaGetLastError();
}

}  // namespace hb
#endif  // HB_DEVICE_BUILD
