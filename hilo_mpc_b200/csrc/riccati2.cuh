// Riccati sweep, sm_100a version 2.
//
// The whole batch is resident in one wave with ~2 warps per scheduler, so the kernel is bound by (1) how fast stage
// blocks reach shared memory and (2) the dependent-latency chain of one stage.  Version 1 (one cp.async.bulk per block, removed) lost on both:
//   (1) one cp.async.bulk per 224/320-byte block is a poor fit for the TMA unit - 64 tiny requests per SM and stage
//       round are processed at ~30 cycles each, i.e. ~1900 cycles per round whatever the arithmetic does (ncu:
//       profiles/ncu_riccati2_r1_summary.csv), and every per-instance mbarrier wait is G-way divergent;
//   (2) two shared-memory exchanges, an LDL^T with two dependent reciprocals and a symmetrisation per stage.
// Here:
//   * stage blocks travel with 16-byte cp.async (LDGSTS) issued by the LPI lanes of the instance itself - 5 warp
//     instructions per backward stage for 4 instances - into the same NBUF-deep ring, tracked by commit groups
//     (wait_group is a counter check, not a 90-cycle mbarrier probe),
//   * ONE exchange per stage: every lane publishes its column of H = [Q S^T; S R] + M^T P M and then forms the feedback
//     gain K = -Huu^{-1} Hux and the lower triangle of P' = Hxx + Hux^T K redundantly (mirrored, hence exactly symmetric:
//     no second exchange for a symmetrisation),
//   * Huu^{-1} in closed form for nu <= 2 (one Newton-refined reciprocal of the determinant),
//   * forward sweep with the closed-loop matrices A + B K, b + B k prepared off the dx-chain: 4 dependent FMAs and one
//     shuffle round per stage.
// Layout and traffic are those of version 1: LPI lanes per instance, lane g < NE owns column g, lane NE the gradient
// column; gains stay in shared memory.
#pragma once
#include "ocp_kernels.cuh"

#if HB_DEVICE_BUILD
namespace hb {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
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

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPENDING) : "memory"); }

template <class P>
struct RiccatiTma2 {
  using T = OcpT<P>;
  static constexpr int NX = T::NX, NU = T::NU, NE = T::NE, N = T::N;
  static constexpr int LPI = (NE + 1 <= 4) ? 4 : (NE + 1 <= 8) ? 8 : (NE + 1 <= 16) ? 16 : 32;  // lanes per instance
  static constexpr int G = 32 / LPI;          // instances per warp
  static constexpr int CW = 4;                // warps per CTA
  static constexpr int THREADS = CW * 32;
  static constexpr int NBUF = 4;              // ring slots per instance (one stage each)
  static constexpr int SLOT = T::NAp + T::NHp;  // doubles per slot
  static constexpr int CS = (NE + 1) & ~1;    // column stride of the exchanged H (even: 128-bit loads)
  static constexpr int XCH = NE * CS;
  static constexpr int FBD = (N * T::KD + 1) & ~1;  // feedback-gain store, padded to keep 16-byte alignment behind it
  static constexpr int RAW = NBUF * SLOT + FBD + XCH;
  static constexpr int STRIDE = ((RAW + 15) / 16) * 16 + 4;  // doubles; instances of a warp start 8 banks apart
  static constexpr int SMEM_BYTES = CW * G * STRIDE * 8;
  static constexpr int CHA = T::NAp / 2, CHH = T::NHp / 2;   // 16-byte chunks of the dynamics / Hessian part of a stage
  static constexpr bool SUPPORTED = (NE + 1 <= 32) && (NX <= 6) && (NU >= 1) && (NU <= 6) && (CW * G <= 32) &&
                                    (SMEM_BYTES <= 110 * 1024) && (N >= NBUF) && !T::FRAG;

  static __device__ __forceinline__ double gshfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src, LPI); }
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
    const int first = blockIdx.x * CW * G;   // first instance of this CTA
    // =============================================================== compute warps
    const int grp = lane / LPI, g = lane % LPI;
    const int b = first + warp * G + grp;
    const bool valid = b < a.B;
    const int bb = valid ? b : (a.B - 1);
    double* st = a.st + (long long)bb * ST_COUNT;
    const bool active = valid && st[ST_ACTIVE] != 0.0;
    if (!__any_sync(0xffffffffu, active)) return;

    double* ring = smem + (size_t)(warp * G + grp) * STRIDE;
    double* fbs = ring + NBUF * SLOT;
    double* xch = fbs + FBD;
    const double* qa_g = a.qa + (long long)bb * N * T::NAp;
    const double* qh_g = a.qh + (long long)bb * N * T::NHp;
    // sequence j = 0..2N-1: backward stage N-1-j (full block) for j < N, forward stage j-N (dynamics part) after;
    // lane g copies the 16-byte chunks g, g + LPI, ... of the stage (slot layout = [qa | qh], so chunk c lands at 2c)
    auto issue = [&](int j) {
      if (j < 2 * N) {
        double* dst = ring + (j % NBUF) * SLOT;
        if (j < N) {
          const double* sa = qa_g + (long long)(N - 1 - j) * T::NAp;
          const double* sh = qh_g + (long long)(N - 1 - j) * T::NHp;
#pragma unroll
          for (int c0 = 0; c0 < CHA + CHH; c0 += LPI) {
            const int c = c0 + g;
            if (c < CHA) cp_async16(dst + 2 * c, sa + 2 * c);
            else if (c < CHA + CHH) cp_async16(dst + 2 * c, sh + 2 * (c - CHA));
          }
        } else {
          const double* sa = qa_g + (long long)(j - N) * T::NAp;
#pragma unroll
          for (int c0 = 0; c0 < CHA; c0 += LPI) {
            const int c = c0 + g;
            if (c < CHA) cp_async16(dst + 2 * c, sa + 2 * c);
          }
        }
      }
      cp_async_commit();   // (possibly empty) group: the group count stays uniform
    };
    const double* tn = a.qpN + (long long)bb * T::SDN;
    double Pm[NX][NX], pv[NX], lam[NX];
    double edual = 0.0;
    bool ok = true;
    // Inertia correction (IPOPT's algorithm IC on the stage structure, see InertiaSchedule): the backward sweep is
    // repeated with delta_w on the diagonals of Q_k, R_k, Q_N until every Huu (and P_0 when x_0 is free) is positive
    // definite.  The instances of a warp repeat together (the ring is shared machinery); one that was fine keeps its dw.
    InertiaSchedule ic(st[ST_DW]);
    double dw = 0.0;
    double dx[NX];
    double L0[P::FIXED_X0 ? 1 : NX][P::FIXED_X0 ? 1 : NX];
    bool given_up = false;
    for (int attempt = 0;; ++attempt) {
    if (attempt > 0) { cp_async_wait<0>(); __syncwarp(); }
#pragma unroll
    for (int j = 0; j < NBUF - 1; ++j) issue(j);

    // ---- terminal data: P = Q_N (replicated in every lane), p = q_N, lam = qbar_N
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
      for (int j = 0; j < NX; ++j) Pm[i][j] = tn[i * NX + j] + ((i == j) ? dw : 0.0);
      pv[i] = tn[NX * NX + i];
      lam[i] = tn[NX * NX + NX + i];
    }
    if (active && g < NX) a.lam[((long long)b * (N + 1) + N) * NX + g] = pick<NX>(lam, g);
    edual = 0.0;
    ok = true;

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
      const double* slot = ring + s * SLOT;
      issue(j + NBUF - 1);          // into the slot read in the previous iteration (all lanes are past that __syncwarp)
      cp_async_wait<NBUF - 1>();    // stage j has landed (this lane's chunks) ...
      __syncwarp();                 // ... and everybody else's
      // whole dynamics block [A B] (broadcast 128-bit loads), the lane's own column, its cost column
      double Mf[NX * NE + (NX * NE & 1)];
      {
        const double2* v = reinterpret_cast<const double2*>(slot);
#pragma unroll
        for (int i = 0; i < (NX * NE + 1) / 2; ++i) { const double2 t2 = v[i]; Mf[2 * i] = t2.x; if (2 * i + 1 < NX * NE + (NX * NE & 1)) Mf[2 * i + 1] = t2.y; }
      }
      double m[NX], h[NE];
#pragma unroll
      for (int l = 0; l < NX; ++l) m[l] = slot[moff[l]];
#pragma unroll
      for (int i = 0; i < NE; ++i) h[i] = slot[coff[i]] + ((i == g) ? dw : 0.0);
      const double gbar = (g < NE) ? slot[goff] : 0.0;
      __syncwarp();                       // all lanes have the slot in registers (and are done with last stage's H)

      // adjoint / reduced gradient with lam_{k+1}
      {
        double acc = gbar;
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
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        double acc = h[i];
#pragma unroll
        for (int l = 0; l < NX; ++l) {
          const double mli = (i < NX) ? Mf[T::oA + l * NX + i] : Mf[T::oB + l * NU + (i - NX)];
          acc += mli * w[l];
        }
        h[i] = acc;
      }
      // the one exchange: column g of H, contiguous
      if (g < NE) {
#pragma unroll
        for (int i = 0; i < NE; ++i) xch[g * CS + i] = h[i];
      }
      __syncwarp();
      double Hc[NX][CS];    // columns 0..NX-1:  Hc[j][i] = H[i][j]  (rows < NX: Hxx, rows >= NX: Hux)
      double Huu[NU][NU];   // Huu[c][d] = H[NX + c][NX + d]
#pragma unroll
      for (int jc = 0; jc < NX; ++jc) {
        const double2* v = reinterpret_cast<const double2*>(xch + jc * CS);
#pragma unroll
        for (int i = 0; i < CS / 2; ++i) { const double2 t2 = v[i]; Hc[jc][2 * i] = t2.x; Hc[jc][2 * i + 1] = t2.y; }
      }
#pragma unroll
      for (int d = 0; d < NU; ++d)
#pragma unroll
        for (int c = 0; c < NU; ++c) Huu[c][d] = xch[(NX + d) * CS + NX + c];
      // Huu^{-1} (symmetrised): closed form for nu <= 2, Cholesky + triangular inverse otherwise (every lane, redundantly)
      double inv[NU][NU];
      if constexpr (NU == 1) {
        const double a0 = Huu[0][0];
        const bool pd = a0 > 0.0;
        ok = ok && pd;
        inv[0][0] = fast_rcp(pd ? a0 : 1.0);
      } else if constexpr (NU == 2) {
        const double a0 = Huu[0][0], d0 = Huu[1][1], b0 = 0.5 * (Huu[0][1] + Huu[1][0]);
        const double det = a0 * d0 - b0 * b0;
        const bool pd = (a0 > 0.0) && (det > 0.0);
        ok = ok && pd;
        const double idet = fast_rcp(pd ? det : 1.0);
        inv[0][0] = d0 * idet; inv[1][1] = a0 * idet; inv[0][1] = -b0 * idet; inv[1][0] = inv[0][1];
      } else {
        double Lc[NU][NU], Li[NU][NU];   // Huu = Lc Lc^T, Li = Lc^{-1} (lower), inv = Li^T Li
#pragma unroll
        for (int jj = 0; jj < NU; ++jj) {
          double d = Huu[jj][jj];
#pragma unroll
          for (int l = 0; l < jj; ++l) d -= Lc[jj][l] * Lc[jj][l];
          const bool pd = d > 0.0;
          ok = ok && pd;
          const double rs = rsqrt(pd ? d : 1.0);   // full-precision device rsqrt
          Lc[jj][jj] = rs;                         // the reciprocal of the pivot is what is needed below
#pragma unroll
          for (int i = jj + 1; i < NU; ++i) {
            double sacc = 0.5 * (Huu[i][jj] + Huu[jj][i]);
#pragma unroll
            for (int l = 0; l < jj; ++l) sacc -= Lc[i][l] * Lc[jj][l];
            Lc[i][jj] = sacc * rs;
          }
        }
#pragma unroll
        for (int jj = 0; jj < NU; ++jj) {          // column jj of Li by forward substitution
#pragma unroll
          for (int i = 0; i < NU; ++i) {
            if (i < jj) { Li[i][jj] = 0.0; continue; }
            double sacc = (i == jj) ? 1.0 : 0.0;
#pragma unroll
            for (int l = jj; l < i; ++l) sacc -= Lc[i][l] * Li[l][jj];
            Li[i][jj] = sacc * Lc[i][i];
          }
        }
#pragma unroll
        for (int c = 0; c < NU; ++c)
#pragma unroll
          for (int d = 0; d <= c; ++d) {
            double sacc = 0.0;
#pragma unroll
            for (int l = c; l < NU; ++l) sacc += Li[l][c] * Li[l][d];
            inv[c][d] = sacc; inv[d][c] = sacc;
          }
      }
      // feedback gain K = -Huu^{-1} Hux (all of it, every lane) and the lower triangle of P' = Hxx + Hux^T K
      double K[NU][NX];
#pragma unroll
      for (int c = 0; c < NU; ++c)
#pragma unroll
        for (int jc = 0; jc < NX; ++jc) {
          double acc = 0.0;
#pragma unroll
          for (int d = 0; d < NU; ++d) acc -= inv[c][d] * Hc[jc][NX + d];
          K[c][jc] = acc;
        }
#pragma unroll
      for (int i = 0; i < NX; ++i)
#pragma unroll
        for (int jc = 0; jc <= i; ++jc) {
          double acc = Hc[jc][i];
#pragma unroll
          for (int c = 0; c < NU; ++c) acc += Hc[i][NX + c] * K[c][jc];
          Pm[i][jc] = acc; Pm[jc][i] = acc;
        }
      // affine part (lane NE): k = -Huu^{-1} h_u,  p' = h_x + Hux^T k
      double kv[NU];
#pragma unroll
      for (int c = 0; c < NU; ++c) {
        double acc = 0.0;
#pragma unroll
        for (int d = 0; d < NU; ++d) acc -= inv[c][d] * h[NX + d];
        kv[c] = acc;
      }
      if (g == NE) {
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double acc = h[i];
#pragma unroll
          for (int c = 0; c < NU; ++c) acc += Hc[i][NX + c] * kv[c];
          pv[i] = acc;
        }
      }
      double* fb = fbs + k * T::KD;
      if (g < NX) {
#pragma unroll
        for (int c = 0; c < NU; ++c) {
          double kc[NX];
#pragma unroll
          for (int jc = 0; jc < NX; ++jc) kc[jc] = K[c][jc];
          fb[c * NX + g] = pick<NX>(kc, g);
        }
      } else if (g == NE) {
#pragma unroll
        for (int c = 0; c < NU; ++c) fb[NU * NX + c] = kv[c];
      }
    }

    // =============================================================== factor P_0 (free x_0)
    if constexpr (!P::FIXED_X0) {   // free x_0: P_0 is the last pivot block of the reduced Hessian
#pragma unroll
      for (int jj = 0; jj < NX; ++jj) {
        double d = Pm[jj][jj];
#pragma unroll
        for (int l = 0; l < jj; ++l) d -= L0[jj][l] * L0[jj][l];
        const bool pd = d > 0.0;
        ok = ok && pd;
        d = sqrt(pd ? d : 1.0);
        L0[jj][jj] = d;
        const double rd = 1.0 / d;
#pragma unroll
        for (int i = jj + 1; i < NX; ++i) {
          double sacc = Pm[i][jj];
#pragma unroll
          for (int l = 0; l < jj; ++l) sacc -= L0[i][l] * L0[jj][l];
          L0[i][jj] = sacc * rd;
        }
      }
    }
    // ---- inertia check of the whole warp
    {
      const bool fail = active && !ok;
      if (!__any_sync(0xffffffffu, fail)) break;
      if (fail) {
        if (!ic.next()) given_up = true;
        dw = ic.dw;
      }
      if (__all_sync(0xffffffffu, !fail || given_up)) break;
    }
    }   // attempt loop
    if (given_up) ok = false;
    if (P::FIXED_X0) {
#pragma unroll
      for (int i = 0; i < NX; ++i) dx[i] = 0.0;
    } else {
      // lam_0 must vanish; dx_0 = -P_0^{-1} p_0 (p_0 lives in lane NE)
#pragma unroll
      for (int i = 0; i < NX; ++i) { edual = dmax(edual, dabs(lam[i])); pv[i] = gshfl(pv[i], NE); }
      if constexpr (!P::FIXED_X0) {
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double sacc = -pv[i];
#pragma unroll
          for (int l = 0; l < i; ++l) sacc -= L0[i][l] * dx[l];
          dx[i] = sacc / L0[i][i];
        }
#pragma unroll
        for (int i = NX - 1; i >= 0; --i) {
          double sacc = dx[i];
#pragma unroll
          for (int l = i + 1; l < NX; ++l) sacc -= L0[l][i] * dx[l];
          dx[i] = sacc / L0[i][i];
        }
      }
    }

    // =============================================================== forward sweep
    // lane r < NX owns row r of the closed loop  dx+ = (A + B K) dx + (b + B k); lanes NX.. the input rows for the output
    __syncwarp();   // gains of every stage are in shared memory
    const int r = g < NX ? g : 0, cu = (g >= NX && g < NE) ? (g - NX) : 0;
    for (int j = N; j < 2 * N; ++j) {
      const int k = j - N, s = j % NBUF;
      const double* qa = ring + s * SLOT;
      const double* fb = fbs + k * T::KD;
      issue(j + NBUF - 1);
      cp_async_wait<NBUF - 1>();
      __syncwarp();
      double Ar[NX], Br[NU], Kf[NU][NX], kf[NU];
#pragma unroll
      for (int l = 0; l < NX; ++l) Ar[l] = qa[T::oA + r * NX + l];
#pragma unroll
      for (int c = 0; c < NU; ++c) Br[c] = qa[T::oB + r * NU + c];
      double bcl = qa[T::ob + r];
#pragma unroll
      for (int c = 0; c < NU; ++c) {
#pragma unroll
        for (int l = 0; l < NX; ++l) Kf[c][l] = fb[c * NX + l];
        kf[c] = fb[NU * NX + c];
      }
      __syncwarp();
      // closed-loop row (independent of dx)
#pragma unroll
      for (int c = 0; c < NU; ++c) {
        bcl += Br[c] * kf[c];
#pragma unroll
        for (int l = 0; l < NX; ++l) Ar[l] += Br[c] * Kf[c][l];
      }
      // outputs of this stage
      if (active) {
        if (g < NX) {
          a.dx[((long long)b * (N + 1) + k) * NX + g] = pick<NX>(dx, g);
        } else if (g < NE) {
          double du = 0.0;
#pragma unroll
          for (int c = 0; c < NU; ++c) {
            double acc = kf[c];
#pragma unroll
            for (int l = 0; l < NX; ++l) acc += Kf[c][l] * dx[l];
            du = (cu == c) ? acc : du;
          }
          a.du[((long long)b * N + k) * NU + cu] = du;
        }
      }
      double acc = bcl;
#pragma unroll
      for (int l = 0; l < NX; ++l) acc += Ar[l] * dx[l];
#pragma unroll
      for (int i = 0; i < NX; ++i) dx[i] = gshfl(acc, i);
    }
    if (active && g < NX) a.dx[((long long)b * (N + 1) + N) * NX + g] = pick<NX>(dx, g);
    // ---- per-instance results
    double em = edual;
#pragma unroll
    for (int off = LPI / 2; off > 0; off >>= 1) em = dmax(em, __shfl_xor_sync(0xffffffffu, em, off, LPI));
    if (active && g == 0) {
      st[ST_EDUAL] = em;
      if (!ok) { a.status[b] = -1; st[ST_ACTIVE] = 0.0; }
      else if (dw > 0.0) st[ST_DW] = dw;
    }
  }
};

#ifndef HB_RIC2_MINB
#define HB_RIC2_MINB 2   // CTAs per SM the register allocation aims at
#endif
template <class P>
__global__ void __launch_bounds__(RiccatiTma2<P>::THREADS, HB_RIC2_MINB) riccati_tma2_kernel(OcpArgs a) { RiccatiTma2<P>::run(a); }

template <class P>
inline cudaError_t launch_riccati_tma2(const OcpArgs& a, cudaStream_t s) {
  using R = RiccatiTma2<P>;
  cudaError_t e = cudaFuncSetAttribute(riccati_tma2_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, R::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  const int per_cta = R::CW * R::G;
  const int grid = (a.B + per_cta - 1) / per_cta;
  riccati_tma2_kernel<P><<<grid, R::THREADS, R::SMEM_BYTES, s>>>(a);
  return cudaGetLastError();
}

}  // namespace hb
#endif  // HB_DEVICE_BUILD
