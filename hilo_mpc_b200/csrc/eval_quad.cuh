// Node evaluation, sm_100a version for the collocation transcription: 4 lanes per (instance, interval).
//
// The one-thread-per-node kernel (OcpEval) keeps the deg*nx x (deg*nx + nx + nu + 1) elimination matrix in local
// memory (3 KB per thread, 253 registers) and is bound by local-memory latency.  Here the matrix is row-distributed over
// the four lanes of a quad and lives in registers (row r belongs to lane r % 4); Gauss-Jordan elimination with
// (implicit) partial pivoting broadcasts the normalised pivot row through shared memory; everything else that several
// lanes need (model derivatives at the collocation points, the solved T block) is exchanged through a small per-quad
// shared-memory scratch pad.  Work split:
//   lanes 0..deg-1 : model + cost derivatives at "their" collocation point (generated device functions)
//   lane  3        : barrier terms of x_k and u_k
//   all lanes      : their rows of the (transposed and plain) collocation systems, rows of A/B/b, a share of the
//                    reduced Hessian entries, a share of the output copy.
// Supported: collocation, NMPC-type problems (x_0 pinned, continuous objective, no additive input, no arrival cost)
// with deg <= 4 and rows_per_lane * columns <= 64 doubles; everything else uses OcpEval.  Terminal nodes are evaluated
// by OcpEvalTail (one thread per node, same launch as the re-evaluation of flagged nodes).
#pragma once
#include "ocp_kernels.cuh"

namespace hb {

// terminal node only (the k == N branch of OcpEval), launched over instances
// What the quad kernel leaves to the one-thread-per-node body, as ONE launch: the terminal node (k = N) and the nodes
// whose statically pivoted elimination met a degenerate pivot (recomputed with partial pivoting).  The body needs 255
// registers, so a thread per node would cost five waves of threads that only read a flag: one thread takes CHUNK
// consecutive nodes of an instance instead (index space (instance, chunk)); flagged nodes are rare and handled in turn.
template <class P>
struct OcpEvalTail {
  static constexpr int CHUNK = 8;
  static constexpr int NCHUNK = (OcpT<P>::N + 1 + CHUNK - 1) / CHUNK;
  static HB_HD void run(const OcpArgs& a, int i, int c) {
    constexpr int N = OcpT<P>::N;
    const int b = a.instance(i);
    if (a.st[(long long)b * ST_COUNT + ST_ACTIVE] == 0.0) return;
    const int k1 = (c + 1) * CHUNK < N + 1 ? (c + 1) * CHUNK : N + 1;
    for (int k = c * CHUNK; k < k1; ++k)
      if (k == N || a.nflag[(long long)b * N + k] != 0) OcpEval<P>::run(a, b, k);
  }
};

#if HB_DEVICE_BUILD

// tuning knobs of the 8-lanes-per-node variant (see NCOPY and the launch bounds)
#ifndef HB_EVQ_PAIRRED
#define HB_EVQ_PAIRRED 1
#endif
#ifndef HB_EVQ_MINB8
#define HB_EVQ_MINB8 2   // 3 (168 registers, 450 bytes of spills) measured 9 % slower at config 5
#endif
#ifndef HB_EVQ_MINB4
#define HB_EVQ_MINB4 3
#endif

template <class P>
struct EvalQuad {
  using T = OcpT<P>;
  static constexpr int NX = T::NX, NU = T::NU, NC = T::NC, NE = T::NE, N = T::N, D = P::DEG;
  static constexpr int W = NC + NE + 1;      // columns of the plain system  [G_X | G_x G_u | c]
  static constexpr int WT = NC + 1;          // columns of the transposed system
  // lanes per node: 4 while the lane's rows of the plain system fit the register budget (the BASELINE CSTR: 3 rows x 19
  // columns), 8 for larger nodes (nx = 5..6 with deg 3: 2 rows x 24..28 columns)
  static constexpr int LPN = (((NC + 3) / 4) * W <= 64) ? 4 : 8;
  static constexpr int NPW = 32 / LPN;       // nodes per warp
  static constexpr int RPL = (NC + LPN - 1) / LPN;   // rows per lane
  static constexpr bool NXL = (NX % LPN == 0);       // then the collocation point of a local row is a compile-time constant
  static constexpr int NV = NX + NU + NC;    // bounded variables of a node: x_k | u_k | X_k
  // cost terms that live on the node itself rather than on the collocation points: the discrete-type objective
  // l(x_k, u_k) (MHE, 'objective_function': 'discrete') and the arrival cost on x_0; evaluated by lane D
  static constexpr bool NODECOST = P::OBJ_DISCRETE || P::HAS_ARR;
  static constexpr bool SUPPORTED = (P::METHOD == METHOD_COLLOCATION) && (D <= 3) && (D < LPN) && (RPL * W <= 64) &&
                                    (NC >= 4);
  // per-node scratch (doubles).  Lifetimes:  F (model derivatives) until the plain system is
  // assembled, then the same words hold T (its solution);  PB (pivot row) during the two eliminations, the collocation
  // multipliers NU in between;  L (weighted cost derivatives + constraint curvature) until the partial sums of the
  // reduced Hessian exist - from then on [F/T | PB | L] is the reduction area RED and only the sums over the
  // collocation points of lu / luu (SUM) are still needed.
  static constexpr int FSZ = NX + NX * NX + NX * NU;                             // f | fx | fu of one point
  static constexpr int LSZ = 1 + NX + NU + NX * NX + NU * NX + NU * NU;          // l | lx | lu | lxx | lux | luu (weighted)
  // partial sums of one lane: lower triangle of the reduced Hessian [NE][NE] row by row, the Newton gradient [NE], the
  // multiplier-form gradient [NE]; ODD stride between the lanes' copies: the nodes of a warp sit on the even 8-byte bank
  // pairs (SCR below), two lanes of a node share a half-warp (lane mapping in run()), so the second lane must land on the
  // odd ones
  static constexpr int NTRI = NE * (NE + 1) / 2;
  static HB_HD constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }   // j <= i
  static constexpr int NRED = (NTRI + 2 * NE) | 1;
  static constexpr int FT = (((D * FSZ > NC * (NE + 1)) ? D * FSZ : NC * (NE + 1)) + 1) & ~1;   // even: PB holds 128-bit words
  static constexpr int PBN = (2 * ((W + 1) & ~1) > NC) ? 2 * ((W + 1) & ~1) : NC;   // 2 pivot-row buffers
  // copies of the partial sums that meet in shared memory: with 8 lanes per node the lanes q and q ^ 4 add theirs by a
  // shuffle first (HB_EVQ_PAIRRED, default on), which takes the reduction area below the size of [F/T | PB | L] and
  // the CTA under a third of the SM's shared memory
  static constexpr int NCOPY = (LPN == 8 && HB_EVQ_PAIRRED) ? 4 : LPN;
  static constexpr int A1 = FT + PBN + D * LSZ, A2 = NCOPY * NRED;
  static constexpr int NCSZ = NODECOST ? (NX + NU + NX * NX + NU * NX + NU * NU) : 0;  // gx | gu | Qn | Sn | Rn of the node cost
  static constexpr int oF = 0, oT = 0, oPB = FT, oNU = oPB, oL = oPB + PBN, oRED = 0,
                       oSUM = (A1 > A2 ? A1 : A2),                       // sum_i lu_i [NU] | sum_i luu_i [NU*NU]
                       oBG = oSUM + NU + NU * NU, oBD = oBG + NV, oBS = oBD + NV,   // barrier gradient | dz | sigma of x|u|X
                       oNC = oBS + NV, oEND = oNC + NCSZ;
  static constexpr int oNCu = oNC + NX, oNCQ = oNCu + NU, oNCS = oNCQ + NX * NX, oNCR = oNCS + NU * NX;
  static constexpr int oGX = oBG + NX + NU, oDGX = oBD + NX + NU, oSGX = oBS + NX + NU;
  // stride = 2 (mod 16) doubles: 16-byte aligned, and the nodes of a warp start 4 banks apart (conflict-free 64- and
  // 128-bit accesses to the same offset)
  static constexpr int SCR = ((oEND + 13) / 16) * 16 + 2;
  static constexpr int QUADS = 128 / LPN;            // nodes per CTA, 128 threads
  static constexpr int SMEM_BYTES = QUADS * SCR * 8;
  static constexpr bool EL_VEC2 = (NE % 2 == 0) && (T::oT % 2 == 0) && (T::TD % 2 == 0);

  // lanes q, q^m of the same node (lane = q * NPW + node-in-warp, see run())
  static __device__ __forceinline__ double qshfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m * NPW); }
  static __device__ __forceinline__ int qshfl_xor(int v, int m) { return __shfl_xor_sync(0xffffffffu, v, m * NPW); }

  // Gauss-Jordan on rows R[RPL][NCOLS] (row lr of lane q is global row lr*LPN + q) with the STATIC pivot order
  // "row c for column c": no search, no row selection; the owner lane and its local row are compile-time constants.
  // The owner publishes its (unnormalised) pivot row, every lane forms its own multipliers with a Newton-refined
  // reciprocal, and the rows are scaled by their inverse pivots once at the end.
  // The collocation matrix is dt*blkdiag(f_x) - C^T (x) I, whose unpivoted pivots are about (-3.2, -2.1, -9.0) per
  // state for the Radau-3 constants; the largest |1 / pivot| met is returned so that the caller can hand nodes with
  // a degenerate pivot (stiff dt*f_x) to the partially pivoted one-thread-per-node kernel.
  // On return row r holds, in columns >= NPIV, the solution of variable r.
  template <int NCOLS, int NPIV>
  static __device__ __forceinline__ double gauss_jordan(double (&R)[RPL][NCOLS], int q, double* pb) {
    constexpr int PS = (NCOLS + 1) & ~1;   // even stride: the pivot rows travel in 128-bit words
    double maxinv = 0.0;                   // largest |1 / pivot| met
    double ipiv[RPL];
#pragma unroll
    for (int lr = 0; lr < RPL; ++lr) ipiv[lr] = 1.0;
#pragma unroll
    for (int c = 0; c < NPIV; ++c) {
      const int po = c % LPN, plr = c / LPN;   // owner lane / local row of the pivot row
      const int j0 = c & ~1;                 // first (even) column of the broadcast; column j0 < c is dead weight
      double* prow = pb + (c & 1) * PS;      // double-buffered broadcast row
      if (q == po) {
#pragma unroll
        for (int j = j0; j < NCOLS; j += 2)
          *reinterpret_cast<double2*>(prow + j) = make_double2(R[plr][j], (j + 1 < NCOLS) ? R[plr][j + 1] : 0.0);
      }
      __syncwarp();
      double pr[PS];
#pragma unroll
      for (int j = j0; j < NCOLS; j += 2) {
        const double2 t2 = *reinterpret_cast<const double2*>(prow + j);
        pr[j] = t2.x; pr[j + 1] = t2.y;
      }
      const double inv = rcp_nr(pr[c]);
      ipiv[plr] = (q == po) ? inv : ipiv[plr];
#pragma unroll
      for (int lr = 0; lr < RPL; ++lr) {
        const bool is_piv = (lr == plr) && (q == po);
        const double f = is_piv ? 0.0 : R[lr][c] * inv;
#pragma unroll
        for (int j = c + 1; j < NCOLS; ++j) R[lr][j] -= f * pr[j];
      }
    }
#pragma unroll
    for (int lr = 0; lr < RPL; ++lr) {
      maxinv = fmax(maxinv, fabs(ipiv[lr]));
#pragma unroll
      for (int j = NPIV; j < NCOLS; ++j) R[lr][j] *= ipiv[lr];
    }
    __syncwarp();
    return maxinv;
  }

  static __device__ void run(const OcpArgs& a) {
    extern __shared__ __align__(16) double smem[];
    // lane = q * NPW + node-in-warp: the lanes with the same role q of the nodes of a warp are adjacent.  The pivot-row
    // store of the elimination (one lane per node, 128 bits) then comes from ONE quarter-warp = one shared-memory
    // wavefront instead of four, and two (not four) lanes of a node meet in a half-warp for the 64-bit accesses.
    const int quad = ((threadIdx.x >> 5) * NPW) | (threadIdx.x & (NPW - 1)), q = (threadIdx.x & 31) / NPW;
    const long long node = (long long)blockIdx.x * QUADS + quad;
    const long long total = (long long)a.n_instances() * N;   // late iterations run over the list of active instances
    const bool valid = node < total;
    const long long nn = valid ? node : (total - 1);
    const int b = a.instance((int)(nn / N)), k = (int)(nn % N);
    const double* st = a.st + (long long)b * ST_COUNT;
    const bool active = valid && st[ST_ACTIVE] != 0.0;
    // a whole warp leaves together; partially active warps compute everything and mask the stores
    if (!__any_sync(0xffffffffu, active)) return;
    double* S = smem + (size_t)quad * SCR;
    const double mu = st[ST_MU];
    const long long nb = (long long)b * (N + 1) + k, sb = (long long)b * N + k;
    const double t = a.time0 + k * a.dt, dt = a.dt;
    const double* sp = a.sp + (long long)b * a.sp_stride + (long long)k * T::NSP;
    const double* uo = a.uold + (long long)b * a.uold_stride;
    const bool with_lag = k >= P::LAG_FROM;
    const bool first = (k == 0) && (T::NUOLD > 0);
    const bool x_free = (k > 0) || !P::FIXED_X0;

    // run-time indexed accesses go to global memory (L1-resident); register arrays are only indexed statically
    const double* xg = a.x + nb * NX;
    const double* xng = a.x + (nb + 1) * NX;
    const double* Xg = a.xc + sb * NC;
    const double* ug = a.u + sb * NU;
    double u[NU], Xq[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) Xq[i] = Xg[(q < D ? q : 0) * NX + i];
#pragma unroll
    for (int i = 0; i < NU; ++i) u[i] = ug[i];

    double phi = 0.0, comp = 0.0, c1 = 0.0, cinf = 0.0, maxinv = 0.0;
    // ---------------------------------------------------------------- 1a. barrier terms, LPN variables at a time
    {
      BarrierSum bs;
#pragma unroll
      for (int it = 0; it < (NV + LPN - 1) / LPN; ++it) {
        const int e = it * LPN + q;
        const bool ok = (it * LPN + LPN - 1 < NV) || (e < NV);
        const double *vp, *lp, *up, *zlp, *zup;
        bool use = ok;
        if (e < NX) {
          vp = xg + e; lp = a.xlb + e; up = a.xub + e; zlp = a.zlx + nb * NX + e; zup = a.zux + nb * NX + e;
          use = use && x_free;
        } else if (e < NX + NU) {
          const int i = e - NX;
          vp = ug + i; lp = a.ulb + i; up = a.uub + i; zlp = a.zlu + sb * NU + i; zup = a.zuu + sb * NU + i;
        } else {
          const int r = ok ? e - NX - NU : 0, j = r % NX;
          vp = Xg + r; lp = a.xlb + j; up = a.xub + j; zlp = a.zlc + sb * NC + r; zup = a.zuc + sb * NC + r;
        }
        double g = 0.0, dg = 0.0, sg = 0.0;
        if (use) {
          const double v = *vp, lb = *lp, ub = *up, zl = *zlp, zu = *zup;
          const BarrierTerm bt = barrier_term(v, lb, ub, zl, zu, mu);
          g = bt.grad; dg = bt.dz; sg = bt.sigma;
          bs.add(v, lb, ub);
          if (has_lb(lb)) comp = dmax(comp, dabs((v - lb) * zl - mu));
          if (has_ub(ub)) comp = dmax(comp, dabs((ub - v) * zu - mu));
        }
        if (ok) { S[oBG + e] = g; S[oBD + e] = dg; S[oBS + e] = sg; }
      }
      phi += bs.value(mu);
    }
    // ---------------------------------------------------------------- 1b. derivatives at the collocation points
    // outputs of the model's call nodes (GP / MLP loops) at this lane's point: filled by the Jacobian pass, read by the
    // Hessian pass further down (thread-local memory; one walk over the table instead of two)
    double cc[P::NCALLCACHE > 0 ? P::NCALLCACHE : 1];
    if (q < D) {
      double cb = 0.0, ctau = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i)
        if (q == i) { cb = P::CB(i + 1); ctau = P::CTAU(i); }
      const double ti = t + dt * ctau;
      double* F = S + oF + q * FSZ;
      if constexpr (P::NCALLCACHE > 0) P::f_jac_c(Xq, u, sp, ti, dt, F, F + NX, F + NX + NX * NX, cc);
      else P::f_jac(Xq, u, sp, ti, dt, F, F + NX, F + NX + NX * NX);
      double* L = S + oL + q * LSZ;
      double l = 0.0, lx[NX], lu[NU], lxx[NX * NX], lux[NU * NX], luu[NU * NU];
      if (with_lag && !P::OBJ_DISCRETE) {
        P::lag_all(Xq, u, sp, ti, dt, &l, lx, lu, lxx, lux, luu);
        if (first) {
          double l2, lu2[NU], luu2[NU * NU];
          P::lagdu_all(u, uo, &l2, lu2, luu2);
          l += l2;
#pragma unroll
          for (int j = 0; j < NU; ++j) lu[j] += lu2[j];
#pragma unroll
          for (int j = 0; j < NU * NU; ++j) luu[j] += luu2[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < NX; ++j) lx[j] = 0.0;
#pragma unroll
        for (int j = 0; j < NU; ++j) lu[j] = 0.0;
#pragma unroll
        for (int j = 0; j < NX * NX; ++j) lxx[j] = 0.0;
#pragma unroll
        for (int j = 0; j < NU * NX; ++j) lux[j] = 0.0;
#pragma unroll
        for (int j = 0; j < NU * NU; ++j) luu[j] = 0.0;
      }
      const double w = dt * cb;
      phi += w * l;
      L[0] = w * l;
#pragma unroll
      for (int j = 0; j < NX; ++j) L[1 + j] = w * lx[j];
#pragma unroll
      for (int j = 0; j < NU; ++j) L[1 + NX + j] = w * lu[j];
#pragma unroll
      for (int j = 0; j < NX * NX; ++j) L[1 + NX + NU + j] = w * lxx[j];
#pragma unroll
      for (int j = 0; j < NU * NX; ++j) L[1 + NX + NU + NX * NX + j] = w * lux[j];
#pragma unroll
      for (int j = 0; j < NU * NU; ++j) L[1 + NX + NU + NX * NX + NU * NX + j] = w * luu[j];
    }
    if constexpr (NODECOST) {
      if (q == D) {   // cost terms on the node itself: l(x_k, u_k) of a discrete-type objective, arrival cost on x_0
        double xk[NX], gxn[NX], gun[NU], Qn[NX * NX], Sn[NU * NX], Rn[NU * NU];
#pragma unroll
        for (int j = 0; j < NX; ++j) { xk[j] = xg[j]; gxn[j] = 0.0; }
#pragma unroll
        for (int j = 0; j < NU; ++j) gun[j] = 0.0;
#pragma unroll
        for (int j = 0; j < NX * NX; ++j) Qn[j] = 0.0;
#pragma unroll
        for (int j = 0; j < NU * NX; ++j) Sn[j] = 0.0;
#pragma unroll
        for (int j = 0; j < NU * NU; ++j) Rn[j] = 0.0;
        if (P::OBJ_DISCRETE && with_lag) {
          double l;
          P::lag_all(xk, u, sp, t, dt, &l, gxn, gun, Qn, Sn, Rn);
          if (first) {
            double l2, lu2[NU], luu2[NU * NU];
            P::lagdu_all(u, uo, &l2, lu2, luu2);
            l += l2;
#pragma unroll
            for (int j = 0; j < NU; ++j) gun[j] += lu2[j];
#pragma unroll
            for (int j = 0; j < NU * NU; ++j) Rn[j] += luu2[j];
          }
          phi += l;
        }
        if (P::HAS_ARR && k == 0) {
          double val, g[NX], H[NX * NX];
          P::arr_all(xk, a.ap + (long long)b * a.ap_stride, &val, g, H);
          phi += val;
#pragma unroll
          for (int j = 0; j < NX; ++j) gxn[j] += g[j];
#pragma unroll
          for (int j = 0; j < NX * NX; ++j) Qn[j] += H[j];
        }
#pragma unroll
        for (int j = 0; j < NX; ++j) S[oNC + j] = gxn[j];
#pragma unroll
        for (int j = 0; j < NU; ++j) S[oNCu + j] = gun[j];
#pragma unroll
        for (int j = 0; j < NX * NX; ++j) S[oNCQ + j] = Qn[j];
#pragma unroll
        for (int j = 0; j < NU * NX; ++j) S[oNCS + j] = Sn[j];
#pragma unroll
        for (int j = 0; j < NU * NU; ++j) S[oNCR + j] = Rn[j];
      }
    }
    __syncwarp();
    // gX (cost + barrier) for every collocation state: gX[r] = w*lx_i[j] + barrier grad
#pragma unroll
    for (int lr = 0; lr < RPL; ++lr) {
      const int r = lr * LPN + q;
      if (r < NC) S[oGX + r] += S[oL + (r / NX) * LSZ + 1 + (r % NX)];
    }
    __syncwarp();

    // ---------------------------------------------------------------- 2. exact Hessian: collocation multipliers
    if (P::EXACT_HESS) {
      const double* lam1 = a.lam + (nb + 1) * NX;
      double Rt[RPL][WT];
#pragma unroll
      for (int lr = 0; lr < RPL; ++lr) {
        const int r = lr * LPN + q;       // row (i', cc) of G_X^T  = column of G_X
        const int ip = NXL ? (lr * LPN) / NX : r / NX, cc = NXL ? ((lr * LPN) % NX + q) : r % NX;  // ip = i' - 1
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int i = c / NX, rr = c % NX;   // column (i, rr): row of G_X
          double v = 0.0;
          if (r < NC) {
            if (i == ip) v = dt * S[oF + ip * FSZ + NX + rr * NX + cc];
            if (rr == cc) v -= P::CC((ip + 1) * (D + 1) + (i + 1));
          }
          Rt[lr][c] = v;
        }
        Rt[lr][NC] = (r < NC) ? -(S[oGX + r] + S[oDGX + r] + P::CD(ip + 1) * lam1[cc]) : 0.0;
      }
      maxinv = fmax(maxinv, gauss_jordan<WT, NC>(Rt, q, S + oPB));
#pragma unroll
      for (int lr = 0; lr < RPL; ++lr)
        if (lr * LPN + q < NC) S[oNU + lr * LPN + q] = Rt[lr][NC];
      __syncwarp();
      if (q < D) {
        double ctau = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i)
          if (q == i) ctau = P::CTAU(i);
        double nu[NX], hxx[NX * NX], hux[NU * NX], huu[NU * NU];
#pragma unroll
        for (int rr = 0; rr < NX; ++rr) nu[rr] = dt * S[oNU + q * NX + rr];
        if constexpr (P::NCALLCACHE > 0) P::f_hess_c(Xq, u, sp, t + dt * ctau, dt, nu, hxx, hux, huu, cc);
        else P::f_hess(Xq, u, sp, t + dt * ctau, dt, nu, hxx, hux, huu);
        double* L = S + oL + q * LSZ;
#pragma unroll
        for (int j = 0; j < NX * NX; ++j) L[1 + NX + NU + j] += hxx[j];
#pragma unroll
        for (int j = 0; j < NU * NX; ++j) L[1 + NX + NU + NX * NX + j] += hux[j];
#pragma unroll
        for (int j = 0; j < NU * NU; ++j) L[1 + NX + NU + NX * NX + NU * NX + j] += huu[j];
      }
      __syncwarp();
    }

    // sums over the collocation points of the input gradient / input Hessian of the cost (the L area is recycled later)
#pragma unroll
    for (int it = 0; it < (NU + NU * NU + LPN - 1) / LPN; ++it) {
      const int e = it * LPN + q;
      if (e < NU + NU * NU) {
        const int off = (e < NU) ? (1 + NX + e) : (1 + NX + NU + NX * NX + NU * NX + (e - NU));
        double sacc = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) sacc += S[oL + i * LSZ + off];
        S[oSUM + e] = sacc;
      }
    }
    // ---------------------------------------------------------------- 3. plain system  G_X dX + G_x dx + G_u du + c = 0
    double Tr[RPL][NE + 1];   // own T rows (negated solution)
    {
      double R[RPL][W];
#pragma unroll
      for (int lr = 0; lr < RPL; ++lr) {
        const int r = lr * LPN + q;  // row (i, rr)
        const int ii = NXL ? (lr * LPN) / NX : r / NX, rr = NXL ? ((lr * LPN) % NX + q) : r % NX;  // ii = i - 1
        const bool rv = r < NC;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int jp = c / NX, cc = c % NX;  // column (j, cc), jp = j - 1
          double v = 0.0;
          if (rv) {
            if (jp == ii) v = dt * S[oF + ii * FSZ + NX + rr * NX + cc];
            if (cc == rr) v -= P::CC((jp + 1) * (D + 1) + (ii + 1));
          }
          R[lr][c] = v;
        }
#pragma unroll
        for (int c = 0; c < NX; ++c) R[lr][NC + c] = (rv && c == rr) ? -P::CC(0 * (D + 1) + (ii + 1)) : 0.0;
#pragma unroll
        for (int c = 0; c < NU; ++c) R[lr][NC + NX + c] = rv ? dt * S[oF + ii * FSZ + NX + NX * NX + rr * NU + c] : 0.0;
        double res = 0.0;
        if (rv) {
          double xp = P::CC(0 * (D + 1) + (ii + 1)) * xg[rr];
#pragma unroll
          for (int j = 1; j <= D; ++j) xp += P::CC(j * (D + 1) + (ii + 1)) * Xg[(j - 1) * NX + rr];
          res = dt * S[oF + ii * FSZ + rr] - xp;
          c1 += dabs(res); cinf = dmax(cinf, dabs(res));
        }
        R[lr][NC + NE] = res;
      }
      maxinv = fmax(maxinv, gauss_jordan<W, NC>(R, q, S + oPB));
      // own T rows stay in registers and are published to the scratch pad for the other lanes
#pragma unroll
      for (int lr = 0; lr < RPL; ++lr) {
        const int r = lr * LPN + q;
#pragma unroll
        for (int j = 0; j <= NE; ++j) {
          Tr[lr][j] = (r < NC) ? -R[lr][NC + j] : 0.0;
          if (r < NC) S[oT + r * (NE + 1) + j] = Tr[lr][j];
        }
      }
    }
    __syncwarp();

    // ---------------------------------------------------------------- 4. reduced stage data
    double* qa = a.qa + ((long long)b * T::N + k) * T::NAp;
    double* qh = a.qh + ((long long)b * T::N + k) * T::NHp;
    double* el = a.el + sb * T::TD;
    const double* Tm = S + oT;
    // dynamics rows (lane q -> state rows q, q+LPN, ...)
#pragma unroll
    for (int it = 0; it < (NX + LPN - 1) / LPN; ++it) {
      const int rr = it * LPN + q;
      if (rr < NX) {
        double xf = P::CD(0) * xg[rr], bsum = 0.0;
#pragma unroll
        for (int i = 1; i <= D; ++i) { xf += P::CD(i) * Xg[(i - 1) * NX + rr]; bsum += P::CD(i) * Tm[((i - 1) * NX + rr) * (NE + 1) + NE]; }
        if (P::HAS_GADD) {
#pragma unroll
          for (int c = 0; c < NU; ++c) xf += P::GADD(rr * NU + c) * u[c];
        }
        const double res = xf - xng[rr];
        c1 += dabs(res); cinf = dmax(cinf, dabs(res));
        if (active) {
#pragma unroll
          for (int c = 0; c < NX; ++c) {
            double s = (rr == c) ? P::CD(0) : 0.0;
#pragma unroll
            for (int i = 1; i <= D; ++i) s += P::CD(i) * Tm[((i - 1) * NX + rr) * (NE + 1) + c];
            qa[T::iA(rr, c)] = s;
          }
#pragma unroll
          for (int c = 0; c < NU; ++c) {
            double s = P::HAS_GADD ? P::GADD(rr * NU + c) : 0.0;
#pragma unroll
            for (int i = 1; i <= D; ++i) s += P::CD(i) * Tm[((i - 1) * NX + rr) * (NE + 1) + NX + c];
            qa[T::iB(rr, c)] = s;
          }
          qa[T::ib(rr)] = bsum + res;
        }
      }
    }
    if (active) T::fill_pads(qa, qh, q, LPN);   // operand-order layout: zero column of M, pads of small problems
    // elimination block for Expand, own rows straight from the registers: T | t0 | gX
    if (active) {
#pragma unroll
      for (int lr = 0; lr < RPL; ++lr) {
        const int r = lr * LPN + q;
        if (r < NC) {
          if constexpr (EL_VEC2) {
            double2* dst = reinterpret_cast<double2*>(el + T::oT + r * NE);
#pragma unroll
            for (int j = 0; j < NE / 2; ++j) dst[j] = make_double2(Tr[lr][2 * j], Tr[lr][2 * j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < NE; ++j) el[T::oT + r * NE + j] = Tr[lr][j];
          }
          el[T::ot0 + r] = Tr[lr][NE];
          el[T::ogX + r] = S[oGX + r];
        }
      }
    }
    // Partial sums over the OWN rows r = (i, l), all operands in registers; only the LOWER triangle of the (symmetric)
    // reduced Hessian is formed:
    //   Pt[ii][jj] += T[r][ii] * HT[r][jj]        HT[r][:] = (H_i + diag sigma)[l][:] T_i[:][:]         (jj <= ii)
    //   Pt[NX+a][jj] += Hu_i[a][l] * T[r][jj]     (jj <= NX+a)
    //   Pt[NX+b][NX+a] += Hu_i[a][l] * T[r][NX+b] (a <= b: the transposed twin of the line above inside the R block)
    //   gq[ii]      += T[r][ii] * HT[r][NE]  (+ Hu_i[a][l] T[r][NE] for input rows)   Newton gradient, HT[r][NE] incl. gX[r]
    //   gb[ii]      += T[r][ii] * (gX + dz)[r]                                         multiplier-form gradient
    double Pt[NTRI], gq[NE], gbp[NE];
#pragma unroll
    for (int ii = 0; ii < NE; ++ii) { gbp[ii] = 0.0; gq[ii] = 0.0; }
#pragma unroll
    for (int e = 0; e < NTRI; ++e) Pt[e] = 0.0;
#pragma unroll
    for (int lr = 0; lr < RPL; ++lr) {
      const int r = lr * LPN + q;
      if (r < NC) {
        const int ii0 = NXL ? (lr * LPN) / NX : r / NX, l = NXL ? ((lr * LPN) % NX + q) : r % NX;
        const double* Hx = S + oL + ii0 * LSZ + 1 + NX + NU + l * NX;
        const double* Hu = S + oL + ii0 * LSZ + 1 + NX + NU + NX * NX;
        const double gXr = S[oGX + r];
        double HTr[NE + 1];
#pragma unroll
        for (int c = 0; c <= NE; ++c) HTr[c] = (c == NE) ? gXr : 0.0;
#pragma unroll
        for (int l2 = 0; l2 < NX; ++l2) {
          double hv = Hx[l2];
          if (l2 == l) hv += S[oSGX + r];
          const double* T2 = Tm + (ii0 * NX + l2) * (NE + 1);
#pragma unroll
          for (int c = 0; c <= NE; ++c) HTr[c] += hv * T2[c];
        }
        const double gbx = gXr + S[oDGX + r];
#pragma unroll
        for (int ii = 0; ii < NE; ++ii) {
          gbp[ii] += Tr[lr][ii] * gbx;
          gq[ii] += Tr[lr][ii] * HTr[NE];
#pragma unroll
          for (int jj = 0; jj <= ii; ++jj) Pt[tri(ii, jj)] += Tr[lr][ii] * HTr[jj];
        }
#pragma unroll
        for (int aa = 0; aa < NU; ++aa) {
          const double hu = Hu[aa * NX + l];
          gq[NX + aa] += hu * Tr[lr][NE];
#pragma unroll
          for (int jj = 0; jj <= NX + aa; ++jj) Pt[tri(NX + aa, jj)] += hu * Tr[lr][jj];
#pragma unroll
          for (int bb = aa; bb < NU; ++bb) Pt[tri(NX + bb, NX + aa)] += hu * Tr[lr][NX + bb];
        }
      }
    }
    __syncwarp();   // every lane is done with F / T / PB / L: the area becomes the reduction buffer
    if constexpr (NCOPY < LPN) {
#pragma unroll
      for (int e = 0; e < NTRI; ++e) Pt[e] += qshfl_xor(Pt[e], NCOPY);
#pragma unroll
      for (int ii = 0; ii < NE; ++ii) { gq[ii] += qshfl_xor(gq[ii], NCOPY); gbp[ii] += qshfl_xor(gbp[ii], NCOPY); }
    }
    if (q < NCOPY) {
      double* red = S + oRED + q * NRED;
#pragma unroll
      for (int e = 0; e < NTRI; ++e) red[e] = Pt[e];
#pragma unroll
      for (int ii = 0; ii < NE; ++ii) { red[NTRI + ii] = gq[ii]; red[NTRI + NE + ii] = gbp[ii]; }
    }
    __syncwarp();
    // every lane finalises whole rows (lane q -> rows q, q+LPN, ...): the column index is static
    auto rsum = [&](int e) {   // sum of entry e over the lanes' copies
      double v = S[oRED + e];
#pragma unroll
      for (int c = 1; c < NCOPY; ++c) v += S[oRED + c * NRED + e];
      return v;
    };
#pragma unroll
    for (int it = 0; it < (NE + LPN - 1) / LPN; ++it) {
      const int ii = it * LPN + q;
      if (ii < NE) {
        const double sg = rsum(NTRI + NE + ii);
        const double sq = rsum(NTRI + ii);
        const bool xrow = NXL ? (it * LPN < NX) : (ii < NX);
        if (xrow) {
          double g = S[oBG + ii];
          if constexpr (NODECOST) g += S[oNC + ii];
          if (active) {
#pragma unroll
            for (int jj = 0; jj < NX; ++jj) {
              double v = rsum(jj <= ii ? tri(ii, jj) : tri(jj, ii));
              if (jj == ii) v += S[oBS + ii];
              if constexpr (NODECOST) v += S[oNCQ + ii * NX + jj];
              qh[T::iQ(ii, jj)] = v;
            }
            qh[T::iq(ii)] = g + sq; el[T::ogx + ii] = g;
            qh[T::iqb(ii)] = g + S[oBD + ii] + sg;
          }
        } else {
          const int aa = ii - NX;
          double g = S[oBG + ii] + S[oSUM + aa];
          if constexpr (NODECOST) g += S[oNCu + aa];
          if (active) {
#pragma unroll
            for (int jj = 0; jj < NX; ++jj) {
              double v = rsum(tri(ii, jj));
              if constexpr (NODECOST) v += S[oNCS + aa * NX + jj];
              qh[T::iS(aa, jj)] = v;
              if constexpr (T::FRAG) qh[T::iSt(aa, jj)] = v;
            }
#pragma unroll
            for (int bb = 0; bb < NU; ++bb) {
              double v = rsum(bb <= aa ? tri(ii, NX + bb) : tri(NX + bb, ii));
              v += S[oSUM + NU + aa * NU + bb];
              if (bb == aa) v += S[oBS + ii];
              if constexpr (NODECOST) v += S[oNCR + aa * NU + bb];
              qh[T::iR(aa, bb)] = v;
            }
            qh[T::ir(aa)] = g + sq; el[T::ogu + aa] = g;
            qh[T::irb(aa)] = g + S[oBD + ii] + sg;
          }
        }
      }
    }
    // ---------------------------------------------------------------- 5. per-node scalars
#pragma unroll
    for (int m = 1; m < LPN; m <<= 1) {
      phi += qshfl_xor(phi, m); c1 += qshfl_xor(c1, m);
      comp = dmax(comp, qshfl_xor(comp, m)); cinf = dmax(cinf, qshfl_xor(cinf, m));
      maxinv = fmax(maxinv, qshfl_xor(maxinv, m));
    }
    if (active && q == 0) {
      a.s_cinf[nb] = cinf; a.s_c1[nb] = c1; a.s_phi[nb] = phi; a.s_comp[nb] = comp;
      a.nflag[sb] = (maxinv > 20.0 || !(maxinv == maxinv)) ? 1 : 0;  // degenerate static pivot (|pivot| < 0.05): redo with partial pivoting
    }
  }
};

template <class P>
__global__ void __launch_bounds__(128, EvalQuad<P>::LPN == 4 ? HB_EVQ_MINB4 : HB_EVQ_MINB8) eval_quad_kernel(OcpArgs a) { EvalQuad<P>::run(a); }

template <class P>
inline cudaError_t launch_eval_quad(const OcpArgs& a, cudaStream_t s) {
  using E = EvalQuad<P>;
  cudaError_t e = cudaFuncSetAttribute(eval_quad_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, E::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  const long long nodes = (long long)a.n_instances() * OcpT<P>::N;
  const int grid = (int)((nodes + E::QUADS - 1) / E::QUADS);
  eval_quad_kernel<P><<<grid, 128, E::SMEM_BYTES, s>>>(a);
  launch_2d<OcpEvalTail<P>>(a, a.n_instances(), OcpEvalTail<P>::NCHUNK, s);
  return cudaGetLastError();
}

#endif  // HB_DEVICE_BUILD
}  // namespace hb
