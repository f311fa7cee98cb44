// Riccati sweep for sm_100a: ONE WARP PER INSTANCE on the fp64 tensor pipe (mma.sync.m8n8k4.f64), written for a
// minimal instruction count.
//
// Why a warp per instance: the BASELINE batch has 4096 instances.  With 1-8 lanes per instance (riccati2.cuh and its
// predecessors, all 55-60 us) a scheduler holds one or two warps and every stage exposes its dependent latency chain.
// With a warp per instance all 4096 warps are resident at once (28 per SM, 7 per scheduler) and latency is hidden by
// thread-level parallelism; what bounds the kernel then is the number of issued instructions and the fp64 pipe, so the
// stage is cast as three 8x8x4 tensor-core products whose fragment layout spreads the work over the 32 lanes:
//   W      = P [A B b 0]             + [0 .. 0 p lam+]        A-operand P, B-operand M = [A B b 0]
//   H      = M^T W                   + [Q S^T q qbar; S R r rbar]
//            -> Hxx, Hux, Huu, Newton gradient (column 6) and, for free, the adjoint lam_k / reduced gradient
//               (column 7 = [A B]^T lam_{k+1} + [qbar; rbar]: the dual infeasibility)
//   P'|p'  = Hux^T [K | k | 0]       + [Hxx | h_x | lam_k],   K = -Huu^{-1} Hux, k = -Huu^{-1} h_u (closed-form 2x2)
// Column 7 rides through all three products untouched by the feedback (its B-operand column is exactly zero), so
// lam_k needs no lane roles of its own.  Between the products the accumulator fragment passes through a small
// shared-memory tile and is read back in the operand layout of the next product (one 128-bit store and one load per
// lane; the three tiles are distinct, so three __syncwarp per stage order everything).
//
// The stage blocks arrive in operand order (OcpT::FRAG layout, written that way by the evaluation kernels): lane l
// reads M[l & 3][l >> 2] = qa[l] and the cost entries qh[2l], qh[2l + 1] - two coalesced loads per stage straight from
// HBM into a register ring PF stages deep; no staging buffers, no barriers.  Per-lane shared-memory addresses are set up
// once so that lanes which must see a zero read it from a zero word instead of selecting.
//
// Forward sweep (x_0 fixed): [dx+; du] = [A; K] dx + [b; k] as one product (A-operand rows 0-3 from HBM, rows 4-5 =
// the gains from shared memory, through one generic pointer per lane), one shuffle hands du to the operand layout,
// dx+ += [B | 0] [du; 0] as a second product, one shuffle hands dx+ on.
//
// A non-positive-definite Huu (non-convex problem far from the solution) is only flagged in the sweep (ST_ICREQ); the
// small kernel OcpRiccatiFix that follows redoes such instances with the inertia-correcting one-thread-per-instance
// sweep (OcpRiccati: delta_w * I on the primal variables, IPOPT's schedule).
// Every primitive goes through warp_emu.cuh, so the same source runs lane-by-lane on the host in the CPU logic tests.
#pragma once
#include "ocp_kernels.cuh"
#include "warp_emu.cuh"

// tools/ric_bench.cu only: 1 = every stage re-reads the first block (no HBM traffic), 2 = loads without the stage
// arithmetic.  Both give wrong results by construction; the product always builds mode 0.
#ifndef HB_R6_MODE
#define HB_R6_MODE 0
#endif
#ifndef HB_R6_PFMODE      // L2 prefetch: 0 none, 1 one lane per line, 2 all lanes (surplus lanes repeat lines)
#define HB_R6_PFMODE 1
#endif
#ifndef HB_R6_HINT        // 1: evict-first L2 policy on the cost-tile loads
#define HB_R6_HINT 1
#endif
#ifndef HB_R6_PD
#define HB_R6_PD 6
#endif
#ifndef HB_R6_SKIPFWD     // 1: an instance that the step kernel will find converged skips the forward sweep (tuning runs)
#define HB_R6_SKIPFWD 0
#endif

namespace hb {

template <class P, int PF_, int PD_ = HB_R6_PD>
struct Riccati6 {
  using T = OcpT<P>;
  static constexpr int NX = T::NX, NU = T::NU, N = T::N;
  static constexpr int CW = 4;                        // warps (= instances) per CTA
  static constexpr int THREADS = CW * 32;
  static constexpr int PF = PF_;                      // stages of register loads in flight per warp
  static constexpr int PD = PD_;                      // stages the L2 prefetch runs ahead of the sweep
  // gains of a stage in shared memory: [K0(4) k0 | K1(4) k1]
  static constexpr int GD = 10;
  // W and P' (rows 0-3; 32 bytes of padding behind every pair of rows, entry (r, c) at rof(r) + c: the row-major 128-bit
  // fragment stores and the transposed / mirrored 64-bit operand loads are then both free of bank conflicts), rows 4, 5
  // of H (all the feedback needs: Huu, Hux, h_u), a zero word on a bank pair that no operand load of either half-warp
  // touches (lanes that must feed zeros into a product read it instead of selecting), gains
  static HB_HD constexpr int rof(int r) { return r * 8 + 4 * (r >> 1); }
  static constexpr int oW = 0, oP = 40, oH = 80, oZ = oH + 23, oG = oH + 24;
  static constexpr int WSTRIDE = oG + N * GD + ((N * GD) & 1);
  static constexpr int SMEM_BYTES = CW * WSTRIDE * 8;
  // 7 CTAs per SM (all 4096 instances of the BASELINE batch resident) up to N = 56; longer horizons keep their gains in
  // shared memory at a lower occupancy (one CTA per SM at N ~ 700), still far ahead of one thread per instance
  static constexpr bool SUPPORTED = T::FRAG && (NX >= 1) && (NU >= 1) && (N >= 2 * PF) && (PF >= 1) &&
                                    (SMEM_BYTES + 1024 <= 227 * 1024);

  static HB_WP void store2(double* dst, double v0, double v1) {   // dst 16-byte aligned
#if HB_DEVICE_BUILD
    *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
#else
    dst[0] = v0; dst[1] = v1;
#endif
  }
  static HB_WP void load2(const double* src, double& v0, double& v1) {
#if HB_DEVICE_BUILD
    const double2 t = *reinterpret_cast<const double2*>(src);
    v0 = t.x; v1 = t.y;
#else
    v0 = src[0]; v1 = src[1];
#endif
  }
  static HB_WP void ldg2(const double* src, double& v0, double& v1) {
#if HB_DEVICE_BUILD
    const double2 t = __ldg(reinterpret_cast<const double2*>(src));
    v0 = t.x; v1 = t.y;
#else
    v0 = src[0]; v1 = src[1];
#endif
  }
  static HB_WP void ldg2p(const double* src, double& v0, double& v1, wp::l2pol_t pol) {
    if (HB_R6_HINT) wp::ldg2_pol(src, v0, v1, pol);
    else ldg2(src, v0, v1);
  }
  static HB_WP double ldg(const double* p) {
#if HB_DEVICE_BUILD
    return __ldg(p);
#else
    return *p;
#endif
  }

  // Per-lane shared-memory addresses and one flag word, set up once.  The kernel is bound by the load/store data pipe
  // of the SM (one 128-byte wavefront per cycle for the shared AND global traffic of all four schedulers), so every
  // access is issued by exactly the lanes that carry data: rows 0-3 of W / P', rows 4-5 of H, no dump slots.
  struct Lane {
    wp::sa_t wst, hst;                // fragment stores: W at wst, P' at wst + oP (lanes gid < 4); H rows 4, 5 at hst
    wp::sa_t wrd, prd;                // operand loads: W[tig][gid], P'[max][min]
    wp::sa_t own, oth, sel;           // H[4 + t][gid], H[5 - t][gid] (t = tig & 1; the zero word in the lanes that feed
                                      // zeros), H[5 - t][5 - t]
    wp::sa_t huu;                     // Huu = H[4..5][4..5] (warp-uniform)
    int f;                            // flag bits below
  };
  enum { F_ROW03 = 1, F_ROW45 = 2, F_LAM = 4, F_GAIN = 8, F_COL05 = 16 };

  // one backward stage; (c0, c1) = cost fragment, mB = M operand, state (pA, w0, w1) carried between stages;
  // gk = this lane's gain slot of the stage, lamk = its lam_k slot, ed = running max |column 7| (dual infeasibility in
  // the lanes of rows 4, 5)
  template <bool SLOW>
  static HB_WP void bstage(const Lane& L, double mB, double c0, double c1, double& pA, double& w0, double& w1,
                           wp::sa_t gk, double* lamk, double& ed, int& ok) {
    const int f = L.f;
    wp::dmma(w0, w1, pA, mB);                              // W = P M (+ p, lam+ in columns 6, 7)
    wp::sts2_if<F_ROW03, 0>(f, L.wst, w0, w1);
    wp::syncwarp();
    const double wB = wp::lds<0>(L.wrd);
    double h0, h1;
    wp::dmma4(h0, h1, mB, wB, c0, c1);                     // H = M^T W + cost tile
    wp::stg_if<F_LAM>(f, lamk, h1);                        // lam_k (column 7, rows 0..nx-1)
    wp::sts2_if<F_ROW45, 0>(f, L.hst, h0, h1);
    { const double ah = fabs(h1); ed = ah > ed ? ah : ed; }   // rows 4, 5 of column 7: the reduced gradient
    wp::syncwarp();
    double h00, h01, h10, h11;
    wp::lds2<0>(L.huu, h00, h01);
    wp::lds2<8>(L.huu, h10, h11);
    const double hown = wp::lds<0>(L.own), hoth = wp::lds<0>(L.oth);   // row tig of [Hux h_u], the other row
    double dsel = wp::lds<0>(L.sel);                       // d0 for the first row of the inverse, a0 for the second
    (void)h10;
    double det = h00 * h11 - h01 * h01;
    if constexpr (SLOW) {
      if (!((h00 > 0.0) && (det > 0.0))) {                 // identical in all lanes: regularise until positive definite
        double reg = 1e-8;
        bool pd = false;
        for (int attempt = 0; attempt < 11 && !pd; ++attempt) {
          const double a0 = h00 + reg, d0 = h11 + reg;
          det = a0 * d0 - h01 * h01;
          pd = (a0 > 0.0) && (det > 0.0);
          if (pd) dsel += reg;
          reg *= 100.0;
        }
        if (!pd) { ok = 0; det = 1.0; }
      }
    } else {
      ok &= (int)(h00 > 0.0) & (int)(det > 0.0);
    }
    const double rdet = wp::rcp(det);
    const double kk = (h01 * hoth - dsel * hown) * rdet;   // K[tig][gid] (gid < 4), k[tig] (gid == 6); exactly 0 in the
                                                           // lanes of B-operand rows 2, 3 and column 7 (lam rides through)
    wp::sts_if<F_GAIN, 0>(f, gk, kk);
    wp::dmma4(w0, w1, hown, kk, h0, h1);                   // P' | p' | lam = Hux^T [K | k | 0] + [Hxx | h_x | lam]
    wp::sts2_if<F_ROW03, oP>(f, L.wst, w0, w1);
    wp::zero_if<F_COL05>(f, w0, w1);                       // accumulator input of the next stage: [0 .. 0 | p' | lam]
    wp::syncwarp();
    pA = wp::lds<0>(L.prd);                                // P'[max][min]: exactly symmetric
  }

  template <bool SLOW>
  static HB_WP void backward(const OcpArgs& a, const Lane& L, int b, int lane, double* gs, double& ed, int& ok) {
    const int gid = lane >> 2, tig = lane & 3, g4 = gid & 3;
    const int f = L.f;
    // everything of an instance is contiguous: stage k at + k * block, i.e. compile-time offsets after unrolling.
    // Rows 6, 7 of the cost fragment are never used: those lanes re-read rows 4, 5 (finite, bounded).
    const int lc = lane < 24 ? lane : lane - 8;
    const double* pM = a.qa + ((long long)b * N + (N - 1)) * T::NAp + lane;
    const double* pC = a.qh + ((long long)b * N + (N - 1)) * T::NHp + 2 * lc;
    // terminal data: P = Q_N, p = q_N, lam_N = qbar_N
    const double* tn = a.qpN + (long long)b * T::SDN;
    double pA = (g4 < NX && tig < NX) ? tn[g4 * NX + tig] : 0.0;
    double w0 = (tig == 3 && gid < NX) ? tn[NX * NX + gid] : 0.0;
    double w1 = (tig == 3 && gid < NX) ? tn[NX * NX + NX + gid] : 0.0;
    double* lamk = a.lam + ((long long)b * (N + 1) + N) * NX + (gid < NX ? gid : 0);
    wp::stg_if<F_LAM>(f, lamk, w1);
    lamk -= NX;                                         // -> stage N-1
    // gain slot K[tig][gid] | k[tig] (gid 6)
    wp::sa_t gk = wp::keep(wp::sa(gs + (N - 1) * GD + (tig & 1) * 5 + (gid < 4 ? gid : 4)));
    ed = 0.0;
    // L2 prefetch PD stages ahead of the register loads, one instruction per PF stages: lane l takes the 128-byte line
    // l % LPB of the 2 * PF qa lines and 3 * PF qh lines of those stages (surplus lanes repeat a line); the register
    // ring then sees L2 latency instead of HBM latency
    constexpr int LPB = 5 * PF;
    const int pl = lane % LPB;
    const double* qa_b = a.qa + (long long)b * N * T::NAp;
    const double* qh_b = a.qh + (long long)b * N * T::NHp;
    // lines of the block of stages [k - PF + 1, k] are contiguous in both arrays; pp points into the block that ends at
    // stage N - 1 - PF - PD (the first one the loop prefetches), pstep = doubles per PF stages
    const bool pqa = pl < 2 * PF;
    const int pstep = pqa ? PF * T::NAp : PF * T::NHp;
    const double* pp = pqa ? qa_b + (long long)(N - 2 * PF - PD) * T::NAp + 16 * pl
                           : qh_b + (long long)(N - 2 * PF - PD) * T::NHp + 16 * (pl - 2 * PF);
#if HB_R6_MODE != 1
#pragma unroll
    for (int d = 1; d <= PD / PF; ++d)
      if (N - PF - d * PF >= 0 && HB_R6_PFMODE != 0 && (HB_R6_PFMODE == 2 || lane < LPB))
        wp::prefetch_l2<false>(pp + (long long)(PD / PF + 1 - d) * pstep);   // stages before the loop's
#endif
    const wp::l2pol_t pol = wp::l2_evict_first();       // the cost tiles are read once: first out of L2, so that the
                                                        // dynamics blocks survive until the forward sweep reads them again
    double qM[PF], qC0[PF], qC1[PF];
#pragma unroll
    for (int d = 0; d < PF; ++d) { qM[d] = ldg(pM - d * T::NAp); ldg2(pC - d * T::NHp, qC0[d], qC1[d]); }
    // stages j = 0 .. N-1 (k = N-1-j), slot j % PF; the refill of a slot fetches stage j + PF
    constexpr int MAIN = (N / PF - 1) * PF;            // stages whose refill is always in range
    constexpr int PFEND = N - 2 * PF - PD;             // blocks starting at stage j0 < PFEND + 1 prefetch inside the arrays
    for (int j0 = 0; j0 < MAIN; j0 += PF) {
#pragma unroll
      for (int d = 0; d < PF; ++d) {
        const double mB = qM[d], h0 = qC0[d], h1 = qC1[d];
#if HB_R6_MODE != 1
        if (d == 0 && HB_R6_PFMODE != 0) { if (j0 <= PFEND && (HB_R6_PFMODE == 2 || lane < LPB)) wp::prefetch_l2<false>(pp); pp -= pstep; }
#endif
        qM[d] = ldg(pM - (d + PF) * T::NAp);
        ldg2p(pC - (d + PF) * T::NHp, qC0[d], qC1[d], pol);
#if HB_R6_MODE == 2
        w0 += mB + h0; w1 += h1;
#else
        bstage<SLOW>(L, mB, h0, h1, pA, w0, w1, wp::sadd(gk, -d * GD), lamk - d * NX, ed, ok);
#endif
      }
#if HB_R6_MODE != 1
      pM -= PF * T::NAp; pC -= PF * T::NHp;
#endif
      gk = wp::sadd(gk, -PF * GD); lamk -= PF * NX;
    }
#pragma unroll
    for (int j = MAIN; j < N; ++j) {
      const int d = j % PF, r = j - MAIN;
      const double mB = qM[d], h0 = qC0[d], h1 = qC1[d];
      if (j + PF < N) {
        qM[d] = ldg(pM - (r + PF) * T::NAp);
        ldg2p(pC - (r + PF) * T::NHp, qC0[d], qC1[d], pol);
      }
      bstage<SLOW>(L, mB, h0, h1, pA, w0, w1, wp::sadd(gk, -r * GD), lamk - r * NX, ed, ok);
    }
  }

  static HB_WP void run(const OcpArgs& a) {
    double* smem = wp::dyn_smem();
    const int warp = wp::tid() >> 5, lane = wp::tid() & 31;
    const int gid = lane >> 2, tig = lane & 3, g4 = gid & 3;
    const int bi = wp::bid() * CW + warp;
    if (bi >= a.n_instances()) return;                  // warp-uniform
    const int b = a.instance(bi);                       // late iterations run over the list of active instances
    double* st = a.st + (long long)b * ST_COUNT;
    if (st[ST_ACTIVE] == 0.0) return;

    double* Wsm = smem + (size_t)warp * WSTRIDE;
    double* gs = Wsm + oG;

    Lane L;
    L.wst = wp::keep(wp::sa(Wsm + oW + rof(g4) + 2 * tig));
    L.hst = wp::keep(wp::sa(Wsm + oH + (gid & 1) * 8 + 2 * tig));
    L.wrd = wp::keep(wp::sa(Wsm + oW + rof(tig) + gid));
    L.prd = wp::keep(wp::sa(Wsm + oP + (g4 > tig ? rof(g4) + tig : rof(tig) + g4)));
    const int t1 = tig & 1;
    const bool feeds = (tig < 2) && (gid != 7);         // lanes whose K entry / Hux operand is not identically zero
    L.own = wp::keep(wp::sa(Wsm + (feeds ? oH + t1 * 8 + gid : oZ)));
    L.oth = wp::keep(wp::sa(Wsm + (feeds ? oH + (1 - t1) * 8 + gid : oZ)));
    L.sel = wp::keep(wp::sa(Wsm + oH + (1 - t1) * 8 + (5 - t1)));
    L.huu = wp::keep(wp::sa(Wsm + oH + 4));
    {
      int f = 0;
      if (gid < 4) f |= F_ROW03;
      if (gid == 4 || gid == 5) f |= F_ROW45;
      if (tig == 3 && gid < NX) f |= F_LAM;
      if (tig < 2 && (gid < 4 || gid == 6)) f |= F_GAIN;
      if (tig != 3) f |= F_COL05;
      L.f = wp::keep(f);
    }
    if (lane == 0) Wsm[oZ] = 0.0;
    wp::syncwarp();

    double ed = 0.0;
    int ok = 1;
    backward<false>(a, L, b, lane, gs, ed, ok);
    if (wp::any(ok == 0)) {                             // rare: some Huu was not positive definite
      // The instance is handed to OcpRiccatiFix (next launch: inertia-correcting sweep, delta_w * I with IPOPT's schedule),
      // which overwrites everything written here.  The stage-wise regularising repetition below only keeps the numbers of
      // this launch finite; it also happens to be the code shape for which the 72-register allocation of the hot path
      // above is spill-free (an early return or a call here costs 30-60 bytes of spills in the stage loop: 47 -> 57 us).
      if (lane == 0) st[ST_ICREQ] = 1.0;
      ok = 1;
      wp::syncwarp();
      backward<true>(a, L, b, lane, gs, ed, ok);
      ok = 1;
    }
    wp::syncwarp();
#if HB_R6_SKIPFWD
    {  // the convergence test of the step kernel (ls_prelude): when it will stop the instance, the step is never used
      double cinf = 0.0, comp = 0.0;
      for (int k = lane; k <= N; k += 32) {
        cinf = dmax(cinf, a.s_cinf[(long long)b * (N + 1) + k]);
        comp = dmax(comp, a.s_comp[(long long)b * (N + 1) + k]);
      }
      double e = dmax(cinf, comp);
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) e = dmax(e, wp::shfl_xor(e, m, 32));
      const double e4 = wp::shfl(ed, 19, 32), e5 = wp::shfl(ed, 23, 32);
      const double edual = NU > 1 ? dmax(e4, e5) : e4;
      if (ok && st[ST_MU] <= a.mu_min * (1.0 + 1e-12) && dmax(e, edual) <= a.tol) {
        if (lane == 0) st[ST_EDUAL] = edual;
        return;
      }
    }
#endif

    // =============================================================== forward sweep (x_0 fixed: dx_0 = 0)
    // product 1, A-operand rows 0-3: A[gid][tig] (global), rows 4-7: K[gid & 1][tig] (shared memory; rows 6, 7 idle copies)
    //            accumulator init  : b[gid]                ,           k[gid & 1]         (column 0 is the one that counts)
    // product 2, A-operand: [B | 0 0][gid][tig] for rows 0-3 (the zero column of M for tig >= 2), 0 for rows 4-7
    const bool glob = gid < 4;
    const double* pq = a.qa + (long long)b * N * T::NAp;
    const double* pg1 = pq + (glob ? tig * 4 + gid : 0);
    const double* pgc = pq + (glob ? 24 + gid : 0);
    const double* pg2 = pq + (tig < 2 ? (4 + tig) * 4 + g4 : 28 + g4);
    const double* ps1 = gs + (gid & 1) * 5 + tig;
    const double* psc = gs + (gid & 1) * 5 + 4;
    auto ld1 = [&](int k) { return glob ? ldg(pg1 + k * T::NAp) : ps1[k * GD]; };
    auto ldc = [&](int k) { return glob ? ldg(pgc + k * T::NAp) : psc[k * GD]; };
    double q1[PF], qc[PF], q2[PF];
#pragma unroll
    for (int d = 0; d < PF; ++d) { q1[d] = ld1(d); qc[d] = ldc(d); q2[d] = glob ? ldg(pg2 + d * T::NAp) : 0.0; }
    double* pdx = a.dx + (long long)b * (N + 1) * NX + (tig < NX ? tig : 0);
    double* pdu = a.du + (long long)b * N * NU + ((gid - 4) >= 0 && (gid - 4) < NU ? gid - 4 : 0);
    const bool st_dx = (gid == 0) && (tig < NX);
    const bool st_du = (tig == 0) && (gid >= 4) && (gid - 4 < NU);
    const int src_u = 16 + 4 * (tig & 1), src_x = 4 * tig;
    double vx = 0.0;                                     // dx_k[tig]
    auto fstage = [&](double a1, double c0, double a2, int k) {
      double c1 = 0.0;
      wp::dmma(c0, c1, a1, vx);                          // rows 0-3: A dx + b, rows 4-5: du = K dx + k
      const double v2 = wp::shfl(c0, src_u, 32);         // [du0 du1 . .][tig] (rows 2, 3 meet the zero columns of product 2)
      if (st_dx) pdx[k * NX] = vx;
      if (st_du) pdu[k * NU] = c0;
      wp::dmma(c0, c1, a2, v2);                          // rows 0-3 += B du
      vx = wp::shfl(c0, src_x, 32);                      // dx_{k+1}[tig]
    };
    constexpr int MAIN = (N / PF - 1) * PF;
    for (int j0 = 0; j0 < MAIN; j0 += PF) {
#pragma unroll
      for (int d = 0; d < PF; ++d) {
        const double a1 = q1[d], c0 = qc[d], a2 = q2[d];
        q1[d] = ld1(d + PF); qc[d] = ldc(d + PF);
        if (glob) q2[d] = ldg(pg2 + (d + PF) * T::NAp);
#if HB_R6_MODE == 2
        vx += a1 + c0 + a2;
#else
        fstage(a1, c0, a2, d);
#endif
      }
#if HB_R6_MODE != 1
      pg1 += PF * T::NAp; pgc += PF * T::NAp; pg2 += PF * T::NAp;
#endif
      ps1 += PF * GD; psc += PF * GD;
      pdx += PF * NX; pdu += PF * NU;
    }
#pragma unroll
    for (int j = MAIN; j < N; ++j) {
      const int d = j % PF, r = j - MAIN;
      const double a1 = q1[d], c0 = qc[d], a2 = q2[d];
      if (j + PF < N) {
        q1[d] = ld1(r + PF); qc[d] = ldc(r + PF);
        if (glob) q2[d] = ldg(pg2 + (r + PF) * T::NAp);
      }
      fstage(a1, c0, a2, r);
    }
    if (st_dx) pdx[(N - MAIN) * NX] = vx;
    // ---- per-instance results: dual infeasibility = max |reduced gradient| over the stages (lanes tig = 3 of rows 4, 5)
    {
      const double e4 = wp::shfl(ed, 19, 32), e5 = wp::shfl(ed, 23, 32);
      ed = NU > 1 ? dmax(e4, e5) : e4;
    }
    if (lane == 0) {
      st[ST_EDUAL] = ed;
      if (!ok) { a.status[b] = -1; st[ST_ACTIVE] = 0.0; }
    }
  }
};

#if HB_DEVICE_BUILD
template <class R>
__global__ void __launch_bounds__(R::THREADS, 7) riccati6_kernel(OcpArgs a) { R::run(a); }
#endif

template <class R>
inline int launch_riccati6(const OcpArgs& a, cudaStream_t s) {
  const int grid = (a.n_instances() + R::CW - 1) / R::CW;
#if HB_DEVICE_BUILD
  cudaError_t e = cudaFuncSetAttribute(riccati6_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, R::SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  riccati6_kernel<R><<<grid, R::THREADS, R::SMEM_BYTES, s>>>(a);
  return (int)cudaGetLastError();
#else
  (void)s;
  wp::launch<R>(a, grid, R::THREADS, (size_t)R::SMEM_BYTES);
  return 0;
#endif
}

}  // namespace hb
