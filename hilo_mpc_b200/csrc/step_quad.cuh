// Step computation of one interior-point iteration as ONE kernel, sm_100a version for the collocation transcription:
// one CTA per instance, 4 lanes per node.
//
//   Expand -> Prep -> [Trial(alpha) -> Decide]* -> Apply -> Commit        (ocp_module.cuh runs these as 6-8 launches)
//
// Every bounded variable of a node (x_k | u_k | X_k, NV = nx + nu + deg*nx of them) is owned by one lane of the node's
// quad (variable e belongs to lane e % 4) and lives in that lane's registers from the first load to the final store:
// value, step, bounds, both bound multipliers.  The iterate therefore crosses HBM once per iteration in each direction
// instead of once per phase, consecutive lanes touch consecutive addresses (the one-thread-per-node kernels walk the
// 816-byte elimination blocks with a 816-byte stride between threads), and dX never goes to memory at all.
//   expand : own rows of  dX = t0 + T [dx; du],  fraction-to-the-boundary limits, directional derivative
//   trial  : own variables at v + alpha d into a shared-memory copy of the instance; lanes 0..deg-1 evaluate the model
//            and the Lagrange term at "their" collocation point, lane deg the continuity residual
//   decide : thread 0, the same filter line-search helpers as OcpDecide, one candidate at a time (alpha_max 2^-i)
//   apply  : own variables and multipliers updated in registers and stored.
// Reductions over the instance are warp shuffles + one shared-memory round.  Supported: collocation, deg <= 3, continuous
// transcriptions incl. the MHE shape (free x_0, arrival cost, additive stage input, discrete-type objective); explicit RK
// and discrete problems use the thread-per-node bodies (step_fused.cuh or separate launches).
#pragma once
#include "ocp_kernels.cuh"

namespace hb {
#if HB_DEVICE_BUILD

template <class P>
struct StepQuad {
  using T = OcpT<P>;
  static constexpr int NX = T::NX, NU = T::NU, NC = T::NC, NE = T::NE, N = T::N, D = P::DEG;
  static constexpr int NV = NX + NU + NC, NVIT = (NV + 3) / 4;
  static constexpr int THREADS = (((N + 1) * 4 + 31) / 32) * 32, WARPS = THREADS / 32;
  static constexpr bool SUPPORTED = (P::METHOD == METHOD_COLLOCATION) && (D <= 3) && (THREADS <= 1024);
  // shared memory (doubles): trial copy of the instance | warp partials | control block
  static constexpr int oTRIAL = 0, oRED = oTRIAL + (N + 1) * NV, oCTRL = oRED + 8 * 32, oST = oCTRL + 8,
                       SMEM = oST + ST_COUNT;
  enum { C_FLAG = 0, C_ALPHA, C_AD, C_MU };   // control block entries
  enum { F_STOP = 0, F_EVAL = 1, F_APPLY = 2 };

  static __device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
  }
  static __device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = dmin(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
  }
  static __device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = dmax(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
  }

  static __device__ void run(const OcpArgs& a) {
    __shared__ __align__(16) double smem[SMEM];
    const int b = a.instance(blockIdx.x), tid = threadIdx.x, k = tid >> 2, q = tid & 3, warp = tid >> 5, lane = tid & 31;
    double* stg = a.st + (long long)b * ST_COUNT;
    if (stg[ST_ACTIVE] == 0.0) return;   // uniform over the CTA
    // the per-instance solver state is edited by thread 0 between the phases: work on a shared-memory copy (one
    // coalesced load, no global round trips on the serial path) and write it back at the end
    double* st = smem + oST;
    for (int i = tid; i < ST_COUNT; i += THREADS) st[i] = stg[i];   // THREADS may be < ST_COUNT (N <= 7)
    const double mu = stg[ST_MU];
    const double tau = dmax(0.99, 1.0 - mu);
    const bool node = k <= N, inter = k < N;
    const long long nb = (long long)b * (N + 1) + (node ? k : N), sb = (long long)b * N + (inter ? k : N - 1);
    const double* el = a.el + sb * T::TD;
    double* red = smem + oRED;
    double* ctrl = smem + oCTRL;

    // ------------------------------------------------------------------ early exit of a converged instance
    // One launch in four (3-4 iterations per solve) only has to find out that its instance is done: the errors of the
    // current point are known from the evaluation and the sweep, so the test comes before the 900 bytes per node
    // of step data are touched.  Same decision and same stored state as the regular path below (ls_prelude).
    {
      double c1 = 0.0, phi = 0.0, cinf = 0.0, comp = 0.0;
      if (node && q == 0) { c1 = a.s_c1[nb]; phi = a.s_phi[nb]; cinf = a.s_cinf[nb]; comp = a.s_comp[nb]; }
      c1 = warp_sum(c1); phi = warp_sum(phi); cinf = warp_max(cinf); comp = warp_max(comp);
      if (lane == 0) { red[3 * 32 + warp] = c1; red[4 * 32 + warp] = phi; red[5 * 32 + warp] = cinf; red[6 * 32 + warp] = comp; }
      __syncthreads();
      if (tid == 0) {
        double rc1 = 0.0, rphi = 0.0, rcinf = 0.0, rcomp = 0.0;
        for (int w = 0; w < WARPS; ++w) {
          rc1 += red[96 + w]; rphi += red[128 + w]; rcinf = dmax(rcinf, red[160 + w]); rcomp = dmax(rcomp, red[192 + w]);
        }
        const double err = dmax(dmax(st[ST_EDUAL], rcinf), rcomp);
        const bool done = (mu <= a.mu_min * (1.0 + 1e-12)) && (err <= a.tol);
        if (done) { prep_store(st, 1.0, 1.0, 0.0, rc1, rphi, rcinf, rcomp); ls_prelude(a, st, b, mu, err); }
        ctrl[C_FLAG] = done ? (double)F_STOP : (double)F_EVAL;
      }
      __syncthreads();
      if ((int)ctrl[C_FLAG] == F_STOP) {
        for (int i = tid; i < ST_COUNT; i += THREADS) stg[i] = st[i];
        return;
      }
      __syncthreads();   // ctrl / red are rewritten below
    }
#ifndef HB_STEPQ_PREFETCH
#define HB_STEPQ_PREFETCH 1
#endif
    // the 816-byte elimination block of the node (written two kernels ago, mostly in HBM by now): its seven 128-byte
    // lines are requested up front by the four lanes of the node, so that the dependent loads below find them on chip
    if (HB_STEPQ_PREFETCH && inter) {
      const char* eb = reinterpret_cast<const char*>(el);
      constexpr int ELB = T::TD * 8;
      if (q * 256 < ELB) asm volatile("prefetch.global.L2 [%0];" ::"l"(eb + q * 256));
      if (q * 256 + 128 < ELB) asm volatile("prefetch.global.L2 [%0];" ::"l"(eb + q * 256 + 128));
    }
    // ------------------------------------------------------------------ load the own variables
    // slot it holds variable e = 4 it + q of the node:  e < NX: x_k[e] | e < NX+NU: u_k[e-NX] | else X_k[e-NX-NU]
    double v[NVIT], d[NVIT], lb[NVIT], ub[NVIT], zl[NVIT], zu[NVIT];
    int kind[NVIT];   // 0: no variable, 1: value only (x_0 of a pinned initial state), 2: bounded variable
    double ap = 1.0, ad = 1.0, dphi = 0.0;
    double dxu[NE];
    if (inter) {
#pragma unroll
      for (int j = 0; j < NX; ++j) dxu[j] = a.dx[nb * NX + j];
#pragma unroll
      for (int j = 0; j < NU; ++j) dxu[NX + j] = a.du[sb * NU + j];
    }
    double Ex[NX];
    if (node && !inter) {   // terminal node: gradient of the Mayer term
      const double* tp = a.tp + (long long)b * a.tp_stride;
      double E, Exx[NX * NX];
      P::may_all(a.x + nb * NX, tp, a.time0 + k * a.dt, &E, Ex, Exx);
    }
#pragma unroll
    for (int it = 0; it < NVIT; ++it) {
      const int e = it * 4 + q;
      v[it] = 0.0; d[it] = 0.0; lb[it] = -kInfBound; ub[it] = kInfBound; zl[it] = 0.0; zu[it] = 0.0; kind[it] = 0;
      if (!node) continue;
      if (e < NX) {
        const bool x_free = (k > 0) || !P::FIXED_X0;
        v[it] = a.x[nb * NX + e]; d[it] = a.dx[nb * NX + e];
        kind[it] = x_free ? 2 : 1;
        if (x_free) { lb[it] = a.xlb[e]; ub[it] = a.xub[e]; zl[it] = a.zlx[nb * NX + e]; zu[it] = a.zux[nb * NX + e]; }
        if (inter) {
          dphi += el[T::ogx + e] * d[it];
        } else {
          double exi = 0.0;
#pragma unroll
          for (int j = 0; j < NX; ++j)
            if (j == e) exi = Ex[j];
          const BarrierTerm bt = barrier_term(v[it], lb[it], ub[it], 0.0, 0.0, mu);
          dphi += (exi + bt.grad) * d[it];
        }
      } else if (inter && e < NX + NU) {
        const int i = e - NX;
        v[it] = a.u[sb * NU + i]; d[it] = a.du[sb * NU + i];
        lb[it] = a.ulb[i]; ub[it] = a.uub[i]; zl[it] = a.zlu[sb * NU + i]; zu[it] = a.zuu[sb * NU + i];
        kind[it] = 2;
        dphi += el[T::ogu + i] * d[it];
      } else if (inter && e < NV) {
        const int r = e - NX - NU, j = r % NX;
        v[it] = a.xc[sb * NC + r];
        lb[it] = a.xlb[j]; ub[it] = a.xub[j]; zl[it] = a.zlc[sb * NC + r]; zu[it] = a.zuc[sb * NC + r];
        kind[it] = 2;
        double s = el[T::ot0 + r];
        if constexpr ((NE % 2 == 0) && (T::oT % 2 == 0) && (T::TD % 2 == 0)) {
          const double2* Tr = reinterpret_cast<const double2*>(el + T::oT + r * NE);
#pragma unroll
          for (int c = 0; c < NE / 2; ++c) { const double2 tv = Tr[c]; s += tv.x * dxu[2 * c] + tv.y * dxu[2 * c + 1]; }
        } else {
          const double* Tr = el + T::oT + r * NE;
#pragma unroll
          for (int c = 0; c < NE; ++c) s += Tr[c] * dxu[c];
        }
        d[it] = s;
        dphi += el[T::ogX + r] * s;
      }
      if (kind[it] == 2) {
        if (has_lb(lb[it])) {
          const double s = v[it] - lb[it];
          ap = frac_to_boundary(s, d[it], tau, ap);
          ad = frac_to_boundary(zl[it], dual_step(zl[it], s, d[it], mu), tau, ad);
        }
        if (has_ub(ub[it])) {
          const double s = ub[it] - v[it];
          ap = frac_to_boundary(s, -d[it], tau, ap);
          ad = frac_to_boundary(zu[it], dual_step(zu[it], s, -d[it], mu), tau, ad);
        }
      }
    }
    // ------------------------------------------------------------------ instance-wide sums of the current point
    {
      double c1 = 0.0, phi = 0.0, cinf = 0.0, comp = 0.0;
      if (node && q == 0) { c1 = a.s_c1[nb]; phi = a.s_phi[nb]; cinf = a.s_cinf[nb]; comp = a.s_comp[nb]; }
      ap = warp_min(ap); ad = warp_min(ad); dphi = warp_sum(dphi); c1 = warp_sum(c1); phi = warp_sum(phi);
      cinf = warp_max(cinf); comp = warp_max(comp);
      if (lane == 0) {
        red[0 * 32 + warp] = ap; red[1 * 32 + warp] = ad; red[2 * 32 + warp] = dphi; red[3 * 32 + warp] = c1;
        red[4 * 32 + warp] = phi; red[5 * 32 + warp] = cinf; red[6 * 32 + warp] = comp;
      }
    }
    __syncthreads();
    // per-instance line-search context, thread 0 only
    double theta0 = 0.0, phi0 = 0.0, dphi0 = 0.0, err = 0.0, last_finite = -1.0, alpha_full = 0.0;
    bool full_finite = false;
    int nf = 0;
    if (tid == 0) {
      double rap = 1.0, rad = 1.0, rd = 0.0, rc1 = 0.0, rphi = 0.0, rcinf = 0.0, rcomp = 0.0;
      for (int w = 0; w < WARPS; ++w) {
        rap = dmin(rap, red[w]); rad = dmin(rad, red[32 + w]); rd += red[64 + w]; rc1 += red[96 + w];
        rphi += red[128 + w]; rcinf = dmax(rcinf, red[160 + w]); rcomp = dmax(rcomp, red[192 + w]);
      }
      prep_store(st, rap, rad, rd, rc1, rphi, rcinf, rcomp);
      err = dmax(dmax(st[ST_EDUAL], rcinf), rcomp);
      const bool go = ls_prelude(a, st, b, mu, err);
      theta0 = rc1; phi0 = rphi; dphi0 = rd;
      if (go) nf = ls_filter_entries(st, mu);
      ctrl[C_FLAG] = go ? (double)F_EVAL : (double)F_STOP;
      ctrl[C_ALPHA] = rap; ctrl[C_AD] = rad;
    }
    __syncthreads();

    // ------------------------------------------------------------------ line search, one candidate at a time
    const double t = a.time0 + k * a.dt, dt = a.dt;
    const double* sp = a.sp + (long long)b * a.sp_stride + (long long)(inter ? k : 0) * T::NSP;
    const double* uo = a.uold + (long long)b * a.uold_stride;
    const bool with_lag = k >= P::LAG_FROM;
    const bool first = (k == 0) && (T::NUOLD > 0);
    double* S = smem + oTRIAL + (node ? k : 0) * NV;
    for (int ai = 0;; ++ai) {
      const int flag = (int)ctrl[C_FLAG];
      if (flag == F_STOP) {
        for (int i = tid; i < ST_COUNT; i += THREADS) stg[i] = st[i];
        return;
      }
      if (flag == F_APPLY) break;
      const double alpha = ctrl[C_ALPHA];
      // own variables at the trial point
      BarrierSum bs;
#pragma unroll
      for (int it = 0; it < NVIT; ++it) {
        if (kind[it] != 0) {
          const double vt = v[it] + alpha * d[it];
          S[it * 4 + q] = vt;
          if (kind[it] == 2) bs.add(vt, lb[it], ub[it]);
        }
      }
      __syncthreads();
      double phi = 0.0, c1 = 0.0;
      if (node) {
        phi = bs.value(mu);
        if (inter) {
          const double* xs = S;
          const double* us = S + NX;
          const double* Xs = S + NX + NU;
          if (q < D) {
            double cb = 0.0, ctau = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i)
              if (q == i) { cb = P::CB(i + 1); ctau = P::CTAU(i); }
            const double ti = t + dt * ctau;
            double Xi[NX], ut[NU > 0 ? NU : 1], f[NX];
#pragma unroll
            for (int r = 0; r < NX; ++r) Xi[r] = Xs[q * NX + r];
#pragma unroll
            for (int r = 0; r < NU; ++r) ut[r] = us[r];
            P::f_val(Xi, ut, sp, ti, dt, f);
            double cc[D + 1];   // column (q + 1) of the collocation matrix
#pragma unroll
            for (int j = 0; j <= D; ++j) {
              cc[j] = 0.0;
#pragma unroll
              for (int i = 0; i < D; ++i)
                if (q == i) cc[j] = P::CC(j * (D + 1) + (i + 1));
            }
#pragma unroll
            for (int r = 0; r < NX; ++r) {
              double xp = cc[0] * xs[r];
#pragma unroll
              for (int j = 1; j <= D; ++j) xp += cc[j] * Xs[(j - 1) * NX + r];
              c1 += dabs(dt * f[r] - xp);
            }
            if (with_lag && !P::OBJ_DISCRETE) {
              double l = P::lag_val(Xi, ut, sp, ti, dt);
              if (first) l += P::lagdu_val(ut, uo);
              phi += dt * cb * l;
            }
          } else if (q == D) {
            const double* xn = S + NV;   // node k + 1
#pragma unroll
            for (int r = 0; r < NX; ++r) {
              double xf = P::CD(0) * xs[r];
#pragma unroll
              for (int i = 1; i <= D; ++i) xf += P::CD(i) * Xs[(i - 1) * NX + r];
              if (P::HAS_GADD) {   // additive stage input (MHE state noise)
#pragma unroll
                for (int c = 0; c < NU; ++c) xf += P::GADD(r * NU + c) * us[c];
              }
              c1 += dabs(xf - xn[r]);
            }
            if ((k == 0 && P::HAS_ARR) || (P::OBJ_DISCRETE && with_lag)) {   // cost terms on the node itself
              double xk[NX], uk[NU > 0 ? NU : 1];
#pragma unroll
              for (int r = 0; r < NX; ++r) xk[r] = xs[r];
#pragma unroll
              for (int r = 0; r < NU; ++r) uk[r] = us[r];
              if (k == 0 && P::HAS_ARR) phi += P::arr_val(xk, a.ap + (long long)b * a.ap_stride);
              if (P::OBJ_DISCRETE && with_lag) {
                double l = P::lag_val(xk, uk, sp, t, dt);
                if (first) l += P::lagdu_val(uk, uo);
                phi += l;
              }
            }
          }
        } else if (q == 0) {
          double xN[NX];
#pragma unroll
          for (int r = 0; r < NX; ++r) xN[r] = S[r];
          phi += P::may_val(xN, a.tp + (long long)b * a.tp_stride, t);
        }
      }
      phi = warp_sum(phi); c1 = warp_sum(c1);
      if (lane == 0) { red[warp] = phi; red[32 + warp] = c1; }
      __syncthreads();
      if (tid == 0) {
        double pht = 0.0, tht = 0.0;
        for (int w = 0; w < WARPS; ++w) { pht += red[w]; tht += red[32 + w]; }
        const LsVerdict vd = ls_candidate(a, st, alpha, pht, tht, theta0, phi0, dphi0, nf);
        if (vd.finite && (ai > 0 || a.n_alpha == 1)) last_finite = alpha;
        if (ai == 0) { full_finite = vd.finite; alpha_full = alpha; }
        double chosen = -1.0;
        bool decided = false, augment = false;
        if (vd.accept) { chosen = alpha; augment = vd.augment; decided = true; }
        else if (ai + 1 >= a.n_alpha) {
          // no candidate acceptable: smallest finite backtracking candidate, else the full step if only that was
          // finite, else the instance fails with status -1 (same rule as OcpDecide; no restoration phase)
          chosen = last_finite;
          if (chosen < 0.0 && a.n_alpha > 1 && full_finite) chosen = alpha_full;
          decided = true;
        }
        if (decided) {
          ls_finish(a, st, b, chosen, augment, theta0, phi0, nf, mu, err);
          ctrl[C_FLAG] = (chosen < 0.0) ? (double)F_STOP : (double)F_APPLY;
          ctrl[C_ALPHA] = chosen;
        } else {
          ctrl[C_ALPHA] = 0.5 * alpha;
        }
      }
      __syncthreads();
    }

    // ------------------------------------------------------------------ apply the step to the own variables
    {
      const double alp = ctrl[C_ALPHA], ald = ctrl[C_AD];
      if (alp != 0.0) {
#pragma unroll
        for (int it = 0; it < NVIT; ++it) {
          if (kind[it] != 2) continue;
          const int e = it * 4 + q;
          const double vn = v[it] + alp * d[it];
          double *pv, *pzl, *pzu;
          if (e < NX) { pv = a.x + nb * NX + e; pzl = a.zlx + nb * NX + e; pzu = a.zux + nb * NX + e; }
          else if (e < NX + NU) { const int i = e - NX; pv = a.u + sb * NU + i; pzl = a.zlu + sb * NU + i; pzu = a.zuu + sb * NU + i; }
          else { const int r = e - NX - NU; pv = a.xc + sb * NC + r; pzl = a.zlc + sb * NC + r; pzu = a.zuc + sb * NC + r; }
          if (has_lb(lb[it])) *pzl = z_update(zl[it], v[it] - lb[it], d[it], mu, ald, vn - lb[it]);
          if (has_ub(ub[it])) *pzu = z_update(zu[it], ub[it] - v[it], -d[it], mu, ald, ub[it] - vn);
          *pv = vn;
        }
      }
    }
    // write the state back; the barrier parameter chosen by the line search is committed on the way
    for (int i = tid; i < ST_COUNT; i += THREADS) stg[i] = (i == ST_MU) ? st[ST_NACC] : st[i];
  }
};

// CTAs per SM the register allocation aims at.  The kernel waits on global loads (44 % of its warp samples), so warps
// per SM matter, but under about HB_STEPQ_REGS registers per thread the spills cost more than the extra warps bring
// (measured: 224-thread CTAs 226 us at 128 registers / 275 us at 80; 128-thread CTAs 332 us at 255 / 259 us at 168):
// as many CTAs as still leave that many registers.
#ifndef HB_STEPQ_REGS
#define HB_STEPQ_REGS 128
#endif
template <class P>
struct StepQuadLaunch {
#ifdef HB_STEPQ_MINB
  static constexpr int MINB = HB_STEPQ_MINB;
#else
  static constexpr int FIT = 65536 / (HB_STEPQ_REGS * StepQuad<P>::THREADS);
  static constexpr int MINB = FIT < 2 ? 2 : (FIT > 8 ? 8 : FIT);
#endif
};
template <class P>
__global__ void __launch_bounds__(StepQuad<P>::THREADS, StepQuadLaunch<P>::MINB) step_quad_kernel(const OcpArgs a) { StepQuad<P>::run(a); }

template <class P>
inline cudaError_t launch_step_quad(const OcpArgs& a, cudaStream_t s) {
  step_quad_kernel<P><<<a.n_instances(), StepQuad<P>::THREADS, 0, s>>>(a);
  return cudaGetLastError();
}

#endif  // HB_DEVICE_BUILD
}  // namespace hb
