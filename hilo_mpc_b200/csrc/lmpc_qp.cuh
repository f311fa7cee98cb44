// Linear MPC as the north star states it: the horizon is condensed on the host once per controller (LMPC.setup), every
// instance then solves the dense QP in the inputs
//     min_u  1/2 u^T H u + (F x0)^T u     s.t.  u_lb <= u <= u_ub,   x_lb <= Phi x0 + Gam u <= x_ub
// (replaces ca.conic('solver', 'qpoases', ...) + solver(h=, g=, a=, lbx=, ubx=, lba=, uba=), mpc.py:2268-2276, :2374).
// n = N nu inputs, mx = N nx predicted states; H = Gam^T Qbar Gam + Rbar, F = Gam^T Qbar Phi.  The state box becomes a set of
// general linear inequalities after condensing, so the QP is solved by a primal-dual interior-point method (Mehrotra
// predictor-corrector) on the inequality form  G u <= h,  G = [I; -I; Gam; -Gam]  restricted to the finite bounds:
// per iteration one n x n Cholesky of H + G^T diag(z / s) G and two back-substitutions.
//
// One thread per instance; all per-instance arrays live in a batch-fastest workspace (element e of instance b at
// work[e * B + b]): the threads of a warp run the same instruction stream (only the iteration count is data dependent)
// and touch consecutive addresses.  Sizes are run-time values (n <= 256): the condensed matrices are shared by the batch
// and are read through the read-only path.
#pragma once
#include "hb_common.cuh"

namespace hb {

struct LmpcArgs {
  int B, n, mx, nx, max_iter;
  double tol;
  const double *H, *F, *Gam, *Phi;       // [n][n], [n][nx], [mx][n], [mx][nx]
  const double *ulb, *uub, *xlb, *xub;   // [n], [n], [mx], [mx]   (|bound| >= 1e19: none)
  const double* x0;                      // [B][nx]
  double* u;                             // [B][n]   out
  double* xpred;                         // [B][mx]  out (or NULL): Phi x0 + Gam u
  int *status, *iters;                   // [B]: 1 optimal, 5 iteration cap, -1 numerical failure
  double* work;
};

HB_HD size_t lmpc_work_per_instance(int n, int mx) {
  const size_t m = 2 * (size_t)n + 2 * (size_t)mx;
  return (size_t)n * n + 5 * (size_t)n + 2 * (size_t)mx + 6 * m;
}

struct LmpcDenseQp {
  static HB_HD void run(const LmpcArgs& a, int b) {
    const int n = a.n, mx = a.mx, nx = a.nx, m = 2 * n + 2 * mx;
    const size_t B = (size_t)a.B;
    double* w = a.work + b;
    size_t off = 0;
    auto take = [&](size_t cnt) { double* p = w + off * B; off += cnt; return p; };
    double* M = take((size_t)n * n);      // KKT matrix / Cholesky factor
    double* g = take(n);                  // linear term F x0
    double* u = take(n);
    double* du = take(n);
    double* rhs = take(n);
    double* rd = take(n);
    double* d = take(mx);                 // Phi x0
    double* Gu = take(mx);                // Gam u
    double* s = take(m);
    double* z = take(m);
    double* ds = take(m);
    double* dz = take(m);
    double* hh = take(m);                 // right-hand side h of G u <= h (+inf rows are skipped through `on`)
    double* rp = take(m);
#define WV(p, i) (p)[(size_t)(i) * B]
    const double* x0 = a.x0 + (size_t)b * nx;
    for (int i = 0; i < n; ++i) {
      double acc = 0.0;
      for (int j = 0; j < nx; ++j) acc += a.F[(size_t)i * nx + j] * x0[j];
      WV(g, i) = acc;
      WV(u, i) = 0.0;
    }
    for (int r = 0; r < mx; ++r) {
      double acc = 0.0;
      for (int j = 0; j < nx; ++j) acc += a.Phi[(size_t)r * nx + j] * x0[j];
      WV(d, r) = acc;
    }
    // rows: [0, n) u <= uub; [n, 2n) -u <= -ulb; [2n, 2n+mx) Gam u <= xub - d; [2n+mx, m) -Gam u <= -(xlb - d)
    auto bound = [&](int r) -> double {
      if (r < n) return a.uub[r];
      if (r < 2 * n) return -a.ulb[r - n];
      if (r < 2 * n + mx) return a.xub[r - 2 * n] - WV(d, r - 2 * n);
      return -(a.xlb[r - 2 * n - mx] - WV(d, r - 2 * n - mx));
    };
    auto on = [&](int r) -> bool {
      if (r < n) return a.uub[r] < kInfBound;
      if (r < 2 * n) return a.ulb[r - n] > -kInfBound;
      if (r < 2 * n + mx) return a.xub[r - 2 * n] < kInfBound;
      return a.xlb[r - 2 * n - mx] > -kInfBound;
    };
    // start: u = projection of 0 on the input box (strictly inside), s = max(h - G u, 1), z = 1
    for (int i = 0; i < n; ++i) {
      double v = 0.0;
      const double lo = a.ulb[i], hi = a.uub[i];
      if (lo > -kInfBound && hi < kInfBound) { if (v <= lo || v >= hi) v = 0.5 * (lo + hi); }
      else if (lo > -kInfBound) { if (v <= lo) v = lo + 1.0; }
      else if (hi < kInfBound) { if (v >= hi) v = hi - 1.0; }
      WV(u, i) = v;
    }
    auto gam_u = [&](const double* vec, double* out) {   // out = Gam vec
      for (int r = 0; r < mx; ++r) {
        double acc = 0.0;
        for (int j = 0; j < n; ++j) acc += a.Gam[(size_t)r * n + j] * WV(vec, j);
        WV(out, r) = acc;
      }
    };
    auto row_dot = [&](int r, const double* vec, const double* gvec) -> double {   // (G vec)_r with Gam vec precomputed
      if (r < n) return WV(vec, r);
      if (r < 2 * n) return -WV(vec, r - n);
      if (r < 2 * n + mx) return WV(gvec, r - 2 * n);
      return -WV(gvec, r - 2 * n - mx);
    };
    gam_u(u, Gu);
    int n_on = 0;
    for (int r = 0; r < m; ++r) {
      WV(hh, r) = bound(r);
      if (on(r)) {
        const double slack = WV(hh, r) - row_dot(r, u, Gu);
        WV(s, r) = slack > 1.0 ? slack : 1.0;
        WV(z, r) = 1.0;
        ++n_on;
      } else {
        WV(s, r) = 1.0; WV(z, r) = 0.0;
      }
    }
    int status = 5, it = 0;
    double scale = 1.0;
    for (int i = 0; i < n; ++i) scale = dmax(scale, dabs(WV(g, i)));
    for (it = 0; it < a.max_iter; ++it) {
      // residuals
      gam_u(u, Gu);
      double mu = 0.0, e_p = 0.0, e_d = 0.0;
      for (int r = 0; r < m; ++r) {
        if (!on(r)) { WV(rp, r) = 0.0; continue; }
        const double v = row_dot(r, u, Gu) + WV(s, r) - WV(hh, r);
        WV(rp, r) = v;
        e_p = dmax(e_p, dabs(v));
        mu += WV(s, r) * WV(z, r);
      }
      mu = n_on > 0 ? mu / n_on : 0.0;
      for (int i = 0; i < n; ++i) {
        double acc = WV(g, i);
        for (int j = 0; j < n; ++j) acc += a.H[(size_t)i * n + j] * WV(u, j);
        if (on(i)) acc += WV(z, i);
        if (on(n + i)) acc -= WV(z, n + i);
        WV(rd, i) = acc;
      }
      for (int r = 0; r < mx; ++r) {   // + Gam^T (z_up - z_lo)
        const double zz = (on(2 * n + r) ? WV(z, 2 * n + r) : 0.0) - (on(2 * n + mx + r) ? WV(z, 2 * n + mx + r) : 0.0);
        if (zz != 0.0)
          for (int i = 0; i < n; ++i) WV(rd, i) += a.Gam[(size_t)r * n + i] * zz;
      }
      for (int i = 0; i < n; ++i) e_d = dmax(e_d, dabs(WV(rd, i)));
      if (e_p <= a.tol * scale && e_d <= a.tol * scale && mu <= a.tol * scale) { status = 1; break; }
      // M = H + G^T diag(z / s) G, Cholesky
      for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) WV(M, (size_t)i * n + j) = a.H[(size_t)i * n + j];
      for (int i = 0; i < n; ++i) {
        double dd = 0.0;
        if (on(i)) dd += WV(z, i) / WV(s, i);
        if (on(n + i)) dd += WV(z, n + i) / WV(s, n + i);
        WV(M, (size_t)i * n + i) += dd;
      }
      for (int r = 0; r < mx; ++r) {
        double dd = 0.0;
        if (on(2 * n + r)) dd += WV(z, 2 * n + r) / WV(s, 2 * n + r);
        if (on(2 * n + mx + r)) dd += WV(z, 2 * n + mx + r) / WV(s, 2 * n + mx + r);
        if (dd == 0.0) continue;
        for (int i = 0; i < n; ++i) {
          const double gi = a.Gam[(size_t)r * n + i] * dd;
          if (gi == 0.0) continue;
          for (int j = 0; j <= i; ++j) WV(M, (size_t)i * n + j) += gi * a.Gam[(size_t)r * n + j];
        }
      }
      bool pd = true;
      for (int j = 0; j < n && pd; ++j) {
        double dj = WV(M, (size_t)j * n + j);
        for (int l = 0; l < j; ++l) dj -= WV(M, (size_t)j * n + l) * WV(M, (size_t)j * n + l);
        if (!(dj > 0.0)) { pd = false; break; }
        dj = sqrt(dj);
        WV(M, (size_t)j * n + j) = dj;
        for (int i = j + 1; i < n; ++i) {
          double sacc = WV(M, (size_t)i * n + j);
          for (int l = 0; l < j; ++l) sacc -= WV(M, (size_t)i * n + l) * WV(M, (size_t)j * n + l);
          WV(M, (size_t)i * n + j) = sacc / dj;
        }
      }
      if (!pd) { status = -1; break; }
      // two solves: predictor (sigma = 0, no correction), then corrector with Mehrotra's sigma and second-order term
      double sigma = 0.0;
      for (int pass = 0; pass < 2; ++pass) {
        // Delta z_r = -z_r + (sigma mu - corr_r) / s_r + (z_r / s_r) (rp_r + (G du)_r);  t_r := that without the G du part
        for (int i = 0; i < n; ++i) WV(rhs, i) = -WV(rd, i);
        for (int r = 0; r < m; ++r) {
          if (!on(r)) continue;
          const double corr = pass == 1 ? WV(ds, r) * WV(dz, r) : 0.0;    // predictor's ds * dz
          const double t = -WV(z, r) + (sigma * mu - corr) / WV(s, r) + WV(z, r) / WV(s, r) * WV(rp, r);
          WV(hh, r) = t;                                                   // hh is free until the next iteration
          if (r < n) WV(rhs, r) -= t;
          else if (r < 2 * n) WV(rhs, r - n) += t;
        }
        for (int r = 0; r < mx; ++r) {
          const double tt = (on(2 * n + r) ? WV(hh, 2 * n + r) : 0.0) - (on(2 * n + mx + r) ? WV(hh, 2 * n + mx + r) : 0.0);
          if (tt != 0.0)
            for (int i = 0; i < n; ++i) WV(rhs, i) -= a.Gam[(size_t)r * n + i] * tt;
        }
        for (int i = 0; i < n; ++i) {   // L y = rhs
          double sacc = WV(rhs, i);
          for (int l = 0; l < i; ++l) sacc -= WV(M, (size_t)i * n + l) * WV(du, l);
          WV(du, i) = sacc / WV(M, (size_t)i * n + i);
        }
        for (int i = n - 1; i >= 0; --i) {   // L^T du = y
          double sacc = WV(du, i);
          for (int l = i + 1; l < n; ++l) sacc -= WV(M, (size_t)l * n + i) * WV(du, l);
          WV(du, i) = sacc / WV(M, (size_t)i * n + i);
        }
        gam_u(du, Gu);
        double ap = 1.0, ad = 1.0, mu_aff = 0.0;
        for (int r = 0; r < m; ++r) {
          if (!on(r)) continue;
          const double gdu = row_dot(r, du, Gu);
          const double dsr = -WV(rp, r) - gdu;
          const double dzr = WV(hh, r) + WV(z, r) / WV(s, r) * gdu;
          WV(ds, r) = dsr; WV(dz, r) = dzr;
          if (dsr < 0.0) ap = dmin(ap, -WV(s, r) / dsr);
          if (dzr < 0.0) ad = dmin(ad, -WV(z, r) / dzr);
        }
        if (pass == 0) {
          for (int r = 0; r < m; ++r)
            if (on(r)) mu_aff += (WV(s, r) + ap * WV(ds, r)) * (WV(z, r) + ad * WV(dz, r));
          mu_aff = n_on > 0 ? mu_aff / n_on : 0.0;
          const double ratio = mu > 0.0 ? mu_aff / mu : 0.0;
          sigma = ratio * ratio * ratio;
          if (n_on == 0) { ap = ad = 1.0; pass = 1; }   // unconstrained: the Newton step is the solution
          else continue;
        }
        const double tau = 0.995;
        ap = dmin(1.0, tau * ap); ad = dmin(1.0, tau * ad);
        if (n_on == 0) { ap = 1.0; ad = 1.0; }
        for (int i = 0; i < n; ++i) WV(u, i) += ap * WV(du, i);
        for (int r = 0; r < m; ++r)
          if (on(r)) { WV(s, r) += ap * WV(ds, r); WV(z, r) += ad * WV(dz, r); }
      }
      for (int r = 0; r < m; ++r) WV(hh, r) = bound(r);   // restore h
    }
    gam_u(u, Gu);
    for (int i = 0; i < n; ++i) a.u[(size_t)b * n + i] = WV(u, i);
    if (a.xpred)
      for (int r = 0; r < mx; ++r) a.xpred[(size_t)b * mx + r] = WV(d, r) + WV(Gu, r);
    a.status[b] = status;
    a.iters[b] = it;
#undef WV
  }
};

}  // namespace hb
