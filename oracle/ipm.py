"""Oracle (test infrastructure): sparse primal-dual interior-point NLP solver.

Stands in for the ``ca.nlpsol('solver', 'ipopt', ...)`` object of the reference
(``hilo_mpc/modules/controller/mpc.py:1778-1781``, ``hilo_mpc/modules/estimator/mhe.py:782-784``).  IPOPT (+MUMPS) is
an un-vendored third-party dependency (``setup.py:52``: ``casadi>=3.5``, no upper pin) that is not installed here, so
its published algorithm (Waechter & Biegler 2006, "On the implementation of an interior-point filter line-search
algorithm for large-scale nonlinear programming") is restated with the IPOPT default constants that matter for
*where* it converges: ``tol=1e-8``, ``mu_init=0.1``, monotone barrier update (``kappa_mu=0.2``, ``theta_mu=1.5``,
``kappa_eps=10``), fraction-to-boundary ``tau_min=0.99``, ``bound_push=bound_frac=1e-2``, fixed variables removed
(``fixed_variable_treatment=make_parameter``).  The globalisation is a backtracking l1-merit line search instead of
the filter; both accept full Newton steps near a solution, so the KKT point reached from a given start is the same
unless the problem has several nearby local minima (tests flag that case explicitly).

Problem form:  min f(v)  s.t.  g(v) = 0,  lb <= v <= ub   (inequalities are slacked by the caller).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class IpmResult(dict):
    __getattr__ = dict.get


def solve_nlp(callback, v0, lb, ub, tol=1e-8, max_iter=3000, mu_init=.1, acceptable_tol=1e-6,
              acceptable_iter=15, verbose=False, mu_min=None, lam0=None, psd_hessian=False,
              finish_barrier=False):
    """``callback(v, lam) -> f, grad_f (n,), g (m,), J (m x n sparse), W (n x n sparse Hessian of f + lam^T g)``.

    ``finish_barrier=True`` additionally drives the barrier parameter down to ``mu_min`` before declaring optimality
    (IPOPT stops as soon as the unperturbed KKT error is below ``tol``, whatever mu is; components held by weakly
    active bounds then carry an O(mu/z) offset that depends on the final mu — tests comparing two solvers use this
    switch so that both stop on the same central-path point).

    Returns IpmResult(x, f, lam, z_l, z_u, status, iterations, kkt) with ``status`` mapped like
    ``hilo_mpc/modules/optimizer.py:1091-1104``: 1 optimal, 2 acceptable, 5 max-iter, -1 failure.
    """
    v0 = np.asarray(v0, dtype=float).copy()
    lb = np.asarray(lb, dtype=float)
    ub = np.asarray(ub, dtype=float)
    n_full = v0.size
    fixed = lb == ub
    free = ~fixed
    v_full = v0.copy()
    v_full[fixed] = lb[fixed]
    idx = np.flatnonzero(free)
    n = idx.size
    lbf, ubf = lb[idx], ub[idx]
    has_l = np.isfinite(lbf)
    has_u = np.isfinite(ubf)

    # IPOPT initial point: push into the interior (bound_push = bound_frac = 1e-2)
    x = v_full[idx].copy()
    k1 = k2 = 1e-2
    pl = np.where(has_l & has_u, np.minimum(k1 * np.maximum(1., np.abs(lbf)), k2 * (ubf - lbf)),
                  k1 * np.maximum(1., np.abs(lbf)))
    pu = np.where(has_l & has_u, np.minimum(k1 * np.maximum(1., np.abs(ubf)), k2 * (ubf - lbf)),
                  k1 * np.maximum(1., np.abs(ubf)))
    x = np.where(has_l, np.maximum(x, lbf + pl), x)
    x = np.where(has_u, np.minimum(x, ubf - pu), x)

    def evaluate(xf, lam):
        v = v_full.copy()
        v[idx] = xf
        f, gf, g, J, W = callback(v, lam)
        J = sp.csr_matrix(J)[:, idx]
        W = None if W is None else sp.csr_matrix(W)[idx][:, idx]
        return f, np.asarray(gf)[idx], np.asarray(g, dtype=float), J, W

    m = evaluate(x, None)[2].size
    lam = np.zeros(m) if lam0 is None else np.asarray(lam0, dtype=float).copy()
    z_l = np.where(has_l, 1., 0.)
    z_u = np.where(has_u, 1., 0.)
    mu = mu_init
    if mu_min is None:
        mu_min = tol / 10.
    nu = 0.
    filt = []
    filt_mu = None
    theta_max = theta_min = None
    delta_w_last = 0.
    n_acc = 0
    status = 5
    s_max = 100.
    it = 0
    kkt = np.inf
    dbg = {}

    def slacks(xf):
        sl = np.where(has_l, xf - np.where(has_l, lbf, 0.), 1.)
        su = np.where(has_u, np.where(has_u, ubf, 0.) - xf, 1.)
        return sl, su

    def errors(gf, g, J, lam, z_l, z_u, sl, su, mu_):
        nz = has_l.sum() + has_u.sum()
        s_d = max(s_max, (np.abs(lam).sum() + z_l.sum() + z_u.sum()) / max(1, m + nz)) / s_max
        s_c = max(s_max, (z_l.sum() + z_u.sum()) / max(1, nz)) / s_max
        dual = np.abs(gf + J.T @ lam - z_l + z_u).max(initial=0.) / s_d
        prim = np.abs(g).max(initial=0.)
        comp = max(np.abs(np.where(has_l, sl * z_l - mu_, 0.)).max(initial=0.),
                   np.abs(np.where(has_u, su * z_u - mu_, 0.)).max(initial=0.)) / s_c
        return max(dual, prim, comp), dual, prim, comp

    f, gf, g, J, W = evaluate(x, lam)
    # least-squares multiplier initialisation (IPOPT: lambda from min ||grad f + J^T lam - z||)
    if lam0 is None and m > 0:
        try:
            K = sp.bmat([[sp.identity(n), J.T], [J, None]], format='csc')
            sol = spla.splu(K).solve(np.concatenate([-(gf - z_l + z_u), np.zeros(m)]))
            cand = sol[n:]
            if np.all(np.isfinite(cand)) and np.abs(cand).max(initial=0.) <= 1e3:
                lam = cand
        except RuntimeError:
            pass
        f, gf, g, J, W = evaluate(x, lam)

    for it in range(max_iter + 1):
        sl, su = slacks(x)
        e0, dual, prim, comp = errors(gf, g, J, lam, z_l, z_u, sl, su, 0.)
        kkt = e0
        if verbose:
            print(f"it {it:3d} f={f:.8e} prim={prim:.2e} dual={dual:.2e} comp={comp:.2e} mu={mu:.1e} "
                  f"alpha={dbg.get('alpha', 0):.2e} a_pr={dbg.get('a_pr', 0):.2e} a_du={dbg.get('a_du', 0):.2e} "
                  f"dw={dbg.get('dw', 0):.1e} nu={nu:.2e}")
        if e0 <= tol and (not finish_barrier or mu <= mu_min):
            status = 1
            break
        if e0 <= acceptable_tol:
            n_acc += 1
            if n_acc >= acceptable_iter:
                status = 2
                break
        else:
            n_acc = 0
        if it == max_iter:
            status = 5
            break
        # barrier parameter update (monotone Fiacco-McCormick)
        while True:
            emu = errors(gf, g, J, lam, z_l, z_u, sl, su, mu)[0]
            if emu <= 10. * mu and mu > mu_min:
                mu = max(mu_min, min(.2 * mu, mu ** 1.5))
            else:
                break
        tau = max(.99, 1. - mu)

        sig = np.where(has_l, z_l / sl, 0.) + np.where(has_u, z_u / su, 0.)
        rhs_x = -(gf + J.T @ lam - np.where(has_l, mu / sl, 0.) + np.where(has_u, mu / su, 0.))
        rhs = np.concatenate([rhs_x, -g])
        Wm = W if W is not None else sp.csr_matrix((n, n))
        delta_w = 0.
        dx = dlam = None
        for attempt in range(40):
            # IPOPT's inertia correction, with the inertia test replaced by a curvature test along the step
            # (no sparse LDL^T with inertia in SciPy): regularise until the step has positive curvature.
            K = sp.bmat([[Wm + sp.diags(sig + delta_w), J.T], [J, -1e-14 * sp.identity(m)]], format='csc')
            ok = True
            try:
                sol = spla.splu(K).solve(rhs)
                dx, dlam = sol[:n], sol[n:]
                ok = bool(np.all(np.isfinite(sol)))
            except RuntimeError:
                ok = False
            if ok and not psd_hessian:
                curv = dx @ (Wm @ dx) + dx @ ((sig + delta_w) * dx)
                ok = curv > 1e-14 * (dx @ dx) or (dx @ dx) == 0.
            if ok:
                break
            if delta_w == 0.:
                delta_w = 1e-4 if delta_w_last == 0. else max(1e-20, delta_w_last / 3.)
            else:
                delta_w *= 8. if delta_w_last > 0. else 100.
        delta_w_last = delta_w
        dz_l = np.where(has_l, mu / sl - z_l - z_l / sl * dx, 0.)
        dz_u = np.where(has_u, mu / su - z_u + z_u / su * dx, 0.)

        def max_step(val, dval, t):
            neg = dval < 0.
            return min(1., (t * val[neg] / -dval[neg]).min(initial=1.))
        a_pr = min(max_step(sl[has_l], dx[has_l], tau), max_step(su[has_u], -dx[has_u], tau))
        a_du = min(max_step(z_l[has_l], dz_l[has_l], tau), max_step(z_u[has_u], dz_u[has_u], tau))

        # ---- filter line search (Waechter & Biegler 2006, Alg. A, steps A-5.x; IPOPT constants)
        def barrier(fv, xf):
            s1, s2 = slacks(xf)
            if (s1[has_l] <= 0.).any() or (s2[has_u] <= 0.).any():
                return np.inf
            return fv - mu * (np.log(s1[has_l]).sum() + np.log(s2[has_u]).sum())
        theta0 = np.abs(g).sum()
        phi0 = barrier(f, x)
        dphi = (gf - np.where(has_l, mu / sl, 0.) + np.where(has_u, mu / su, 0.)) @ dx
        if filt_mu != mu:
            filt.clear()
            filt_mu = mu
        if theta_max is None:
            theta_max = 1e4 * max(1., theta0)
            theta_min = 1e-4 * max(1., theta0)
        g_theta, g_phi, eta_phi, s_th, s_ph, delta = 1e-5, 1e-5, 1e-8, 1.1, 2.3, 1.
        alpha = a_pr
        accepted = False
        eps_m = 10. * np.finfo(float).eps
        for ls in range(40):
            xt = x + alpha * dx
            ft, gft, gt, Jt, Wt = evaluate(xt, lam + alpha * dlam)
            tht = np.abs(gt).sum()
            pht = barrier(ft, xt)
            ok = np.isfinite(pht) and np.isfinite(tht) and tht <= theta_max
            if ok:
                ok = not any(tht >= tf_ and pht >= pf_ for tf_, pf_ in filt)
            if ok:
                switching = dphi < 0. and alpha * (-dphi) ** s_ph > delta * theta0 ** s_th
                if theta0 <= theta_min and switching:
                    ok = pht <= phi0 + eta_phi * alpha * dphi + eps_m * abs(phi0)
                    augment = False
                else:
                    ok = tht <= (1. - g_theta) * theta0 or pht <= phi0 - g_phi * theta0 + eps_m * abs(phi0)
                    augment = True
            if ok:
                accepted = True
                if augment:
                    filt.append(((1. - g_theta) * theta0, phi0 - g_phi * theta0))
                break
            alpha *= .5
        if not accepted:
            # restoration phase is not implemented: take the tiny step; the iteration cap reports failure
            xt = x + alpha * dx
            ft, gft, gt, Jt, Wt = evaluate(xt, lam + alpha * dlam)
            if not np.isfinite(ft):
                status = -1
                break
        dbg = dict(alpha=alpha, a_pr=a_pr, a_du=a_du, dw=delta_w)
        x = xt
        lam = lam + alpha * dlam
        z_l = z_l + a_du * dz_l
        z_u = z_u + a_du * dz_u
        # IPOPT eq. (16): keep z within [mu/(k s), k mu/s], k = 1e10
        sl, su = slacks(x)
        z_l = np.where(has_l, np.clip(z_l, mu / (1e10 * sl), 1e10 * mu / sl), 0.)
        z_u = np.where(has_u, np.clip(z_u, mu / (1e10 * su), 1e10 * mu / su), 0.)
        f, gf, g, J, W = ft, gft, gt, Jt, Wt

    v = v_full.copy()
    v[idx] = x
    zl_full = np.zeros(n_full)
    zu_full = np.zeros(n_full)
    zl_full[idx] = z_l
    zu_full[idx] = z_u
    return IpmResult(x=v, f=float(f), lam=lam, z_l=zl_full, z_u=zu_full, status=status, iterations=it, kkt=float(kkt),
                     mu=mu)
