#!/usr/bin/env python
"""Micro-benchmark of the estimator kernels (north star: EKF / UKF / PF predict-update): B = 16384 filters on the CSTR
(nx = 4, ny = 2, nu = 2, 1 parameter), continuous model with 10 RK4 sub-steps per sampling interval.

Prints one JSON object: per filter type the time per batched step (CUDA events, L2 flushed by a 256 MB write between
steps), filter steps/s, model propagations/s (UKF: 2 nx + 1 sigma points x sub-steps x 4 RK stages; PF: particles x ...)
and the algorithmic HBM bytes per step / achieved GB/s (UKF reads x, P, Q, R, u, y and writes x, P, y_pred)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from hilo_mpc_b200 import EKF, PF, UKF
    from hilo_mpc_b200.library import cstr_model, cstr_steady_state
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    reps = 20
    dev = torch.device('cuda', 0)
    x_ss, u_ss = cstr_steady_state()
    rng = np.random.default_rng(0)
    x0 = x_ss * (1. + .01 * rng.standard_normal((B, 4)))
    y = torch.tensor(x0[:, 1:3] + np.array([.01, .1]) * rng.standard_normal((B, 2)), device=dev)
    u = torch.tensor(np.tile(u_ss, (B, 1)), device=dev)
    p = torch.ones(B, 1, dtype=torch.float64, device=dev)
    flush = torch.empty(32 * 1024 * 1024, dtype=torch.float64, device=dev)
    out = {'B': B, 'model': 'CSTR nx=4 ny=2 nu=2 np=1, continuous, 10 RK4 sub-steps per interval'}
    nx, ny, sub = 4, 2, 10

    def run(step):
        for _ in range(3):
            step()
        ms = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return float(np.median(ms))

    def make(cls, **kw):
        f = cls(cstr_model(dt=.005, name='cstr_k1', with_parameter=True), **kw)
        f.setup()
        f.Q = np.diag((1e-2 * x_ss) ** 2 * .005)
        f.R = np.diag([1e-4, 1e-2])
        f.set_initial_guess(x0, P0=np.diag([1e-2, 1e-2, 1., 1.]))
        return f
    ekf = make(EKF)
    ms = run(lambda: ekf.estimate(y=y, u=u, p=p))
    byts = B * 8 * (nx + nx * nx + nx + nx * nx + ny + ny + 2 + 1)
    out['ekf'] = {'ms_per_step': ms, 'filter_steps_per_s': B / ms * 1e3,
                  'rhs_evaluations_per_s': B * sub * 4 / ms * 1e3, 'algorithmic_bytes': byts, 'gb_per_s': byts / ms * 1e-6}
    ukf = make(UKF, alpha=.5)
    ms = run(lambda: ukf.estimate(y=y, u=u, p=p))
    ns = 2 * nx + 1
    byts = B * 8 * (nx + nx * nx + nx + nx * nx + ny + ny + 2 + 1 + 2 * ns * nx)
    out['ukf'] = {'ms_per_step': ms, 'filter_steps_per_s': B / ms * 1e3, 'sigma_point_propagations_per_s': B * ns / ms * 1e3,
                  'rhs_evaluations_per_s': B * ns * sub * 4 / ms * 1e3, 'algorithmic_bytes': byts, 'gb_per_s': byts / ms * 1e-6}
    n_s = 64
    # PF on the scalar benchmark model of the reference's tests (n_y = 1: the reference's likelihood is univariate)
    from hilo_mpc_b200.model import Model
    sm = Model(discrete=True)
    sm.set_dynamical_states('x')
    sm.set_measurements('y')
    sm.set_dynamical_equations('x/2 + 25*dt*x/(1 + x^2)')
    sm.set_measurement_equations('x^2/20')
    sm.setup(dt=1.)
    pf = PF(sm)
    pf.setup(n_samples=n_s)
    pf.Q, pf.R = 10., 1.
    pf.set_initial_guess(np.full((B, 1), 9.))
    Xp = 9. + rng.standard_normal((B, n_s, 1))
    pf.set_particles(Xp)
    w = torch.tensor(rng.normal(0., np.sqrt(10.), size=(B, n_s, 1)), device=dev)
    v = torch.tensor(rng.normal(0., 1., size=(B, n_s, 1)), device=dev)
    un = torch.tensor(rng.random((B, n_s)), device=dev)
    yp = torch.tensor(np.full((B, 1), 3.2), device=dev)
    ms = run(lambda: pf.estimate(y=yp, w=w, v=v, uniforms=un))
    pf.kernel_variant = 0
    ms_thread = run(lambda: pf.estimate(y=yp, w=w, v=v, uniforms=un))
    pf.kernel_variant = 1
    byts = B * n_s * 8 * (1 + 1 + 1 + 1 + 1 + 1)   # particles in/out, w, v, uniforms, weights (+ int64 indices)
    out['pf'] = {'ms_per_step': ms, 'ms_per_step_one_thread_per_filter': ms_thread, 'particles': n_s, 'filter_steps_per_s': B / ms * 1e3,
                 'particle_propagations_per_s': B * n_s / ms * 1e3, 'algorithmic_bytes': byts + B * n_s * 8,
                 'gb_per_s': (byts + B * n_s * 8) / ms * 1e-6, 'model': 'scalar benchmark model of tests/test_KFs.py:548-566'}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
