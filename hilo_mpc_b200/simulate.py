"""``Model.simulate`` on the GPU (``hilo_mpc/modules/dynamic_model/dynamic_model.py:3911-4212``).

The stored state advances by ``steps`` sampling intervals through the generated ``hb_sim_step`` kernel (discrete map or
RK4 sub-stepping; ``solver_options={'substeps': n}``, default 10).  ``u``/``p`` follow the reference's conventions:
scalars / vectors are held constant, ``(n_u, steps)`` arrays give one column per step.
"""
from __future__ import annotations

import numpy as np

from .estgen import EstEngine


def _engine(model) -> EstEngine:
    if model._sim_module is None:
        model._sim_module = EstEngine(model)
    return model._sim_module


def simulate_model(model, u=None, p=None, steps: int = 1, **kwargs):
    from . import _lib
    torch = _lib.require_cuda()
    eng = _engine(model)
    eng.open()
    nx, nu, npar, ny = model.n_x, model.n_u, model.n_p, model.n_y
    if nu > 0 and u is None:
        raise RuntimeError(f"The model has {nu} input(s) but no input was supplied (simulate(u=...))")
    if p is None and npar > 0:
        p = model._p_values
        if p is None:
            raise RuntimeError(f"The model has {npar} parameter(s) but no values were supplied")
    u_arr = None if nu == 0 else np.asarray(u.cpu() if torch.is_tensor(u) else u, dtype=float)
    if u_arr is not None:
        u_arr = u_arr.reshape(nu, -1) if u_arr.size != nu else u_arr.reshape(nu, 1)
        if u_arr.shape[1] not in (1, steps):
            steps = u_arr.shape[1]
    substeps = int(model._solver_options.get('substeps', 10))
    x = eng.tensor(model.solution['x:f'].reshape(-1), (1, nx))
    y = torch.zeros(1, max(ny, 1), dtype=torch.float64, device=eng.device)
    t = float(model.solution['t:f'].reshape(-1)[0]) if not model.solution.is_empty('t') else model._t0
    dt = model.dt_value
    pt = None if npar == 0 else eng.tensor(p, (npar,))
    for k in range(steps):
        uk = None if nu == 0 else eng.tensor(u_arr[:, min(k, u_arr.shape[1] - 1)], (nu,))
        eng.call('sim', 1, dt, x, u=uk, p=pt, y_pred=y if ny else None, substeps=substeps, t0=t)
        t += dt
        xk = x[0].cpu().numpy()
        model.solution.update(t=[t], x=xk)
        if nu:
            model.solution.add('u', uk.cpu().numpy())
        if ny:
            model.solution.add('y', y[0, :ny].cpu().numpy())
    return model.solution['x:f']
