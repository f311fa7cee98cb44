"""``SimpleControlLoop`` (``hilo_mpc/modules/control_loop.py:41-431``) with the whole loop on the device.

The reference runs ``controller -> plant -> observer`` for ONE instance and a Python object per step
(``control_loop.py:343-398``: ``controller.optimize(x0, cp=...)``, ``plant.simulate(u=u, p=p)``, ``observer.estimate()``).
Here the plant state is a ``(B, n_x)`` device tensor: the controller is the batched engine (``NMPC`` / ``LMPC``), the plant
advances through the generated ``hb_sim_step`` kernel with per-instance inputs, an optional Kalman-type observer runs
beside it, and nothing returns to the host inside the loop except the controller's own convergence poll.

``run(steps)`` without batched initial states keeps the reference behaviour: the plant's stored initial condition is used,
the trajectory is written to ``plant.solution`` (and to ``controller.solution`` by ``optimize`` itself).
"""
from __future__ import annotations

import numpy as np

from .simulate import _engine


class SimpleControlLoop:
    """Feedback loop of a plant :class:`Model`, a controller with ``optimize(x0, cp=..., **refs)`` and an optional observer
    with ``estimate(y=, u=, p=)``.

    :param plant: the (set-up) plant model; its dynamical states must contain the controller model's states by name
    :param controller: NMPC / LMPC (anything with ``optimize``), or an ANN-type object with ``predict``
    :param observer: KF / EKF / UKF (optional).  As in the reference the controller is fed the PLANT state
        (``control_loop.py:428-431``); ``feedback='observer'`` closes the loop over the estimate instead.
    """

    def __init__(self, plant, controller, observer=None, feedback: str = 'plant') -> None:
        if not plant.is_setup():
            plant.setup()
        if hasattr(controller, 'is_setup') and not controller.is_setup():
            controller.setup()
        if observer is not None and not observer.is_setup():
            observer.setup()
        if feedback not in ('plant', 'observer'):
            raise ValueError("feedback must be 'plant' or 'observer'")
        if feedback == 'observer' and observer is None:
            raise ValueError("feedback='observer' needs an observer")
        self._plant, self._controller, self._observer, self._feedback = plant, controller, observer, feedback
        self._has_optimize = hasattr(controller, 'optimize')
        if not self._has_optimize and not hasattr(controller, 'predict'):
            raise TypeError("Controller must provide optimize(x0, ...) (MPC) or predict(x) (learned controller)")
        # states of the controller model may be a subset of the plant's (control_loop.py:353-355)
        names = plant.dynamical_state_names
        cmodel = getattr(controller, '_model_orig', None) or getattr(controller, '_model', None)
        cnames = cmodel.dynamical_state_names if cmodel is not None else names
        missing = [n for n in cnames if n not in names]
        if missing:
            raise ValueError(f"Controller states {missing} are not states of the plant")
        self._ind_states = [names.index(n) for n in cnames]
        self.x_history = self.u_history = self.y_history = self.x_estimate_history = self.status_history = None

    # ------------------------------------------------------------------------------------------------------------
    def _control(self, x, p, kwargs):
        import torch
        xc = x[:, self._ind_states].contiguous()
        if self._has_optimize:
            cp = p
            tvp_names = getattr(self._controller, '_tvp_names', None) or []
            if p is not None and len(tvp_names) > 0:   # control_loop.py:357-362
                keep = [i for i, n in enumerate(self._plant.parameter_names) if n not in tvp_names]
                cp = [np.asarray(p, dtype=float).reshape(-1)[i] for i in keep]
            u = self._controller.optimize(xc, cp=cp, **kwargs)
        else:
            u = self._controller.predict(xc)
        if not torch.is_tensor(u):
            u = torch.as_tensor(np.asarray(u, dtype=float), dtype=torch.float64, device=x.device)
        return u.reshape(x.shape[0], -1).contiguous()

    def run(self, steps: int, p=None, x0=None, disturbance=None, **kwargs):
        """Run ``steps`` closed-loop steps.

        :param x0: ``(B, n_x)`` initial plant states (array or device tensor) for a batched run; ``None`` = the plant's
            stored initial condition, one instance, results also written to ``plant.solution`` like the reference
        :param p: plant parameters (constant)
        :param disturbance: optional callable ``(k, x) -> dx`` added to the plant state after each step (device tensor)
        :returns: dict with the trajectories ``x (steps+1, B, n_x)``, ``u (steps, B, n_u)``, ``y``, ``status`` (device)
        """
        from . import _lib
        torch = _lib.require_cuda()
        plant = self._plant
        eng = _engine(plant)
        eng.open()
        nx, nu, npar, ny = plant.n_x, plant.n_u, plant.n_p, plant.n_y
        single = x0 is None
        if single:
            if plant.solution.is_empty('x'):
                raise RuntimeError("No initial dynamical states found. Please set initial conditions of the plant "
                                   "before running the control loop!")
            x = eng.tensor(plant.solution['x:f'].reshape(-1), (1, nx))
            t = float(plant.solution['t:f'].reshape(-1)[0]) if not plant.solution.is_empty('t') else plant._t0
        else:
            x = eng.tensor(x0).reshape(-1, nx).clone()
            t = plant._t0
        B = x.shape[0]
        if p is None and npar > 0:
            p = plant._p_values
            if p is None:
                raise RuntimeError(f"The plant has {npar} parameter(s) but no values were supplied")
        pt = None if npar == 0 else eng.tensor(p, (npar,))
        substeps = int(plant._solver_options.get('substeps', 10))
        dt = plant.dt_value
        y = torch.zeros(B, max(ny, 1), dtype=torch.float64, device=eng.device)
        xs = torch.empty(steps + 1, B, nx, dtype=torch.float64, device=eng.device)
        us = torch.empty(steps, B, max(nu, 1), dtype=torch.float64, device=eng.device)
        ys = torch.empty(steps, B, max(ny, 1), dtype=torch.float64, device=eng.device)
        xe = torch.empty(steps, B, nx, dtype=torch.float64, device=eng.device) if self._observer is not None else None
        status = []
        xs[0] = x
        x_fb = x
        for k in range(steps):
            u = self._control(x_fb, p, kwargs)
            st = getattr(self._controller, 'solver_status_code', None)
            if st is not None:
                status.append(st.clone() if torch.is_tensor(st) else torch.full((B,), int(st), device=eng.device))
            eng.call('sim', B, dt, x, u=u if nu else None, p=pt, y_pred=y if ny else None, substeps=substeps, t0=t)
            if disturbance is not None:
                x += disturbance(k, x)
            t += dt
            xs[k + 1] = x
            us[k] = u
            if ny:
                ys[k] = y
            if self._observer is not None:
                self._observer.estimate(y=y[:, :ny] if B > 1 else y[0, :ny], u=u if B > 1 else u[0], p=p)
                xe[k] = self._observer._x.reshape(B, -1)[:, :nx]   # device-resident estimate of the Kalman-type filters
            x_fb = xe[k] if self._feedback == 'observer' else x
            if single:   # keep the reference's bookkeeping for the one-instance loop
                plant.solution.update(t=[t], x=x[0].cpu().numpy())
                if nu:
                    plant.solution.add('u', u[0].cpu().numpy())
                if ny:
                    plant.solution.add('y', y[0, :ny].cpu().numpy())
        self.x_history, self.u_history, self.y_history, self.x_estimate_history = xs, us[:, :, :nu], ys[:, :, :ny], xe
        self.status_history = torch.stack(status) if status else None
        return {'x': xs, 'u': self.u_history, 'y': self.y_history, 'x_estimate': xe, 'status': self.status_history}
