"""Cost and constraint holders with the reference's builder API.

``QuadraticCost`` mirrors ``hilo_mpc/util/modeling.py:89-450`` (``add_states / add_inputs / add_measurements /
add_inputs_change`` with ``weights``, ``ref``, ``path_following``, ``trajectory_tracking``), ``GenericCost`` mirrors
``modeling.py:35-86`` (``cost`` property).  Besides the scalar cost expression, every quadratic term keeps its residual
and weight so that the CUDA generator can emit a Gauss-Newton Hessian (exact for affine residuals).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from . import symbolic as sx


def _as_list(v):
    if v is None:
        return None
    if isinstance(v, str):
        return [v]
    if isinstance(v, (int, float)):
        return [v]
    return list(v)


def create_weight_matrix(arg, name: str = 'weights') -> np.ndarray:
    """modeling.py:165-185: square matrix as is, 1-D array / list -> diagonal, scalar -> 1x1."""
    if isinstance(arg, np.ndarray):
        if arg.ndim == 2 and arg.shape[0] == arg.shape[1]:
            return arg.astype(float)
        if arg.ndim == 1:
            return np.diag(arg.astype(float))
        raise TypeError(f"{name} must be a square matrix,a 1-D array or a list of real numbers.")
    if isinstance(arg, (list, tuple)):
        return np.diag(np.asarray(arg, dtype=float))
    if isinstance(arg, (float, int)):
        return np.diag([float(arg)])
    raise TypeError(f"The {name} must be a list of floats, numpy array or casadi DM.")


class GenericCost:
    def __init__(self, model):
        self._model = model
        self._cost = sx.Expr.const(0.)
        self._is_set = False

    @property
    def cost(self):
        return self._cost

    @cost.setter
    def cost(self, expr):
        # the reference accumulates when a path variable adds its own term (mpc.py:1199-1201)
        self._cost = self._cost + sx.Expr.wrap(expr) if self._is_set else sx.Expr.wrap(expr)
        self._is_set = True

    @property
    def is_set(self):
        return self._is_set


class QuadraticCost(GenericCost):
    def __init__(self, model):
        super().__init__(model)
        self._terms: List[dict] = []
        self._has_path_following = False
        self._has_trajectory_following = False

    # -- public builders ---------------------------------------------------------------------------------------
    def add_states(self, names, weights, ref=None, path_following=False, trajectory_tracking=False):
        self._add('states', names, weights, ref, path_following, trajectory_tracking)

    def add_inputs(self, names, weights, ref=None, path_following=False, trajectory_tracking=False):
        self._add('inputs', names, weights, ref, path_following, trajectory_tracking)

    def add_measurements(self, names, weights, ref=None, path_following=False, trajectory_tracking=False):
        self._add('measurements', names, weights, ref, path_following, trajectory_tracking)

    def add_inputs_change(self, names, weights):
        self._add('inputs_change', names, weights, None, False, False)

    def _add(self, kind, names, weights, ref, path_following, trajectory_tracking):
        names = _as_list(names)
        m = self._model
        pool = {'states': m.dynamical_state_names, 'inputs': m.input_names, 'inputs_change': m.input_names,
                'measurements': m.measurement_names}[kind]
        for n in names:
            if n not in pool:
                raise ValueError(f"The {kind.rstrip('s')} {n} does not exist. The available {kind} are {pool}")
        if weights is None:
            raise ValueError(f"You passed the following {kind}: {names} to the cost function, but I do not have any "
                             f"weights for it/them. Please pass me the weights.")
        W = create_weight_matrix(weights if not isinstance(weights, (list, tuple, np.ndarray)) or
                                 isinstance(weights, np.ndarray) else list(weights))
        if W.shape[0] != len(names):
            raise ValueError(f"{kind} and weights dimensions must be compatible. The {kind} vector you passed me is "
                             f"{len(names)} long while cost {W.shape}.")
        if trajectory_tracking and path_following:
            raise TypeError("You cannot have both trajectory tracking and path following activated. Please chose one "
                            "of them.")
        ref_is_sym = ref is not None and sx.is_symbolic(ref)
        if path_following and not ref_is_sym:
            raise TypeError(f"If path_following=True you need to pass an casadi SX or MX expression to the "
                            f"{kind}_ref  parameter.")
        if ref_is_sym and not (path_following or trajectory_tracking):
            raise ValueError("You passed an symbolic expression to the reference. In this case you need to tell me if "
                             "this is a path following or trajectory tracking problem. \n Please set either "
                             "'path_following=True' or trajectory_tracking=True.")
        if ref is not None and not ref_is_sym:
            ref = np.asarray(_as_list(ref) if not isinstance(ref, np.ndarray) else ref, dtype=float).reshape(-1)
            if ref.size != len(names):
                raise ValueError(f"{kind} and reference dimensions must be compatible. The states vector you passed "
                                 f"me is {len(names)} long while cost {ref.size}.")
        elif ref_is_sym:
            ref = sx.as_vector(ref)
            if len(ref) != len(names):
                raise ValueError(f"{kind} and reference dimensions must be compatible.")
        ind = [pool.index(n) for n in names]
        self._terms.append(dict(kind=kind, names=names, ind=ind, W=W, ref=ref, path_following=bool(path_following),
                                trajectory_tracking=bool(trajectory_tracking)))
        self._has_path_following |= bool(path_following)
        self._has_trajectory_following |= bool(trajectory_tracking)
        self._is_set = True

    @property
    def terms(self):
        return self._terms

    @property
    def n_of_trajectories(self):
        return sum(1 for t in self._terms if t['trajectory_tracking'])

    @property
    def name_varying_trajectories(self):
        return [n for t in self._terms if t['trajectory_tracking'] and t['ref'] is None for n in t['names']]
