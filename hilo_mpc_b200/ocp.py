"""Stage-structured optimal-control problem description produced by ``NMPC.setup()`` / ``MHE.setup()``.

This is the hand-over point between the API shell (which mirrors the symbolic part of the reference's
``NMPC._setup``, ``hilo_mpc/modules/controller/mpc.py:1133-1437``) and the two consumers of the problem:

* the CUDA code generator (``hilo_mpc_b200/codegen.py``), which turns the expressions into sm_100a device functions,
* the CPU oracle (``oracle/nmpc.py``, tests only), which restates the reference's flat NLP from the same expressions.

All expressions are in *scaled* coordinates (``mpc.py:248-263``, ``base.py:1562-1591``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from . import symbolic as sx


@dataclass
class OcpSpec:
    nx: int
    nu: int
    N: int
    # symbols
    t: sx.Expr
    dt: sx.Expr
    x: np.ndarray
    u: np.ndarray
    sp: np.ndarray  # per-stage parameters: [tvp_k | cp | stage references | ...]
    tp: np.ndarray  # terminal parameters (terminal references)
    uold: np.ndarray  # previous input placeholders entering the input-change cost (stage 0 only)
    # expressions
    ode: np.ndarray  # continuous: dx/dt ; discrete: x_{k+1}
    lag: sx.Expr  # Lagrange / stage cost l(t, x, u, sp)
    lag_du: sx.Expr  # input-change cost (u - uold)^T W (u - uold), active at stage 0 only
    may: sx.Expr  # Mayer / terminal cost E(t, x, tp)
    # least-squares structure of the cost (for the Gauss-Newton Hessian): lag = sum_i r_i^T W_i r_i + lag_rest
    lag_ls: List[Tuple[np.ndarray, np.ndarray]] = field(default_factory=list)
    lag_rest: Optional[sx.Expr] = None
    may_ls: List[Tuple[np.ndarray, np.ndarray]] = field(default_factory=list)
    may_rest: Optional[sx.Expr] = None
    # discretisation (optimizer.py:1410-1456)
    method: str = 'collocation'  # 'collocation' | 'erk' | 'discrete'
    degree: int = 3
    points: str = 'radau'
    order: int = 4
    objective: str = 'continuous'  # 'continuous' (quadrature of lag) | 'discrete' (lag(x_k,u_k), no dt factor)
    # bounds / guesses (scaled), each of length nx / nu
    x_lb: np.ndarray = None
    x_ub: np.ndarray = None
    u_lb: np.ndarray = None
    u_ub: np.ndarray = None
    x_guess: np.ndarray = None
    u_guess: np.ndarray = None
    # additive post-integration input x_{k+1} = F(x_k, ., sp) + gadd @ u_k  (MHE state noise, mhe.py:719-735)
    gadd: Optional[np.ndarray] = None
    # whether x_0 is pinned to the measured state (NMPC, mpc.py:797-807) or free (MHE)
    fixed_x0: bool = True
    # first stage index that carries the Lagrange term (MHE: stage cost only for k >= 1, mhe.py:738-743)
    lag_from: int = 0
    # arrival cost on (x_0) for MHE, in terms of x and ap (arrival parameters)
    arr: Optional[sx.Expr] = None
    arr_ls: List[Tuple[np.ndarray, np.ndarray]] = field(default_factory=list)
    ap: Optional[np.ndarray] = None
    # exact Hessian of the Lagrangian (constraint curvature through the collocation multipliers) like IPOPT; False =
    # Gauss-Newton (always used for the explicit RK transcription)
    exact_hessian: bool = True
    name: str = 'ocp'

    @property
    def nsp(self) -> int:
        return len(self.sp)

    @property
    def ntp(self) -> int:
        return len(self.tp)

    @property
    def nuold(self) -> int:
        return len(self.uold)

    @property
    def nap(self) -> int:
        return 0 if self.ap is None else len(self.ap)

    @property
    def ncol(self) -> int:
        return self.degree * self.nx if self.method == 'collocation' else 0

    @property
    def n_v(self) -> int:
        """Size of the reference's flat decision vector (mpc.py:1440-1453) for the supported feature set."""
        return (self.N + 1) * self.nx + self.N * self.nu + self.N * self.ncol

    @property
    def n_g(self) -> int:
        return self.N * (self.ncol + self.nx)
