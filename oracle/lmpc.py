"""Oracle (test infrastructure): the reference's LMPC QP, restated entry by entry, and solved on the CPU.

Follows ``hilo_mpc/modules/controller/mpc.py:2195-2266`` (``LMPC._setup``): decision vector ``v = [x_0..x_N | u_0..u_{N-1}]``
(stage-major), ``H = blkdiag(I_N (x) Q, P, I_N (x) R)``, objective ``1/2 v^T H v`` (CasADi conic convention, ``g = 0``),
equalities ``Aeq v = 0`` with ``Aeq = [kron(I_N, A) 0] + kron([0 -I_N], I_nx) | input block]``, bounds tiled, ``x_0`` pinned
through ``lbx = ubx`` (mpc.py:2361-2362).  The input block of the reference is ``kron(B, I_N)`` (mpc.py:2243, SURVEY.md
Appendix B-1); ``bug_compat=False`` builds the mathematically intended ``kron(I_N, B)`` instead.  Solved with the oracle's
interior-point method on the FLAT QP (no condensing), i.e. independently of the product's condensed formulation."""
import numpy as np
import scipy.sparse as sp

from .ipm import solve_nlp


def lmpc_flat_qp(A, B, Q, R, P, N, bug_compat=True):
    A, B, Q, R, P = (np.atleast_2d(np.asarray(v, dtype=float)) for v in (A, B, Q, R, P))
    nx, nu = B.shape
    Abar1 = np.hstack([np.kron(np.eye(N), A), np.zeros((N * nx, nx))])         # mpc.py:2209-2213
    aux2 = np.zeros((N, N + 1))
    for i in range(N):
        aux2[i, i + 1] = -1.                                                    # mpc.py:2228
    Abar2 = np.kron(aux2, np.eye(nx))                                           # mpc.py:2234
    Abar3 = np.kron(B, np.eye(N)) if bug_compat else np.kron(np.eye(N), B)      # mpc.py:2243 vs. the intended block
    Aeq = np.hstack([Abar1 + Abar2, Abar3])                                     # mpc.py:2245
    H = np.zeros(((N + 1) * nx + N * nu,) * 2)
    H[:N * nx, :N * nx] = np.kron(np.eye(N), Q)                                 # mpc.py:2251-2255
    H[N * nx:(N + 1) * nx, N * nx:(N + 1) * nx] = P
    H[(N + 1) * nx:, (N + 1) * nx:] = np.kron(np.eye(N), R)
    return H, Aeq


def solve_lmpc(A, B, Q, R, P, N, x0, x_lb, x_ub, u_lb, u_ub, bug_compat=True, tol=1e-12):
    """Returns ``x [N+1, nx]``, ``u [N, nu]`` of the reference's QP for the measured state ``x0``."""
    H, Aeq = lmpc_flat_qp(A, B, Q, R, P, N, bug_compat)
    nx, nu = np.atleast_2d(B).shape
    lb = np.concatenate([np.tile(x_lb, N + 1), np.tile(u_lb, N)]).astype(float)
    ub = np.concatenate([np.tile(x_ub, N + 1), np.tile(u_ub, N)]).astype(float)
    lb[:nx] = ub[:nx] = x0                                                      # mpc.py:2361-2362
    Hs, As = sp.csr_matrix(H), sp.csr_matrix(Aeq)

    def cb(v, lam):
        return .5 * v @ H @ v, H @ v, Aeq @ v, As, Hs
    v0 = np.zeros(H.shape[0])
    v0[:nx] = x0
    res = solve_nlp(cb, v0, lb, ub, tol=tol, finish_barrier=True, max_iter=500)
    if res.status != 1:
        raise RuntimeError(f"oracle LMPC solve failed with status {res.status}")
    v = res.x
    return v[:(N + 1) * nx].reshape(N + 1, nx), v[(N + 1) * nx:].reshape(N, nu)
