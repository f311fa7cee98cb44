"""Oracle (test infrastructure): finite-horizon discrete LQR, restating ``hilo_mpc/modules/controller/lqr.py:236-244``.

``P_{k+1} = A^T P_k A - (A^T P_k B + N)(R + B^T P_k B)^{-1}(B^T P_k A + N^T) + Q`` for ``horizon`` steps starting from the
terminal weight ``P`` (default: ``Q``, lqr.py ``call``), then ``K = (R + B^T P B)^{-1}(B^T P A + N^T)``, ``u = -K x``.
Pinned by tests/test_LQR.py:314-342 (``K = [[1.39207671, 1.35221712, 0], [0, 0, .61802575]]``).
"""
import numpy as np


def lqr_gain(A, B, Q, R, horizon, P=None, N=None):
    A, B, Q, R = (np.atleast_2d(np.asarray(m, dtype=float)) for m in (A, B, Q, R))
    P = Q.copy() if P is None else np.atleast_2d(np.asarray(P, dtype=float))
    N = np.zeros((A.shape[0], B.shape[1])) if N is None else np.asarray(N, dtype=float)
    for _ in range(horizon):
        APBN = A.T @ P @ B + N
        RBPB = R + B.T @ P @ B
        P = A.T @ P @ A - np.linalg.solve(RBPB.T, APBN.T).T @ (B.T @ P @ A + N.T) + Q
    K = np.linalg.solve(R + B.T @ P @ B, B.T @ P @ A + N.T)
    return K, P
