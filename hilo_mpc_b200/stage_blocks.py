"""Stage-QP block layout of the generated modules (``OcpT`` in ``csrc/ocp_kernels.cuh``), restated on the host: packs the
matrices of the block-banded KKT system into the ``qa`` / ``qh`` blocks the Riccati sweep reads (``hb_ocp_riccati``).

Two layouts (``hb_ocp_dims.frag``):
  * operand order of ``mma.sync.m8n8k4.f64`` for nx <= 4, nu <= 2, x0 fixed: ``qa`` = the 4 x 8 matrix ``[A B b 0]``
    column-major (entry (row l, column c) at ``c * 4 + l``), ``qh`` = the 6 x 8 tile ``[Q S^T q qbar; S R r rbar]`` row-major;
    unused rows / columns zero, ones on the diagonal of the R pad;
  * compact row-major ``[A | B | b]`` and ``[Q | S | R | q | r | qbar | rbar]``, padded to an even number of doubles.
"""
from __future__ import annotations

import numpy as np


def pack_stage_blocks(dims, A, Bm, b, Q, S, R, q, r, qbar=None, rbar=None):
    """``A [..., nx, nx]``, ``Bm [..., nx, nu]``, ``b [..., nx]``, ``Q [..., nx, nx]``, ``S [..., nu, nx]`` (d2L/du dx),
    ``R [..., nu, nu]``, ``q [..., nx]``, ``r [..., nu]``, multiplier-form gradients ``qbar``, ``rbar`` (default: q, r).
    Returns ``qa [..., block_a]``, ``qh [..., block_h]`` (leading dimensions = batch x stages)."""
    nx, nu = dims.nx, dims.nu
    A, Bm, b, Q, S, R, q, r = (np.asarray(v, dtype=float) for v in (A, Bm, b, Q, S, R, q, r))
    qbar = q if qbar is None else np.asarray(qbar, dtype=float)
    rbar = r if rbar is None else np.asarray(rbar, dtype=float)
    lead = A.shape[:-2]
    qa = np.zeros(lead + (dims.block_a,))
    qh = np.zeros(lead + (dims.block_h,))
    if dims.frag:
        M = np.zeros(lead + (8, 4))                     # [column][row]
        M[..., :nx, :nx] = np.swapaxes(A, -1, -2)
        M[..., 4:4 + nu, :nx] = np.swapaxes(Bm, -1, -2)
        M[..., 6, :nx] = b
        qa[...] = M.reshape(lead + (32,))
        T = np.zeros(lead + (6, 8))
        T[..., :nx, :nx] = Q
        T[..., :nx, 4:4 + nu] = np.swapaxes(S, -1, -2)
        T[..., 4:4 + nu, :nx] = S
        T[..., 4:4 + nu, 4:4 + nu] = R
        for c in range(nu, 2):
            T[..., 4 + c, 4 + c] = 1.
        T[..., :nx, 6], T[..., 4:4 + nu, 6] = q, r
        T[..., :nx, 7], T[..., 4:4 + nu, 7] = qbar, rbar
        qh[...] = T.reshape(lead + (48,))
    else:
        flat = lambda v: v.reshape(lead + (-1,))
        ca = np.concatenate([flat(A), flat(Bm), b], axis=-1)
        ch = np.concatenate([flat(Q), flat(S), flat(R), q, r, qbar, rbar], axis=-1)
        qa[..., :ca.shape[-1]] = ca
        qh[..., :ch.shape[-1]] = ch
    return qa, qh


def riccati_reference(A, Bm, b, Q, S, R, q, r, QN, qN):
    """Plain NumPy Riccati recursion for ONE instance (stages on the leading axis): returns ``dx [N+1, nx]``,
    ``du [N, nu]`` of the equality-constrained QP with dx_0 = 0 (host-side helper for tests and examples)."""
    N = A.shape[0]
    P, p = np.array(QN, dtype=float), np.array(qN, dtype=float)
    K, k = [None] * N, [None] * N
    for i in range(N - 1, -1, -1):
        Hxx = Q[i] + A[i].T @ P @ A[i]
        Hux = S[i] + Bm[i].T @ P @ A[i]
        Huu = R[i] + Bm[i].T @ P @ Bm[i]
        hx = q[i] + A[i].T @ (P @ b[i] + p)
        hu = r[i] + Bm[i].T @ (P @ b[i] + p)
        K[i] = -np.linalg.solve(Huu, Hux)
        k[i] = -np.linalg.solve(Huu, hu)
        P = Hxx + Hux.T @ K[i]
        P = .5 * (P + P.T)
        p = hx + Hux.T @ k[i]
    dx = np.zeros((N + 1, A.shape[1]))
    du = np.zeros((N, Bm.shape[2]))
    for i in range(N):
        du[i] = K[i] @ dx[i] + k[i]
        dx[i + 1] = A[i] @ dx[i] + Bm[i] @ du[i] + b[i]
    return dx, du
