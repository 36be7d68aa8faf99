"""CPU restatement (numpy fp64) of the reference's MD pre-processing layer and of its Jacobian.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference: `MD_data_set.align` (src/data_set.py:126-153) — Kabsch alignment of every configuration to a reference
configuration: centre on the selected atoms, H = sum_s y_s^T q_s, SVD H = U S V^T, rotation
R = U diag(1, 1, sign det(U V^T)) V^T, output (x - c) R.  The layer has no parameters and is applied to the raw
coordinates inside every net's forward (src/network_arch.py:62), so the input gradient the loss uses is
grad_x (F o A) = J^T grad F with J = dA/dx, and the energy of SURVEY 8 becomes u0^T M u0 with the per-sample metric
M = J diag(c) J^T (u0 = grad F at the aligned point).

The Jacobian is written in closed form (no autograd): with H = R P, P = R^T H symmetric,
dR = R [omega]_x  and  (tr(P) I - P) omega = axial(R^T dH - dH^T R)   (derivative of the orthogonal polar factor;
the sign matrix is locally constant, as in the reference where it is detached).
tests/test_oracle_golden.py checks it against torch.autograd on the reference-style SVD expression and against the
aligned minibatch stored by the reference run.
"""
from __future__ import annotations

import numpy as np


def kabsch_rotation(H):
    """R = U diag(1,1,sign det(U V^T)) V^T for H = U S V^T (numpy SVD, singular values descending)."""
    U, _, Vt = np.linalg.svd(H)
    D = np.eye(3)
    D[2, 2] = np.sign(np.linalg.det(U @ Vt))
    return U @ D @ Vt


def _skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def align_one(x, ref_x, ref_index, jacobian=True):
    """x [3N] raw coordinates -> (aligned [3N], J [3N, 3N] with J[o, i] = d aligned_o / d x_i)."""
    p = np.asarray(x, np.float64).reshape(-1, 3)
    N, m = p.shape[0], len(ref_index)
    q = np.asarray(ref_x, np.float64)
    c = p[ref_index].mean(axis=0)
    y = p[ref_index] - c
    H = y.T @ q
    R = kabsch_rotation(H)
    z = p - c
    out = (z @ R).reshape(-1)
    if not jacobian:
        return out, None
    P = R.T @ H
    T = np.trace(P) * np.eye(3) - P
    qc = q - q.mean(axis=0)
    J = np.zeros((3 * N, 3 * N))
    sel = {a: s for s, a in enumerate(ref_index)}
    for b in range(N):
        for i in range(3):
            col = np.zeros((N, 3))
            col[b] += R[i]                                   # d p_b = e_i  ->  e_i R
            if b in sel:
                col -= R[i][None, :] / m                     # - d c R for every atom
                dH = np.zeros((3, 3)); dH[i] = qc[sel[b]]    # e_i^T (q_b - mean q)
                A = R.T @ dH
                K = A - A.T
                omega = np.linalg.solve(T, np.array([K[2, 1], K[0, 2], K[1, 0]]))
                col += z @ (R @ _skew(omega))
            J[:, 3 * b + i] = col.reshape(-1)
    return out, J


def align_batch(X, ref_x, ref_index, diag_coeff=None):
    """Aligned coordinates [B, d] and, if diag_coeff is given, the metrics M[B, d, d] = J diag(c) J^T."""
    X = np.asarray(X, np.float64)
    outs, Ms = [], []
    for x in X:
        o, J = align_one(x, ref_x, list(ref_index), jacobian=diag_coeff is not None)
        outs.append(o)
        if diag_coeff is not None:
            Ms.append((J * np.asarray(diag_coeff, np.float64)[None, :]) @ J.T)
    return np.stack(outs), (np.stack(Ms) if diag_coeff is not None else None)
