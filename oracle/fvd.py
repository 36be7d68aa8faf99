"""Finite-volume known-answer oracle (TEST INFRASTRUCTURE ONLY, like the rest of oracle/).

CPU restatement of the reference's finite-volume eigen solvers, used as the accuracy oracle for BASELINE config 2
("2-D double-well potential, k = 2 eigenpairs, checked against utils/FVD-2d.py finite-volume eigenvalues"):

  * `fvd2d_matrix` / `fvd2d_eigs`  follow /root/reference/utils/FVD-2d.py:17-84 (matrix of the conjugated generator,
    one 5-point row per cell, symmetric) and :155-165 (Hermitian problem, LARGEST_REAL, k + 1 pairs).  The reference
    assembles a PETSc matrix row by row and solves with SLEPc; both are absent here, so the same stencil is assembled
    vectorised into a scipy CSR matrix and solved with scipy.sparse.linalg.eigsh (shift-invert at 0).
  * `fvd1d_eigs` follows /root/reference/utils/FVD-1d.py:37-58 (dense tridiagonal matrix, numpy eigh).
  * `potential` / `grad_potential` restate /root/reference/src/potentials.py (pot_id 1..7: :19-64 modified_dw1d, :66-92,
    :94-122, :124-187 three-well, :189-201, :204-244 dispatch) in vectorised numpy.
  * `sample_trajectory` restates the Euler-Maruyama generator /root/reference/src/data_generator.py:17-86 (single chain,
    10 000 burn-in steps, weights exp(-(beta - SDE_beta) V) normalised to mean 1) with an explicit seed (the
    reference never seeds).

Pinned in tests/test_fvd_oracle.py against (i) the reference's own potentials.PotClass / PrepareData when
/root/reference is present, (ii) the eigenvalues recorded in BASELINE.md section 3.
"""
from __future__ import annotations

import math

import numpy as np


# ---- potentials (src/potentials.py) -----------------------------------------------------------
def _modified_dw1d(x, a=1.0, b=1.0, c=0.7):
    x = np.asarray(x, dtype=np.float64)
    out = np.where(x > a, b * math.pi ** 2 / (4 * a * a) * (x - a) ** 2,
                   np.where(x < -a, b * math.pi ** 2 / (4 * a * a) * (x + a) ** 2,
                            b * 0.5 * (np.cos(x / a * math.pi) + 1.0)))
    return out - c * x


def _grad_modified_dw1d(x, a=1.0, b=1.0, c=0.7):
    x = np.asarray(x, dtype=np.float64)
    out = np.where(x > a, b * math.pi ** 2 / (2 * a * a) * (x - a),
                   np.where(x < -a, b * math.pi ** 2 / (2 * a * a) * (x + a),
                            -b * math.pi / (2 * a) * np.sin(x * math.pi / a)))
    return out - c


def _v_2d_3well(x, stiff_eps):
    theta = np.arctan2(x[:, 1], x[:, 0])
    r = np.sqrt(x[:, 0] ** 2 + x[:, 1] ** 2)
    v1 = np.zeros(len(x))
    hi, lo = theta > math.pi / 3, theta < -math.pi / 3
    mid = (theta > -math.pi / 3) & (theta < math.pi / 3)          # the boundary angles keep 0, as in the reference
    v1[hi] = (1 - (theta[hi] * 3 / math.pi - 1.0) ** 2) ** 2
    v1[lo] = (1 - (theta[lo] * 3 / math.pi + 1.0) ** 2) ** 2
    v1[mid] = 3.0 / 5.0 - 2.0 / 5.0 * np.cos(3 * theta[mid])
    return v1 + (r - 1) ** 2 / stiff_eps + 5.0 * np.exp(-5.0 * r ** 2)


def _grad_v_2d_3well(x, stiff_eps):
    theta = np.arctan2(x[:, 1], x[:, 0])
    r = np.sqrt(x[:, 0] ** 2 + x[:, 1] ** 2)
    dv = np.zeros(len(x))
    hi, lo = theta > math.pi / 3, theta < -math.pi / 3
    mid = (theta > -math.pi / 3) & (theta < math.pi / 3)
    dv[hi] = 12 / math.pi * (theta[hi] * 3 / math.pi - 1) * ((theta[hi] * 3 / math.pi - 1.0) ** 2 - 1)
    dv[lo] = 12 / math.pi * (theta[lo] * 3 / math.pi + 1) * ((theta[lo] * 3 / math.pi + 1.0) ** 2 - 1)
    dv[mid] = 1.2 * np.sin(3 * theta[mid])
    dr = 2.0 * (r - 1.0) / stiff_eps - 50.0 * r * np.exp(-r ** 2 / 0.2)
    return np.column_stack((-dv * x[:, 1] / (r * r) + dr * x[:, 0] / r, dv * x[:, 0] / (r * r) + dr * x[:, 1] / r))


def potential(x, pot_id, stiff_eps=0.5):
    """V(x) for x [n, dim] (PotClass.V, src/potentials.py:204-222)."""
    x = np.asarray(x, dtype=np.float64)
    if pot_id == 1:
        return 0.5 * x[:, 0] ** 2
    if pot_id == 2:
        return _modified_dw1d(x[:, 0], b=7.0, c=1.0)
    if pot_id == 3:
        return 0.5 * x[:, 0] ** 2 + 2.0 * x[:, 1] ** 2
    if pot_id == 4:
        return _modified_dw1d(x[:, 0]) + 0.5 * x[:, 1] ** 2
    if pot_id == 5:
        return _v_2d_3well(x, stiff_eps)
    if pot_id == 6:
        return (x[:, 0] ** 2 - 1.0) ** 2 + 1.0 / stiff_eps * (x[:, 0] ** 2 + x[:, 1] - 1.0) ** 2
    if pot_id == 7:
        return _v_2d_3well(x[:, 0:2], stiff_eps) + 5.0 * np.sum(x[:, 2:] ** 2, axis=1)
    raise ValueError("pot_id %r" % (pot_id,))


def grad_potential(x, pot_id, stiff_eps=0.5):
    """grad V(x) for x [n, dim] (PotClass.grad_V, src/potentials.py:224-244)."""
    x = np.asarray(x, dtype=np.float64)
    if pot_id == 1:
        return x[:, :1].copy()
    if pot_id == 2:
        return _grad_modified_dw1d(x[:, 0], b=7.0, c=1.0).reshape(-1, 1)
    if pot_id == 3:
        return np.column_stack((x[:, 0], 4.0 * x[:, 1]))
    if pot_id == 4:
        return np.column_stack((_grad_modified_dw1d(x[:, 0]), x[:, 1]))
    if pot_id == 5:
        return _grad_v_2d_3well(x, stiff_eps)
    if pot_id == 6:
        q = x[:, 0] ** 2 + x[:, 1] - 1.0
        return np.column_stack((4.0 * (x[:, 0] ** 2 - 1.0) * x[:, 0] + 4.0 / stiff_eps * q * x[:, 0], 2.0 / stiff_eps * q))
    if pot_id == 7:
        return np.concatenate((_grad_v_2d_3well(x[:, 0:2], stiff_eps), 10.0 * x[:, 2:]), axis=1)
    raise ValueError("pot_id %r" % (pot_id,))


# ---- finite volumes ----------------------------------------------------------------------------
def fvd2d_matrix(pot_id, beta=1.0, stiff_eps=0.5, xmin=-3.0, xmax=3.0, nx=500, ymin=-3.0, ymax=3.0, ny=300):
    """The matrix of utils/FVD-2d.py:17-84 as scipy CSR: cell Ii = j * nx + i with centre ((i+.5)dx, (j+.5)dy);
    off-diagonal towards a neighbour = exp(-beta (V_face - V_i)) / (beta h^2) * exp(-beta/2 (V_i - V_nb)),
    diagonal = -sum of the first factors (:33-35, :80-81)."""
    import scipy.sparse as sp
    dx, dy = (xmax - xmin) / nx, (ymax - ymin) / ny
    V = lambda X, Y: potential(np.column_stack((X.ravel(), Y.ravel())), pot_id, stiff_eps).reshape(X.shape)
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")          # [ny, nx]; flat index = j * nx + i
    xc, yc = (ii + 0.5) * dx + xmin, (jj + 0.5) * dy + ymin
    vi = V(xc, yc)
    idx = (jj * nx + ii)
    rows, cols, vals = [], [], []
    diag = np.zeros_like(vi)

    def side(mask, xf, yf, xn, yn, h, nb_idx):
        val = np.exp(-beta * (V(xf, yf) - vi)) / (beta * h * h)
        diag[mask] += val[mask]
        off = val * np.exp(-0.5 * beta * (vi - V(xn, yn)))
        rows.append(idx[mask]); cols.append(nb_idx[mask]); vals.append(off[mask])

    side(ii > 0, ii * dx + xmin, yc, (ii - 0.5) * dx + xmin, yc, dx, idx - 1)            # :37-47
    side(ii < nx - 1, (ii + 1.0) * dx + xmin, yc, (ii + 1.5) * dx + xmin, yc, dx, idx + 1)     # :48-58
    side(jj > 0, xc, jj * dy + ymin, xc, (jj - 0.5) * dy + ymin, dy, idx - nx)           # :59-69
    side(jj < ny - 1, xc, (jj + 1.0) * dy + ymin, xc, (jj + 1.5) * dy + ymin, dy, idx + nx)    # :70-80
    rows.append(idx.ravel()); cols.append(idx.ravel()); vals.append(-diag.ravel())
    n = nx * ny
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))


def fvd2d_eigs(pot_id, k, beta=1.0, stiff_eps=0.5, **grid):
    """The k + 1 eigenvalues of A closest to 0 from below (:155-165: HEP, LARGEST_REAL, k + 1 pairs), returned as the
    eigenvalues of -L in ascending order (the first is the trivial ~0)."""
    from scipy.sparse.linalg import eigsh
    A = fvd2d_matrix(pot_id, beta, stiff_eps, **grid)
    w = eigsh(A, k=k + 1, sigma=1e-6, which="LM", return_eigenvectors=False)
    return np.sort(-w)


def fvd1d_eigs(pot_id, k, beta=1.0, stiff_eps=0.5, xmin=-3.0, xmax=3.0, nx=500):
    """utils/FVD-1d.py:37-58: dense tridiagonal matrix of the conjugated generator, all eigenvalues by eigh,
    the first k + 1 returned (ascending; the first is the trivial ~0)."""
    dx = (xmax - xmin) / nx
    V = lambda x: potential(np.asarray(x, dtype=np.float64).reshape(-1, 1), pot_id, stiff_eps)
    xc = xmin + (np.arange(nx) + 0.5) * dx
    vc = V(xc)
    A = np.zeros((nx, nx))
    i = np.arange(1, nx)                                            # left neighbours (:40-46)
    vf = V(xmin + i * dx)
    A[i, i - 1] = -np.exp(beta * 0.5 * (vc[i - 1] + vc[i] - 2 * vf)) / (beta * dx ** 2)
    A[i, i] += np.exp(beta * (vc[i] - vf)) / (beta * dx ** 2)
    i = np.arange(0, nx - 1)                                        # right neighbours (:47-52)
    vf = V(xmin + (i + 1) * dx)
    A[i, i + 1] = -np.exp(beta * 0.5 * (vc[i + 1] + vc[i] - 2 * vf)) / (beta * dx ** 2)
    A[i, i] += np.exp(beta * (vc[i] - vf)) / (beta * dx ** 2)
    return np.linalg.eigh(A)[0][: k + 1]


# ---- trajectory data (src/data_generator.py:17-86) ---------------------------------------------
def sample_trajectory(pot_id, dim, K, delta_t, beta=1.0, sde_beta=None, stiff_eps=0.5, how_often=1, seed=0,
                      burn_step=10000):
    """Euler-Maruyama chain X <- X - grad V(X) dt + sqrt(2 dt / SDE_beta) xi (:53-69), every `how_often`-th state kept
    with weight exp(-(beta - SDE_beta) V) normalised to mean 1 (:72-82).  Returns (X [K // how_often, dim], w).
    The legacy global numpy RNG is seeded and consumed exactly as the reference does (randn(dim) per step), so with
    np.random.seed(seed) the reference's own generator produces the same chain."""
    sde_beta = beta if sde_beta is None else sde_beta
    how_often = max(int(how_often), 1)
    rs = np.random.RandomState(seed)
    x = rs.randn(dim)
    sig = math.sqrt(2 * delta_t / sde_beta)
    for _ in range(burn_step):
        xi = rs.randn(dim)
        x = x - grad_potential(x.reshape(1, dim), pot_id, stiff_eps)[0] * delta_t + sig * xi
    X = np.zeros((K // how_often, dim))
    w = np.zeros(K // how_often)
    kk = 0
    for i in range(K):
        xi = rs.randn(dim)
        x = x - grad_potential(x.reshape(1, dim), pot_id, stiff_eps)[0] * delta_t + sig * xi
        if i % how_often == 0 and kk < len(X):
            X[kk] = x
            kk += 1
    X = X[:kk]
    w = np.exp(-(beta - sde_beta) * potential(X, pot_id, stiff_eps))
    return X, w / w.mean()


def sample_dw2d_fast(K, delta_t, beta=1.0, how_often=1, seed=0, burn_step=10000):
    """The same chain as sample_trajectory(pot_id=4, dim=2, ...) with the double-well gradient written out in scalar
    Python (about 1 us per step instead of 30): used to make the 10^5-10^6-state data set of the known-answer run."""
    rs = np.random.RandomState(seed)
    x0, x1 = (float(v) for v in rs.randn(2))
    sig = math.sqrt(2 * delta_t / beta)
    n = burn_step + K
    xi = rs.randn(n, 2) * sig
    keep = max(int(how_often), 1)
    X = np.empty((K // keep + 1, 2))
    kk = 0
    pi, sin = math.pi, math.sin
    c2 = pi * pi / 2.0
    for s in range(n):
        if x0 > 1.0:
            g = c2 * (x0 - 1.0)
        elif x0 < -1.0:
            g = c2 * (x0 + 1.0)
        else:
            g = -pi / 2.0 * sin(x0 * pi)
        x0 = x0 - (g - 0.7) * delta_t + xi[s, 0]
        x1 = x1 - x1 * delta_t + xi[s, 1]
        if s >= burn_step and (s - burn_step) % keep == 0:
            X[kk, 0] = x0; X[kk, 1] = x1
            kk += 1
    X = X[:kk]
    return X, np.ones(kk)
