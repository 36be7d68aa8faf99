"""numpy fp64 closed-form restatement of the EigenPDE-NN loss+gradient step.

Oracle (test infrastructure).  Follows, without autograd, what the reference
computes in

  * src/network_arch.py:52-69,108-117   k independent MLPs, y = cat_i f_i(x)
  * src/pytorch_eigen_solver.py:164     input gradients  grad_x f_i
  * src/pytorch_eigen_solver.py:166-173 W, mean_i, var_i
  * src/pytorch_eigen_solver.py:215     Rayleigh quotients (eig_vals)
  * src/pytorch_eigen_solver.py:217-221 ascending sort -> cvec
  * src/pytorch_eigen_solver.py:224     non_penalty_loss
  * src/pytorch_eigen_solver.py:177-189 penalty_term
  * src/pytorch_eigen_solver.py:229-233 loss and d loss / d theta

The backward pass is the hand-derived reverse sweep of SURVEY.md section 8(a).
Parameters are a list (one entry per net) of lists of (W_l, b_l) with
W_l of shape [n_l, n_{l-1}] (torch.nn.Linear layout).
"""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------------------
# activations: value, first and second derivative as functions of z
# (torch.nn defaults: ELU alpha=1, CELU alpha=1, Softplus beta=1 threshold=20)
# ---------------------------------------------------------------------------

def _sigmoid(z):
    return 0.5 * (1.0 + np.tanh(0.5 * z))


def act_tanh(z):
    t = np.tanh(z)
    d1 = 1.0 - t * t
    return t, d1, -2.0 * t * d1


def act_sigmoid(z):
    s = _sigmoid(z)
    d1 = s * (1.0 - s)
    return s, d1, d1 * (1.0 - 2.0 * s)


def act_softplus(z):
    s = _sigmoid(z)
    val = np.where(z > 20.0, z, np.logaddexp(0.0, np.minimum(z, 20.0)))
    d1 = np.where(z > 20.0, 1.0, s)
    d2 = np.where(z > 20.0, 0.0, s * (1.0 - s))
    return val, d1, d2


def act_logsigmoid(z):
    s = _sigmoid(z)
    return -np.logaddexp(0.0, -z), 1.0 - s, -s * (1.0 - s)


def act_elu(z):  # identical to CELU for alpha = 1
    e = np.exp(np.minimum(z, 0.0))
    pos = z > 0.0
    return np.where(pos, z, e - 1.0), np.where(pos, 1.0, e), np.where(pos, 0.0, e)


def act_relu(z):
    pos = z > 0.0
    return np.where(pos, z, 0.0), np.where(pos, 1.0, 0.0), np.zeros_like(z)


def act_tanhshrink(z):
    t = np.tanh(z)
    return z - t, t * t, 2.0 * t * (1.0 - t * t)


ACTIVATIONS = {
    "Tanh": act_tanh,
    "Sigmoid": act_sigmoid,
    "Softplus": act_softplus,
    "LogSigmoid": act_logsigmoid,
    "ELU": act_elu,
    "CELU": act_elu,
    "ReLU": act_relu,
    "Tanhshrink": act_tanhshrink,
}


def forward_net(layers, x, act):
    """One MLP (network_arch.py:52-69).  Returns y[B] and the per-layer cache."""
    a = x
    zs, acts, d1s, d2s = [], [x], [], []
    L = len(layers)
    for l, (Wl, bl) in enumerate(layers):
        z = a @ Wl.T + bl
        zs.append(z)
        if l < L - 1:
            a, d1, d2 = act(z)
            acts.append(a)
            d1s.append(d1)
            d2s.append(d2)
    return zs[-1][:, 0], (acts, d1s, d2s)


def input_jacobian(layers, cache):
    """grad_x f (pytorch_eigen_solver.py:164) by the chain g_L=1, u_{l-1}=W_l^T g_l."""
    acts, d1s, _ = cache
    L = len(layers)
    B = acts[0].shape[0]
    g = [None] * (L + 1)   # g[l] = d y / d z_l   (l = 1..L)
    u = [None] * (L + 1)   # u[l] = d y / d a_l   (l = 0..L-1)
    g[L] = np.ones((B, 1))
    for l in range(L, 0, -1):
        u[l - 1] = g[l] @ layers[l - 1][0]
        if l - 1 >= 1:
            g[l - 1] = d1s[l - 2] * u[l - 1]
    return g, u


def batch_sums(y, e, w):
    """W, S1, S2, E of SURVEY 8 (fp64)."""
    W = w.sum()
    S1 = (w[:, None] * y).sum(axis=0)
    S2 = (w[:, None, None] * y[:, :, None] * y[:, None, :]).sum(axis=0)
    E = (w[:, None] * e).sum(axis=0)
    return W, S1, S2, E


def coefficients(W, S1, S2, E, beta, alpha, eig_w, sort=True):
    """Everything that depends only on the batch-global sums.

    Returns dict with eig_vals (sorted ascending if sort), cvec, loss,
    non_penalty, penalty, and the per-sample seed coefficients
    ebar_coef[i] (ebar_{n,i} = w_n * ebar_coef[i]) and the symmetric matrix C
    (ybar_n = (w_n / W) * C (y_n - m)).
    """
    k = S1.shape[0]
    m = S1 / W
    cov = S2 / W - np.outer(m, m)
    var = np.diag(cov).copy()
    R = E / (W * beta)
    lam = R / var
    if sort:
        # np.argsort default (quicksort, not stable) is what the reference
        # calls (pytorch_eigen_solver.py:219); for distinct values any sort agrees.
        cvec = np.argsort(lam)
    else:
        cvec = np.arange(k)
    eig_sorted = lam[cvec]
    wt = np.zeros(k)
    for p in range(k):
        wt[cvec[p]] = eig_w[p]
    non_penalty = float((wt * lam).sum())
    penalty = float(((var - 1.0) ** 2).sum())
    for i in range(k):
        for j in range(i + 1, k):
            penalty += cov[i, j] ** 2
    loss = non_penalty + alpha * penalty
    C = 2.0 * alpha * cov.copy()
    for i in range(k):
        C[i, i] = 2.0 * (-wt[i] * R[i] / var[i] ** 2 + 2.0 * alpha * (var[i] - 1.0))
    ebar_coef = wt / (W * beta * var)
    return dict(eig_vals=eig_sorted, cvec=cvec.astype(np.int64), loss=loss,
                non_penalty=non_penalty, penalty=penalty, mean=m, var=var, cov=cov,
                C=C, ebar_coef=ebar_coef, lam=lam, W=W)


def step(params, x, w, diag_coeff, beta, alpha, eig_w, activation="Tanh", sort=True, global_sums=None,
         sums_only=False, metric=None):
    """Loss, eigenvalue estimates and parameter gradients for one batch.

    x [B,d] fp64 (already gathered), w [B] fp64.
    Returns dict; grads has the same nesting as params.
    global_sums=(W,S1,S2,E): use these batch-global sums instead of the local ones (the samples
    given are then one shard of the batch and `grads` is that shard's additive share).
    sums_only: stop after the local sums (phase 1 of a sharded step).
    metric [B,d,d] (symmetric): per-sample metric replacing diag(c) in the energy, e = u0^T M u0 - the MD alignment
    layer (oracle/md_align.py: x is then the ALIGNED minibatch and M = J diag(c) J^T).
    """
    act = ACTIVATIONS[activation]
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    c = np.asarray(diag_coeff, dtype=np.float64)
    k = len(params)
    B = x.shape[0]
    ys, caches, gs, us = [], [], [], []
    for layers in params:
        y, cache = forward_net(layers, x, act)
        g, u = input_jacobian(layers, cache)
        ys.append(y); caches.append(cache); gs.append(g); us.append(u)
    y = np.stack(ys, axis=1)
    if metric is None:
        mu0 = [c[None, :] * u[0] for u in us]                       # M u0 with M = diag(c)
    else:
        mu0 = [np.einsum("bij,bj->bi", np.asarray(metric, np.float64), u[0]) for u in us]
    e = np.stack([(m_ * u[0]).sum(axis=1) for m_, u in zip(mu0, us)], axis=1)
    W, S1, S2, E = batch_sums(y, e, w)
    if sums_only:
        return dict(sums=dict(W=W, S1=S1, S2=S2, E=E), y=y, e=e)
    if global_sums is not None:
        W, S1, S2, E = global_sums
    co = coefficients(W, S1, S2, E, beta, alpha, eig_w, sort)
    ybar = (w / W)[:, None] * ((y - co["mean"][None, :]) @ co["C"])  # C symmetric
    grads = []
    for i, layers in enumerate(params):
        L = len(layers)
        acts, d1s, d2s = caches[i]
        g, u = gs[i], us[i]
        ebar = w * co["ebar_coef"][i]
        gW = [np.zeros_like(Wl) for Wl, _ in layers]
        gb = [np.zeros_like(bl) for _, bl in layers]
        zbar = [None] * (L + 1)
        ubar = 2.0 * mu0[i] * ebar[:, None]                         # ubar_0 = 2 ebar M u0
        for l in range(1, L + 1):
            gW[l - 1] += g[l].T @ ubar
            gbar = ubar @ layers[l - 1][0].T                       # W_l ubar_{l-1}
            if l < L:
                ubar = d1s[l - 1] * gbar
                zbar[l] = d2s[l - 1] * u[l] * gbar
        zbar[L] = ybar[:, i:i + 1].copy()
        for l in range(L, 0, -1):
            gW[l - 1] += zbar[l].T @ acts[l - 1]
            gb[l - 1] += zbar[l].sum(axis=0)
            if l > 1:
                zbar[l - 1] = zbar[l - 1] + d1s[l - 2] * (zbar[l] @ layers[l - 1][0])
        grads.append([(gW[l], gb[l]) for l in range(L)])
    out = dict(co)
    out.update(grads=grads, y=y, e=e, sums=dict(W=W, S1=S1, S2=S2, E=E))
    return out


def flatten(nested):
    """Flat fp64 vector in model.parameters() order: nets x [W_1,b_1,...,W_L,b_L]."""
    return np.concatenate([np.concatenate([Wl.ravel(), bl.ravel()]) for net in nested for (Wl, bl) in net])


def forward_only(params, x, activation="Tanh"):
    """y[B,k] (network_arch.py:108-117); used by the stage-end re-normalisation."""
    act = ACTIVATIONS[activation]
    return np.stack([forward_net(layers, np.asarray(x, np.float64), act)[0] for layers in params], axis=1)
