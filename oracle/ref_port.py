"""torch fp64 autograd restatement of the reference's update_step (CPU "port").

Oracle / CPU baseline (test infrastructure; see oracle/__init__.py).  It issues
the same sequence of ATen operations as the reference so that timing it is a
fair stand-in for the reference's CPU path where /root/reference is absent
(the GPU box):

  * nets: k independent Linear->act->...->Linear(->1) stacks, outputs
    concatenated                                   (src/network_arch.py:21-39,52-69,108-117)
  * k calls of autograd.grad(y[:,i], x, ones, create_graph=True)
                                                    (src/pytorch_eigen_solver.py:164)
  * W, mean_i, var_i                               (src/pytorch_eigen_solver.py:166-173)
  * detached Rayleigh quotients + argsort          (src/pytorch_eigen_solver.py:215-221)
  * differentiable Rayleigh sum, penalty, loss     (src/pytorch_eigen_solver.py:224-229,177-189)
  * loss.backward()                                (src/pytorch_eigen_solver.py:232-233)
  * minibatch by random.choices + row gather       (src/data_set.py:67,77-80)
  * Adam step                                      (src/pytorch_eigen_solver.py:234,379)
"""
from __future__ import annotations

import itertools
import random

import numpy as np
import torch


class PortNet(torch.nn.Module):
    """k MLPs with the reference's parameter naming nets.{i}.linears.{l}.{weight,bias}."""

    class _One(torch.nn.Module):
        def __init__(self, sizes, activation_name):
            super().__init__()
            self.linears = torch.nn.ModuleList(
                [torch.nn.Linear(sizes[i], sizes[i + 1]) for i in range(len(sizes) - 1)])
            act = getattr(torch.nn, activation_name, torch.nn.CELU)
            self.activations = torch.nn.ModuleList([act() for _ in range(len(sizes) - 2)])
            for lin in self.linears:                      # network_arch.py:37-39
                torch.nn.init.normal_(lin.weight, 0, 0.5)
                torch.nn.init.normal_(lin.bias, 0, 0.5)

        def forward(self, x):
            n = len(self.linears)
            for i, lin in enumerate(self.linears):
                x = lin(x)
                if i < n - 1:
                    x = self.activations[i](x)
            return x

    def __init__(self, sizes, activation_name, k):
        super().__init__()
        self.nets = torch.nn.ModuleList([PortNet._One(sizes, activation_name) for _ in range(k)])

    def forward(self, x):
        return torch.cat([net(x) for net in self.nets], dim=1)


def params_to_nested(model):
    """-> list[k] of list[L] of (W, b) numpy fp64 (the closed_form layout)."""
    return [[(lin.weight.detach().double().numpy().copy(), lin.bias.detach().double().numpy().copy())
             for lin in net.linears] for net in model.nets]


def nested_to_model(nested, activation_name):
    sizes = [nested[0][0][0].shape[1]] + [Wl.shape[0] for Wl, _ in nested[0]]
    model = PortNet(sizes, activation_name, len(nested)).double()
    with torch.no_grad():
        for net, layers in zip(model.nets, nested):
            for lin, (Wl, bl) in zip(net.linears, layers):
                lin.weight.copy_(torch.from_numpy(np.asarray(Wl, dtype=np.float64)))
                lin.bias.copy_(torch.from_numpy(np.asarray(bl, dtype=np.float64)))
    return model


def loss_and_backward(model, x_batch, b_weights, diag_coeff, beta, alpha, eig_w, sort=True):
    """One loss evaluation + backward on a given batch; fills p.grad.

    x_batch [B,d] fp64 tensor (leaf, requires_grad is set here), b_weights [B] fp64.
    Returns (eig_vals sorted np[k], cvec np.int64[k], loss, non_penalty, penalty) as the
    reference's update_step does (src/pytorch_eigen_solver.py:236).
    """
    k = len(model.nets)
    x_batch = x_batch.detach().clone().requires_grad_(True)
    ones = torch.ones(x_batch.shape[0], dtype=x_batch.dtype)
    y = model(x_batch)
    y_grad = [torch.autograd.grad(y[:, i], x_batch, ones, create_graph=True)[0] for i in range(k)]
    tot_w = b_weights.sum()
    mean = [(y[:, i] * b_weights).sum() / tot_w for i in range(k)]
    var = [(y[:, i] ** 2 * b_weights).sum() / tot_w - mean[i] ** 2 for i in range(k)]
    energy = [torch.sum((y_grad[i] ** 2 * diag_coeff).sum(dim=1) * b_weights) for i in range(k)]
    eig_vals = torch.tensor([float(1.0 / (tot_w * beta) * energy[i] / var[i]) for i in range(k)],
                            dtype=torch.float64)
    if sort:
        cvec = np.argsort(eig_vals.numpy())
    else:
        cvec = np.arange(k)
    eig_sorted = eig_vals.numpy()[cvec]
    non_penalty = 1.0 / (tot_w * beta) * sum(eig_w[p] * energy[cvec[p]] / var[cvec[p]] for p in range(k))
    penalty = sum((var[i] - 1.0) ** 2 for i in range(k))
    for i, j in itertools.combinations(range(k), 2):
        penalty = penalty + ((y[:, i] * y[:, j] * b_weights).sum() / tot_w - mean[i] * mean[j]) ** 2
    loss = non_penalty + alpha * penalty
    for p in model.parameters():
        p.grad = None
    loss.backward()
    return (eig_sorted, np.asarray(cvec, dtype=np.int64), float(loss.detach()),
            float(non_penalty.detach()), float(penalty.detach()))


def grads_nested(model):
    return [[(lin.weight.grad.numpy().copy(), lin.bias.grad.numpy().copy()) for lin in net.linears]
            for net in model.nets]


class PortSolver:
    """Minimal trainer around the port: minibatch selection + loss/backward + Adam.

    Mirrors the per-step behaviour of eigen_solver.update_step
    (src/pytorch_eigen_solver.py:191-236) on a data set held as fp64 tensors.
    """

    def __init__(self, model, X, weights, diag_coeff, beta, eig_w, lr=5e-3):
        self.model = model
        self.X = torch.as_tensor(X, dtype=torch.float64)
        self.w = torch.as_tensor(weights, dtype=torch.float64)
        self.K = self.X.shape[0]
        self.cum_weights = np.arange(1, self.K + 1)
        self.diag_coeff = torch.as_tensor(diag_coeff, dtype=torch.float64)
        self.beta = beta
        self.eig_w = eig_w
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=lr)

    def minibatch(self, bsz):
        idx = random.choices(range(self.K), cum_weights=self.cum_weights, k=bsz)   # data_set.py:67
        return idx, self.X[idx].reshape(bsz, -1), self.w[idx]

    def loss_grad(self, x_batch, b_weights, alpha):
        return loss_and_backward(self.model, x_batch, b_weights, self.diag_coeff, self.beta, alpha, self.eig_w)

    def update_step(self, bsz, alpha):
        _, xb, wb = self.minibatch(bsz)
        out = self.loss_grad(xb, wb, alpha)
        self.optimizer.step()
        return out
