"""TEST INFRASTRUCTURE: a CPU stand-in for eigenpde_nn_b200.engine.StepEngine built on the fp64 oracle, so that the host
side of the drop-in (module names, Param fields, stage schedule, files, log format) can be driven by the reference's
UNCHANGED utils/train_nn.py in the build container, which has no GPU.  Never imported by the package."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class OracleStepEngine:
    def __init__(self, X, weights, arch, k, activation_name, diag_coeff, beta, eig_w, sort=True, device=None,
                 process_group=None, deterministic=False, metric=None, ffma2=False):
        from oracle import closed_form
        self.cf = closed_form
        self.X = np.asarray(X, dtype=np.float64)
        self.wv = np.asarray(weights, dtype=np.float64)
        self.w = torch.from_numpy(self.wv)
        self.arch, self.k, self.act = [int(a) for a in arch], int(k), activation_name
        self.diag, self.beta, self.eig_w = np.asarray(diag_coeff, dtype=np.float64), float(beta), list(eig_w)[:k]
        self.K, self.device, self.fused, self.metric, self.pg = self.X.shape[0], torch.device("cpu"), False, None, None
        self.n_params = k * sum(self.arch[l + 1] * (self.arch[l] + 1) for l in range(len(self.arch) - 1))

    def _nested(self, flat):
        v = flat.detach().cpu().double().numpy()
        out, off = [], 0
        for _ in range(self.k):
            net = []
            for l in range(len(self.arch) - 1):
                n, m = self.arch[l + 1], self.arch[l]
                W = v[off:off + n * m].reshape(n, m); off += n * m
                b = v[off:off + n]; off += n
                net.append((W, b))
            out.append(net)
        return out

    def flat_params(self, model):
        return torch.cat([p.detach().reshape(-1) for p in model.parameters()]).double()

    def step(self, idx, params, alpha):
        idx = np.arange(self.K) if idx is None else np.asarray(idx)
        r = self.cf.step(self._nested(params), self.X[idx], self.wv[idx], self.diag, self.beta, alpha, self.eig_w, self.act)
        return dict(loss=r["loss"], non_penalty=r["non_penalty"], penalty=r["penalty"], eig_vals=np.asarray(r["eig_vals"]),
                    cvec=np.asarray(r["cvec"], dtype=np.int64), grad=torch.from_numpy(self.cf.flatten(r["grads"])))

    def forward(self, idx_dev, B, params):
        nets = self._nested(params)
        f = {"Tanh": np.tanh}[self.act]
        cols = []
        for net in nets:
            a = self.X
            for l, (W, b) in enumerate(net):
                a = a @ W.T + b
                if l < len(net) - 1:
                    a = f(a)
            cols.append(a[:, 0])
        return torch.from_numpy(np.stack(cols, axis=1))


def install(fixed_time=None):
    os.environ["EPDE_HOST_LOOP"] = "1"
    import eigenpde_nn_b200.engine as eng
    eng.StepEngine = OracleStepEngine
    if fixed_time is not None:
        import time
        time.time = lambda: float(fixed_time)
