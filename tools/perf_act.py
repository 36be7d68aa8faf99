"""Step time of a 50-D, k = 3, 20-20-20 net with a non-Tanh activation: fused FFMA2 kernels vs the layer-wise family."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from eigenpde_nn_b200.engine import StepEngine
rng = np.random.default_rng(0)
K, d, k = 1 << 20, 50, 3
X = torch.from_numpy((0.5 * rng.standard_normal((K, d))).astype(np.float32))
w = torch.ones(K)
for act in ("Softplus", "ELU"):
    for B in (5000, 1 << 18):
        for generic in (False, True):
            eng = StepEngine(X, w, [d, 20, 20, 20, 1], k, act, np.ones(d), 1.0, [1.0, 0.8, 0.6], generic=generic)
            p = (0.3 * torch.randn(eng.n_params, generator=torch.Generator().manual_seed(1))).cuda()
            idx = torch.randint(0, K, (B,), device="cuda")
            for _ in range(3): eng.step_device(idx, B, p, 20.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): eng.step_device(idx, B, p, 20.0)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print("%-9s B = %7d  %-22s %.3f ms per step  %.1f M samples/s" % (act, B, "layer-wise family" if generic else "fused FFMA2 kernels", ms, B / ms / 1e3))
