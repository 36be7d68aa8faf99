"""Step time of the two fused kernel families at small and medium batch sizes (50-D, k = 3)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from eigenpde_nn_b200 import _lib
from eigenpde_nn_b200.engine import StepEngine
L = _lib.lib()
K, d, k = 1 << 20, 50, 3
X = torch.zeros((K, 52), device="cuda"); X[:, :d] = torch.randn((K, d), device="cuda")
w = torch.ones(K, device="cuda")
p = None
for B in [512, 2000, 5000, 10000, 20000, 50000, 200000]:
    idx = torch.randint(0, K, (B,), device="cuda")
    row = []
    for impl in (0, 1):
        eng = StepEngine(X, w, [d, 20, 20, 20, 1], k, "Tanh", np.ones(d), 1.0, [1.0, 0.8, 0.6], ffma2=(impl == 0))
        if p is None: p = (0.5 * torch.randn(eng.n_params)).cuda()
        for _ in range(5): eng.step_device(idx, B, p, 20.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): eng.step_device(idx, B, p, 20.0)
        e1.record(); torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) / 50 * 1e3)
    print("B=%7d  ffma2 %8.1f us   tcgen05 %8.1f us   ratio %.2f" % (B, row[0], row[1], row[0] / row[1]))
