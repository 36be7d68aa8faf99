"""Timing of the generic path on BASELINE config 5 (4x256, 50-D, k=4)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eigenpde_nn_b200.engine import StepEngine
d, k = 50, 4
arch = [d, 256, 256, 256, 256, 1]
Kd = 1 << 20
X = torch.zeros((Kd, 52), device="cuda"); X[:, :d] = 0.3 * torch.randn((Kd, d), device="cuda")
w = torch.ones(Kd, device="cuda")
eng = StepEngine(X, w, arch, k, "Tanh", np.ones(d), 1.0, [1.0, 0.8, 0.6, 0.3])
params = (0.09 * torch.randn(eng.n_params)).cuda()
M = sum(arch[i] * arch[i + 1] for i in range(len(arch) - 1))
flop = 2 * k * (6 * M - arch[0] * arch[1])
for B in [int(b) for b in os.environ.get("BS", "16384,65536,262144").split(",")]:
    idx = torch.randint(0, Kd, (B,), device="cuda", dtype=torch.int64)
    for _ in range(2):
        eng.step_device(idx, B, params, 20.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        eng.step_device(idx, B, params, 20.0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("wide 4x256 k=4  B=%7d  %.2f ms/step  %.2f Msamples/s  alg %.1f TFLOP/s" % (B, ms, B / ms / 1e3, flop * B / ms / 1e9))
