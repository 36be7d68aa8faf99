"""Quick device-side timing of the fused step at a given batch size (diagnostics)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eigenpde_nn_b200.engine import StepEngine

d, k = int(os.environ.get("D", 50)), int(os.environ.get("KNETS", 3))
arch = [d, 20, 20, 20, 1]
Kdata = int(os.environ.get("KDATA", 1 << 22))
rng = np.random.default_rng(0)
X = rng.standard_normal((Kdata, d)).astype(np.float32) * 0.3
w = np.ones(Kdata, np.float32)
eng = StepEngine(X, w, arch, k, "Tanh", np.ones(d), 1.0, [1.0, 0.8, 0.6, 0.3, 0.2, 0.1][:k])
g = torch.Generator(device="cpu").manual_seed(1)
params = (0.5 * torch.randn(eng.n_params, generator=g)).cuda()
for B in [int(b) for b in os.environ.get("BS", "5000,20000,262144,1048576,4194304").split(",")]:
    idx = torch.randint(0, Kdata, (B,), device="cuda", dtype=torch.int64)
    for _ in range(3):
        eng.step_device(idx, B, params, 20.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        eng.step_device(idx, B, params, 20.0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flop = 2 * k * (6 * (d * 20 + 400 + 400 + 20) - d * 20)
    print("B=%8d  %.3f ms/step  %.1f Msamples/s  alg %.2f TFLOP/s (%.1f%% of 73)" % (
        B, ms, B / ms / 1e3, flop * B / ms / 1e9, flop * B / ms / 1e9 / 73 * 100))
