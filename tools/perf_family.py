"""Step time of one fused kernel family on the headline shape (50-D, k = 3): python tools/perf_family.py [ffma2|tc] [B]."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from eigenpde_nn_b200.engine import StepEngine

fam = sys.argv[1] if len(sys.argv) > 1 else "ffma2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 19
rng = np.random.default_rng(0)
K, d, k = 1 << 21, 50, 3
X = torch.from_numpy((0.5 * rng.standard_normal((K, d))).astype(np.float32))
w = torch.ones(K)
eng = StepEngine(X, w, [d, 20, 20, 20, 1], k, "Tanh", np.ones(d), 1.0, [1.0, 0.8, 0.6], ffma2=(fam == "ffma2"))
p = (0.3 * torch.randn(eng.n_params)).cuda()
idx = torch.randint(0, K, (B,), device="cuda")
for _ in range(3): eng.step_device(idx, B, p, 20.0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): eng.step_device(idx, B, p, 20.0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("%s family: %.3f ms per step at B = %d -> %.1f M samples/s (lib %s)" % (fam, ms, B, B / ms / 1e3, os.environ.get("EPDE_LIB", "in-tree")))
