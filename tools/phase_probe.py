"""Per-phase cycle breakdown of k_pass2 (needs a -DEPDE_PHASE_TIMING build selected with EPDE_LIB)."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eigenpde_nn_b200.engine import StepEngine
from eigenpde_nn_b200 import _lib
d, k, B, Kd = 50, 3, 1 << 20, 1 << 22
X = torch.zeros((Kd, 52), device="cuda"); X[:, :d] = 0.3 * torch.randn((Kd, d), device="cuda")
w = torch.ones(Kd, device="cuda")
eng = StepEngine(X, w, [d, 20, 20, 20, 1], k, "Tanh", np.ones(d), 1.0, [1.0, 0.8, 0.6])
params = (0.5 * torch.randn(eng.n_params)).cuda()
idx = torch.randint(0, Kd, (B,), device="cuda", dtype=torch.int64)
L = _lib.lib()
L.epde_debug_phase_cycles.argtypes = [C.c_void_p, C.c_int]
h = (C.c_ulonglong * 8)()
for _ in range(3):
    eng.step_device(idx, B, params, 20.0)
torch.cuda.synchronize()
L.epde_debug_phase_cycles(h, 1)
reps = 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    eng.step_device(idx, B, params, 20.0)
e1.record(); torch.cuda.synchronize()
L.epde_debug_phase_cycles(h, 0)
ncta = (148 // k) * k
tiles_per_cta = (B / 256) / (148 // k)
names = ["MV1", "sync", "GM1", "sync", "MV2", "sync", "GM2", "sync"]
tot = sum(h)
print("lib", os.environ.get("EPDE_LIB"), " %.3f ms/step" % (e0.elapsed_time(e1) / reps))
for n, v in zip(names, h):
    print("  %-5s %8.0f cycles/tile  %5.1f%%" % (n, v / reps / ncta / tiles_per_cta, 100.0 * v / tot))
print("  total %.0f cycles/tile" % (tot / reps / ncta / tiles_per_cta))
