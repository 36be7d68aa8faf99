"""Cycle breakdown of the tensor-core fused kernels (needs a -DEPDE_PHASE_TIMING build selected with EPDE_LIB)."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eigenpde_nn_b200.engine import StepEngine
from eigenpde_nn_b200 import _lib
d, k, B, Kd = 50, 3, 1 << 20, 1 << 22
X = torch.zeros((Kd, 52), device="cuda"); X[:, :d] = 0.3 * torch.randn((Kd, d), device="cuda")
w = torch.ones(Kd, device="cuda")
L = _lib.lib()
_lib.check(L.epde_set_fused_impl(1), "impl")
eng = StepEngine(X, w, [d, 20, 20, 20, 1], k, "Tanh", np.ones(d), 1.0, [1.0, 0.8, 0.6])
params = (0.5 * torch.randn(eng.n_params)).cuda()
idx = torch.randint(0, Kd, (B,), device="cuda", dtype=torch.int64)
L.epde_debug_tc_cycles.argtypes = [C.c_void_p, C.c_int]
h = (C.c_ulonglong * 16)()
for _ in range(3):
    eng.step_device(idx, B, params, 20.0)
torch.cuda.synchronize()
L.epde_debug_tc_cycles(h, 1)
reps = 5
for _ in range(reps):
    eng.step_device(idx, B, params, 20.0)
torch.cuda.synchronize()
L.epde_debug_tc_cycles(h, 0)
gx = 148 // k
tiles = (B / 128) / (gx * 3)            # tiles per group (thread 0 = group 0 of each CTA)
names = sys.argv[1:] or ["alu", "bar", "issue", "mma-wait"] + ["s%d" % i for i in range(4, 16)]
tot = sum(h)
for n, v in zip(names, h):
    if v:
        print("  %-9s %8.0f cycles/tile  %5.1f%%" % (n, v / reps / (gx * k) / tiles, 100.0 * v / tot))
print("  total %.0f cycles/tile(128 samples, one net)" % (tot / reps / (gx * k) / tiles))
