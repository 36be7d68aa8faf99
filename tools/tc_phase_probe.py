"""Cycle breakdown of the tensor-core fused kernels (needs a -DEPDE_PHASE_TIMING build selected with EPDE_LIB).
TC_PHASE=1: k_pass1_tc, TC_PHASE=2: k_pass2_tc.  Slot names are given on the command line."""
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
eng = StepEngine(X, w, [d, 20, 20, 20, 1], k, "Tanh", np.ones(d), 1.0, [1.0, 0.8, 0.6])
params = (0.5 * torch.randn(eng.n_params)).cuda()
idx = torch.randint(0, Kd, (B,), device="cuda", dtype=torch.int64)
L.epde_debug_tc_cycles.argtypes = [C.c_void_p, C.c_int]
for _ in range(3):
    eng.step_device(idx, B, params, 20.0)
torch.cuda.synchronize()
which = os.environ.get("TC_PHASE", "1")
ws = eng._workspace(B)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
args = (C.byref(eng.desc), X.data_ptr(), w.data_ptr(), idx.data_ptr(), C.c_int64(B), params.data_ptr(), eng.diag.data_ptr())
def snap():
    h = (C.c_ulonglong * 16)(); L.epde_debug_tc_cycles(h, 0); return np.array(list(h), dtype=np.float64)
reps, acc = 5, np.zeros(16)
for _ in range(reps):
    s0 = snap()
    _lib.check(L.epde_step_partial(*args, ws.data_ptr(), C.c_size_t(ws.numel()), eng.sums.data_ptr(), st), "p")
    torch.cuda.synchronize()
    s1 = snap()
    _lib.check(L.epde_step_finish(*args, C.c_double(1.0), C.c_double(20.0), eng._eigw_c, eng.sums.data_ptr(), ws.data_ptr(),
                                  C.c_size_t(ws.numel()), eng.out.data_ptr(), eng.grad.data_ptr(), st), "f")
    torch.cuda.synchronize()
    s2 = snap()
    acc += (s1 - s0) if which == "1" else (s2 - s1)
gx = 148 // k
ngroups = 3 if which == "1" else 2
tiles = (B / 128) / (gx * ngroups)      # tiles per group (thread 0 = group 0 of each CTA)
names = sys.argv[1:] + ["s%d" % i for i in range(len(sys.argv) - 1, 16)]
tot = acc.sum()
for n, v in zip(names, acc):
    if v:
        print("  %-10s %8.0f cycles/tile  %5.1f%%" % (n, v / reps / (gx * k) / tiles, 100.0 * v / tot))
print("  total %.0f cycles/tile (128 samples, one net, one group)" % (tot / reps / (gx * k) / tiles))
