"""MD mode (BASELINE config 3 shape WITH the alignment layer): one-time epde_md_align cost and the
per-step time of the metric step on the fused FFMA2 kernels and on the generic family.  d = 30 (10 atoms), k = 2, K = 1e5 rows, B = 10 000."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from eigenpde_nn_b200.engine import StepEngine, md_align

rng = np.random.default_rng(0)
K, d, k, B = 100000, 30, 2, 10000
tmpl = 1.5 * rng.standard_normal((1, d))
X = (tmpl + np.sqrt(0.1) * rng.standard_normal((K, d))).astype(np.float32)
w = np.exp(2.0 * rng.standard_normal(K)).astype(np.float32); w /= w.mean()
ri = [0, 2, 5, 7, 8]
ref = X[0].reshape(-1, 3)[ri].astype(np.float64); ref -= ref.mean(0, keepdims=True)
beta = 1.0 / (0.001987191 * 300.0)
diag = np.full(d, 1e-5 * 1e7 * beta)
torch.cuda.synchronize()
md_align(X[:100], ref, ri, diag)                      # warm-up (context, module load)
t0 = time.perf_counter()
Xa, M = md_align(X, ref, ri, diag)
t_align = time.perf_counter() - t0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
Xd = torch.zeros((K, 32), device="cuda"); Xd[:, :d] = torch.from_numpy(X).cuda()
dg = torch.from_numpy(diag).float().cuda()
import ctypes as C
from eigenpde_nn_b200 import _lib
L = _lib.lib()
refc = np.ascontiguousarray(ref); ric = np.asarray(ri, np.int32)
e0.record()
for _ in range(5):
    L.epde_md_align(C.c_int64(K), 10, 32, Xd.data_ptr(), refc.ctypes.data_as(C.c_void_p), ric.ctypes.data_as(C.c_void_p), 5,
                    dg.data_ptr(), Xa.data_ptr(), M.data_ptr(), torch.cuda.current_stream().cuda_stream)
e1.record(); torch.cuda.synchronize()
print("epde_md_align: %.2f ms for K = %d rows (kernel only; %.1f ms incl. upload/allocation), metric %.0f MB"
      % (e0.elapsed_time(e1) / 5, K, t_align * 1e3, M.numel() * 4 / 1e6))
p = None
idx = torch.randint(0, K, (B,), device="cuda")
for generic in (False, True):
    eng = StepEngine(Xa, torch.from_numpy(w), [d, 20, 20, 20, 1], k, "Tanh", diag, beta, [1.0, 0.8], metric=M, generic=generic)
    if p is None: p = (0.5 * torch.randn(eng.n_params)).cuda()
    for _ in range(3): eng.step_device(idx, B, p, 20.0)
    e0.record()
    for _ in range(20): eng.step_device(idx, B, p, 20.0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("MD-mode step (per-row metric, %s): %.3f ms per step at B = %d -> %.2f M samples/s"
          % ("generic layer-wise family" if generic else "fused FFMA2 kernels", ms, B, B / ms / 1e3))
eng2 = StepEngine(Xa, torch.from_numpy(w), [d, 20, 20, 20, 1], k, "Tanh", diag, beta, [1.0, 0.8])
for _ in range(3): eng2.step_device(idx, B, p, 20.0)
e0.record()
for _ in range(20): eng2.step_device(idx, B, p, 20.0)
e1.record(); torch.cuda.synchronize()
print("same shape without the layer (fused tensor-core family): %.3f ms per step" % (e0.elapsed_time(e1) / 20))
