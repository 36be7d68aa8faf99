"""Dev probe: FFMA vs tensor-core fused family against the fp64 oracle + timing of the two phases."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch, ctypes as C
from eigenpde_nn_b200 import _lib
from eigenpde_nn_b200.engine import StepEngine
from oracle import closed_form
import test_gpu_parity as T
from conftest import check_step_against, unflatten

L = _lib.lib()
cases = [(50, 3, 3000, 1000), (50, 3, 3000, 4099), (50, 3, 3000, 130), (2, 2, 3000, 2000), (30, 2, 3000, 1500), (50, 6, 3000, 2500),
         (7, 1, 500, 300), (60, 3, 2000, 900)]
worst = 0
for (d, k, K, B) in cases:
    c = T._rand_case(d * 7 + k, d, k, [20, 20, 20], K, B)
    ref = T._oracle(c)
    for impl in (0, 1):
        eng = T._engine(c, ffma2=(impl == 0))
        res = eng.step(c["idx"], T._flat(c["params"]), c["alpha"])
        grads = unflatten(res["grad"].cpu().numpy().astype(np.float64), c["arch"], c["k"])
        try:
            err = check_step_against(ref, res, 1e-5, grads)
            print("d=%d k=%d B=%d impl=%d OK %s" % (d, k, B, impl, err))
        except AssertionError as e:
            print("d=%d k=%d B=%d impl=%d FAIL %s" % (d, k, B, impl, str(e)[:300]))
        # direct comparison of scalars
        print("    loss %.9e ref %.9e  eig %s" % (res["loss"], ref["loss"], np.array2string(np.asarray(res["eig_vals"]), precision=8)))

# timing at the bench shape
if "--time" in sys.argv:
    d, k, K, B = 50, 3, 1 << 22, 1 << 20
    rng = np.random.default_rng(0)
    X = torch.randn(K, 52, device="cuda"); X[:, 50:] = 0
    w = torch.rand(K, device="cuda") + 0.5
    c = T._rand_case(1, d, k, [20, 20, 20], 10, 10)
    p = T._flat(c["params"])
    idx = torch.randint(0, K, (B,), device="cuda")
    for impl in (0, 1):
        eng = StepEngine(X, w, c["arch"], k, "Tanh", np.ones(d), 1.0, c["eig_w"], ffma2=(impl == 0))
        for _ in range(3): eng.step_device(idx, B, p, 20.0)
        torch.cuda.synchronize()
        ws = eng._workspace(B)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        args = (C.byref(eng.desc), X.data_ptr(), w.data_ptr(), idx.data_ptr(), C.c_int64(B), p.data_ptr(), eng.diag.data_ptr())
        e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tp = tf = 0
        for _ in range(10):
            e0.record()
            _lib.check(L.epde_step_partial(*args, ws.data_ptr(), C.c_size_t(ws.numel()), eng.sums.data_ptr(), st), "p")
            e1.record()
            _lib.check(L.epde_step_finish(*args, C.c_double(1.0), C.c_double(20.0), eng._eigw_c, eng.sums.data_ptr(), ws.data_ptr(),
                                          C.c_size_t(ws.numel()), eng.out.data_ptr(), eng.grad.data_ptr(), st), "f")
            e2.record(); torch.cuda.synchronize()
            tp += e0.elapsed_time(e1); tf += e1.elapsed_time(e2)
        print("impl=%d partial %.3f ms finish %.3f ms  -> %.1f M samples/s" % (impl, tp / 10, tf / 10, B / (tp + tf) * 10 / 1e3))
