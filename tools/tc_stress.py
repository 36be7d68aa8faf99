"""Stress the tensor-core family: random shapes against the oracle, and repeated large batches against the
fixed-order result (races in the ticket / staging protocol would show up as sporadic mismatches)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import test_gpu_parity as T
from conftest import check_step_against, unflatten

rng = np.random.default_rng(int(os.environ.get("SEED", "0")))
bad = 0
for it in range(int(os.environ.get("N_SMALL", "40"))):
    d = int(rng.integers(1, 61)); k = int(rng.integers(1, 7)); B = int(rng.integers(8, 6000))
    c = T._rand_case(1000 + it, d, k, [20, 20, 20], 4000, B)
    eng = T._engine(c)
    res = eng.step(c["idx"], T._flat(c["params"]), c["alpha"])
    grads = unflatten(res["grad"].cpu().numpy().astype(np.float64), c["arch"], c["k"])
    try:
        w = check_step_against(T._oracle(c), res, 1e-5, grads)
    except AssertionError as e:
        bad += 1; print("FAIL d=%d k=%d B=%d: %s" % (d, k, B, str(e)[:200]))
print("small cases done, failures:", bad)
for (d, k, B) in [(50, 3, 1 << 20), (50, 6, 1 << 19), (30, 2, 700001), (2, 2, 300000)]:
    c = T._rand_case(5 + d, d, k, [20, 20, 20], 1 << 18, B)
    p = T._flat(c["params"])
    ref = T._engine(c).step(c["idx"], p, c["alpha"])
    g0 = ref["grad"].clone()
    eng = T._engine(c)
    worst = 0.0
    for rep in range(int(os.environ.get("N_BIG", "12"))):
        r = eng.step(c["idx"], p, c["alpha"])
        e = float((r["grad"] - g0).norm() / g0.norm())
        worst = max(worst, e, abs(r["loss"] - ref["loss"]) / abs(ref["loss"]))
    ok = worst <= 2e-6
    bad += 0 if ok else 1
    print("d=%d k=%d B=%d: worst run-to-run deviation from the fixed-order result %.2e %s" % (d, k, B, worst, "ok" if ok else "FAIL"))
print("TOTAL FAILURES", bad)
