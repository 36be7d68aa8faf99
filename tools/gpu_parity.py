"""GPU parity sweep over tests/golden (diagnostics; prints per-tensor errors)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden_names, load_golden, unflatten
from eigenpde_nn_b200.engine import StepEngine
from oracle import closed_form

names = sys.argv[1:] or golden_names()
for name in names:
    g = load_golden(name)
    try:
        eng = StepEngine(g["X"], g["w"], g["arch"], g["k"], g["act"], g["diag_coeff"], g["beta"], g["eig_w"])
    except Exception as e:
        print(name, "ENGINE ERROR", e); continue
    flat = torch.from_numpy(closed_form.flatten(g["params"])).to("cuda", torch.float32)
    try:
        res = eng.step(g["idx"], flat, g["alpha"])
    except Exception as e:
        print(name, "STEP ERROR", e); continue
    torch.cuda.synchronize()
    print("==", name, "fused" if eng.fused else "generic")
    print("  loss %.9g ref %.9g rel %.2e" % (res["loss"], g["loss"], abs(res["loss"] - g["loss"]) / abs(g["loss"])))
    print("  nonpen rel %.2e  pen %.6g ref %.6g" % (abs(res["non_penalty"] - g["non_penalty"]) / abs(g["non_penalty"]), res["penalty"], g["penalty"]))
    print("  eig", res["eig_vals"], "ref", g["eig_vals"], "cvec", res["cvec"], g["cvec"])
    gn = unflatten(res["grad"].cpu().numpy().astype(np.float64), g["arch"], g["k"])
    for i in range(g["k"]):
        errs = []
        for l, ((gW, gb), (rW, rb)) in enumerate(zip(gn[i], g["grads"][i])):
            errs.append("W%d %.1e b%d %.1e" % (l, np.linalg.norm(gW - rW) / np.linalg.norm(rW), l,
                                              np.linalg.norm(gb - rb) / max(np.linalg.norm(rb), 1e-30) if l < len(gn[i]) - 1 else np.abs(gb - rb).max()))
        print("  net%d: %s" % (i, "  ".join(errs)))
