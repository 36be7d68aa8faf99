import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from conftest import load_golden
from eigenpde_nn_b200.engine import StepEngine
from oracle import closed_form
g = load_golden("cfg2_dw_2d_k2")
eng = StepEngine(g["X"], g["w"], g["arch"], g["k"], g["act"], g["diag_coeff"], g["beta"], g["eig_w"])
flat = torch.from_numpy(closed_form.flatten(g["params"])).to("cuda", torch.float32)
res = eng.step(g["idx"], flat, g["alpha"]); torch.cuda.synchronize(); print("ok", res["loss"])
