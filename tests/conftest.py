import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


MD_ALIGNED = "cfg3_md_30d_k2_aligned"       # MD mode with the Kabsch alignment layer in the forward: its own tests


def golden_names():
    skip = {"index_streams.npz", "dropin_run.npz", "dropin_run_md.npz", MD_ALIGNED + ".npz"}
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f not in skip)


def load_golden(name):
    """Fixture written by oracle/gen_golden.py from the unmodified reference."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    k = int(z["k"])
    arch = [int(v) for v in z["arch"]]
    L = len(arch) - 1
    params = [[(z[f"W_{i}_{l}"].astype(np.float64), z[f"b_{i}_{l}"].astype(np.float64)) for l in range(L)]
              for i in range(k)]
    grads = [[(z[f"gW_{i}_{l}"].astype(np.float64), z[f"gb_{i}_{l}"].astype(np.float64)) for l in range(L)]
             for i in range(k)]
    extra = {}
    if "ref_x" in z.files:                       # MD-aligned fixture: reference configuration + the aligned minibatch
        extra = dict(ref_x=z["ref_x"].astype(np.float64), ref_index=[int(v) for v in z["ref_index"]],
                     x_aligned=z["x_aligned"].astype(np.float64))
    return dict(**extra, name=name, k=k, arch=arch, act=str(z["act"]), X=z["X"].astype(np.float64), w=z["w"].astype(np.float64),
                idx=z["idx"], params=params, grads=grads, eig_vals=z["eig_vals"], cvec=z["cvec"],
                loss=float(z["loss"]), non_penalty=float(z["non_penalty"]), penalty=float(z["penalty"]),
                diag_coeff=z["diag_coeff"], beta=float(z["beta"]), alpha=float(z["alpha"]),
                eig_w=[float(v) for v in z["eig_w"]], minibatch_seed=int(z["minibatch_seed"]))


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def check_step_against(ref, got, tol, grads_nested=None, flat=None):
    """Parity metric of SURVEY.md section 8(c).

    scalars: relative; penalty: absolute against max(|ref|, 1e-2*|loss|);
    gradients: per tensor ||d||2/||g_ref||2, last-layer biases (identically zero
    in exact arithmetic) against tol * max|g_ref(net)|.
    """
    assert list(got["cvec"]) == list(ref["cvec"]), (got["cvec"], ref["cvec"])
    assert rel(got["loss"], ref["loss"]) <= tol, ("loss", got["loss"], ref["loss"])
    assert rel(got["non_penalty"], ref["non_penalty"]) <= tol, ("non_penalty", got["non_penalty"], ref["non_penalty"])
    for a, b in zip(got["eig_vals"], ref["eig_vals"]):
        assert rel(a, b) <= tol, ("eig", a, b)
    pen_scale = max(abs(ref["penalty"]), 1e-2 * abs(ref["loss"]))
    assert abs(got["penalty"] - ref["penalty"]) <= tol * pen_scale, ("penalty", got["penalty"], ref["penalty"])
    worst = 0.0
    if grads_nested is not None:
        for i, (gnet, rnet) in enumerate(zip(grads_nested, ref["grads"])):
            ginf = max(np.abs(rW).max() for rW, _ in rnet)
            L = len(rnet)
            for l, ((gW, gb), (rW, rb)) in enumerate(zip(gnet, rnet)):
                eW = np.linalg.norm(gW - rW) / np.linalg.norm(rW)
                worst = max(worst, eW)
                assert eW <= tol, (f"gW net{i} layer{l}", eW)
                if l == L - 1:
                    assert np.abs(gb - rb).max() <= tol * ginf, (f"gb(last) net{i}", gb, rb, ginf)
                else:
                    eb = np.linalg.norm(gb - rb) / np.linalg.norm(rb)
                    worst = max(worst, eb)
                    assert eb <= tol, (f"gb net{i} layer{l}", eb)
    return worst


def unflatten(flat, arch, k):
    """Flat model.parameters()-ordered vector -> nested [(W,b)]."""
    out, off = [], 0
    for _ in range(k):
        net = []
        for l in range(len(arch) - 1):
            nW = arch[l + 1] * arch[l]
            W = np.asarray(flat[off:off + nW]).reshape(arch[l + 1], arch[l]); off += nW
            b = np.asarray(flat[off:off + arch[l + 1]]); off += arch[l + 1]
            net.append((W, b))
        out.append(net)
    assert off == len(flat)
    return out
