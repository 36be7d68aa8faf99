"""CPU: pin the oracle against outputs of the unmodified reference (tests/golden)."""
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN, MD_ALIGNED, check_step_against, golden_names, load_golden
from oracle import closed_form, md_align, minibatch, ref_port


@pytest.mark.parametrize("name", golden_names())
def test_closed_form_matches_reference(name):
    g = load_golden(name)
    act = g["act"] if g["act"] in closed_form.ACTIVATIONS else "CELU"   # network_arch.py:27-32 fallback
    xb, wb = g["X"][g["idx"]], g["w"][g["idx"]]
    out = closed_form.step(g["params"], xb, wb, g["diag_coeff"], g["beta"], g["alpha"], g["eig_w"], act)
    tol = 2e-7 if "wide" in name else 1e-9      # the wide fixture stores its gradients as float32
    check_step_against(g, out, tol, out["grads"])


@pytest.mark.parametrize("name", [n for n in golden_names() if "wide" not in n])
def test_torch_port_matches_reference(name):
    g = load_golden(name)
    model = ref_port.nested_to_model(g["params"], g["act"])
    xb = torch.from_numpy(g["X"][g["idx"]]); wb = torch.from_numpy(g["w"][g["idx"]])
    eig, cvec, loss, nonpen, pen = ref_port.loss_and_backward(
        model, xb, wb, torch.from_numpy(g["diag_coeff"]), g["beta"], g["alpha"], g["eig_w"])
    got = dict(eig_vals=eig, cvec=cvec, loss=loss, non_penalty=nonpen, penalty=pen)
    check_step_against(g, got, 1e-11, ref_port.grads_nested(model))


def test_minibatch_streams_bit_exact():
    z = np.load(GOLDEN + "/index_streams.npz")
    n = 0
    while f"s{n}_meta" in z:
        seed, K, B, reps = (int(v) for v in z[f"s{n}_meta"])
        gen = minibatch.MT19937(seed)
        for r in range(reps):
            assert np.array_equal(gen.minibatch(K, B), z[f"s{n}_idx"][r]), (seed, K, B, r)
        n += 1
    assert n >= 4


@pytest.mark.parametrize("name", ["cfg1_ex1_50d_k3", "cfg3_md_30d_k2"])
def test_golden_indices_are_the_reference_stream(name):
    g = load_golden(name)
    got = minibatch.MT19937(g["minibatch_seed"]).minibatch(g["X"].shape[0], len(g["idx"]))
    assert np.array_equal(got, g["idx"])


def test_minibatch_matches_cpython_random_live():
    for seed, K, B in [(5, 1000, 300), (123456789012345, 77777, 1000)]:
        random.seed(seed)
        ref = random.choices(range(K), cum_weights=np.arange(1, K + 1), k=B)
        assert np.array_equal(minibatch.MT19937(seed).minibatch(K, B), np.asarray(ref))


def test_md_alignment_layer_matches_reference():
    """the reference run with `MD_data_set` + reference configuration (alignment inside every net's
    forward, differentiated by autograd) against the oracle's restatement: aligned rows, closed-form Jacobian, and the
    whole step with the per-sample metric M = J diag(c) J^T."""
    g = load_golden(MD_ALIGNED)
    xb, wb = g["X"][g["idx"]], g["w"][g["idx"]]
    xa, M = md_align.align_batch(xb, g["ref_x"], g["ref_index"], g["diag_coeff"])
    assert np.abs(xa - g["x_aligned"]).max() <= 1e-12
    out = closed_form.step(g["params"], xa, wb, g["diag_coeff"], g["beta"], g["alpha"], g["eig_w"], g["act"], metric=M)
    check_step_against(g, out, 1e-9, out["grads"])


def test_md_alignment_jacobian_matches_autograd():
    """The closed-form derivative of the Kabsch rotation (no autograd) against torch.autograd on the reference's
    SVD expression (src/data_set.py:133-153)."""
    g = load_golden(MD_ALIGNED)
    ref, ri = torch.from_numpy(g["ref_x"]), g["ref_index"]

    def align_like_reference(x):
        p = x.reshape(-1, 3)
        sel = p[ri]
        c = sel.mean(0, keepdim=True)
        H = (sel - c).T @ ref
        u, s, vh = torch.linalg.svd(H)
        D = torch.eye(3, dtype=torch.float64)
        D[2, 2] = torch.sign(torch.linalg.det(u @ vh)).detach()
        return ((p - c) @ (u @ D @ vh)).reshape(-1)

    for n in (0, 7, 123):
        x = g["X"][g["idx"][n]]
        Jt = torch.autograd.functional.jacobian(align_like_reference, torch.from_numpy(x)).numpy()
        xa, J = md_align.align_one(x, g["ref_x"], ri)
        assert np.abs(J - Jt).max() <= 1e-12
        assert np.abs(xa - align_like_reference(torch.from_numpy(x)).numpy()).max() <= 1e-12
