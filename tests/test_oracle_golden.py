"""CPU: pin the oracle against outputs of the unmodified reference (tests/golden)."""
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN, check_step_against, golden_names, load_golden
from oracle import closed_form, minibatch, ref_port


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
