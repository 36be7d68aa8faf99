"""Drop-in facade: module / class names, RNG-stream parity of the parameter initialisation, the
data-set loader, and (GPU) a short training run against the log of the unmodified reference."""
import os
import random
import sys
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, load_golden

DROPIN = os.path.join(ROOT, "eigenpde_nn_b200", "dropin")


@pytest.fixture()
def dropin(monkeypatch):
    monkeypatch.syspath_prepend(DROPIN)
    for name in ("data_set", "network_arch", "pytorch_eigen_solver"):
        sys.modules.pop(name, None)
    import data_set, network_arch, pytorch_eigen_solver  # noqa: E401
    yield types.SimpleNamespace(data_set=data_set, network_arch=network_arch, pes=pytorch_eigen_solver)
    for name in ("data_set", "network_arch", "pytorch_eigen_solver"):
        sys.modules.pop(name, None)


def test_network_init_consumes_rng_like_reference(dropin):
    """torch.manual_seed(seed) -> MyNet(...) gives the reference's parameters (golden cfg1 was
    initialised by the reference's own network_arch under seed 11, rounded to fp32)."""
    g = load_golden("cfg1_ex1_50d_k3")
    torch.manual_seed(11)
    net = dropin.network_arch.MyNet(g["arch"], "Tanh", g["k"])
    assert type(net).__module__ == "network_arch" and type(net).__name__ == "MyNet"
    names = [n for n, _ in net.named_parameters()]
    assert names[0] == "nets.0.linears.0.weight" and names[-1] == "nets.2.linears.3.bias"
    for i, sub in enumerate(net.nets):
        for l, lin in enumerate(sub.linears):
            W, b = g["params"][i][l]
            assert np.array_equal(lin.weight.detach().numpy().astype(np.float32), W.astype(np.float32))
            assert np.array_equal(lin.bias.detach().numpy().astype(np.float32), b.astype(np.float32))


def test_shift_and_normalize(dropin):
    torch.manual_seed(0)
    net = dropin.network_arch.MyNet([4, 6, 1], "Tanh", 2).double()
    ds = dropin.data_set.data_set(np.random.default_rng(0).standard_normal((50, 4)), np.ones(50))
    ds.generate_minibatch(50, False)
    y0 = net(ds).detach()
    net.shift_and_normalize([y0[:, 0].mean(), y0[:, 1].mean()], [y0[:, 0].var(unbiased=False), y0[:, 1].var(unbiased=False)])
    y1 = net(ds).detach()
    assert torch.allclose(y1.mean(dim=0), torch.zeros(2, dtype=torch.float64), atol=1e-12)
    assert torch.allclose(y1.var(dim=0, unbiased=False), torch.ones(2, dtype=torch.float64), atol=1e-12)


def test_data_set_loader_and_minibatch(dropin, tmp_path, monkeypatch):
    rng = np.random.default_rng(1)
    X, w = rng.standard_normal((40, 3)), rng.uniform(0.5, 2.0, 40)
    f = tmp_path / "states.txt"
    with open(f, "w") as fh:
        fh.write("40 3\n")
        np.savetxt(fh, np.concatenate([X, w[:, None]], axis=1), fmt="%.10f")
    ds = dropin.data_set.data_set.from_file(str(f))
    assert ds.K == 40 and ds.tot_dim == 3 and ds.X_vec.dtype == torch.float64
    assert np.allclose(ds.X_vec.numpy(), X, atol=1e-10) and np.allclose(ds.weights.numpy(), w, atol=1e-10)
    random.seed(77)
    want = random.choices(range(40), cum_weights=np.arange(1, 41), k=25)
    random.seed(77)
    ds.generate_minibatch(25)                                  # native MT19937 path through the C ABI
    assert list(ds.active_indices) == want
    assert ds.x_batch.shape == (25, 3) and ds.x_batch.requires_grad
    assert torch.equal(ds.weights_minibatch(), ds.weights[want])
    follow = random.random()
    random.seed(77); random.choices(range(40), cum_weights=np.arange(1, 41), k=25)
    assert random.random() == follow                           # Python's stream continues unbroken
    monkeypatch.setenv("EPDE_PY_MINIBATCH", "1")
    random.seed(77)
    ds.generate_minibatch(25)
    assert list(ds.active_indices) == want
    ds.generate_minibatch(40, False)
    assert ds.batch_size == 40 and ds.x_batch.shape == (40, 3)


def _param_namespace():
    return types.SimpleNamespace(
        k=3, eig_w=[1.0, 0.8, 0.6, 0.3, 0.2, 0.1], md_data_flag=False, md_beta=1.0, beta=1.0, diffusion_coeff=0.0,
        stage_list=[0, 6, 10], batch_size_list=[500, 800], sort_eigvals_in_training=True,
        learning_rate_list=[5e-3, 2e-3], alpha_list=[20.0, 10.0], inner_arch_size_list=[20, 20, 20],
        activation_name="Tanh", train_max_step=10, load_init_model=False, init_model_name="x.pt",
        print_every_step=1, data_filename_prefix="states", eig_file_name_prefix="eig", log_filename="log.txt")


def test_constructor_matches_reference_contract(dropin, tmp_path, monkeypatch):
    z = np.load(os.path.join(GOLDEN, "dropin_run.npz"))
    (tmp_path / "data").mkdir()
    with open(tmp_path / "data" / "states.txt", "w") as fh:
        fh.write("%d %d\n" % z["X"].shape)
        np.savetxt(fh, np.concatenate([z["X"][:3], z["w"][:3, None]], axis=1), fmt="%.10f")
    monkeypatch.chdir(tmp_path)
    s = dropin.pes.eigen_solver(_param_namespace(), 5)
    assert s.tot_dim == 50 and s.k == 3 and s.beta == 1.0
    assert s.ij_list == [(0, 1), (0, 2), (1, 2)] and s.num_ij_pairs == 3
    assert torch.equal(s.diag_coeff, torch.ones(50, dtype=torch.float64))
    bad = _param_namespace(); bad.eig_w = [1.0, 1.0, 0.5]
    with pytest.raises(SystemExit):
        dropin.pes.eigen_solver(bad, 5)


@pytest.mark.gpu
@pytest.mark.parametrize("device_loop", [True, False], ids=["default_device_loop", "host_loop"])
def test_training_run_matches_reference_log(dropin, tmp_path, monkeypatch, device_loop):
    """Same seed, same data file: 10 Adam steps over two stages.  The reference log is fp64 CPU; ours is the fp32
    CUDA step with fp64 master parameters / Adam moments, so the trajectories agree to ~1e-5.
    device_loop=True is the DEFAULT `train()`: minibatch draw, step, fp64 Adam, averaging and eigenvalue statistics in
    one CUDA graph per step; False (EPDE_HOST_LOOP=1) is the statement-by-statement mirror of the reference's loop.
    Both are held to the same tolerance against the same reference log and checkpoints."""
    if device_loop:
        monkeypatch.delenv("EPDE_HOST_LOOP", raising=False)
    else:
        monkeypatch.setenv("EPDE_HOST_LOOP", "1")
    z = np.load(os.path.join(GOLDEN, "dropin_run.npz"))
    (tmp_path / "data").mkdir()
    with open(tmp_path / "data" / "states.txt", "w") as fh:
        fh.write("%d %d\n" % z["X"].shape)
        np.savetxt(fh, np.concatenate([z["X"], z["w"][:, None]], axis=1), fmt="%.10f")
    monkeypatch.chdir(tmp_path)
    solver = dropin.pes.eigen_solver(_param_namespace(), int(z["seed"]))
    solver.run()
    # Python's `random` ends where the reference leaves it: 6 draws of 500 + 4 of 800 from random.seed(seed)
    after = random.random()
    random.seed(int(z["seed"]))
    K = z["X"].shape[0]
    for b in [500] * 6 + [800] * 4:
        random.choices(range(K), cum_weights=np.arange(1, K + 1), k=b)
    assert random.random() == after
    ref = np.array([[float(v) for v in line.split()] for line in str(z["log"]).strip().splitlines()])
    got = np.loadtxt(tmp_path / "data" / "log.txt", ndmin=2)
    assert got.shape == ref.shape == (10, 7)
    assert np.array_equal(got[:, 0], ref[:, 0])
    scale = np.maximum(np.abs(ref[:, 1:]), 1e-2 * np.abs(ref[:, 1:2]))
    assert (np.abs(got[:, 1:] - ref[:, 1:]) / scale).max() < 2e-4
    for name in ["eig", "eig_averaged", "eig_stage1", "eig_stage2_averaged"]:
        m = torch.load(tmp_path / "data" / (name + ".pt"), weights_only=False)
        assert type(m).__module__ == "network_arch"
        flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).numpy()
        want = z["ck_" + name]
        assert flat.dtype == np.float64 and flat.shape == want.shape
        assert np.linalg.norm(flat - want) / np.linalg.norm(want) < 1e-4, name


@pytest.mark.gpu
def test_equal_stages_replay_the_graph_across_a_stage_end_forward(dropin, tmp_path, monkeypatch):
    """Two stages with the SAME (batch size, alpha) and K > B: the stage-end full-data forward
    (update_mean_and_var_of_model) runs between two replays of the same CUDA graph.  It must not move the workspace the
    graph captured (forward has its own; a re-allocation invalidates the graph) -- the device loop has to reproduce the
    host loop on this schedule."""
    z = np.load(os.path.join(GOLDEN, "dropin_run.npz"))
    (tmp_path / "data").mkdir()
    with open(tmp_path / "data" / "states.txt", "w") as fh:
        fh.write("%d %d\n" % z["X"].shape)
        np.savetxt(fh, np.concatenate([z["X"], z["w"][:, None]], axis=1), fmt="%.10f")
    monkeypatch.chdir(tmp_path)
    logs = {}
    for mode in ("device", "host"):
        if mode == "host":
            monkeypatch.setenv("EPDE_HOST_LOOP", "1")
        else:
            monkeypatch.delenv("EPDE_HOST_LOOP", raising=False)
        P = _param_namespace()
        P.stage_list, P.batch_size_list, P.learning_rate_list, P.alpha_list = [0, 5, 10], [400, 400], [5e-3, 5e-3], [20.0, 20.0]
        assert z["X"].shape[0] > 400
        solver = dropin.pes.eigen_solver(P, 31)
        solver.run()
        logs[mode] = np.loadtxt(tmp_path / "data" / "log.txt", ndmin=2)
        if mode == "device":
            assert solver.device_trainer._graph is not None                 # the graph path was actually taken
    a, b = logs["device"], logs["host"]
    assert a.shape == b.shape == (10, 7)
    scale = np.maximum(np.abs(b[:, 1:]), 1e-2 * np.abs(b[:, 1:2]))
    assert (np.abs(a[:, 1:] - b[:, 1:]) / scale).max() < 1e-4


@pytest.mark.gpu
def test_md_mode_training_run_matches_reference_log(dropin, tmp_path, monkeypatch):
    """`md_data_flag=True`: the reference builds an `MD_data_set`, reads ./data/ref_state.txt and runs
    the Kabsch alignment layer inside every net's forward (training steps AND the stage-end full-data passes).  The
    drop-in evaluates the layer once per data set on the GPU (`epde_md_align`) and trains on the aligned rows with the
    per-row Jacobian metric: 8 Adam steps over two stages against the log and checkpoints of the unmodified reference."""
    monkeypatch.delenv("EPDE_HOST_LOOP", raising=False)
    monkeypatch.delenv("EPDE_MD_NO_ALIGN", raising=False)
    z = np.load(os.path.join(GOLDEN, "dropin_run_md.npz"))
    (tmp_path / "data").mkdir()
    with open(tmp_path / "data" / "states.txt", "w") as fh:
        fh.write("%d %d\n" % z["X"].shape)
        np.savetxt(fh, np.concatenate([z["X"], z["w"][:, None]], axis=1), fmt="%.10f")
    with open(tmp_path / "data" / "ref_state.txt", "w") as fh:
        fh.write("%d\n%s\n" % (len(z["ref_index"]), " ".join(str(int(i)) for i in z["ref_index"])))
        np.savetxt(fh, z["ref_x"], fmt="%.8f")
    monkeypatch.chdir(tmp_path)
    md_beta = float(z["md_beta"])
    P = types.SimpleNamespace(
        k=2, eig_w=[1.0, 0.8], md_data_flag=True, md_beta=md_beta, beta=md_beta, diffusion_coeff=1e-5,
        stage_list=[0, 4, 8], batch_size_list=[300, 400], sort_eigvals_in_training=True,
        learning_rate_list=[5e-3, 2e-3], alpha_list=[20.0, 10.0], inner_arch_size_list=[20, 20, 20],
        activation_name="Tanh", train_max_step=8, load_init_model=False, init_model_name="x.pt",
        print_every_step=1, data_filename_prefix="states", eig_file_name_prefix="eig", log_filename="log.txt")
    solver = dropin.pes.eigen_solver(P, int(z["seed"]))
    solver.run()
    assert solver.engine.metric is not None and solver.engine.fused      # d = 30: the per-row metric runs on the fused FFMA2 kernels
    ref = np.array([[float(v) for v in line.split()] for line in str(z["log"]).strip().splitlines()])
    got = np.loadtxt(tmp_path / "data" / "log.txt", ndmin=2)
    assert got.shape == ref.shape == (8, 6)
    assert np.array_equal(got[:, 0], ref[:, 0])
    scale = np.maximum(np.abs(ref[:, 1:]), 1e-2 * np.abs(ref[:, 1:2]))
    assert (np.abs(got[:, 1:] - ref[:, 1:]) / scale).max() < 2e-4
    for name in ["eig", "eig_averaged", "eig_stage1", "eig_stage2_averaged"]:
        m = torch.load(tmp_path / "data" / (name + ".pt"), weights_only=False)
        flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).numpy()
        want = z["ck_" + name]
        assert flat.shape == want.shape
        assert np.linalg.norm(flat - want) / np.linalg.norm(want) < 1e-4, name
