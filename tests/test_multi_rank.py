"""CPU, world_size = 2, gloo: the data-parallel protocol of the step (shard the index list, allreduce
the fp64 sums, allreduce the gradient shares) reproduces the single-rank result.  The per-rank compute
is the oracle here; on GPUs the same `parallel.two_phase_step` drives the CUDA phases over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden
from eigenpde_nn_b200.parallel import shard_bounds, two_phase_step
from oracle import closed_form


def test_shard_bounds_cover_the_batch():
    for n in [0, 1, 7, 1000, 1 << 20]:
        for world in [1, 2, 3, 8]:
            cuts = [shard_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, name, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden(name)
    k = g["k"]
    lo, hi = shard_bounds(len(g["idx"]), world, rank)
    idx = g["idx"][lo:hi]
    xb, wb = g["X"][idx], g["w"][idx]
    ns = 1 + 2 * k + k * (k + 1) // 2
    sums = torch.zeros(ns, dtype=torch.float64)
    n_par = sum(W.size + b.size for net in g["params"] for W, b in net)
    grad = torch.zeros(n_par, dtype=torch.float64)
    state = {}
    tri = [(i, j) for i in range(k) for j in range(i + 1)]

    def partial():      # layout of epde_num_sums: [W, S1, E, S2 lower triangle]
        p = closed_form.step(g["params"], xb, wb, g["diag_coeff"], g["beta"], g["alpha"], g["eig_w"], g["act"],
                             sums_only=True)["sums"]
        sums.copy_(torch.from_numpy(np.concatenate([[p["W"]], p["S1"], p["E"], [p["S2"][i, j] for i, j in tri]])))

    def finish():
        v = sums.numpy()
        S2 = np.zeros((k, k))
        for t, (i, j) in enumerate(tri):
            S2[i, j] = S2[j, i] = v[1 + 2 * k + t]
        out = closed_form.step(g["params"], xb, wb, g["diag_coeff"], g["beta"], g["alpha"], g["eig_w"], g["act"],
                               global_sums=(v[0], v[1:1 + k], S2, v[1 + k:1 + 2 * k]))
        state.update(out)
        grad.copy_(torch.from_numpy(closed_form.flatten(out["grads"])))

    two_phase_step(partial, finish, sums, grad)
    if rank == 0:
        q.put((state["loss"], state["eig_vals"], state["cvec"], grad.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["cfg1_ex1_50d_k3_weighted", "cfg2_dw_2d_k2"])
def test_two_rank_protocol_matches_single_rank(name):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    loss, eig, cvec, grad = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    g = load_golden(name)
    assert abs(loss - g["loss"]) <= 1e-10 * abs(g["loss"])
    assert np.allclose(eig, g["eig_vals"], rtol=1e-10) and list(cvec) == list(g["cvec"])
    ref = closed_form.flatten(g["grads"])
    assert np.linalg.norm(grad - ref) / np.linalg.norm(ref) < 1e-9


def test_peer_exchange_declines_without_nccl():
    """PeerExchange.create returns None (-> the two collective calls) when symmetric memory cannot be used: a gloo
    group on CPU here; `two_phase_step` then takes the dist.all_reduce path exercised above."""
    import torch.distributed as dist
    from eigenpde_nn_b200.parallel import PeerExchange
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29611", rank=0, world_size=1)
        made = True
    else:
        made = False
    try:
        assert PeerExchange.create(None, dist.group.WORLD, "cpu") is None
    finally:
        if made:
            dist.destroy_process_group()
