"""Worker of tests/test_gpu_multi_rank.py (launched with torch.distributed.run, one rank per GPU): the N-rank sharded
step -- both exchanges: the one-shot NVLink peer-memory kernel and the NCCL fallback -- against a single-rank step on
the union batch and against the fp64 oracle.  Rank 0 writes a JSON verdict to the path in argv[1]."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from conftest import check_step_against, unflatten
    from eigenpde_nn_b200.engine import StepEngine
    from eigenpde_nn_b200.parallel import shard_bounds
    from oracle import closed_form
    import test_gpu_parity as T
    verdict = {"world": world, "cases": []}
    # "p2p": both exchanges inside the step's own kernels (epde_step_sharded); "p2p-unfolded": the separate one-shot exchange
    # kernel between the phases; "nccl": two NCCL all-reduces
    for (d, k, B, force) in [(50, 3, 40001, "p2p"), (50, 3, 40001, "p2p-unfolded"), (50, 3, 40001, "nccl"), (30, 2, 9001, "p2p"),
                             (50, 6, 20000, "p2p"), (2, 2, 777, "p2p")]:
        os.environ["EPDE_EXCHANGE"] = force.split("-")[0]
        os.environ["EPDE_EXCHANGE_FOLD"] = "0" if force.endswith("unfolded") else "1"
        c = T._rand_case(11 * d + k, d, k, [20, 20, 20], 50000, B)
        eng = StepEngine(c["X"], c["w"], c["arch"], k, c["act"], c["diag_coeff"], c["beta"], c["eig_w"], device=dev,
                         process_group=dist.group.WORLD)
        flat = torch.from_numpy(closed_form.flatten(c["params"])).to(dev, torch.float32)
        lo, hi = shard_bounds(B, world, rank)
        idx_dev = torch.from_numpy(c["idx"][lo:hi]).to(dev)
        worst = 0.0
        for rep in range(3):                        # repeated: the peer exchange alternates two parities
            out, grad = eng.step_device(idx_dev, hi - lo, flat, c["alpha"])
        h = out.cpu().numpy()
        res = dict(loss=h[0], non_penalty=h[1], penalty=h[2], eig_vals=h[4:4 + k], cvec=h[4 + k:4 + 2 * k].astype(np.int64))
        ref = T._oracle(c)
        grads = unflatten(grad.cpu().numpy().astype(np.float64), c["arch"], k)
        worst = check_step_against(ref, res, 1e-5, grads)
        single = StepEngine(c["X"], c["w"], c["arch"], k, c["act"], c["diag_coeff"], c["beta"], c["eig_w"], device=dev)
        o1, g1 = single.step_device(torch.from_numpy(c["idx"]).to(dev), B, flat, c["alpha"])
        rel = float((grad - g1).norm() / g1.norm())
        same = torch.stack([grad.double().sum(), out[0]])
        gathered = [torch.zeros_like(same) for _ in range(world)]
        dist.all_gather(gathered, same)
        identical = all(torch.equal(gathered[0], t) for t in gathered)      # every rank holds bit-identical results
        used = ("peer-folded" if eng._fold_exchange() else "peer") if eng._exchange is not None else "nccl"
        ahead_rel = None
        if eng._ahead_ok():
            # the gather look-ahead on every rank: a chain of steps that announce the next slice; the last one must give
            # what the plain sharded step gave on the same slice
            g_plain, o_plain = grad.clone(), out.clone()      # (the engine owns `grad` / `out`: the chain overwrites them)
            rng2 = np.random.default_rng(1000 + rank)
            others = [torch.from_numpy(rng2.integers(0, c["X"].shape[0], hi - lo)).to(dev) for _ in range(2)]
            chain = others + [idx_dev]
            for i, ix in enumerate(chain):
                oa, ga = eng.step_device(ix, hi - lo, flat, c["alpha"], next_idx=chain[i + 1] if i + 1 < len(chain) else None)
            ahead_rel = float((ga - g_plain).norm() / g_plain.norm())
            assert ahead_rel <= 2e-6 and float((oa[0] - o_plain[0]).abs() / o_plain[0].abs()) <= 1e-9, (ahead_rel, oa[0], o_plain[0])
        verdict["cases"].append(dict(d=d, k=k, B=B, requested=force, exchange=used, worst_vs_oracle=float(worst),
                                     grad_vs_single_rank=rel, identical_on_all_ranks=bool(identical), look_ahead_vs_plain=ahead_rel))
        assert rel <= 1e-5 and identical, verdict["cases"][-1]
        del eng, single
    if rank == 0:
        json.dump(verdict, open(sys.argv[1], "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
