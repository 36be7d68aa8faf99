"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

Run:  python oracle/gen_golden.py            (needs /root/reference)

The reference (src/pytorch_eigen_solver.py) imports matplotlib, which is not
installed; an empty stub package is put on sys.path.  Nothing else is patched:
`eigen_solver.update_step` (src/pytorch_eigen_solver.py:191-236) is called as
shipped, on a solver whose `dataset`, `model`, `optimizer` and `v_in_jac`
attributes are set the way `run`/`train` set them (:335-379, :262), with an
SGD(lr=0) optimiser so the parameters stay fixed and `p.grad` can be read.
The minibatch indices come from the reference's own
`data_set.generate_minibatch` after `random.seed(seed)`.

Each fixture stores inputs (X, weights, parameters, scalars) and the
reference's outputs (eig_vals, cvec, loss, non_penalty, penalty, gradients,
active indices).  Test infrastructure only.
"""
from __future__ import annotations

import os
import random
import sys
import tempfile
import types

import numpy as np

REF = os.environ.get("EPDE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def import_reference():
    stub = tempfile.mkdtemp(prefix="mplstub_")
    os.makedirs(os.path.join(stub, "matplotlib"))
    for f in ("__init__.py", "pyplot.py"):
        open(os.path.join(stub, "matplotlib", f), "w").close()
    sys.path.insert(0, os.path.join(REF, "src"))
    sys.path.insert(0, stub)
    import data_set, network_arch, pytorch_eigen_solver  # noqa: E401
    return pytorch_eigen_solver, network_arch, data_set


def synth_data(case, rng):
    K, d = case["K"], case["d"]
    kind = case["data"]
    if kind == "ring50":       # (x0,x1) near the unit ring, rest N(0, 0.1)
        th = rng.uniform(0, 2 * np.pi, K)
        r = 1.0 + 0.3 * rng.standard_normal(K)
        X = np.sqrt(0.1) * rng.standard_normal((K, d))
        X[:, 0] = r * np.cos(th); X[:, 1] = r * np.sin(th)
    elif kind == "dw2":        # 2-D double well-ish cloud
        X = np.stack([np.sign(rng.standard_normal(K)) + 0.35 * rng.standard_normal(K),
                      0.6 * rng.standard_normal(K)], axis=1)
    elif kind == "md30":       # 10-atom template + noise (Angstrom)
        tmpl = 1.5 * rng.standard_normal((1, d))
        X = tmpl + np.sqrt(0.1) * rng.standard_normal((K, d))
    else:
        X = rng.standard_normal((K, d))
    wk = case["weights"]
    if wk == "ones":
        w = np.ones(K)
    elif wk == "lognormal":
        w = np.exp(0.5 * rng.standard_normal(K)); w /= w.mean()
    elif wk == "heavy":
        w = np.exp(2.0 * rng.standard_normal(K)); w /= w.mean()
    # inputs are rounded to fp32-representable values so that the fp32 device path and the
    # fp64 reference start from bit-identical numbers (and fixtures store them as float32)
    return X.astype(np.float32).astype(np.float64), w.astype(np.float32).astype(np.float64)


CASES = [
    dict(name="cfg1_ex1_50d_k3", d=50, k=3, arch=[20, 20, 20], act="Tanh", K=1500, B=1000, data="ring50",
         weights="ones", beta=1.0, alpha=20.0, eig_w=[1.0, 0.8, 0.6, 0.3, 0.2, 0.1], md=False, seed=11),
    dict(name="cfg1_ex1_50d_k3_weighted", d=50, k=3, arch=[20, 20, 20], act="Tanh", K=1500, B=777, data="ring50",
         weights="lognormal", beta=1.0, alpha=20.0, eig_w=[1.0, 0.8, 0.6], md=False, seed=12),
    dict(name="cfg2_dw_2d_k2", d=2, k=2, arch=[20, 20, 20], act="Tanh", K=3000, B=600, data="dw2",
         weights="ones", beta=1.0, alpha=20.0, eig_w=[1.0, 0.8], md=False, seed=13),
    dict(name="cfg3_md_30d_k2", d=30, k=2, arch=[20, 20, 20], act="Tanh", K=3000, B=1000, data="md30",
         weights="heavy", beta=1.0 / (0.001987191 * 300.0), alpha=20.0, eig_w=[1.0, 0.8], md=True,
         diffusion_coeff=1e-5, seed=14),
    # MD mode WITH the Kabsch alignment layer (src/data_set.py:126-160) inside every net's forward
    dict(name="cfg3_md_30d_k2_aligned", d=30, k=2, arch=[20, 20, 20], act="Tanh", K=1200, B=500, data="md30",
         weights="heavy", beta=1.0 / (0.001987191 * 300.0), alpha=20.0, eig_w=[1.0, 0.8], md=True,
         diffusion_coeff=1e-5, seed=24, align=True, ref_index=[0, 2, 5, 7, 8]),
    dict(name="cfg4_50d_k6", d=50, k=6, arch=[20, 20, 20], act="Tanh", K=1500, B=1024, data="ring50",
         weights="lognormal", beta=1.0, alpha=20.0, eig_w=[1.0, 0.8, 0.6, 0.3, 0.2, 0.1], md=False, seed=15),
    dict(name="cfg5_wide128_50d_k4", d=50, k=4, arch=[128, 128, 128, 128], act="Tanh", K=600, B=192, data="ring50",
         weights="ones", beta=1.0, alpha=20.0, eig_w=[1.0, 0.8, 0.6, 0.3], md=False, seed=16, init_scale=0.17, grad_f32=True),
    dict(name="k1_50d_single", d=50, k=1, arch=[20, 20, 20], act="Tanh", K=2000, B=300, data="ring50",
         weights="lognormal", beta=1.0, alpha=20.0, eig_w=[1.0], md=False, seed=17),
    dict(name="arch_10d_k2_two_hidden", d=10, k=2, arch=[16, 12], act="Tanh", K=2000, B=257, data="gauss",
         weights="lognormal", beta=2.0, alpha=5.0, eig_w=[1.0, 0.5], md=False, seed=18),
    dict(name="act_softplus_8d_k2", d=8, k=2, arch=[20, 20, 20], act="Softplus", K=2000, B=400, data="gauss",
         weights="ones", beta=1.0, alpha=20.0, eig_w=[1.0, 0.8], md=False, seed=19),
    dict(name="act_celu_fallback_8d_k2", d=8, k=2, arch=[12, 12], act="NoSuchAct", K=1000, B=300, data="gauss",
         weights="ones", beta=1.0, alpha=10.0, eig_w=[1.0, 0.8], md=False, seed=20),
    dict(name="act_sigmoid_6d_k3", d=6, k=3, arch=[20, 20, 20], act="Sigmoid", K=1000, B=300, data="gauss",
         weights="lognormal", beta=1.0, alpha=20.0, eig_w=[1.0, 0.8, 0.6], md=False, seed=21),
]


def make_solver(pes, na, ds, case, X, w, workdir):
    os.makedirs(os.path.join(workdir, "data"), exist_ok=True)
    with open(os.path.join(workdir, "data", "states.txt"), "w") as f:
        f.write("%d %d\n" % X.shape)
        np.savetxt(f, np.concatenate([X[:2], w[:2, None]], axis=1), fmt="%.10f")
    P = types.SimpleNamespace(
        k=case["k"], eig_w=list(case["eig_w"]), md_data_flag=case["md"], md_beta=case["beta"], beta=case["beta"],
        diffusion_coeff=case.get("diffusion_coeff", 0.0),
        stage_list=[0, 10], batch_size_list=[case["B"]], sort_eigvals_in_training=True,
        learning_rate_list=[0.0], alpha_list=[case["alpha"]], inner_arch_size_list=list(case["arch"]),
        activation_name=case["act"], train_max_step=10, load_init_model=False, init_model_name="x.pt",
        print_every_step=1, data_filename_prefix="states", eig_file_name_prefix="eig", log_filename="log.txt")
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        solver = pes.eigen_solver(P, case["seed"])
    finally:
        os.chdir(cwd)
    import torch
    solver.dataset = ds.data_set(X, w)            # plain data_set also for the MD-shaped case
    arch = [X.shape[1]] + list(case["arch"]) + [1]
    solver.model = na.MyNet(arch, case["act"], case["k"])
    if "init_scale" in case:                        # keep wide nets out of tanh saturation
        with torch.no_grad():
            for p in solver.model.parameters():
                p.mul_(case["init_scale"] / 0.5)
    with torch.no_grad():                           # fp32-representable parameters (see synth_data)
        for p in solver.model.parameters():
            p.copy_(p.float())
    solver.model.double()
    solver.optimizer = torch.optim.SGD(solver.model.parameters(), lr=0.0)
    solver.v_in_jac = torch.ones(case["B"], dtype=torch.float64)
    return solver


def run_case(pes, na, ds, case):
    import torch
    rng = np.random.default_rng(case["seed"])
    X, w = synth_data(case, rng)
    with tempfile.TemporaryDirectory() as wd:
        solver = make_solver(pes, na, ds, case, X, w, wd)
    extra = {}
    if case.get("align"):                           # MD_data_set + reference configuration, as `run` does (:343-346)
        solver.dataset = ds.MD_data_set(X, w)
        ri = list(case["ref_index"])
        ref = X[0].reshape(-1, 3)[ri]
        ref = (ref - ref.mean(axis=0, keepdims=True)).astype(np.float32).astype(np.float64)
        solver.dataset.ref_num_atoms = len(ri)
        solver.dataset.ref_index = ri
        solver.dataset.ref_x = torch.from_numpy(ref).double()
        extra = dict(ref_x=ref, ref_index=np.asarray(ri, np.int64))
    random.seed(case["seed"] + 1000)
    eig, cvec, loss, nonpen, pen = solver.update_step(case["B"], case["alpha"])
    idx = np.asarray(solver.dataset.active_indices, dtype=np.int64)
    gdt = np.float32 if case.get("grad_f32") else np.float64
    out = dict(X=X.astype(np.float32), w=w.astype(np.float32), idx=idx, eig_vals=np.asarray(eig, np.float64), cvec=np.asarray(cvec, np.int64),
               loss=np.float64(loss), non_penalty=np.float64(nonpen), penalty=np.float64(pen),
               diag_coeff=solver.diag_coeff.numpy().copy(), beta=np.float64(solver.beta),
               alpha=np.float64(case["alpha"]), eig_w=np.asarray(case["eig_w"], np.float64),
               k=np.int64(case["k"]), act=np.str_(case["act"]), minibatch_seed=np.int64(case["seed"] + 1000),
               arch=np.asarray([X.shape[1]] + list(case["arch"]) + [1], np.int64), **extra)
    if case.get("align"):                           # the aligned minibatch the nets saw (for the oracle's own check)
        with torch.no_grad():
            out["x_aligned"] = solver.dataset.align().numpy().copy()
    for i, net in enumerate(solver.model.nets):
        for l, lin in enumerate(net.linears):
            out[f"W_{i}_{l}"] = lin.weight.detach().numpy().astype(np.float32)
            out[f"b_{i}_{l}"] = lin.bias.detach().numpy().astype(np.float32)
            out[f"gW_{i}_{l}"] = lin.weight.grad.numpy().astype(gdt)
            out[f"gb_{i}_{l}"] = lin.bias.grad.numpy().astype(gdt)
    return out


def index_streams():
    """Reference minibatch streams: consecutive generate_minibatch calls after random.seed."""
    import data_set as ds
    out = {}
    for n, (seed, K, B, reps) in enumerate([(3214, 20000, 500, 3), (7, 13, 50, 2), (1792447379, 5000000, 4000, 2),
                                            (2 ** 40 + 17, 12345677, 1000, 1), (0, 1, 10, 1)]):
        data = ds.data_set(np.zeros((K, 1)), np.ones(K))
        random.seed(seed)
        streams = []
        for _ in range(reps):
            data.generate_minibatch(B)
            streams.append(np.asarray(data.active_indices, dtype=np.int64))
        out[f"s{n}_meta"] = np.asarray([seed, K, B, reps], dtype=np.int64)
        out[f"s{n}_idx"] = np.stack(streams)
    return out


def dropin_run(pes):
    """A short end-to-end training run of the UNMODIFIED reference (eigen_solver.run) on a small
    data file: its log.txt lines and final checkpoints pin the drop-in facade."""
    import torch
    rng = np.random.default_rng(99)
    case = dict(K=3000, d=50, data="ring50", weights="lognormal")
    X, w = synth_data(case, rng)
    X, w = np.round(X, 10), np.round(w, 10)          # what "%.10f" text round-trips to
    seed = 4321
    P = types.SimpleNamespace(
        k=3, eig_w=[1.0, 0.8, 0.6, 0.3, 0.2, 0.1], md_data_flag=False, md_beta=1.0, beta=1.0, diffusion_coeff=0.0,
        stage_list=[0, 6, 10], batch_size_list=[500, 800], sort_eigvals_in_training=True,
        learning_rate_list=[5e-3, 2e-3], alpha_list=[20.0, 10.0], inner_arch_size_list=[20, 20, 20],
        activation_name="Tanh", train_max_step=10, load_init_model=False, init_model_name="x.pt",
        print_every_step=1, data_filename_prefix="states", eig_file_name_prefix="eig", log_filename="log.txt")
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as wd:
        os.makedirs(os.path.join(wd, "data"))
        with open(os.path.join(wd, "data", "states.txt"), "w") as f:
            f.write("%d %d\n" % X.shape)
            np.savetxt(f, np.concatenate([X, w[:, None]], axis=1), fmt="%.10f")
        os.chdir(wd)
        try:
            solver = pes.eigen_solver(P, seed)
            solver.run()
            log = open(os.path.join(wd, "data", "log.txt")).read()
            ck = {}
            for name in ["eig", "eig_averaged", "eig_stage1", "eig_stage2_averaged"]:
                m = torch.load(os.path.join(wd, "data", name + ".pt"), weights_only=False)
                ck[name] = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).numpy()
        finally:
            os.chdir(cwd)
    return dict(X=X, w=w, seed=np.int64(seed), log=np.str_(log), **{"ck_" + k: v for k, v in ck.items()})


def dropin_run_md(pes):
    """The same for MD mode (`md_data_flag=True`): `run` builds an `MD_data_set`, reads ./data/ref_state.txt and the
    Kabsch alignment layer sits inside every net's forward, also in the stage-end full-data passes."""
    import torch
    rng = np.random.default_rng(77)
    case = dict(K=900, d=30, data="md30", weights="lognormal")
    X, w = synth_data(case, rng)
    X, w = np.round(X, 10), np.round(w, 10)
    ri = [0, 2, 5, 7, 8]
    ref = X[0].reshape(-1, 3)[ri]
    ref = np.round(ref - ref.mean(axis=0, keepdims=True), 8)
    seed = 2468
    md_beta = 1.0 / (0.001987191 * 300.0)
    P = types.SimpleNamespace(
        k=2, eig_w=[1.0, 0.8], md_data_flag=True, md_beta=md_beta, beta=md_beta, diffusion_coeff=1e-5,
        stage_list=[0, 4, 8], batch_size_list=[300, 400], sort_eigvals_in_training=True,
        learning_rate_list=[5e-3, 2e-3], alpha_list=[20.0, 10.0], inner_arch_size_list=[20, 20, 20],
        activation_name="Tanh", train_max_step=8, load_init_model=False, init_model_name="x.pt",
        print_every_step=1, data_filename_prefix="states", eig_file_name_prefix="eig", log_filename="log.txt")
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as wd:
        os.makedirs(os.path.join(wd, "data"))
        with open(os.path.join(wd, "data", "states.txt"), "w") as f:
            f.write("%d %d\n" % X.shape)
            np.savetxt(f, np.concatenate([X, w[:, None]], axis=1), fmt="%.10f")
        with open(os.path.join(wd, "data", "ref_state.txt"), "w") as f:
            f.write("%d\n%s\n" % (len(ri), " ".join(str(i) for i in ri)))
            np.savetxt(f, ref, fmt="%.8f")
        os.chdir(wd)
        try:
            solver = pes.eigen_solver(P, seed)
            solver.run()
            log = open(os.path.join(wd, "data", "log.txt")).read()
            ck = {}
            for name in ["eig", "eig_averaged", "eig_stage1", "eig_stage2_averaged"]:
                m = torch.load(os.path.join(wd, "data", name + ".pt"), weights_only=False)
                ck[name] = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).numpy()
        finally:
            os.chdir(cwd)
    return dict(X=X, w=w, seed=np.int64(seed), log=np.str_(log), ref_x=ref, ref_index=np.asarray(ri, np.int64),
                md_beta=np.float64(md_beta), **{"ck_" + k: v for k, v in ck.items()})


def main():
    pes, na, ds = import_reference()
    os.makedirs(OUT, exist_ok=True)
    only = os.environ.get("EPDE_GOLDEN_ONLY")      # regenerate a single fixture without touching the others
    for case in CASES:
        if only and case["name"] != only:
            continue
        out = run_case(pes, na, ds, case)
        path = os.path.join(OUT, case["name"] + ".npz")
        np.savez_compressed(path, **out)
        print("%-28s loss=%.6f eig=%s cvec=%s  %.0f KB" % (case["name"], out["loss"], np.round(out["eig_vals"], 4),
                                                         out["cvec"], os.path.getsize(path) / 1024))
    if only == "dropin_run_md" or not only:
        np.savez_compressed(os.path.join(OUT, "dropin_run_md.npz"), **dropin_run_md(pes))
        print("MD-mode drop-in run written")
    if only:
        return
    np.savez_compressed(os.path.join(OUT, "index_streams.npz"), **index_streams())
    print("index streams written")
    np.savez_compressed(os.path.join(OUT, "dropin_run.npz"), **dropin_run(pes))
    print("drop-in run written")


if __name__ == "__main__":
    main()
