"""Drop-in contract: the reference's `utils/train_nn.py` runs UNCHANGED, with
the reference's real `read_parameters.Param` reading a real `params.cfg`, against this repo's `pytorch_eigen_solver` /
`network_arch` / `data_set` modules -- and produces the same `log.txt` as the unmodified reference run from the same
directory at the same seed.  Build container only (needs /root/reference; no GPU there, so the numerical core is a CPU
stand-in on the oracle: tests/_cpu_engine_standin.py; the CUDA core is held to the same log by tests/test_dropin.py)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

from conftest import ROOT

REF = "/root/reference"
CFG = textwrap.dedent("""\
    [default]
    eig_file_name_prefix = eigen_vector
    log_filename = log.txt
    md_data_flag = False
    load_data_how_often = 1
    [SDE]
    dim = 6
    pot_id = 7
    beta = 1.0
    stiff_eps = 0.50
    delta_t = 0.001
    N_state = 600
    SDE_beta = 1.0
    data_filename_prefix = states_6d
    data_filename_prefix_validation = states_6d_test
    [NeuralNetArch]
    arch_size_list = 20,20,20
    activation_name = Tanh
    [Training]
    eig_k = 3
    train_max_step = 6
    load_init_model = False
    init_model_name = eigen_vector.pt
    eig_weight = 1.0, 0.8, 0.6, 0.3, 0.2, 0.1
    stage_list = 0, 4
    batch_size_list = 150, 250
    learning_rate_list = 5e-3, 2e-3
    alpha_list = 20.0, 10.0
    sort_eigvals_in_training = True
    print_every_step = 1
    """)


def _run(cwd, pythonpath, site_code):
    site = cwd / "_site"
    site.mkdir(exist_ok=True)
    (site / "sitecustomize.py").write_text(site_code)
    (site / "matplotlib").mkdir(exist_ok=True)                      # the reference imports matplotlib.pyplot and never uses it
    (site / "matplotlib" / "__init__.py").write_text("")
    (site / "matplotlib" / "pyplot.py").write_text("")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(site)] + pythonpath), OMP_NUM_THREADS="2")
    env.pop("EPDE_HOST_LOOP", None)
    res = subprocess.run([sys.executable, os.path.join(REF, "utils", "train_nn.py")], cwd=cwd, env=env, capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    return res.stdout


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "utils", "train_nn.py")), reason="reference tree only exists in the build container")
def test_unchanged_train_nn_with_real_param(tmp_path):
    rng = np.random.default_rng(3)
    X = rng.standard_normal((600, 6)); w = np.exp(0.3 * rng.standard_normal(600)); w /= w.mean()
    logs, ckpts = {}, {}
    for who in ("ours", "reference"):
        d = tmp_path / who
        (d / "data").mkdir(parents=True)
        (d / "params.cfg").write_text(CFG)
        with open(d / "data" / "states_6d.txt", "w") as fh:
            fh.write("600 6\n")
            np.savetxt(fh, np.concatenate([X, w[:, None]], axis=1), fmt="%.10f")
        if who == "ours":
            # our modules first; read_parameters (config parsing, out of scope) comes from the reference, unchanged
            out = _run(d, [os.path.join(ROOT, "eigenpde_nn_b200", "dropin"), ROOT, os.path.join(REF, "src")],
                       "import sys\nsys.path.insert(0, %r)\nimport _cpu_engine_standin as s\ns.install(fixed_time=1000)\n" % os.path.join(ROOT, "tests"))
            assert "B200 step" in out
        else:
            out = _run(d, [os.path.join(REF, "src")], "import time\ntime.time = lambda: 1000.0\n")
        assert "seed = 4905" in out                                  # 3905 + int(time.time()), utils/train_nn.py:12
        logs[who] = np.loadtxt(d / "data" / "log.txt", ndmin=2)
        names = sorted(f for f in os.listdir(d / "data") if f.endswith(".pt"))
        assert names == ["eigen_vector.pt", "eigen_vector_averaged.pt", "eigen_vector_stage1.pt",
                         "eigen_vector_stage1_averaged.pt", "eigen_vector_stage2.pt", "eigen_vector_stage2_averaged.pt"]
        sys.path.insert(0, os.path.join(ROOT, "eigenpde_nn_b200", "dropin") if who == "ours" else os.path.join(REF, "src"))
        try:
            for m in ("network_arch", "data_set"):
                sys.modules.pop(m, None)
            ckpts[who] = {n: torch.cat([p.detach().reshape(-1) for p in torch.load(d / "data" / n, weights_only=False).parameters()])
                          for n in names}
        finally:
            sys.path.pop(0)
            for m in ("network_arch", "data_set"):
                sys.modules.pop(m, None)
    a, b = logs["ours"], logs["reference"]
    assert a.shape == b.shape == (6, 7)                              # step, loss, non-penalty, penalty, 3 eigenvalues
    assert np.array_equal(a[:, 0], b[:, 0])
    assert np.allclose(a[:, 1:], b[:, 1:], rtol=2e-6, atol=2e-6), np.abs(a - b).max()       # %.6f log lines
    for n in ckpts["ours"]:
        x, y = ckpts["ours"][n], ckpts["reference"][n]
        assert x.dtype == y.dtype == torch.float64 and x.shape == y.shape
        assert float((x - y).norm() / y.norm()) < 1e-7, n
