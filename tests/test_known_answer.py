"""GPU, BASELINE config 2: "2-D double-well potential, k = 2 eigenpairs, checked against utils/FVD-2d.py finite-volume
eigenvalues".  A seeded Euler-Maruyama trajectory of the reference's pot_id = 4 (oracle/fvd.py restates
src/data_generator.py:17-86 and src/potentials.py), trained with the drop-in solver's default (device-resident) loop;
the stage means of the eigenvalue estimates must land on the finite-volume eigenvalues (BASELINE.md section 3:
0.68331, 1.02394; reproduced by oracle/fvd.py in tests/test_fvd_oracle.py)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import fvd

DROPIN = os.path.join(ROOT, "eigenpde_nn_b200", "dropin")
FVD_EIGS = (0.68331, 1.02394)


@pytest.mark.gpu
def test_2d_double_well_k2_lands_on_finite_volume_eigenvalues(tmp_path, monkeypatch):
    monkeypatch.syspath_prepend(DROPIN)
    for name in ("data_set", "network_arch", "pytorch_eigen_solver"):
        sys.modules.pop(name, None)
    import pytorch_eigen_solver
    monkeypatch.delenv("EPDE_HOST_LOOP", raising=False)
    # 10^6 Euler-Maruyama steps of size 0.005 (T = 5000 >> 1 / lambda_1 = 1.5), every 5th state kept
    X, w = fvd.sample_dw2d_fast(1_000_000, 0.005, beta=1.0, how_often=5, seed=7)
    assert X.shape == (200_000, 2)
    (tmp_path / "data").mkdir()
    with open(tmp_path / "data" / "states_2d.txt", "w") as fh:
        fh.write("%d %d\n" % X.shape)
        np.savetxt(fh, np.concatenate([X, w[:, None]], axis=1), fmt="%.10f")
    monkeypatch.chdir(tmp_path)
    P = types.SimpleNamespace(
        k=2, eig_w=[1.0, 0.8, 0.6], md_data_flag=False, md_beta=1.0, beta=1.0, diffusion_coeff=0.0,
        stage_list=[0, 2500, 3500], batch_size_list=[5000, 20000], sort_eigvals_in_training=True,
        learning_rate_list=[5e-3, 1e-3], alpha_list=[20.0, 20.0], inner_arch_size_list=[20, 20, 20],
        activation_name="Tanh", train_max_step=3500, load_init_model=False, init_model_name="x.pt",
        print_every_step=250, data_filename_prefix="states_2d", eig_file_name_prefix="eig", log_filename="log.txt")
    solver = pytorch_eigen_solver.eigen_solver(P, 2024)
    solver.run()
    assert solver.engine.fused
    tr = solver.device_trainer
    mean = tr.stats.cpu().numpy()[:2] / tr.steps_in_stage            # stage-2 means of the sorted estimates (:280-297)
    print("NN eigenvalues (stage-2 mean):", mean, " finite volumes:", FVD_EIGS)
    assert abs(mean[0] - FVD_EIGS[0]) / FVD_EIGS[0] < 0.06
    assert abs(mean[1] - FVD_EIGS[1]) / FVD_EIGS[1] < 0.06
    log = np.loadtxt(tmp_path / "data" / "log.txt", ndmin=2)
    assert log[-1, 3] < 5e-3                                          # the orthonormality penalty has decayed
    for name in ("data_set", "network_arch", "pytorch_eigen_solver"):
        sys.modules.pop(name, None)
