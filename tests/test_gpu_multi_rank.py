"""GPU, >= 2 devices: the sharded step on real hardware -- N ranks over NCCL / the NVLink
peer-memory exchange kernel against the single-rank step on the same global batch and against the oracle.  Skipped on a
single-GPU box (bench.py's `parity` key carries the same check into every N > 1 bench run)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
def test_n_rank_step_matches_single_rank_and_oracle(tmp_path):
    world = min(torch.cuda.device_count(), 8)
    out = tmp_path / "verdict.json"
    env = dict(os.environ)
    env.pop("EPDE_EXCHANGE", None)
    env.pop("EPDE_EXCHANGE_FOLD", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "_mp_gpu_worker.py"), str(out)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    v = json.load(open(out))
    assert v["world"] == world and len(v["cases"]) == 6
    for c in v["cases"]:
        assert c["worst_vs_oracle"] <= 1e-5 and c["grad_vs_single_rank"] <= 1e-5 and c["identical_on_all_ranks"], c
    assert any(c["exchange"] == "nccl" for c in v["cases"])          # the fallback path was exercised
    # on NVLink boxes: the exchanges folded into the step's kernels, and the separate exchange kernel
    if any(c["exchange"].startswith("peer") for c in v["cases"]):
        assert {"peer-folded", "peer"} <= {c["exchange"] for c in v["cases"]}
