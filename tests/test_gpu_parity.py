"""GPU: the CUDA step (called through the C ABI) against the reference's golden outputs and the
fp64 oracle.  Tolerance: 1e-5 relative (north_star), metric of SURVEY.md section 8(c)."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import MD_ALIGNED, check_step_against, golden_names, load_golden, unflatten
from oracle import closed_form

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _engine(g, **kw):
    from eigenpde_nn_b200.engine import StepEngine
    return StepEngine(g["X"], g["w"], g["arch"], g["k"], g["act"], g["diag_coeff"], g["beta"], g["eig_w"], **kw)


def _flat(params):
    return torch.from_numpy(closed_form.flatten(params)).to("cuda", torch.float32)


def _rand_case(seed, d, k, arch_inner, K, B, act="Tanh", weights="lognormal", scale=0.5, alpha=20.0, beta=1.0,
               diag=None):
    rng = np.random.default_rng(seed)
    arch = [d] + list(arch_inner) + [1]
    X = (0.7 * rng.standard_normal((K, d))).astype(np.float32).astype(np.float64)
    w = np.ones(K) if weights == "ones" else np.exp(0.5 * rng.standard_normal(K))
    w = w.astype(np.float32).astype(np.float64)
    params = [[((scale * rng.standard_normal((arch[l + 1], arch[l]))).astype(np.float32).astype(np.float64),
                (scale * rng.standard_normal(arch[l + 1])).astype(np.float32).astype(np.float64))
               for l in range(len(arch) - 1)] for _ in range(k)]
    idx = rng.integers(0, K, B) if B else None
    eig_w = [1.0, 0.8, 0.6, 0.3, 0.2, 0.1][:k]
    diag = np.ones(d) if diag is None else diag
    return dict(X=X, w=w, params=params, idx=idx, arch=arch, k=k, act=act, diag_coeff=diag, beta=beta, alpha=alpha,
                eig_w=eig_w)


def _oracle(c):
    idx = c["idx"] if c["idx"] is not None else np.arange(c["X"].shape[0])
    return closed_form.step(c["params"], c["X"][idx], c["w"][idx], c["diag_coeff"], c["beta"], c["alpha"], c["eig_w"],
                            c["act"] if c["act"] in closed_form.ACTIVATIONS else "CELU")


def _check(c, ref, res):
    grads = unflatten(res["grad"].cpu().numpy().astype(np.float64), c["arch"], c["k"])
    return check_step_against(ref, res, TOL, grads)


# The all-positive Sigmoid activations make this case ill-conditioned in fp32 (the gradient is a
# small difference of large sums): the reference's own formulation run in fp32 (oracle/ref_port.py with
# .float()) is off by up to 9.4e-4 on it, the CUDA path by < 1e-4.  Every Tanh configuration named by
# BASELINE.json is held to 1e-5.
ILL_CONDITIONED = {"act_sigmoid_6d_k3": 2e-4}


@pytest.mark.parametrize("name", golden_names())
def test_golden_reference_outputs(name):
    g = load_golden(name)
    eng = _engine(g)
    res = eng.step(g["idx"], _flat(g["params"]), g["alpha"])
    grads = unflatten(res["grad"].cpu().numpy().astype(np.float64), g["arch"], g["k"])
    check_step_against(g, res, ILL_CONDITIONED.get(name, TOL), grads)


@pytest.mark.parametrize("act", ["Tanh", "Softplus", "LogSigmoid", "ELU", "CELU", "ReLU", "Tanhshrink", "NoSuchName"])
def test_generic_path_activations_vs_oracle(act):
    c = _rand_case(len(act) * 13, 9, 2, [16, 12, 8], 1500, 700, act=act, scale=0.6)
    eng = _engine(c)
    assert not eng.fused
    _check(c, _oracle(c), eng.step(c["idx"], _flat(c["params"]), c["alpha"]))


@pytest.mark.parametrize("arch_inner", [[20], [7, 5], [20, 20], [33, 17, 9, 5], [64, 64, 64, 64]])
def test_generic_path_architectures_vs_oracle(arch_inner):
    c = _rand_case(sum(arch_inner), 11, 3, arch_inner, 2000, 900, scale=0.35)
    eng = _engine(c)
    _check(c, _oracle(c), eng.step(c["idx"], _flat(c["params"]), c["alpha"]))


def test_wide_4x256_k4_vs_oracle():
    """BASELINE config 5 shape (4x256, 50-D, k=4) at a batch the oracle finishes in seconds."""
    c = _rand_case(256, 50, 4, [256, 256, 256, 256], 1500, 768, scale=0.09)
    eng = _engine(c)
    assert not eng.fused and eng.n_params == 842756
    _check(c, _oracle(c), eng.step(c["idx"], _flat(c["params"]), c["alpha"]))


def test_generic_path_chunking_large_batch():
    """More samples than one chunk (65536): chunk seams must not show."""
    c = _rand_case(5, 6, 2, [8, 8], 5000, 70001, scale=0.6)
    eng = _engine(c)
    _check(c, _oracle(c), eng.step(c["idx"], _flat(c["params"]), c["alpha"]))


@pytest.mark.parametrize("B", [1, 2, 3, 255, 256, 257, 511, 513, 1000, 4099])
def test_ragged_batch_sizes_vs_oracle(B):
    c = _rand_case(100 + B, 50, 3, [20, 20, 20], 3000, B)
    eng = _engine(c)
    res = eng.step(c["idx"], _flat(c["params"]), c["alpha"])
    if B < 8:
        # tiny batches: covariance is near-singular, compare the well-conditioned outputs only
        ref = _oracle(c)
        assert abs(res["tot_weight"] - ref["W"]) <= 1e-6 * ref["W"]
        assert np.allclose(res["mean"], ref["mean"], rtol=1e-4, atol=1e-5)
        return
    _check(c, _oracle(c), res)


@pytest.mark.parametrize("d,k", [(1, 1), (2, 2), (3, 2), (5, 4), (30, 2), (49, 3), (50, 6), (52, 5), (60, 3), (64, 3), (90, 2)])
def test_dims_and_net_counts_vs_oracle(d, k):
    c = _rand_case(7 * d + k, d, k, [20, 20, 20], 2000, 700, scale=0.4)
    eng = _engine(c)
    assert eng.fused == (d <= 60)            # beyond d = 60 the tile state exceeds 227 KB of shared memory
    _check(c, _oracle(c), eng.step(c["idx"], _flat(c["params"]), c["alpha"]))


def test_identity_index_full_dataset():
    c = _rand_case(5, 50, 3, [20, 20, 20], 1537, 0)
    eng = _engine(c)
    _check(c, _oracle(c), eng.step(None, _flat(c["params"]), c["alpha"]))


def test_md_style_diag_coeff_and_heavy_weights():
    rng = np.random.default_rng(3)
    beta = 1.0 / (0.001987191 * 300.0)
    c = _rand_case(33, 30, 2, [20, 20, 20], 3000, 1000, beta=beta, diag=np.full(30, 1e-5 * 1e7 * beta))
    c["w"] = np.exp(2.0 * rng.standard_normal(3000)).astype(np.float32).astype(np.float64)
    eng = _engine(c)
    _check(c, _oracle(c), eng.step(c["idx"], _flat(c["params"]), c["alpha"]))


def test_unsorted_flag():
    c = _rand_case(8, 50, 3, [20, 20, 20], 2000, 600)
    idx = c["idx"]
    ref = closed_form.step(c["params"], c["X"][idx], c["w"][idx], c["diag_coeff"], c["beta"], c["alpha"], c["eig_w"],
                           "Tanh", sort=False)
    eng = _engine(c, sort=False)
    res = eng.step(idx, _flat(c["params"]), c["alpha"])
    assert list(res["cvec"]) == [0, 1, 2]
    _check(c, ref, res)


def test_partial_finish_split_equals_two_shards():
    """The multi-GPU decomposition on one device: two shards' sums added, then two finishes added."""
    from eigenpde_nn_b200 import _lib
    c = _rand_case(21, 50, 3, [20, 20, 20], 4000, 1501)
    eng = _engine(c)
    flat = _flat(c["params"])
    full = eng.step(c["idx"], flat, c["alpha"])
    full_grad = full["grad"].clone()
    L = eng.lib
    idx_dev = torch.from_numpy(c["idx"]).cuda()
    shards = [idx_dev[:700].contiguous(), idx_dev[700:].contiguous()]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    sums = torch.zeros(eng.n_sums, dtype=torch.float64, device="cuda")
    part = torch.zeros_like(sums)
    wss = []
    for sh in shards:
        need = C.c_size_t()
        _lib.check(L.epde_workspace_bytes(C.byref(eng.desc), sh.numel(), C.byref(need)), "ws")
        ws = torch.empty(need.value + 256, dtype=torch.uint8, device="cuda")
        off = (-ws.data_ptr()) % 256
        ws = ws[off:off + need.value]
        wss.append(ws)
        _lib.check(L.epde_step_partial(C.byref(eng.desc), eng.X.data_ptr(), eng.w.data_ptr(), sh.data_ptr(), sh.numel(),
                                       flat.data_ptr(), eng.diag.data_ptr(), ws.data_ptr(), ws.numel(),
                                       part.data_ptr(), st), "partial")
        sums += part
    grad = torch.zeros(eng.n_params, dtype=torch.float32, device="cuda")
    g1 = torch.zeros_like(grad)
    out = torch.zeros(4 + 4 * eng.k, dtype=torch.float64, device="cuda")
    eigw = (C.c_double * eng.k)(*eng.eig_w)
    for sh, ws in zip(shards, wss):
        _lib.check(L.epde_step_finish(C.byref(eng.desc), eng.X.data_ptr(), eng.w.data_ptr(), sh.data_ptr(), sh.numel(),
                                      flat.data_ptr(), eng.diag.data_ptr(), eng.beta, c["alpha"], eigw,
                                      sums.data_ptr(), ws.data_ptr(), ws.numel(), out.data_ptr(), g1.data_ptr(), st),
                   "finish")
        grad += g1
    torch.cuda.synchronize()
    h = out.cpu().numpy()
    assert abs(h[0] - full["loss"]) <= 1e-12 * abs(full["loss"]) + 1e-9 * abs(full["loss"])
    rel = (grad - full_grad).norm() / full_grad.norm()
    assert rel < 2e-6, rel
    res = dict(loss=h[0], non_penalty=h[1], penalty=h[2], eig_vals=h[4:7], cvec=h[7:10].astype(np.int64), grad=grad)
    _check(c, _oracle(c), res)


def test_torch_operator_and_ctypes_bindings_agree(monkeypatch):
    """The step reaches the library through torch.ops.epde.* (extension shim, default) or through ctypes
    (EPDE_BINDING=ctypes): same C ABI, bit-identical results; the default engine uses the torch operator."""
    c = _rand_case(5, 50, 3, [20, 20, 20], 5000, 3001)
    flat = _flat(c["params"])
    e1 = _engine(c)
    assert e1.ops is not None
    a = e1.step(c["idx"], flat, c["alpha"]); ga = a["grad"].clone()
    monkeypatch.setenv("EPDE_BINDING", "ctypes")
    e2 = _engine(c)
    assert e2.ops is None
    b = e2.step(c["idx"], flat, c["alpha"])
    assert a["loss"] == b["loss"] and torch.equal(ga, b["grad"])
    _check(c, _oracle(c), a)


def test_deterministic_bitwise():
    """Bit-identical results run to run with DEFAULT flags: every group of the tensor-core kernels owns its staging buffer
    and accumulators and adds its tiles in a fixed order, the group images are summed in a fixed order."""
    c = _rand_case(77, 50, 3, [20, 20, 20], 5000, 40000)
    eng = _engine(c)
    flat = _flat(c["params"])
    a = eng.step(c["idx"], flat, c["alpha"]); ga = a["grad"].clone()
    b = eng.step(c["idx"], flat, c["alpha"])
    assert a["loss"] == b["loss"] and torch.equal(ga, b["grad"])


def test_full_size_properties():
    """BASELINE size (2^20 samples, 50-D, k=3): size-independent properties instead of the oracle."""
    rng = np.random.default_rng(1)
    K, B, d, k = 1 << 20, 1 << 20, 50, 3
    c = _rand_case(9, d, k, [20, 20, 20], 4, 4)
    X = (0.5 * rng.standard_normal((K, d))).astype(np.float32)
    w = np.exp(0.3 * rng.standard_normal(K)).astype(np.float32)
    from eigenpde_nn_b200.engine import StepEngine
    eng = StepEngine(X, w, c["arch"], k, "Tanh", np.ones(d), 1.0, c["eig_w"])
    flat = _flat(c["params"])
    idx = rng.integers(0, K, B)
    r1 = eng.step(idx, flat, 20.0); g1 = r1["grad"].clone()
    # (1) permutation invariance of the batch (summation order only)
    perm = rng.permutation(B)
    r2 = eng.step(idx[perm], flat, 20.0); g2 = r2["grad"].clone()
    assert abs(r1["loss"] - r2["loss"]) <= 1e-9 * abs(r1["loss"])
    assert (g1 - g2).norm() / g1.norm() < 2e-6
    # (2) scale invariance in the sample weights: w -> 4 w leaves every output unchanged
    eng.w.mul_(4.0)
    r3 = eng.step(idx, flat, 20.0)
    assert abs(r1["loss"] - r3["loss"]) <= 1e-9 * abs(r1["loss"])
    assert np.allclose(r1["eig_vals"], r3["eig_vals"], rtol=1e-9)
    assert (g1 - r3["grad"]).norm() / g1.norm() < 1e-6
    eng.w.mul_(0.25)
    # (3) the loss is invariant under f_i -> f_i + const: d loss / d b_L == 0
    g = unflatten(g1.cpu().numpy().astype(np.float64), c["arch"], k)
    for net in g:
        ginf = max(np.abs(W).max() for W, _ in net)
        assert abs(net[-1][1][0]) <= 1e-5 * ginf
    # (4) a subsample agrees with the oracle on the same rows
    sub = idx[:4096]
    rs = eng.step(sub, flat, 20.0)
    cs = dict(c); cs.update(X=X.astype(np.float64), w=w.astype(np.float64), idx=sub)
    _check(cs, _oracle(cs), rs)


def test_full_size_k6_shard_properties():
    """BASELINE config 4 shape: 50-D, k = 6, one GPU's shard of 2^21 samples.  Size-independent checks:
    two half-shards combined through partial/finish equal the single call; subsample vs oracle."""
    from eigenpde_nn_b200 import _lib
    from eigenpde_nn_b200.engine import StepEngine
    rng = np.random.default_rng(2)
    K, B, d, k = 1 << 21, 1 << 21, 50, 6
    c = _rand_case(19, d, k, [20, 20, 20], 4, 4)
    X = (0.5 * rng.standard_normal((K, d))).astype(np.float32)
    w = np.exp(0.3 * rng.standard_normal(K)).astype(np.float32)
    eng = StepEngine(X, w, c["arch"], k, "Tanh", np.ones(d), 1.0, c["eig_w"])
    flat = _flat(c["params"])
    idx = torch.from_numpy(rng.integers(0, K, B)).cuda()
    out, grad = eng.step_device(idx, B, flat, 20.0)
    full_out, full_grad = out.clone(), grad.clone()
    L = eng.lib
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    halves = [idx[: B // 2].contiguous(), idx[B // 2:].contiguous()]
    sums = torch.zeros(eng.n_sums, dtype=torch.float64, device="cuda")
    part = torch.zeros_like(sums)
    ws = eng._workspace(B)
    args = lambda sh: (C.byref(eng.desc), eng.X.data_ptr(), eng.w.data_ptr(), sh.data_ptr(), sh.numel(),
                       flat.data_ptr(), eng.diag.data_ptr())
    for sh in halves:
        _lib.check(L.epde_step_partial(*args(sh), ws.data_ptr(), ws.numel(), part.data_ptr(), st), "partial")
        sums += part
    g2 = torch.zeros_like(full_grad); g1 = torch.zeros_like(full_grad)
    o2 = torch.zeros_like(full_out)
    eigw = (C.c_double * k)(*eng.eig_w)
    for sh in halves:
        # finish re-uses the workspace of the matching partial: run partial again for this half first
        _lib.check(L.epde_step_partial(*args(sh), ws.data_ptr(), ws.numel(), part.data_ptr(), st), "partial")
        _lib.check(L.epde_step_finish(*args(sh), eng.beta, 20.0, eigw, sums.data_ptr(), ws.data_ptr(), ws.numel(),
                                      o2.data_ptr(), g1.data_ptr(), st), "finish")
        g2 += g1
    torch.cuda.synchronize()
    assert abs(float(o2[0]) - float(full_out[0])) <= 1e-9 * abs(float(full_out[0]))
    assert torch.equal(o2[4 + k:4 + 2 * k], full_out[4 + k:4 + 2 * k])           # cvec
    assert (g2 - full_grad).norm() / full_grad.norm() < 2e-6
    sub = idx[:4096].cpu().numpy()
    rs = eng.step(sub, flat, 20.0)
    cs = dict(c); cs.update(X=X.astype(np.float64), w=w.astype(np.float64), idx=sub)
    _check(cs, _oracle(cs), rs)


def test_cabi_rejects_bad_arguments_on_device():
    """Error behaviour of the compute entry points: integer codes, never a crash."""
    from eigenpde_nn_b200 import _lib
    c = _rand_case(3, 50, 3, [20, 20, 20], 500, 200)
    eng = _engine(c)
    flat = _flat(c["params"])
    L = eng.lib
    idx = torch.from_numpy(c["idx"]).cuda()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ws = eng._workspace(200)
    eigw = (C.c_double * 3)(*eng.eig_w)
    base = (C.byref(eng.desc), eng.X.data_ptr(), eng.w.data_ptr(), idx.data_ptr(), 200, flat.data_ptr(),
            eng.diag.data_ptr(), 1.0, 20.0, eigw)
    tail = (eng.out.data_ptr(), eng.grad.data_ptr(), st)
    assert L.epde_step(*base, ws.data_ptr(), 1024, *tail) == -4                   # workspace too small
    assert L.epde_step(*base, ws.data_ptr() + 16, ws.numel() - 16, *tail) == -5   # misaligned workspace
    assert L.epde_step(C.byref(eng.desc), eng.X.data_ptr() + 4, *base[2:], ws.data_ptr(), ws.numel(), *tail) == -5
    assert L.epde_step(*base[:4], 0, *base[5:], ws.data_ptr(), ws.numel(), *tail) == -6       # B = 0
    assert L.epde_step(*base[:9], None, ws.data_ptr(), ws.numel(), *tail) == -1               # eig_w NULL
    assert L.epde_step(*base, ws.data_ptr(), ws.numel(), *tail) == 0
    torch.cuda.synchronize()


def test_forward_only_vs_oracle():
    c = _rand_case(4, 50, 3, [20, 20, 20], 3000, 1001)
    eng = _engine(c)
    idx_dev = torch.from_numpy(c["idx"]).cuda()
    y = eng.forward(idx_dev, len(c["idx"]), _flat(c["params"])).cpu().numpy().astype(np.float64)
    ref = closed_form.forward_only(c["params"], c["X"][c["idx"]])
    assert np.abs(y - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())


def test_autograd_function_and_adam_step():
    """optimizer.zero_grad(); loss.backward(); optimizer.step() keeps working (reference :232-234)."""
    from eigenpde_nn_b200.engine import EigenLossFn
    from oracle import ref_port
    c = _rand_case(12, 50, 3, [20, 20, 20], 2000, 800)
    eng = _engine(c)
    model = ref_port.nested_to_model(c["params"], "Tanh")          # fp64 CPU parameters, as in the reference
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    opt.zero_grad()
    loss = EigenLossFn.apply(eng, c["idx"], c["alpha"], *model.parameters())
    loss.backward()
    ref = _oracle(c)
    assert abs(float(loss) - ref["loss"]) <= TOL * abs(ref["loss"])
    got = [[(lin.weight.grad.numpy(), lin.bias.grad.numpy()) for lin in net.linears] for net in model.nets]
    check_step_against(ref, dict(eng.last), TOL, got)
    before = [p.detach().clone() for p in model.parameters()]
    opt.step()
    assert any(not torch.equal(a, b) for a, b in zip(before, model.parameters()))


def test_fused_adam_matches_torch():
    from eigenpde_nn_b200 import _lib
    L = _lib.lib()
    n = 5643
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(n, generator=g)
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=5e-3)
    p = p0.cuda(); m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    for step in range(1, 6):
        grad = torch.randn(n, generator=g)
        ref_p.grad = grad.clone(); opt.step()
        gd = grad.cuda()
        _lib.check(L.epde_adam_step(p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, 5e-3, step,
                                    torch.cuda.current_stream().cuda_stream), "adam")
    assert torch.allclose(p.cpu(), ref_p.detach(), rtol=2e-5, atol=1e-6)


def test_device_side_averaging_and_eig_stats():
    """epde_average_accumulate / epde_eig_stats_accumulate against the host bookkeeping of the reference's
    `train` (record_model_parameters :114-118, running eigenvalue sums :280-282)."""
    import ctypes as C
    from eigenpde_nn_b200 import _lib
    L = _lib.lib()
    k, arch = 3, [50, 20, 20, 20, 1]
    desc = _lib.make_desc(arch, k, "Tanh", 52)
    n = C.c_int64()
    _lib.check(L.epde_param_count(C.byref(desc), C.byref(n)), "count")
    n, per = n.value, n.value // k
    g = torch.Generator().manual_seed(3)
    avg = torch.zeros(n, dtype=torch.float64, device="cuda")
    stats = torch.zeros(2 * k, dtype=torch.float64, device="cuda")
    want_avg = np.zeros(n)
    want_stats = np.zeros(2 * k)
    st = torch.cuda.current_stream().cuda_stream
    for it in range(4):
        p = torch.randn(n, generator=g)
        cvec = torch.randperm(k, generator=g).numpy()
        eig = torch.rand(k, generator=g, dtype=torch.float64).numpy()
        out = np.zeros(4 + 4 * k)
        out[4:4 + k] = eig
        out[4 + k:4 + 2 * k] = cvec
        pd, od = p.cuda(), torch.from_numpy(out).cuda()
        _lib.check(L.epde_average_accumulate(C.byref(desc), avg.data_ptr(), pd.data_ptr(), od.data_ptr(), st), "avg")
        _lib.check(L.epde_eig_stats_accumulate(C.byref(desc), stats.data_ptr(), od.data_ptr(), st), "stats")
        pn = p.numpy().astype(np.float64)
        for i in range(k):
            want_avg[i * per:(i + 1) * per] += pn[cvec[i] * per:(cvec[i] + 1) * per]
        want_stats[:k] += eig
        want_stats[k:] += eig ** 2
    assert np.array_equal(avg.cpu().numpy(), want_avg)          # fp64 sums of fp32 values: exact
    assert np.allclose(stats.cpu().numpy(), want_stats, rtol=1e-15)


@pytest.mark.parametrize("impl", [0, 1], ids=["ffma2", "tcgen05"])
@pytest.mark.parametrize("shape", [(50, 3, 4099), (2, 2, 2000), (30, 2, 1500), (50, 6, 2500), (7, 1, 300), (60, 3, 900)],
                         ids=lambda s: "d%d_k%d_B%d" % s)
def test_both_fused_families_vs_oracle(impl, shape):
    """The FFMA2 kernels (epde_fused.cuh, EPDE_FLAG_FFMA2) and the tcgen05 kernels (epde_tc_fused.cuh / epde_tc_pass2.cuh)
    compute the same step: both are held to 1e-5 against the fp64 oracle on the BASELINE shapes."""
    d, k, B = shape
    c = _rand_case(d * 7 + k, d, k, [20, 20, 20], 3000, B)
    eng = _engine(c, ffma2=(impl == 0))
    assert eng.fused
    _check(c, _oracle(c), eng.step(c["idx"], _flat(c["params"]), c["alpha"]))


@pytest.mark.parametrize("impl", [0, 1], ids=["ffma2", "tcgen05"])
def test_fused_families_full_size_properties(impl):
    """B = 2^20 (the bench shape): the loss decomposes into non-penalty + alpha * penalty, eigenvalue estimates are
    sorted, a second call is bit-identical (both families accumulate in a fixed order), and the two families agree
    with each other to fp32 round-off."""
    c = _rand_case(77, 50, 3, [20, 20, 20], 1 << 18, 1 << 20)
    eng = _engine(c, ffma2=(impl == 0))
    p = _flat(c["params"])
    r1 = eng.step(c["idx"], p, c["alpha"])
    g1 = r1["grad"].clone()
    r2 = eng.step(c["idx"], p, c["alpha"])
    assert torch.equal(g1, r2["grad"]) and r1["loss"] == r2["loss"]          # bit-identical run to run
    assert abs(r1["loss"] - (r1["non_penalty"] + c["alpha"] * r1["penalty"])) <= 1e-9 * abs(r1["loss"])
    assert list(r1["eig_vals"]) == sorted(r1["eig_vals"])
    eng2 = _engine(c, ffma2=(impl != 0))      # the other family
    r3 = eng2.step(c["idx"], p, c["alpha"])
    assert abs(r3["loss"] - r1["loss"]) <= 1e-5 * abs(r1["loss"])
    assert float((r3["grad"] - g1).norm() / g1.norm()) <= 1e-5


def _bench_case(k, K, B, seed):
    """The synthetic trajectory of bench.py (ring in (x0, x1), N(0, 0.1) elsewhere, lognormal weights)."""
    rng = np.random.default_rng(seed)
    d = 50
    X = np.sqrt(0.1) * rng.standard_normal((K, d))
    th = rng.uniform(0, 2 * np.pi, K); r = 1.0 + 0.3 * rng.standard_normal(K)
    X[:, 0], X[:, 1] = r * np.cos(th), r * np.sin(th)
    X = X.astype(np.float32).astype(np.float64)
    w = np.exp(0.5 * rng.standard_normal(K)); w = (w / w.mean()).astype(np.float32).astype(np.float64)
    arch = [d, 20, 20, 20, 1]
    params = [[((0.5 * rng.standard_normal((arch[l + 1], arch[l]))).astype(np.float32).astype(np.float64),
                (0.5 * rng.standard_normal(arch[l + 1])).astype(np.float32).astype(np.float64)) for l in range(4)]
              for _ in range(k)]
    idx = rng.integers(0, K, B)
    return dict(X=X, w=w, arch=arch, k=k, act="Tanh", params=params, idx=idx, diag_coeff=np.ones(d), beta=1.0, alpha=20.0,
                eig_w=[1.0, 0.8, 0.6, 0.3, 0.2, 0.1][:k])


_FULL_REF = {}          # (k, logB) -> (case, oracle result): the fp64 oracle over the full batch is computed once per size


@pytest.mark.parametrize("impl", [1, 0], ids=["tcgen05", "ffma2"])
@pytest.mark.parametrize("k,logB", [(3, 20), (6, 21)], ids=["cfg1_k3_B2e20", "cfg4_k6_shard_B2e21"])
def test_benchmarked_batch_sizes_full_batch_vs_oracle(impl, k, logB):
    """The configurations bench.py measures, at their FULL batch size, against the fp64 oracle on the same batch:
    k = 3, B = 2^20 (the headline line) and k = 6, B = 2^21 (one rank's shard of BASELINE config 4).  SURVEY H3 is about
    accumulation error growing with B: the 1e-5 bar is shown here, not inferred from a subsample."""
    if (k, logB) not in _FULL_REF:
        c = _bench_case(k, 1 << 19, 1 << logB, seed=100 + k)
        _FULL_REF[(k, logB)] = (c, _oracle(c))        # numpy fp64 matmuls over the whole batch: 20 - 80 s on the host
    c, ref = _FULL_REF[(k, logB)]
    eng = _engine(c, ffma2=(impl == 0))
    assert eng.fused
    res = eng.step(c["idx"], _flat(c["params"]), c["alpha"])
    _check(c, ref, res)


def test_md_align_kernel_and_metric_step_vs_reference():
    """SURVEY 8(f) N1 on the GPU: `epde_md_align` (aligned rows + per-row metric of the alignment Jacobian) against the
    oracle, then the step with `epde_desc.metric` against the outputs of the unmodified reference run in MD mode
    (alignment layer inside every net's forward)."""
    from eigenpde_nn_b200.engine import StepEngine, md_align as gpu_md_align
    from oracle import md_align
    g = load_golden(MD_ALIGNED)
    Xa, M = gpu_md_align(g["X"], g["ref_x"], g["ref_index"], g["diag_coeff"])
    d = g["arch"][0]
    rows = g["idx"][:64]
    xa_ref, M_ref = md_align.align_batch(g["X"][rows], g["ref_x"], g["ref_index"], g["diag_coeff"])
    assert np.abs(Xa[rows, :d].cpu().numpy() - xa_ref).max() <= 2e-6 * np.abs(xa_ref).max()       # fp32 storage
    assert np.abs(M[rows].cpu().numpy() - M_ref).max() <= 2e-6 * np.abs(M_ref).max()
    assert float(Xa[:, d:].abs().max()) == 0.0                                                       # padding columns
    for generic in (False, True):               # the fused kernels with the per-row metric, and the layer-wise family
        eng = StepEngine(Xa, torch.from_numpy(g["w"]), g["arch"], g["k"], g["act"], g["diag_coeff"], g["beta"], g["eig_w"],
                         metric=M, generic=generic)
        assert eng.fused == (not generic)
        res = eng.step(g["idx"], _flat(g["params"]), g["alpha"])
        grads = unflatten(res["grad"].cpu().numpy().astype(np.float64), g["arch"], g["k"])
        check_step_against(g, res, TOL, grads)


def _metric_case(seed, d, k, K, B, identity=False):
    """Random symmetric positive definite per-row metrics M_r = A_r A_r^T + 0.1 I on top of a random case."""
    c = _rand_case(seed, d, k, [20, 20, 20], K, 0 if identity else B, scale=0.4)
    rng = np.random.default_rng(seed + 1)
    A = 0.4 * rng.standard_normal((K, d, d))
    M = np.einsum("rij,rkj->rik", A, A) + 0.1 * np.eye(d)[None]
    c["metric"] = M.astype(np.float32).astype(np.float64)
    c["w"] = np.exp(1.0 * rng.standard_normal(K)).astype(np.float32).astype(np.float64)
    return c


def _oracle_metric(c):
    idx = c["idx"] if c["idx"] is not None else np.arange(c["X"].shape[0])
    return closed_form.step(c["params"], c["X"][idx], c["w"][idx], c["diag_coeff"], c["beta"], c["alpha"], c["eig_w"],
                            c["act"], metric=c["metric"][idx])


@pytest.mark.parametrize("d,k,B", [(30, 2, 1000), (30, 2, 257), (30, 3, 37), (6, 1, 300), (33, 2, 513), (40, 2, 300),
                                   (9, 6, 700), (30, 2, 10000)])
def test_per_row_metric_fused_kernels_vs_oracle(d, k, B):
    """The per-row metric (MD alignment layer) on the fused FFMA2 kernels: explicit u0 = W1^T g1, warp-wide
    p = M_r u0, the extra g1 (x) p outer-product phase; ragged batches, d below / above one warp of columns."""
    from eigenpde_nn_b200.engine import StepEngine
    c = _metric_case(900 + d + B, d, k, 1200, B)
    eng = StepEngine(c["X"], c["w"], c["arch"], c["k"], c["act"], c["diag_coeff"], c["beta"], c["eig_w"],
                     metric=torch.from_numpy(c["metric"]))
    assert eng.fused
    _check(c, _oracle_metric(c), eng.step(c["idx"], _flat(c["params"]), c["alpha"]))


def test_per_row_metric_identity_index_and_generic_family_agree():
    from eigenpde_nn_b200.engine import StepEngine
    c = _metric_case(77, 30, 2, 777, 0, identity=True)
    ref = _oracle_metric(c)
    for generic in (False, True):
        eng = StepEngine(c["X"], c["w"], c["arch"], c["k"], c["act"], c["diag_coeff"], c["beta"], c["eig_w"],
                         metric=torch.from_numpy(c["metric"]), generic=generic)
        assert eng.fused == (not generic)
        _check(c, ref, eng.step(None, _flat(c["params"]), c["alpha"]))


def test_per_row_metric_too_wide_for_the_fused_rows_falls_back():
    from eigenpde_nn_b200.engine import StepEngine
    c = _metric_case(5, 42, 2, 400, 300)
    eng = StepEngine(c["X"], c["w"], c["arch"], c["k"], c["act"], c["diag_coeff"], c["beta"], c["eig_w"],
                     metric=torch.from_numpy(c["metric"]))
    assert not eng.fused                        # u0 / p need d <= 40 shared-memory rows
    _check(c, _oracle_metric(c), eng.step(c["idx"], _flat(c["params"]), c["alpha"]))


def test_train_epilogue_equals_separate_calls():
    """epde_train_epilogue (one launch, device-side learning rate / step count) against epde_adam_step +
    epde_average_accumulate + epde_eig_stats_accumulate."""
    import ctypes as C
    from eigenpde_nn_b200 import _lib
    L = _lib.lib()
    k, arch = 3, [50, 20, 20, 20, 1]
    desc = _lib.make_desc(arch, k, "Tanh", 52)
    n = C.c_int64(); _lib.check(L.epde_param_count(C.byref(desc), C.byref(n)), "count"); n = n.value
    g = torch.Generator().manual_seed(5)
    p1 = torch.randn(n, generator=g).cuda(); p2 = p1.clone()
    m1 = torch.zeros(n, device="cuda"); v1 = torch.zeros(n, device="cuda"); m2 = m1.clone(); v2 = v1.clone()
    avg1 = torch.zeros(n, dtype=torch.float64, device="cuda"); avg2 = avg1.clone()
    st1 = torch.zeros(2 * k, dtype=torch.float64, device="cuda"); st2 = st1.clone()
    state = torch.tensor([5e-3, 0.0], dtype=torch.float64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for step in range(1, 5):
        grad = torch.randn(n, generator=g).cuda()
        out = torch.zeros(4 + 4 * k, dtype=torch.float64)
        out[4:4 + k] = torch.rand(k, generator=g, dtype=torch.float64)
        out[4 + k:4 + 2 * k] = torch.randperm(k, generator=g).double()
        out = out.cuda()
        _lib.check(L.epde_adam_step(p1.data_ptr(), grad.data_ptr(), m1.data_ptr(), v1.data_ptr(), n, 5e-3, step, s), "adam")
        _lib.check(L.epde_average_accumulate(C.byref(desc), avg1.data_ptr(), p1.data_ptr(), out.data_ptr(), s), "avg")
        _lib.check(L.epde_eig_stats_accumulate(C.byref(desc), st1.data_ptr(), out.data_ptr(), s), "stats")
        _lib.check(L.epde_train_epilogue(C.byref(desc), p2.data_ptr(), grad.data_ptr(), m2.data_ptr(), v2.data_ptr(),
                                         state.data_ptr(), avg2.data_ptr(), st2.data_ptr(), out.data_ptr(), s), "epilogue")
    assert float(state[1]) == 4.0
    assert torch.allclose(p1, p2, rtol=1e-6, atol=1e-7) and torch.allclose(m1, m2) and torch.allclose(v1, v2)
    assert torch.allclose(avg1, avg2, rtol=1e-6, atol=1e-7) and torch.equal(st1, st2)


def test_minibatch_pipeline_delivers_every_batch():
    from eigenpde_nn_b200.engine import MinibatchPipeline
    B = 4096
    pins = [torch.randint(0, 1 << 20, (B,), dtype=torch.int64).pin_memory() for _ in range(6)]
    pipe = MinibatchPipeline(B, "cuda:0")
    pipe.prefetch(pins[0])
    for i in range(6):
        got = pipe.acquire()
        chk = got.clone()                       # consume on the current stream
        pipe.release()
        if i + 1 < 6:
            pipe.prefetch(pins[i + 1])
        torch.cuda.synchronize()
        assert torch.equal(chk.cpu(), pins[i])


@pytest.mark.parametrize("d,k,B", [(50, 3, 30011), (2, 2, 777), (30, 6, 4099), (60, 1, 129)])
def test_gather_look_ahead_equals_plain_steps(d, k, B):
    """`step_device(..., next_idx=...)`: the rows of the next minibatch are gathered on a second stream while the step
    computes (epde_gather_ahead / epde_step_ahead, two sets of gathered rows, SMs reserved for the gather).  A chain of steps
    with fresh indices must give what the plain steps give (the reserve changes the CTA partition of the sums: round-off
    level differences only), also when the announced minibatch is not the one that comes."""
    c = _rand_case(4242 + d, d, k, [20, 20, 20], 60000, B)
    rng = np.random.default_rng(7)
    idxs = [torch.from_numpy(rng.integers(0, 60000, B)).cuda() for _ in range(6)]
    flat = _flat(c["params"])
    plain = _engine(c)
    ahead = _engine(c)
    assert ahead._ahead_ok()
    want = []
    for ix in idxs:
        o, g = plain.step_device(ix, ix.numel(), flat, c["alpha"])
        want.append((o.clone(), g.clone()))
    seq = [0, 1, 2, 3, 4, 5]
    for i, j in enumerate(seq):
        nxt = idxs[seq[i + 1]] if i + 1 < len(seq) else None
        if i == 2:
            nxt = idxs[0]                      # announce a minibatch that does not come: the step must gather for itself
        o, g = ahead.step_device(idxs[j], idxs[j].numel(), flat, c["alpha"], next_idx=nxt)
        wo, wg = want[j]
        assert float(((o[:4] - wo[:4]).abs() / wo[:4].abs()).max()) < 1e-9, (i, o[:4], wo[:4])
        assert float((g - wg).norm() / wg.norm()) < 2e-6, i
    c["idx"] = idxs[5].cpu().numpy()
    res = ahead.step(c["idx"], flat, c["alpha"])             # host-facing call after a look-ahead chain
    _check(c, _oracle(c), res)


@pytest.mark.parametrize("act", ["Softplus", "LogSigmoid", "ELU", "CELU", "ReLU"])
def test_fused_kernels_with_other_activations_vs_oracle(act):
    """phi, phi', phi'' as a parameter of the fused kernels.  The reference takes any torch.nn activation
    (src/network_arch.py:27-34); for the 20-20-20 architecture the FFMA2 kernels run every activation whose derivatives are
    functions of its value (Act<true> in csrc/epde_fused.cuh), against the oracle and against the layer-wise family."""
    from eigenpde_nn_b200.engine import StepEngine
    c = _rand_case(50 + len(act), 12, 3, [20, 20, 20], 3000, 1531, act=act, scale=0.5)
    ref = _oracle(c)
    tol = TOL
    outs = []
    for generic in (False, True):
        eng = StepEngine(c["X"], c["w"], c["arch"], c["k"], c["act"], c["diag_coeff"], c["beta"], c["eig_w"], generic=generic)
        assert eng.fused == (not generic)
        res = eng.step(c["idx"], _flat(c["params"]), c["alpha"])
        grads = unflatten(res["grad"].cpu().numpy().astype(np.float64), c["arch"], c["k"])
        check_step_against(ref, res, tol, grads)
        outs.append(res)
        y = eng.forward(torch.from_numpy(c["idx"]).cuda(), len(c["idx"]), _flat(c["params"])).cpu().numpy()
        want = closed_form.forward_only(c["params"], c["X"][c["idx"]], c["act"] if c["act"] in closed_form.ACTIVATIONS else "CELU")
        assert np.abs(y - want).max() <= 2e-5 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("act", ["Tanhshrink", "Sigmoid"])
def test_activations_that_stay_on_the_layer_wise_family(act):
    c = _rand_case(91, 12, 2, [20, 20, 20], 2000, 700, act=act, scale=0.5)
    eng = _engine(c)
    assert not eng.fused
    res = eng.step(c["idx"], _flat(c["params"]), c["alpha"])
    grads = unflatten(res["grad"].cpu().numpy().astype(np.float64), c["arch"], c["k"])
    # all-positive Sigmoid units: ill-conditioned in fp32 like the golden act_sigmoid_6d_k3 case (ILL_CONDITIONED)
    check_step_against(_oracle(c), res, 2e-4 if act == "Sigmoid" else TOL, grads)
