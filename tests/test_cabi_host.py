"""CPU: the C-ABI library loads, exports what include/eigenpde_b200.h declares, and its host-side
entry points (descriptor checks, sizes, MT19937 minibatch stream) behave.  No compute calls."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import minibatch


@pytest.fixture(scope="module")
def L():
    from eigenpde_nn_b200 import build, _lib
    build.build()
    return _lib.lib()


def test_exports_match_header(L):
    from eigenpde_nn_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "eigenpde_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(epde_\w+)\s*\(", hdr, flags=re.M))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.epde_abi_version() == _lib.ABI_VERSION == 4
    assert L.epde_error_string(0) == b"ok"
    assert b"workspace" in L.epde_error_string(-4)


def test_param_count_and_sums(L):
    from eigenpde_nn_b200 import _lib
    for arch, k, want in [([50, 20, 20, 20, 1], 3, 5643), ([2, 20, 20, 20, 1], 2, 1842), ([30, 20, 20, 20, 1], 2, 2962),
                          ([50, 20, 20, 20, 1], 6, 11286), ([50, 256, 256, 256, 256, 1], 4, 842756)]:
        D = _lib.make_desc(arch, k, "Tanh", (arch[0] + 3) // 4 * 4)
        n = C.c_int64()
        assert L.epde_param_count(C.byref(D), C.byref(n)) == 0 and n.value == want
        ns = C.c_int32()
        assert L.epde_num_sums(C.byref(D), C.byref(ns)) == 0 and ns.value == 1 + 2 * k + k * (k + 1) // 2
    D = _lib.make_desc([50, 20, 20, 20, 1], 3, "Tanh", 52)
    assert L.epde_has_fused_path(C.byref(D)) == 1
    D = _lib.make_desc([50, 20, 20, 20, 1], 3, "Softplus", 52)
    assert L.epde_has_fused_path(C.byref(D)) == 1          # FFMA2 kernels with the activation as a parameter
    for name in ("Tanhshrink", "Sigmoid"):                  # stay on the layer-wise family
        D = _lib.make_desc([50, 20, 20, 20, 1], 3, name, 52)
        assert L.epde_has_fused_path(C.byref(D)) == 0
    D = _lib.make_desc([50, 256, 256, 256, 256, 1], 4, "Tanh", 52)
    assert L.epde_has_fused_path(C.byref(D)) == 0


def test_descriptor_validation(L):
    from eigenpde_nn_b200 import _lib
    n = C.c_int64()
    D = _lib.make_desc([50, 20, 20, 20, 1], 3, "Tanh", 50)       # d_pad not a multiple of 4
    assert L.epde_param_count(C.byref(D), C.byref(n)) == -5
    D = _lib.make_desc([50, 20, 20, 20, 2], 3, "Tanh", 52)       # output width must be 1
    assert L.epde_param_count(C.byref(D), C.byref(n)) == -2
    D = _lib.make_desc([50, 20, 20, 20, 1], 9, "Tanh", 52)       # k > EPDE_MAX_K
    assert L.epde_param_count(C.byref(D), C.byref(n)) == -2
    D = _lib.make_desc([50, 20, 20, 20, 1], 3, "Tanh", 52)
    assert L.epde_param_count(None, C.byref(n)) == -1
    sz = C.c_size_t()
    assert L.epde_workspace_bytes(C.byref(D), 0, C.byref(sz)) == -6
    assert L.epde_workspace_bytes(C.byref(D), 5000, C.byref(sz)) == 0 and sz.value > 5000 * 3 * 4
    # compute entry points refuse NULL buffers before touching the device
    assert L.epde_step(C.byref(D), None, None, None, 10, None, None, 1.0, 20.0, None, None, 0, None, None, None) == -1
    # unknown activation names fall back to CELU like network_arch.py:27-32
    assert _lib.make_desc([8, 12, 12, 1], 2, "NoSuchAct", 8).act_id == _lib.ACT_IDS["CELU"]


def _seed_state(L, seed):
    key = []
    s = abs(seed)
    while True:
        key.append(s & 0xFFFFFFFF); s >>= 32
        if s == 0:
            break
    karr = (C.c_uint32 * len(key))(*key)
    mt = (C.c_uint32 * 624)()
    pos = C.c_int32()
    assert L.epde_mt_seed(karr, len(key), mt, C.byref(pos)) == 0
    return mt, pos


def test_mt_seed_matches_cpython(L):
    for seed in [0, 1, 3214, 3905 + 1792443474, 2 ** 40 + 17, 2 ** 70 + 12345]:
        mt, pos = _seed_state(L, seed)
        random.seed(seed)
        st = random.getstate()[1]
        assert list(mt) == list(st[:624]) and pos.value == st[624] == 624
        assert np.array_equal(np.array(list(mt), dtype=np.uint32), minibatch.seed_state(seed))


def test_minibatch_stream_bit_exact_vs_reference_fixture(L):
    z = np.load(os.path.join(GOLDEN, "index_streams.npz"))
    n = 0
    while f"s{n}_meta" in z:
        seed, K, B, reps = (int(v) for v in z[f"s{n}_meta"])
        mt, pos = _seed_state(L, seed)
        for r in range(reps):
            out = np.empty(B, dtype=np.int64)
            assert L.epde_minibatch_indices(mt, C.byref(pos), K, B, out.ctypes.data_as(C.c_void_p)) == 0
            assert np.array_equal(out, z[f"s{n}_idx"][r]), (seed, K, B, r)
        n += 1


def test_minibatch_interleaves_with_python_random(L):
    """random.getstate() -> native draw -> random.setstate(): Python's stream continues unbroken."""
    K, B = 123457, 1001
    random.seed(99)
    random.random()                                  # odd position inside the 624 block
    want = random.choices(range(K), cum_weights=np.arange(1, K + 1), k=B)
    after = random.random()
    random.seed(99)
    random.random()
    ver, st, gauss = random.getstate()
    mt = (C.c_uint32 * 624)(*st[:624]); pos = C.c_int32(st[624])
    out = np.empty(B, dtype=np.int64)
    assert L.epde_minibatch_indices(mt, C.byref(pos), K, B, out.ctypes.data_as(C.c_void_p)) == 0
    assert np.array_equal(out, np.asarray(want))
    random.setstate((ver, tuple(mt) + (pos.value,), gauss))
    assert random.random() == after


def test_minibatch_edge_cases(L):
    mt, pos = _seed_state(L, 5)
    out = np.empty(4, dtype=np.int64)
    assert L.epde_minibatch_indices(mt, C.byref(pos), 1, 4, out.ctypes.data_as(C.c_void_p)) == 0
    assert (out == 0).all()                          # K = 1: every draw is row 0
    assert L.epde_minibatch_indices(mt, C.byref(pos), 10, 0, out.ctypes.data_as(C.c_void_p)) == 0   # empty batch
    assert L.epde_minibatch_indices(mt, C.byref(pos), 0, 4, out.ctypes.data_as(C.c_void_p)) == -6


def test_fused_family_is_a_descriptor_flag(L):
    """The kernel family of the fused path travels in epde_desc.flags (EPDE_FLAG_FFMA2): the library has no process-wide
    switch any more (ABI 3), and both settings describe a fused configuration with the same workspace layout."""
    from eigenpde_nn_b200 import _lib
    assert not hasattr(L, "epde_set_fused_impl")
    sizes = []
    for ffma2 in (False, True):
        desc = _lib.make_desc([50, 20, 20, 20, 1], 3, "Tanh", 52, ffma2=ffma2)
        assert bool(desc.flags & _lib.FLAG_FFMA2) == ffma2
        assert L.epde_has_fused_path(C.byref(desc)) == 1
        need = C.c_size_t()
        assert L.epde_workspace_bytes(C.byref(desc), 4096, C.byref(need)) == 0
        sizes.append(need.value)
    assert sizes[0] == sizes[1]
    big = _lib.make_desc([64, 20, 20, 20, 1], 3, "Tanh", 64)       # d + 1 > 64: outside the tensor-core family's x staging
    assert L.epde_has_fused_path(C.byref(big)) == 0


def test_activation_names_follow_the_reference(L):
    """src/network_arch.py:27-32: a name torch.nn does not know falls back to CELU; a torch.nn activation without a
    kernel here must raise instead of silently training CELU."""
    from eigenpde_nn_b200 import _lib
    assert _lib.activation_id("Tanh") == 0
    assert _lib.activation_id("NoSuchActivation") == _lib.ACT_IDS["CELU"]
    for name in ("GELU", "SiLU", "Softsign", "LeakyReLU"):
        with pytest.raises(_lib.EpdeError):
            _lib.activation_id(name)


def test_argument_validation_of_the_newer_entry_points(L):
    """NULL / range checks happen before any launch, so they can be exercised without a GPU."""
    from eigenpde_nn_b200 import _lib
    null = C.c_void_p(0)
    one = C.c_void_p(8)                      # any non-NULL address: the calls below must fail before touching it
    # epde_peer_allreduce: NULL pointers, epoch 0, rank out of range, payload larger than a slot, slot misaligned
    assert L.epde_peer_allreduce(null, 0, 2, 1, one, one, 13, 1, 65536, null) == -1
    assert L.epde_peer_allreduce(one, 0, 2, 0, one, one, 13, 1, 65536, null) == -6
    assert L.epde_peer_allreduce(one, 2, 2, 1, one, one, 13, 1, 65536, null) == -6
    assert L.epde_peer_allreduce(one, 0, 2, 1, one, one, 9000, 1, 65536, null) == -6
    assert L.epde_peer_allreduce(one, 0, 2, 1, one, one, 13, 1, 65530, null) == -6
    assert L.epde_peer_allreduce(one, 0, 33, 1, one, one, 13, 1, 65536, null) == -6           # flag block holds 2 x 32 flags
    # epde_md_align: NULL pointers, too many atoms, reference index out of range
    ref = (C.c_double * 9)(*([0.0] * 9)); ri = (C.c_int32 * 3)(0, 1, 2); bad = (C.c_int32 * 3)(0, 1, 11)
    assert L.epde_md_align(10, 10, 32, null, ref, ri, 3, one, one, null, null) == -1
    assert L.epde_md_align(10, 40, 120, one, ref, ri, 3, one, one, null, null) == -6
    assert L.epde_md_align(10, 10, 32, one, ref, bad, 3, one, one, null, null) == -6
    assert L.epde_md_align(10, 10, 28, one, ref, ri, 3, one, one, null, null) == -6          # d_pad < 3 * n_atoms
    # epde_train_epilogue: descriptor and pointer checks
    desc = _lib.make_desc([50, 20, 20, 20, 1], 3, "Tanh", 52)
    assert L.epde_train_epilogue(C.byref(desc), null, one, one, one, one, one, one, one, null) == -1


def test_torch_extension_shim_loads_and_mirrors_the_c_abi(L):
    """libeigenpde_torch.so: TORCH_LIBRARY ops over the C ABI (north_star: "ships as a PyTorch C++/CUDA extension behind a
    thin C-ABI").  Without a GPU only the host-side op can run; the schemas of the compute ops are checked."""
    from eigenpde_nn_b200 import _lib, build
    build.build_torch_ext()
    ops = _lib.torch_ops()
    desc = _lib.make_desc([50, 20, 20, 20, 1], 3, "Tanh", 52)
    need = C.c_size_t()
    assert L.epde_workspace_bytes(C.byref(desc), 4096, C.byref(need)) == 0
    assert ops.workspace_bytes([50, 20, 20, 20, 1], 3, 0, 52, int(desc.flags), 4096) == need.value
    for name in ("step", "step_partial", "step_finish"):
        schema = str(getattr(ops, name).default._schema)
        assert "Tensor X" in schema and "Tensor(a!) ws" in schema, schema
    with pytest.raises(RuntimeError):
        ops.workspace_bytes([50, 20, 20, 20, 1], 9, 0, 52, 1, 4096)          # k > EPDE_MAX_K: EPDE_ERR_DESC surfaces as an exception
