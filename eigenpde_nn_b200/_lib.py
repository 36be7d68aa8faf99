"""ctypes binding of libeigenpde_b200.so (include/eigenpde_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an
exception is raised.  The library is never built implicitly on import.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EPDE_LIB", os.path.join(HERE, "libeigenpde_b200.so"))   # EPDE_LIB: developer override

EPDE_MAX_LAYERS = 8
EPDE_MAX_K = 8
ACT_IDS = {"Tanh": 0, "Sigmoid": 1, "Softplus": 2, "LogSigmoid": 3, "ELU": 4, "CELU": 4, "ReLU": 5,
           "Tanhshrink": 6}
FLAG_SORT = 1
FLAG_DETERMINISTIC = 2
FLAG_FFMA2 = 4          # fused path: FP32 FFMA2 kernel family instead of the tcgen05 family (include/eigenpde_b200.h)
FLAG_GENERIC = 8        # never take the fused path (layer-wise generic family)
ABI_VERSION = 4

EXPORTS = ["epde_abi_version", "epde_error_string", "epde_param_count", "epde_num_sums", "epde_has_fused_path",
           "epde_workspace_bytes", "epde_step_partial", "epde_step_finish", "epde_step", "epde_step_sharded", "epde_gather_ahead", "epde_step_ahead", "epde_forward",
           "epde_minibatch_indices", "epde_minibatch_indices_device", "epde_minibatch_indices_device_split", "epde_mt_jump_poly",
           "epde_mt_jump_device",
           "epde_mt_seed", "epde_adam_step", "epde_average_accumulate",
           "epde_eig_stats_accumulate", "epde_md_align", "epde_train_epilogue", "epde_train_epilogue64", "epde_peer_allreduce", "epde_peer_allreduce_graph"]


class EpdeDesc(C.Structure):
    _fields_ = [("d", C.c_int32), ("d_pad", C.c_int32), ("k", C.c_int32), ("n_layers", C.c_int32),
                ("widths", C.c_int32 * (EPDE_MAX_LAYERS + 1)), ("act_id", C.c_int32), ("flags", C.c_uint32), ("metric", C.c_void_p)]


class EpdePeer(C.Structure):
    _fields_ = [("d_peer_bufs", C.c_void_p), ("rank", C.c_int32), ("world", C.c_int32), ("d_epoch", C.c_void_p),
                ("slot_bytes", C.c_int64)]


class EpdeError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EpdeError("%s not found: build it with `python -m eigenpde_nn_b200.build` "
                            "(there is no CPU or PyTorch fallback for the step)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, i64, f64, sz = C.c_void_p, C.c_int64, C.c_double, C.c_size_t
        dp = C.POINTER(EpdeDesc)
        L.epde_abi_version.restype = C.c_int
        L.epde_error_string.restype = C.c_char_p
        L.epde_error_string.argtypes = [C.c_int]
        L.epde_param_count.argtypes = [dp, C.POINTER(i64)]
        L.epde_num_sums.argtypes = [dp, C.POINTER(C.c_int32)]
        L.epde_has_fused_path.argtypes = [dp]
        L.epde_workspace_bytes.argtypes = [dp, i64, C.POINTER(sz)]
        L.epde_step_partial.argtypes = [dp, vp, vp, vp, i64, vp, vp, vp, sz, vp, vp]
        L.epde_step_finish.argtypes = [dp, vp, vp, vp, i64, vp, vp, f64, f64, C.POINTER(f64), vp, vp, sz, vp, vp, vp]
        L.epde_step.argtypes = [dp, vp, vp, vp, i64, vp, vp, f64, f64, C.POINTER(f64), vp, sz, vp, vp, vp]
        L.epde_step_sharded.argtypes = [dp, vp, vp, vp, i64, vp, vp, f64, f64, C.POINTER(f64), vp, sz, vp, vp,
                                        C.POINTER(EpdePeer), vp]
        L.epde_gather_ahead.argtypes = [dp, vp, vp, vp, i64, vp, sz, C.c_int32, vp]
        L.epde_step_ahead.argtypes = [dp, vp, vp, vp, i64, vp, vp, f64, f64, C.POINTER(f64), vp, sz, vp, vp, C.c_int32, C.c_int32,
                                      C.c_int32, C.POINTER(EpdePeer), vp, vp]
        L.epde_forward.argtypes = [dp, vp, vp, i64, vp, vp, vp, sz, vp, vp]
        L.epde_minibatch_indices.argtypes = [vp, C.POINTER(C.c_int32), i64, i64, vp]
        L.epde_minibatch_indices_device.argtypes = [vp, i64, i64, i64, i64, vp, vp, vp]
        L.epde_minibatch_indices_device_split.argtypes = [vp, C.c_int32, i64, i64, i64, vp, vp, vp]
        L.epde_mt_jump_poly.argtypes = [i64, vp]
        L.epde_mt_jump_device.argtypes = [vp, vp, C.c_int32, vp, vp, vp]
        L.epde_mt_seed.argtypes = [vp, C.c_int32, vp, C.POINTER(C.c_int32)]
        L.epde_adam_step.argtypes = [vp, vp, vp, vp, i64, f64, i64, vp]
        L.epde_average_accumulate.argtypes = [dp, vp, vp, vp, vp]
        L.epde_eig_stats_accumulate.argtypes = [dp, vp, vp, vp]
        L.epde_peer_allreduce.argtypes = [vp, C.c_int, C.c_int, C.c_uint32, vp, vp, i64, C.c_int, i64, vp]
        L.epde_train_epilogue.argtypes = [dp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.epde_train_epilogue64.argtypes = [dp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.epde_peer_allreduce_graph.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, i64, C.c_int, i64, vp]
        L.epde_md_align.argtypes = [i64, C.c_int32, C.c_int32, vp, vp, vp, C.c_int32, vp, vp, vp, vp]
        for name in EXPORTS:
            if name != "epde_error_string":
                getattr(L, name).restype = C.c_int
        _lib = L
    return _lib


TORCH_LIB_PATH = os.path.join(HERE, "libeigenpde_torch.so")
_ops = None


def torch_ops():
    """torch.ops.epde (the PyTorch extension shim over the C ABI, csrc_torch/epde_torch.cpp), loaded on first use.
    Raises if libeigenpde_torch.so has not been built -- like the C library it is never built implicitly."""
    global _ops
    if _ops is None:
        import torch
        if not os.path.exists(TORCH_LIB_PATH):
            raise EpdeError("%s not found: build it with `python -m eigenpde_nn_b200.build`" % TORCH_LIB_PATH)
        lib()                                   # libeigenpde_b200.so first (the shim links against it)
        torch.ops.load_library(TORCH_LIB_PATH)
        _ops = torch.ops.epde
    return _ops


def check(rc: int, what: str):
    if rc != 0:
        raise EpdeError("%s failed: %d (%s)" % (what, rc, lib().epde_error_string(rc).decode()))


def activation_id(activation_name) -> int:
    """The reference resolves `getattr(torch.nn, name)` and falls back to CELU only when torch.nn has no such attribute
    (src/network_arch.py:27-32).  A torch.nn activation that exists but has no kernel here is an ERROR, never a silent
    substitution: the host MyNet / the saved checkpoints would use the real activation while the step trained CELU."""
    if activation_name in ACT_IDS:
        return ACT_IDS[activation_name]
    import torch
    if getattr(torch.nn, str(activation_name), None) is None:
        return ACT_IDS["CELU"]
    raise EpdeError("activation torch.nn.%s exists but the step kernels implement only %s (EPDE_ERR_UNSUPPORTED)"
                    % (activation_name, sorted(ACT_IDS)))


def make_desc(arch, k, activation_name, d_pad, sort=True, deterministic=False, metric_ptr=None, ffma2=False,
              generic=False) -> EpdeDesc:
    """arch = [d, n_1, ..., n_{L-1}, 1] (src/pytorch_eigen_solver.py:352)."""
    if len(arch) - 1 > EPDE_MAX_LAYERS:
        raise EpdeError("too many layers")
    D = EpdeDesc()
    D.d, D.d_pad, D.k, D.n_layers = int(arch[0]), int(d_pad), int(k), len(arch) - 1
    for i, n in enumerate(arch):
        D.widths[i] = int(n)
    D.act_id = activation_id(activation_name)
    D.flags = (FLAG_SORT if sort else 0) | (FLAG_DETERMINISTIC if deterministic else 0) | (FLAG_FFMA2 if ffma2 else 0) | (FLAG_GENERIC if generic else 0)
    D.metric = metric_ptr
    return D
