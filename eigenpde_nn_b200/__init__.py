"""B200-native loss+gradient step of EigenPDE-NN.

Layout
  csrc/      hand-written sm_100a kernels + the C ABI (include/eigenpde_b200.h)
  _lib.py    ctypes binding (no fallback)
  engine.py  StepEngine: device buffers, streams, multi-GPU reduction; EigenLossFn autograd.Function
  dropin/    pytorch_eigen_solver / network_arch / data_set modules with the reference's names
"""
from ._lib import EpdeError, lib, make_desc  # noqa: F401
