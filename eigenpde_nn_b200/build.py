"""Build libeigenpde_b200.so in-tree with nvcc for sm_100a (no torch headers involved)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libeigenpde_b200.so")
# EPDE_FAST_TANH: tanh via MUFU.EX2 + MUFU.RCP (abs. error ~2e-7, parity-tested) instead of tanhf
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-DEPDE_FAST_TANH", "-shared", "-Xcompiler", "-fPIC"]


def _sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "eigenpde_b200.h"))
    return srcs, deps


def _deps_of(src, seen=None):
    """Transitive closure of the quoted #include files of one source."""
    import re
    seen = set() if seen is None else seen
    if src in seen or not os.path.exists(src):
        return seen
    seen.add(src)
    for m in re.finditer(r'^\s*#include\s+"([^"]+)"', open(src).read(), re.M):
        _deps_of(os.path.normpath(os.path.join(os.path.dirname(src), m.group(1))), seen)
    return seen


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _sources()[1])


def build(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    """Compile every csrc/*.cu to an object (in parallel, only the stale ones) and link the shared library."""
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs, deps = _sources()
    objdir = os.path.join(HERE, "_obj")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + list(extra_flags) + (["-Xptxas", "-v"] if verbose else [])
    tag = os.path.join(objdir, "flags.txt")
    same_flags = os.path.exists(tag) and open(tag).read() == " ".join(flags)
    procs, objs = [], []
    for src in srcs:
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if (not force and same_flags and os.path.exists(obj)
                and os.path.getmtime(obj) > max(os.path.getmtime(p) for p in _deps_of(src))):
            continue
        procs.append((src, subprocess.Popen([nvcc] + flags + ["-c", "-o", obj, src], stdout=subprocess.PIPE,
                                            stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            failed = True
            sys.stderr.write(out)
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc failed building %s" % LIB)
    open(tag, "w").write(" ".join(flags))
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs,
                         capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking %s" % LIB)
    return LIB


TORCH_LIB = os.path.join(HERE, "libeigenpde_torch.so")


def build_torch_ext(force: bool = False) -> str:
    """The PyTorch extension shim (csrc_torch/epde_torch.cpp: TORCH_LIBRARY ops over the C ABI) -> libeigenpde_torch.so,
    in-tree, linked against libeigenpde_b200.so ($ORIGIN rpath) and the installed torch.  Plain g++: the shim has no kernels."""
    src = os.path.join(HERE, "csrc_torch", "epde_torch.cpp")
    hdr = os.path.join(os.path.dirname(HERE), "include", "eigenpde_b200.h")
    if (not force and os.path.exists(TORCH_LIB)
            and os.path.getmtime(TORCH_LIB) > max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(LIB))):
        return TORCH_LIB
    import torch
    from torch.utils import cpp_extension as ce
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = (["g++", "-O2", "-fPIC", "-shared", "-std=c++17",
            "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI), src]
           + ["-I" + i for i in ce.include_paths()] + ["-I/usr/local/cuda/include", "-L" + HERE, "-leigenpde_b200",
              "-Wl,-rpath,$ORIGIN", "-L" + libdir, "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda", "-Wl,-rpath," + libdir,
              "-o", TORCH_LIB])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building %s" % TORCH_LIB)
    return TORCH_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_torch_ext(force="--force" in sys.argv))
