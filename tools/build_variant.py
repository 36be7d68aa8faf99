"""Build a development variant of the library into devlibs/<name>.so (git-ignored, travels to the GPU box):
    python tools/build_variant.py timing -DEPDE_PHASE_TIMING
and select it with EPDE_LIB=devlibs/<name>.so."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name, flags = sys.argv[1], sys.argv[2:]
csrc = os.path.join(ROOT, "eigenpde_nn_b200", "csrc")
out = os.path.join(ROOT, "devlibs", name + ".so")
os.makedirs(os.path.dirname(out), exist_ok=True)
objs, procs = [], []
for f in sorted(os.listdir(csrc)):
    if f.endswith(".cu"):
        o = os.path.join(ROOT, "devlibs", "%s_%s.o" % (name, f[:-3]))
        objs.append(o)
        procs.append(subprocess.Popen(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                                       "-DEPDE_FAST_TANH", "-Xcompiler", "-fPIC"] + flags + ["-c", "-o", o, os.path.join(csrc, f)]))
assert all(p.wait() == 0 for p in procs)
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs)
print(out)
