"""Training-loop throughput at the reference's batch sizes (B = 5000, 50-D, k = 3): the drop-in's default loop
(host fp64 Adam, results read every step) vs the device-resident loop (engine.DeviceTrainer)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from eigenpde_nn_b200.engine import StepEngine, DeviceTrainer, EigenLossFn
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "eigenpde_nn_b200", "dropin"))
import network_arch

K, d, k, B, steps = 200000, 50, 3, 5000, 300
rng = np.random.default_rng(0)
X = rng.standard_normal((K, d)).astype(np.float32); w = np.ones(K, np.float32)
eng = StepEngine(X, w, [d, 20, 20, 20, 1], k, "Tanh", np.ones(d), 1.0, [1.0, 0.8, 0.6])
torch.manual_seed(0)
model = network_arch.MyNet([d, 20, 20, 20, 1], "Tanh", k).double()
opt = torch.optim.Adam(model.parameters(), lr=5e-3)
idxs = [rng.integers(0, K, B) for _ in range(16)]
def host_step(i):
    opt.zero_grad()
    loss = EigenLossFn.apply(eng, idxs[i % 16], 20.0, *model.parameters())
    loss.backward(); opt.step()
for i in range(10): host_step(i)
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(steps): host_step(i)
torch.cuda.synchronize(); t_host = (time.perf_counter() - t0) / steps
tr = DeviceTrainer(eng, model)
tr.start_stage()
for i in range(10): tr.step(B, 5e-3, 20.0)
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(steps): tr.step(B, 5e-3, 20.0)
torch.cuda.synchronize(); t_dev = (time.perf_counter() - t0) / steps
print("B=%d: host-Adam loop %.3f ms/step (%.0f steps/s), device-resident loop %.3f ms/step (%.0f steps/s)"
      % (B, t_host * 1e3, 1 / t_host, t_dev * 1e3, 1 / t_dev))
