"""Timing of the device-side MT19937 minibatch generator (k_mt_generate + k_mt_to_idx)."""
import os, random, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eigenpde_nn_b200.engine import DeviceMinibatch
random.seed(1)
mb = DeviceMinibatch(5_000_000, "cuda:0").seed_from_python()
for logB in (12, 14, 16, 18, 20, 21, 23):
    B = 1 << logB
    out = torch.empty(B, dtype=torch.int64, device="cuda:0")
    for _ in range(2): mb.draw(B, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps): mb.draw(B, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("B = 2^%d: %.3f ms per draw  (%.2f G words/s, %.2f G samples/s)" % (logB, ms, 2 * B / ms / 1e6, B / ms / 1e6))
