"""2+ GPU check of the one-shot peer-memory all-reduce against NCCL (run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch, torch.distributed as dist
from eigenpde_nn_b200 import _lib
from eigenpde_nn_b200.parallel import PeerExchange
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ex = PeerExchange.create(_lib.lib(), dist.group.WORLD, torch.device("cuda", local))
if rank == 0: print("exchange:", ex)
if ex is None: sys.exit(0)
torch.manual_seed(rank)
ok = True
for n, dt in [(13, torch.float64), (5643, torch.float32), (34, torch.float64), (16000, torch.float32)]:
    for rep in range(20):
        a = torch.randn(n, dtype=dt, device="cuda")
        b = a.clone()
        dist.all_reduce(b)
        ex.allreduce_(a)
        torch.cuda.synchronize()
        err = float((a - b).abs().max() / b.abs().max())
        ok = ok and err < (1e-14 if dt == torch.float64 else 1e-6)
t = torch.randn(5643, device="cuda"); s = torch.randn(13, dtype=torch.float64, device="cuda")
for name, fn in [("nccl", lambda: (dist.all_reduce(s), dist.all_reduce(t))), ("p2p", lambda: (ex.allreduce_(s), ex.allreduce_(t)))]:
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(200): fn()
    torch.cuda.synchronize(); dt_ = (time.perf_counter() - t0) / 200
    if rank == 0: print("%s: %.1f us per step (both exchanges)" % (name, dt_ * 1e6))
if rank == 0: print("results match NCCL:", ok)
dist.destroy_process_group()
