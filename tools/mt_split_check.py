import os, sys, random, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from eigenpde_nn_b200.engine import DeviceMinibatch
K = 5_000_000
for (Btot, world) in [(1 << 21, 8), (1 << 23, 8), (3 * 300007, 3)]:
    random.seed(99)
    for _ in range(5): random.random()
    saved = random.getstate()
    want = np.array(random.choices(range(K), cum_weights=None, k=1)[:0])  # placeholder
    random.setstate(saved)
    want1 = np.asarray(random.choices(range(K), k=Btot), dtype=np.int64) if Btot <= (1 << 21) else None
    for rank in (0, world - 1):
        random.setstate(saved)
        mb = DeviceMinibatch(K, "cuda:0").seed_from_python()
        n = Btot // world; lo = rank * n
        a = mb.draw(Btot, lo=lo, n=n).clone()
        b = mb.draw(Btot, lo=lo, n=n).clone()
        os.environ["EPDE_MT_JUMP"] = "0"
        random.setstate(saved)
        ms = DeviceMinibatch(K, "cuda:0").seed_from_python()
        a0 = ms.draw(Btot, lo=lo, n=n).clone(); b0 = ms.draw(Btot, lo=lo, n=n).clone()
        os.environ["EPDE_MT_JUMP"] = "1"
        print(Btot, world, rank, "step1", bool(torch.equal(a, a0)), "step2", bool(torch.equal(b, b0)),
              "first mismatch", int((a != a0).nonzero()[0]) if not torch.equal(a, a0) else -1)
        if want1 is not None: print("   vs python:", bool(np.array_equal(a.cpu().numpy(), want1[lo:lo + n])))
