"""Data-parallel protocol of the step: the batch is cut into contiguous slices,
one per rank; two small allreduces per step.  Device-agnostic on purpose: `StepEngine` plugs in the
CUDA phases (NCCL), the CPU tests plug in the oracle (gloo)."""
from __future__ import annotations


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous, balanced slice [lo, hi) of range(n) for `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def two_phase_step(partial_fn, finish_fn, sums, grad, group=None, exchange=None):
    """partial_fn() fills `sums` (this rank's fp64 partial sums); after the sum-allreduce,
    finish_fn() fills `grad` (this rank's additive gradient share); after the second allreduce every
    rank holds the full gradient.  With group=None and no initialised process group: single rank.
    exchange: optional PeerExchange (one-shot all-reduce over NVLink peer memory) used instead of the two
    collective calls."""
    import torch.distributed as dist
    partial_fn()
    multi = dist.is_available() and dist.is_initialized() and (group is not None or dist.get_world_size() > 1)
    if multi:
        if exchange is not None:
            exchange.allreduce_(sums)
        else:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    finish_fn()
    if multi:
        if exchange is not None and exchange.fits(grad):
            exchange.allreduce_(grad)
        else:
            dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)


class PeerExchange:
    """One-shot all-reduce of the step's two small vectors over NVLink peer memory (`epde_peer_allreduce`).

    Plumbing only: torch.distributed._symmetric_memory allocates one buffer per rank and maps every peer's buffer
    into this process; the exchange itself is a single kernel of the library per call (peer stores, flags, rank-order
    sum).  `create` returns None when symmetric memory is not available (no NVLink peer access, gloo, ...), and the
    caller falls back to the two NCCL collectives."""

    SLOT_BYTES = 64 * 1024          # per rank and parity; larger vectors go through NCCL

    def __init__(self, lib, hdl, buf, rank, world):
        import torch
        self.lib, self.hdl, self.buf, self.rank, self.world = lib, hdl, buf, rank, world
        # the call counter (epoch) lives in device memory and is advanced by the kernel itself: the launch arguments never
        # change, so the exchanges can be captured into a CUDA graph together with the step's kernels
        self.epoch_dev = torch.zeros(1, dtype=torch.int32, device=buf.device)

    @classmethod
    def create(cls, lib, group, device):
        import os
        if os.environ.get("EPDE_EXCHANGE", "p2p") != "p2p":
            return None
        try:
            import torch
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm
            world, rank = dist.get_world_size(group), dist.get_rank(group)
            if world < 2 or world > 32 or dist.get_backend(group) != "nccl":      # the flag block holds 2 x 32 flags
                if os.environ.get("EPDE_EXCHANGE_DEBUG"):
                    print("[eigenpde_nn_b200] peer exchange not used: world", world, "backend", dist.get_backend(group))
                return None
            nbytes = 256 + 2 * world * cls.SLOT_BYTES
            buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
            buf.zero_()
            hdl = symm.rendezvous(buf, group)
            torch.cuda.synchronize(device)
            dist.barrier(group=group)              # every buffer is zero before anybody raises a flag
            if os.environ.get("EPDE_EXCHANGE_DEBUG") and rank == 0:
                print("[eigenpde_nn_b200] peer exchange active, world", world)
            return cls(lib, hdl, buf, rank, world)
        except Exception as exc:                   # not available on this system: NCCL path
            if os.environ.get("EPDE_EXCHANGE_DEBUG"):
                print("[eigenpde_nn_b200] peer exchange unavailable:", repr(exc))
            return None

    def peer_struct(self):
        """The `epde_peer` argument of epde_step_sharded (the exchanges folded into the step's kernels)."""
        from . import _lib
        return _lib.EpdePeer(self.hdl.buffer_ptrs_dev, self.rank, self.world, self.epoch_dev.data_ptr(), self.SLOT_BYTES)

    def fits(self, t):
        return t.numel() * t.element_size() <= self.SLOT_BYTES

    def allreduce_(self, t):
        import ctypes as C
        import torch
        from . import _lib
        assert t.is_contiguous() and t.dtype in (torch.float32, torch.float64) and self.fits(t)
        st = C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
        _lib.check(self.lib.epde_peer_allreduce_graph(C.c_void_p(self.hdl.buffer_ptrs_dev), self.rank, self.world,
                                                      C.c_void_p(self.epoch_dev.data_ptr()), C.c_void_p(t.data_ptr()),
                                                      C.c_void_p(t.data_ptr()), C.c_int64(t.numel()),
                                                      1 if t.dtype == torch.float64 else 0, C.c_int64(self.SLOT_BYTES), st),
                   "epde_peer_allreduce_graph")
        return t
