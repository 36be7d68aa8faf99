"""Host side of the step: device mirrors, workspace, the autograd.Function.

Mirrors the reference's interface for the hot path:
  * `StepEngine.step(idx, params, alpha)` = the numerical content of
    `eigen_solver.update_step` (src/pytorch_eigen_solver.py:191-236) up to and including
    `loss.backward()`: returns eig_vals (ascending), cvec, loss, non_penalty_loss, penalty and the
    flat parameter gradient.
  * `EigenLossFn` exposes it as a `torch.autograd.Function` whose only differentiable output is
    `loss`, so `optimizer.zero_grad(); loss.backward(); optimizer.step()` (:232-234) keeps working.

PyTorch is used for device memory, streams and torch.distributed only; all arithmetic of the step
runs in libeigenpde_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class StepEngine:
    """Owns the fp32 device mirror of a data_set and runs loss+grad steps on it.

    X: [K, d] array-like (fp64 as loaded by data_set.from_file), weights: [K].
    arch: [d, n_1, ..., 1]; diag_coeff: [d] (src/pytorch_eigen_solver.py:76-81).
    """

    def __init__(self, X, weights, arch, k, activation_name, diag_coeff, beta, eig_w, sort=True,
                 device=None, process_group=None, deterministic=False, metric=None, ffma2=False, generic=False):
        if not torch.cuda.is_available():
            raise _lib.EpdeError("StepEngine needs a CUDA device (no CPU fallback)")
        self.lib = _lib.lib()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.k = int(k)
        self.arch = [int(a) for a in arch]
        self.d = self.arch[0]
        self.d_pad = (self.d + 3) // 4 * 4
        self.beta = float(beta)
        self.eig_w = [float(v) for v in eig_w][: self.k]
        if len(self.eig_w) < self.k:
            raise _lib.EpdeError("only %d weights for %d eigenvalues" % (len(self.eig_w), self.k))
        # per-row metric [K, d, d] of the MD alignment layer (see md_align below); X then holds the ALIGNED rows
        self.metric = None
        if metric is not None:
            self.metric = metric.to(self.device, torch.float32).contiguous()
            assert self.metric.dim() == 3 and self.metric.shape[1] == self.d and self.metric.shape[2] == self.d
        self.desc = _lib.make_desc(self.arch, self.k, activation_name, self.d_pad, sort, deterministic,
                                   self.metric.data_ptr() if self.metric is not None else None, ffma2=ffma2, generic=generic)
        self.pg = process_group
        n = C.c_int64()
        _lib.check(self.lib.epde_param_count(C.byref(self.desc), C.byref(n)), "epde_param_count")
        self.n_params = n.value
        ns = C.c_int32()
        _lib.check(self.lib.epde_num_sums(C.byref(self.desc), C.byref(ns)), "epde_num_sums")
        self.n_sums = ns.value
        self.fused = self.lib.epde_has_fused_path(C.byref(self.desc)) == 1
        with torch.cuda.device(self.device):
            if isinstance(X, torch.Tensor) and X.is_cuda:
                # already a device mirror: [K, d_pad] fp32 with zero padding columns
                assert X.dtype == torch.float32 and X.shape[1] == self.d_pad and X.is_contiguous()
                self.X, self.K = X, X.shape[0]
                self.w = weights.to(self.device, torch.float32).contiguous()
            else:
                Xt = torch.as_tensor(np.asarray(X), dtype=torch.float32)
                self.K = Xt.shape[0]
                self.X = torch.zeros((self.K, self.d_pad), dtype=torch.float32, device=self.device)
                self.X[:, : self.d] = Xt.to(self.device)
                self.w = torch.as_tensor(np.asarray(weights), dtype=torch.float32).to(self.device).contiguous()
            self.diag = torch.as_tensor(np.asarray(diag_coeff), dtype=torch.float32).to(self.device).contiguous()
            self.sums = torch.zeros(self.n_sums, dtype=torch.float64, device=self.device)
            # out (fp64 scalars) and grad (fp32) are views of ONE buffer: a single device -> host copy returns both
            nout = 4 + 4 * self.k
            self.result = torch.zeros(8 * nout + 4 * self.n_params, dtype=torch.uint8, device=self.device)
            self.out = self.result[:8 * nout].view(torch.float64)
            self.grad = self.result[8 * nout:].view(torch.float32)
        self._ws = None
        self._ws_B = 0
        self.ws_generation = 0          # bumped whenever the step workspace is re-allocated (captured graphs hold its address)
        self._fws = None                # separate workspace of forward(): a stage-end full-data pass never moves the step's
        self._fws_B = 0
        self._exchange = None
        if self.pg is not None:
            from .parallel import PeerExchange
            self._exchange = PeerExchange.create(self.lib, self.pg, self.device)
        self._eigw_c = (C.c_double * self.k)(*self.eig_w)
        # the step is dispatched as a torch operator (torch.ops.epde.*: the extension shim over the C ABI, current-stream
        # semantics by at::cuda::getCurrentCUDAStream); EPDE_BINDING=ctypes calls the C ABI directly instead
        import os
        self.ops = None if os.environ.get("EPDE_BINDING") == "ctypes" else _lib.torch_ops()
        self._act_id = int(self.desc.act_id)
        self._flags = int(self.desc.flags)

    # -- buffers ------------------------------------------------------------------------------
    def _workspace(self, B):
        if self._ws is None or B > self._ws_B:
            if getattr(self, "_pending", None) is not None or getattr(self, "_await", None) is not None:
                self._side.synchronize()        # a look-ahead gather may still be writing into the workspace that goes away
                self._pending = self._await = None
            need = C.c_size_t()
            _lib.check(self.lib.epde_workspace_bytes(C.byref(self.desc), B, C.byref(need)), "epde_workspace_bytes")
            self._ws = torch.empty(need.value + 256, dtype=torch.uint8, device=self.device)
            self._ws_B = B
            self.ws_generation += 1
            off = (-self._ws.data_ptr()) % 256
            self._ws_view = self._ws[off:off + need.value]
        return self._ws_view

    def _forward_workspace(self, B):
        if self._fws is None or B > self._fws_B:
            need = C.c_size_t()
            _lib.check(self.lib.epde_workspace_bytes(C.byref(self.desc), B, C.byref(need)), "epde_workspace_bytes")
            self._fws = torch.empty(need.value + 256, dtype=torch.uint8, device=self.device)
            self._fws_B = B
            off = (-self._fws.data_ptr()) % 256
            self._fws_view = self._fws[off:off + need.value]
        return self._fws_view

    def flat_params(self, model) -> torch.Tensor:
        """fp32 flat copy of model.parameters() (the reference keeps them in fp64, :374)."""
        return torch.cat([p.detach().reshape(-1) for p in model.parameters()]).to(
            device=self.device, dtype=torch.float32)

    # -- the step -----------------------------------------------------------------------------
    # -- gather look-ahead ---------------------------------------------------------------------
    RESERVE_SMS = 10          # SMs the persistent pass kernels leave to the gather of the next step

    def _ahead_ok(self):
        if getattr(self, "_ahead_on", None) is None:
            import os
            self._ahead_on = bool(self.fused and self.metric is None and not (self._flags & (_lib.FLAG_FFMA2 | _lib.FLAG_GENERIC))
                                  and (self.pg is None or self._fold_exchange())
                                  and os.environ.get("EPDE_GATHER_AHEAD", "1") != "0")
            if self._ahead_on:
                lo, hi = torch.cuda.Stream.priority_range()          # (lowest, highest) = (0, -5)
                self._hp = torch.cuda.Stream(device=self.device, priority=hi)       # the step's kernels
                self._side = torch.cuda.Stream(device=self.device, priority=lo)     # the next step's gather
                self._pending = None
                self._await = None
                self.RESERVE_SMS = int(os.environ.get("EPDE_RESERVE_SMS", self.RESERVE_SMS))
                self._ev = [torch.cuda.Event() for _ in range(8)]
                self._evi = 0
        return self._ahead_on

    def _event(self):
        self._evi = (self._evi + 1) % len(self._ev)
        return self._ev[self._evi]

    def compute_stream(self):
        """The (high-priority) stream the look-ahead steps run on; a caller that makes it the current stream
        (`with torch.cuda.stream(engine.compute_stream()): ...`) saves the two hand-over events per step.  None when the
        gather look-ahead does not apply to this engine."""
        return self._hp if self._ahead_ok() else None

    def _step_ahead(self, idx_dev, B, params_dev, alpha, next_idx):
        """The step with the gather of the NEXT step's rows running beside it (`epde_gather_ahead` on a low-priority stream,
        the step itself on a high-priority one with `RESERVE_SMS` SMs left free by its two persistent kernels).  The rows of
        THIS step were gathered during the previous call if it was given this index tensor as `next_idx`."""
        ws = self._workspace(B)
        cur = torch.cuda.current_stream(self.device)
        hp, side = self._hp, self._side
        pend, self._pending = self._pending, None
        pre = int(pend is not None and pend[0] == idx_dev.data_ptr() and pend[1] == B and pend[4] == self.ws_generation)
        slot = pend[2] if pend is not None else 0
        if pend is not None and not pre:
            slot = 1 - pend[2]              # an unused look-ahead may still be writing its set: take the other one
        same = cur == hp                     # the caller already works on the engine's compute stream: no hand-over events
        if not same:
            e0 = self._event(); e0.record(cur)   # everything the caller has enqueued so far (parameters, previous results read)
            hp.wait_event(e0)
        if pre:
            hp.wait_event(pend[3])
        args = (C.byref(self.desc), _ptr(self.X), _ptr(self.w), _ptr(idx_dev), C.c_int64(B), _ptr(params_dev), _ptr(self.diag))
        peer = C.byref(self._peer) if (self.pg is not None and self._fold_exchange()) else None
        ep = self._event() if next_idx is not None else None
        if ep is not None:
            ep.record(hp)                   # (creates the CUDA event; the library records it again between the two phases)
        _lib.check(self.lib.epde_step_ahead(*args, C.c_double(self.beta), C.c_double(alpha), self._eigw_c, _ptr(ws),
                                            C.c_size_t(ws.numel()), _ptr(self.out), _ptr(self.grad), slot, pre,
                                            self.RESERVE_SMS if next_idx is not None else 0, peer,
                                            C.c_void_p(ep.cuda_event) if ep is not None else None, C.c_void_p(hp.cuda_stream)),
                   "epde_step_ahead")
        if not same:
            e1 = self._event(); e1.record(hp)
            cur.wait_event(e1)
        if next_idx is not None:
            self._await = (ep, 1 - slot, B)
            if next_idx is not LATER:
                self.gather_next(next_idx)
        return self.out, self.grad

    def gather_next(self, next_idx, after=None):
        """Enqueue the look-ahead gather announced by `step_device(..., next_idx=LATER)`: the rows of `next_idx` (device
        int64, same B) for the next call.  `after`: an event the indices become valid with (e.g. a draw on another stream)."""
        ep, slot, B = self._await
        self._await = None
        ws = self._workspace(B)
        side = self._side
        # the set was last read by the step BEFORE the one just enqueued, which is complete.  The gather starts behind phase 1
        # of the step just enqueued (bound by HBM itself) and runs beside its phase 2 on the SMs that kernel leaves free.
        side.wait_event(ep)
        if after is not None:
            side.wait_event(after)
        _lib.check(self.lib.epde_gather_ahead(C.byref(self.desc), _ptr(self.X), _ptr(self.w), _ptr(next_idx), C.c_int64(B),
                                              _ptr(ws), C.c_size_t(ws.numel()), slot, C.c_void_p(side.cuda_stream)),
                   "epde_gather_ahead")
        eg = self._event(); eg.record(side)
        self._pending = (next_idx.data_ptr(), B, slot, eg, self.ws_generation)

    def step_device(self, idx_dev, B, params_dev, alpha, next_idx=None):
        """Asynchronous step on the current stream.  idx_dev: int64 device tensor or None.

        Returns (out_scalars fp64 device tensor, grad fp32 device tensor), both owned by the engine.
        With a process group the batch is the union of all ranks' idx slices.
        next_idx: the index tensor of the NEXT call (same B), if the caller already has it - the minibatch does not depend
        on this step's result (src/data_set.py:67) -: its rows are gathered while this step computes (`_step_ahead`).
        """
        if idx_dev is not None and (next_idx is not None or getattr(self, "_pending", None) is not None) and self._ahead_ok():
            return self._step_ahead(idx_dev, B, params_dev, alpha, next_idx)
        ws = self._workspace(B)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        args = (C.byref(self.desc), _ptr(self.X), _ptr(self.w), _ptr(idx_dev), C.c_int64(B), _ptr(params_dev),
                _ptr(self.diag))
        ops = self.ops
        if self.pg is None:
            if ops is not None:
                ops.step(self.X, self.w, idx_dev, B, params_dev, self.diag, self.beta, float(alpha), self.eig_w, self.arch,
                         self.k, self._act_id, self._flags, self.metric, ws, self.out, self.grad)
            else:
                _lib.check(self.lib.epde_step(*args, C.c_double(self.beta), C.c_double(alpha), self._eigw_c,
                                              _ptr(ws), C.c_size_t(ws.numel()), _ptr(self.out), _ptr(self.grad), st),
                           "epde_step")
        elif self._fold_exchange():
            # both exchanges happen inside the step's own kernels (epde_step_sharded): no exchange launches
            _lib.check(self.lib.epde_step_sharded(*args, C.c_double(self.beta), C.c_double(alpha), self._eigw_c,
                                                  _ptr(ws), C.c_size_t(ws.numel()), _ptr(self.out), _ptr(self.grad),
                                                  C.byref(self._peer), st), "epde_step_sharded")
        else:
            from .parallel import two_phase_step

            def partial():
                if ops is not None:
                    ops.step_partial(self.X, self.w, idx_dev, B, params_dev, self.diag, self.arch, self.k, self._act_id,
                                     self._flags, self.metric, ws, self.sums)
                else:
                    _lib.check(self.lib.epde_step_partial(*args, _ptr(ws), C.c_size_t(ws.numel()), _ptr(self.sums), st),
                               "epde_step_partial")

            def finish():
                if ops is not None:
                    ops.step_finish(self.X, self.w, idx_dev, B, params_dev, self.diag, self.beta, float(alpha), self.eig_w,
                                    self.arch, self.k, self._act_id, self._flags, self.metric, self.sums, ws, self.out,
                                    self.grad)
                else:
                    _lib.check(self.lib.epde_step_finish(*args, C.c_double(self.beta), C.c_double(alpha), self._eigw_c,
                                                         _ptr(self.sums), _ptr(ws), C.c_size_t(ws.numel()),
                                                         _ptr(self.out), _ptr(self.grad), st), "epde_step_finish")

            two_phase_step(partial, finish, self.sums, self.grad, self.pg, self._exchange)
        return self.out, self.grad

    def _fold_exchange(self):
        """True when the sharded step can run as ONE call with the exchanges inside its kernels: peer exchange available,
        tensor-core fused family, gradient fits a slot.  EPDE_EXCHANGE_FOLD=0 keeps the separate exchange kernels."""
        if getattr(self, "_fold", None) is None:
            import os
            ex = self._exchange
            self._fold = bool(ex is not None and self.fused and self.metric is None
                              and not (self._flags & (_lib.FLAG_FFMA2 | _lib.FLAG_GENERIC))
                              and ex.fits(self.grad) and os.environ.get("EPDE_EXCHANGE_FOLD", "1") != "0")
            if self._fold:
                self._peer = ex.peer_struct()
        return self._fold

    def step(self, idx, params_dev, alpha):
        """Host-facing step: idx is a host int64 array/list (or None for rows 0..B-1 with B = K)."""
        if idx is None:
            idx_dev, B = None, self.K
        else:
            idx_t = torch.as_tensor(np.asarray(idx, dtype=np.int64))
            B = idx_t.numel()
            idx_dev = idx_t.to(self.device, non_blocking=True)
        out, grad = self.step_device(idx_dev, B, params_dev, alpha)
        h = out.cpu().numpy()
        k = self.k
        return dict(loss=h[0], non_penalty=h[1], penalty=h[2], tot_weight=h[3], eig_vals=h[4:4 + k].copy(),
                    cvec=h[4 + k:4 + 2 * k].astype(np.int64), mean=h[4 + 2 * k:4 + 3 * k].copy(),
                    var=h[4 + 3 * k:4 + 4 * k].copy(), grad=grad)

    def forward(self, idx_dev, B, params_dev):
        ws = self._forward_workspace(max(min(B, 262144), 1))
        y = torch.empty((B, self.k), dtype=torch.float32, device=self.device)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.epde_forward(C.byref(self.desc), _ptr(self.X), _ptr(idx_dev), C.c_int64(B),
                                         _ptr(params_dev), _ptr(self.diag), _ptr(ws), C.c_size_t(ws.numel()),
                                         _ptr(y), st), "epde_forward")
        return y


class EigenLossFn(torch.autograd.Function):
    """loss = EigenLossFn.apply(engine, idx, alpha, *model.parameters())

    Forward runs the whole fused loss+gradient step and caches d loss / d theta; backward hands the
    cached gradient (times grad_output) back to autograd as one view per parameter, in
    model.parameters() order.  Non-differentiable results are left on `engine.last`.
    """

    @staticmethod
    def forward(ctx, engine, idx, alpha, *params):
        flat = torch.cat([p.detach().reshape(-1) for p in params]).to(device=engine.device, dtype=torch.float32)
        res = engine.step(idx, flat, float(alpha))
        engine.last = res
        ctx.shapes = [p.shape for p in params]
        ctx.meta = (params[0].dtype, params[0].device)
        ctx.save_for_backward(res["grad"].clone())
        return torch.tensor(res["loss"], dtype=params[0].dtype, device=params[0].device)

    @staticmethod
    def backward(ctx, grad_out):
        (g,) = ctx.saved_tensors
        dtype, device = ctx.meta
        g = g.to(device=device, dtype=dtype) * grad_out.to(device=device, dtype=dtype)
        outs, off = [], 0
        for shp in ctx.shapes:
            n = int(np.prod(shp))
            outs.append(g[off:off + n].view(shp))
            off += n
        return (None, None, None) + tuple(outs)


class DeviceMinibatch:
    """The reference's minibatch index stream (src/data_set.py:31,67: `random.choices(range(K), cum_weights=1..K, k=B)`)
    continued ON THE DEVICE: the MT19937 state of Python's `random` is uploaded once (`seed_from_python`), every `draw`
    is two kernels of the library (`epde_minibatch_indices_device`) with no host involvement, and `sync_to_python` hands
    the advanced state back to `random.setstate` -- after it, `random` is exactly where the reference would have left it.
    With world > 1 every rank jumps ahead to its own contiguous slice of the global stream; large draws are split over
    several CTAs by the same jump-ahead (`_draw_slice`)."""

    def __init__(self, K, device, lib=None):
        self.lib = lib if lib is not None else _lib.lib()
        self.K = int(K)
        self.device = torch.device(device)
        self.state = torch.zeros(625, dtype=torch.int32, device=self.device)      # 624 words + position (uint32 bit patterns)
        self._raw = None
        self._py_meta = None
        self.seeded = False
        # jump-ahead (sharded draws): polynomials x^J mod phi per jump distance, scratch states, the 20 560-word window
        self._polys = {}
        self._jump = None

    def seed_from_python(self):
        import random
        version, st, gauss = random.getstate()
        self._py_meta = (version, gauss)
        host = np.array(st, dtype=np.uint32).view(np.int32)
        self.state.copy_(torch.from_numpy(host))
        self.seeded = True
        return self

    def sync_to_python(self):
        """Download the generator state and install it in Python's `random` (synchronises the device)."""
        import random
        assert self.seeded
        st = self.state.cpu().numpy().view(np.uint32)
        version, gauss = self._py_meta
        random.setstate((version, tuple(int(v) for v in st), gauss))

    def _poly(self, J):
        """x^J mod (characteristic polynomial of MT19937) on the device; computed once per jump distance on the host."""
        t = self._polys.get(J)
        if t is None:
            host = np.zeros(624, dtype=np.uint32)
            _lib.check(self.lib.epde_mt_jump_poly(C.c_int64(J), host.ctypes.data_as(C.c_void_p)), "epde_mt_jump_poly")
            t = torch.from_numpy(host.view(np.int32)).to(self.device)
            self._polys[J] = t
        return t

    SPLIT_MIN = 1 << 18          # draws from which a slice is generated by several CTAs at once
    SPLIT_PARTS = 8

    def _draw_slice(self, B_total, out, lo, n):
        """Draw by jump-ahead: from ONE window of the base state, jump to the next step (2 * B_total words) and to the start
        of this rank's slice (2 * lo words) - for a large slice to the starts of its SPLIT_PARTS parts, which are then
        generated by as many CTAs at once (the recurrence is sequential per generator: 1.25 G draws/s on one SM).  Only the
        2 * n words of the slice are generated (csrc/epde_mt_jump.cu, csrc/epde_mt.cuh)."""
        parts = self.SPLIT_PARTS if n >= self.SPLIT_MIN else 1
        n_sub = (n + parts - 1) // parts if n > 0 else 1
        while parts > 1 and (parts - 1) * n_sub >= n:
            parts -= 1
        key = (B_total, lo, n, parts)
        if self._jump is None:
            self._jump = {}
        J = self._jump.get(key)
        if J is None:      # buffers are kept per (batch, slice): a captured CUDA graph may hold their addresses
            dev = self.device
            states = torch.zeros((parts + 1, 625), dtype=torch.int32, device=dev)      # part generators, then the next base state
            polys = torch.cat([self._poly(2 * (lo + q * n_sub)) for q in range(parts)] + [self._poly(2 * B_total)]).contiguous()
            outs = (C.c_void_p * (parts + 1))(*[states[q].data_ptr() for q in range(parts + 1)])
            J = self._jump[key] = dict(states=states, polys=polys, outs=outs, win=torch.empty(20560, dtype=torch.int32, device=dev))
        if self._raw is None or self._raw.numel() < 2 * n:
            self._raw = torch.empty(max(2 * n, 2), dtype=torch.int32, device=self.device)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.epde_mt_jump_device(_ptr(self.state), _ptr(J["polys"]), parts + 1, J["outs"], _ptr(J["win"]), st),
                   "epde_mt_jump_device")
        if n > 0:
            _lib.check(self.lib.epde_minibatch_indices_device_split(_ptr(J["states"]), parts, C.c_int64(self.K), C.c_int64(n_sub),
                                                                    C.c_int64(n), _ptr(self._raw), _ptr(out), st),
                       "epde_minibatch_indices_device_split")
        self.state.copy_(J["states"][parts])        # same address every step: the draw can sit inside a replayed CUDA graph
        return out

    def draw(self, B_total, out=None, lo=0, n=None, split=False):
        """Enqueue B_total draws on the current stream; the indices of draws [lo, lo + n) go to `out` (int64, device).
        A proper slice (a rank of a sharded step) is reached by jump-ahead: only its own 2 * n words are generated.
        split: use the jump-ahead path for a whole-stream draw too (several CTAs generate at once; the generator is then left
        as the 624-word window at the new position instead of CPython's (block, position) pair - the same stream)."""
        assert self.seeded, "call seed_from_python() first"
        n = B_total - lo if n is None else n
        if out is None:
            out = torch.empty(n, dtype=torch.int64, device=self.device)
        if B_total == 0:
            return out
        import os
        if (lo != 0 or n != B_total or split) and os.environ.get("EPDE_MT_JUMP", "1") != "0":      # (a whole-stream draw stays sequential:
            # it leaves the generator exactly as CPython would show it, block and position)
            return self._draw_slice(B_total, out, lo, n)
        if self._raw is None or self._raw.numel() < 2 * B_total:
            self._raw = torch.empty(2 * B_total, dtype=torch.int32, device=self.device)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.epde_minibatch_indices_device(_ptr(self.state), C.c_int64(self.K), C.c_int64(B_total),
                                                          C.c_int64(lo), C.c_int64(n), _ptr(self._raw), _ptr(out), st),
                   "epde_minibatch_indices_device")
        return out


LATER = object()        # step_device(..., next_idx=LATER): the next minibatch is announced afterwards with gather_next()


class DrawAhead:
    """Draws the next step's minibatch on a second stream while the current step computes (the index stream does not
    depend on the step's result; the generator is a single CTA and the step's kernels leave an SM idle).
        ahead = DrawAhead(mb, B_total, lo, n); ahead.start()
        for ...: idx = ahead.take(); <enqueue the step reading idx>; ahead.done_with(idx) ..."""

    def __init__(self, mb: DeviceMinibatch, B_total, lo=0, n=None, split=False, depth=1):
        """depth: draws kept in flight (1: the next step's; 2: also the one after it, so that the next step's indices are
        complete while the current step runs - what the gather look-ahead of StepEngine needs)."""
        self.mb, self.B_total, self.lo, self.split, self.depth = mb, B_total, lo, split, depth
        self.n = B_total - lo if n is None else n
        dev = mb.device
        self.nb = depth + 1
        self.buf = [torch.empty(self.n, dtype=torch.int64, device=dev) for _ in range(self.nb)]
        self.stream = torch.cuda.Stream(device=dev)
        self.ready = [torch.cuda.Event() for _ in range(self.nb)]
        self.free = [torch.cuda.Event() for _ in range(self.nb)]
        for e in self.free:
            e.record(torch.cuda.current_stream(dev))
        self.stream.wait_stream(torch.cuda.current_stream(dev))      # the uploaded state is visible to the side stream
        self.head = self.tail = 0

    def start(self):
        for _ in range(self.depth):
            self._enqueue()
        return self

    def _enqueue(self):
        b = self.head % self.nb
        self.stream.wait_event(self.free[b])
        with torch.cuda.stream(self.stream):
            self.mb.draw(self.B_total, self.buf[b], self.lo, self.n, split=self.split)
            self.ready[b].record(self.stream)
        self.head += 1
        return self.buf[b], self.ready[b]

    def take(self, defer=False):
        """Indices of the current step; the draw of the following step starts right away on the side stream - or, with
        defer=True, when the caller says so (`enqueue_next()`, after it has launched the step: enqueueing the draw costs
        host time that would otherwise sit between the previous step's results and this step's first kernel)."""
        b = self.tail % self.nb
        torch.cuda.current_stream(self.mb.device).wait_event(self.ready[b])
        if not defer:
            self._enqueue()
        return self.buf[b]

    def peek(self, ahead=1):
        """(index buffer, drawn-event) of the step `ahead` steps after the one take() returns next; it must be in flight
        already (ahead < depth before the current step's enqueue_next)."""
        assert self.tail + ahead < self.head
        b = (self.tail + ahead) % self.nb
        return self.buf[b], self.ready[b]

    def enqueue_next(self):
        """Returns (index buffer of the newly enqueued draw, event that marks it drawn)."""
        return self._enqueue()

    def done_with(self, idx):
        self.free[self.tail % self.nb].record(torch.cuda.current_stream(self.mb.device))
        self.tail += 1

    def finish(self):
        """Make the main stream wait for the outstanding look-ahead draw (it has already advanced the generator)."""
        torch.cuda.current_stream(self.mb.device).wait_stream(self.stream)


class MinibatchPipeline:
    """Double-buffered host -> device staging of minibatch index arrays on a copy stream.

    The indices of step i+1 do not depend on the result of step i (the reference draws them from the Python RNG,
    src/data_set.py:67), so their upload can run behind the kernels of step i:
        pipe.prefetch(idx_pinned_0)
        for i in ...:  idx_dev = pipe.acquire(); engine.step_device(idx_dev, B, params, alpha); pipe.release();
                       pipe.prefetch(idx_pinned_{i+1}); ... read results ..."""

    def __init__(self, B, device):
        self.dev = torch.device(device)
        self.buf = [torch.empty(B, dtype=torch.int64, device=self.dev) for _ in range(2)]
        self.stream = torch.cuda.Stream(device=self.dev)
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        for e in self.free:
            e.record(torch.cuda.current_stream(self.dev))
        self.head = 0          # next buffer to fill
        self.tail = 0          # next buffer to consume

    def prefetch(self, idx_pinned):
        b = self.head & 1
        self.stream.wait_event(self.free[b])
        with torch.cuda.stream(self.stream):
            self.buf[b].copy_(idx_pinned, non_blocking=True)
            self.ready[b].record(self.stream)
        self.head += 1

    def acquire(self):
        b = self.tail & 1
        torch.cuda.current_stream(self.dev).wait_event(self.ready[b])
        return self.buf[b]

    def release(self):
        self.free[self.tail & 1].record(torch.cuda.current_stream(self.dev))
        self.tail += 1


def md_align(X, ref_x, ref_index, diag_coeff, device=None, want_metric=True):
    """`MD_data_set.align` (src/data_set.py:126-153) for every row of X on the GPU, once per data set.

    X [K, 3 * n_atoms] raw coordinates (host array / tensor), ref_x [m, 3], ref_index [m] as read by
    `load_ref_state` (:112-124).  Returns (Xa, metric): the aligned rows as a [K, d_pad] fp32 device tensor
    (zero padding columns, the layout StepEngine takes as a device mirror) and the per-row metric
    M = J diag(diag_coeff) J^T [K, d, d] fp32 of the layer's Jacobian (None if want_metric is False)."""
    if not torch.cuda.is_available():
        raise _lib.EpdeError("md_align needs a CUDA device (no CPU fallback)")
    L = _lib.lib()
    dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    Xt = torch.as_tensor(np.asarray(X), dtype=torch.float32)
    K, d = Xt.shape
    assert d % 3 == 0
    d_pad = (d + 3) // 4 * 4
    with torch.cuda.device(dev):
        Xd = torch.zeros((K, d_pad), dtype=torch.float32, device=dev)
        Xd[:, :d] = Xt.to(dev)
        Xa = torch.zeros_like(Xd)
        metric = torch.empty((K, d, d), dtype=torch.float32, device=dev) if want_metric else None
        diag = torch.as_tensor(np.asarray(diag_coeff), dtype=torch.float32).to(dev).contiguous()
        ref = np.ascontiguousarray(np.asarray(ref_x, dtype=np.float64).reshape(-1, 3))
        ri = np.ascontiguousarray(np.asarray(ref_index, dtype=np.int32))
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.epde_md_align(C.c_int64(K), C.c_int32(d // 3), C.c_int32(d_pad), _ptr(Xd),
                                   ref.ctypes.data_as(C.c_void_p), ri.ctypes.data_as(C.c_void_p), C.c_int32(len(ri)),
                                   _ptr(diag), _ptr(Xa), _ptr(metric), st), "epde_md_align")
        torch.cuda.synchronize(dev)
    return Xa, metric


class DeviceTrainer:
    """Device-resident version of the per-step loop body of `eigen_solver.train`
    (src/pytorch_eigen_solver.py:277-284): minibatch draw (the reference's MT19937 stream, DeviceMinibatch), loss+gradient
    step, Adam, model averaging in cvec order and the eigenvalue running sums all stay on the GPU and are replayed as ONE
    CUDA graph per (batch size, alpha).  Like the reference, the master parameters and the Adam moments are fp64
    (`model.double()` :374, torch.optim.Adam :379); the step's kernels read an fp32 copy refreshed by the epilogue.
    Nothing is synchronised until `read_scalars()` / `download()` is called (every print_every_step / at a stage end).
    With a process group every rank draws the global batch, keeps its slice, and the two exchanges (NVLink peer-memory
    kernel) are part of the graph."""

    def __init__(self, engine: StepEngine, model, use_graph=None, minibatch=None):
        self.e = engine
        dev = engine.device
        self.p64 = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).to(device=dev, dtype=torch.float64)
        self.p = self.p64.to(torch.float32)
        self.m = torch.zeros_like(self.p64)
        self.v = torch.zeros_like(self.p64)
        self.avg = torch.zeros(engine.n_params, dtype=torch.float64, device=dev)
        self.stats = torch.zeros(2 * engine.k, dtype=torch.float64, device=dev)
        # [learning rate, optimiser steps taken so far] in device memory: the step's launches then have constant
        # arguments and the whole step is replayed as ONE CUDA graph (the optimiser is never reset, :379)
        self.state = torch.zeros(2, dtype=torch.float64, device=dev)
        self._lr = None
        self.steps_in_stage = 0
        self._idx_dev = None
        self.mb = minibatch if minibatch is not None else DeviceMinibatch(engine.K, dev, engine.lib).seed_from_python()
        self.world, self.rank = 1, 0
        if engine.pg is not None:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(engine.pg), dist.get_rank(engine.pg)
        if use_graph is None:
            import os
            use_graph = os.environ.get("EPDE_GRAPH", "1") != "0"
        # NCCL collectives between the phases are not captured; the peer-memory exchange is a plain kernel and is
        self.use_graph = bool(use_graph) and (engine.pg is None or engine._exchange is not None)
        self._graph, self._graph_key, self._seen_key = None, None, None

    def start_stage(self):
        self.avg.zero_()
        self.stats.zero_()
        self.steps_in_stage = 0

    def _launch(self, B, alpha):
        from .parallel import shard_bounds
        e = self.e
        lo, hi = shard_bounds(B, self.world, self.rank)
        self.mb.draw(B, self._idx_dev, lo, hi - lo)            # every rank advances the global stream, keeps its slice
        out, grad = e.step_device(self._idx_dev, hi - lo, self.p, alpha)
        st = C.c_void_p(torch.cuda.current_stream(e.device).cuda_stream)
        # Adam, then the averaging with the UPDATED parameters (:277, :284), then the eigenvalue running sums
        _lib.check(e.lib.epde_train_epilogue64(C.byref(e.desc), _ptr(self.p64), _ptr(self.p), _ptr(grad), _ptr(self.m),
                                               _ptr(self.v), _ptr(self.state), _ptr(self.avg), _ptr(self.stats), _ptr(out),
                                               st), "epde_train_epilogue64")

    def step(self, B, lr, alpha):
        """One training step with a GLOBAL batch of B samples drawn on the device."""
        from .parallel import shard_bounds
        e = self.e
        B = int(B)
        alpha = float(alpha)
        lo, hi = shard_bounds(B, self.world, self.rank)
        if self._idx_dev is None or self._idx_dev.numel() != hi - lo:
            self._idx_dev = torch.empty(hi - lo, dtype=torch.int64, device=e.device)
            self._graph = self._graph_key = self._seen_key = None
        if lr != self._lr:
            self.state[0:1].fill_(float(lr))
            self._lr = lr
        key = (B, alpha, e.ws_generation)
        if self._graph_key is not None and self._graph_key[:2] == key[:2] and self._graph_key[2] != e.ws_generation:
            self._graph = self._graph_key = self._seen_key = None     # the workspace moved: the captured addresses are stale
        if not self.use_graph:
            self._launch(B, alpha)
        elif self._graph_key == key:
            self._graph.replay()
        elif self._seen_key == key:                          # second step of this (B, alpha): record, then replay
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._launch(B, alpha)
            self._graph, self._graph_key = g, key
            g.replay()
        else:                                                # first step: eager (allocates the workspace, sets attributes)
            self._launch(B, alpha)
            self._seen_key = (B, alpha, e.ws_generation)
        self.steps_in_stage += 1

    @property
    def t(self):
        """Optimiser steps taken so far (synchronises)."""
        return int(self.state[1].item())

    def read_scalars(self):
        """Host copy of the last step's (eig_vals, cvec, loss, non_penalty, penalty); synchronises."""
        h = self.e.out.cpu().numpy()
        k = self.e.k
        return h[4:4 + k].copy(), h[4 + k:4 + 2 * k].astype(np.int64), h[0], h[1], h[2]

    def download(self, model, source="params"):
        """Copy the device parameters (or the accumulated average) into a host MyNet."""
        flat = (self.p64 if source == "params" else self.avg).cpu()
        off = 0
        with torch.no_grad():
            for prm in model.parameters():
                n = prm.numel()
                prm.copy_(flat[off:off + n].view_as(prm).to(prm.dtype))
                off += n
