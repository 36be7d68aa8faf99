#!/usr/bin/env python3
"""Benchmark of the fused loss+gradient step (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B]

Workload (config.workload = "ex1_50d_k3"): 50-D, arch 50-20-20-20-1, Tanh, k = 3, beta = 1,
alpha = 20, eig_w = (1.0, 0.8, 0.6) -- examples/test-ex1-50d/params.cfg -- on a synthetic
trajectory of 2^22 states (ring in (x0,x1), N(0,0.1) elsewhere), lognormal weights.
One step = one pass of the hot path (src/pytorch_eigen_solver.py:212-233) over one minibatch of
B = 2^20 samples per GPU (weak scaling: the global batch is N * 2^20, sharded by contiguous slices).

Prints ONE JSON line (rank 0).  `value` = samples/s with the index batch already on the device;
`e2e` = the same step through the host-facing API: every step draws a FRESH minibatch from the reference's
MT19937 stream on the device (bit-exact with random.choices, src/data_set.py:67; generation inside the timed
region, one step ahead on a second stream), uploads the parameters from pinned host memory and reads the
gradient + scalars back.  With N > 1 a pre-timing self-check compares the N-rank step with a single-rank
step on the union batch (`parity`).  `extra` (N = 1 only) carries BASELINE configs 4 and 5 and config 3 in MD mode.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, KNETS, ARCH = 50, 3, [50, 20, 20, 20, 1]
EIG_W, BETA, ALPHA = [1.0, 0.8, 0.6], 1.0, 20.0
K_DATA = 1 << 22
M_MACS = 50 * 20 + 20 * 20 + 20 * 20 + 20
FLOP_ALG = 2 * KNETS * (6 * M_MACS - 50 * 20)            # 59 520 per sample
FLOP_ALG_P1 = 2 * KNETS * 2 * M_MACS                     # forward + input-Jacobian share (phase 1)
BYTES_ALG = 4 * D + 4 + 8                                # x row + weight + int64 index
FFMA_PEAK_TFLOPS = 72.99     # measured on this pool's B200: profiles/ffma_peak_r1.jsonl (nominal 74.4)


def _tensor_peak():
    """Dense bf16 TFLOP/s from the driver-written MEASURED_PEAKS.json (sustained figure), else the recipe's fallback."""
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(mp["bf16_tflops_sustained"]), "of measured: MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS bf16 8192^3 back to back)"
    except Exception:
        return 1400.0, "of fallback: B200_PROFILING.md sustained figure (MEASURED_PEAKS.json absent)"


TENSOR_PEAK_TFLOPS, TENSOR_PEAK_SOURCE = _tensor_peak()


def synth_data(device, seed, K=K_DATA):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    X = torch.zeros((K, 52), dtype=torch.float32, device=device)
    X[:, :D] = (0.1 ** 0.5) * torch.randn((K, D), generator=g, device=device)
    th = 2 * torch.pi * torch.rand(K, generator=g, device=device)
    r = 1.0 + 0.3 * torch.randn(K, generator=g, device=device)
    X[:, 0] = r * torch.cos(th)
    X[:, 1] = r * torch.sin(th)
    w = torch.exp(0.5 * torch.randn(K, generator=g, device=device))
    w /= w.mean()
    return X, w


def init_params(seed):
    """N(0, 0.5) for every weight and bias, as network_arch.py:37-39."""
    import torch
    g = torch.Generator().manual_seed(seed)
    n = KNETS * sum(ARCH[l + 1] * ARCH[l] + ARCH[l + 1] for l in range(len(ARCH) - 1))
    return 0.5 * torch.randn(n, generator=g)


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.t_lo, self.t_hi = None, None

    def mark_start(self):
        self.t_lo = time.time()

    def mark_stop(self):
        self.t_hi = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        lo = self.t_lo or 0.0
        hi = self.t_hi or time.time()
        rows = [r for (t, r) in self.rows if lo <= t <= hi + 0.05 and len(r) >= 9]
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) >= 9:
                for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        smax = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        busy = [v for v in sm if v > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_port_rate(B, steps, warmup, seed=0):
    """samples/s of the oracle port (oracle/ref_port.py, torch fp64 autograd on the host cores):
    (a) update_step as shipped (minibatch + loss + backward + Adam), (b) loss + gradient only."""
    import numpy as np
    import torch
    from oracle import ref_port
    # all host threads the box offers: torchrun exports OMP_NUM_THREADS=1 to its workers, which would make the CPU arm
    # single-threaded in the N > 1 runs
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    if torch.get_num_threads() < ncpu:
        torch.set_num_threads(ncpu)
    rng = np.random.default_rng(seed)
    Kd = max(4 * B, 100000)
    X = np.sqrt(0.1) * rng.standard_normal((Kd, D))
    th = rng.uniform(0, 2 * np.pi, Kd)
    r = 1.0 + 0.3 * rng.standard_normal(Kd)
    X[:, 0], X[:, 1] = r * np.cos(th), r * np.sin(th)
    w = np.exp(0.5 * rng.standard_normal(Kd)); w /= w.mean()
    torch.manual_seed(seed)
    model = ref_port.PortNet(ARCH, "Tanh", KNETS).double()
    solver = ref_port.PortSolver(model, X, w, np.ones(D), BETA, EIG_W)
    import random
    random.seed(seed)
    t_full, t_lg = [], []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, xb, wb = solver.minibatch(B)
        t1 = time.perf_counter()
        solver.loss_grad(xb, wb, ALPHA)
        t2 = time.perf_counter()
        solver.optimizer.step()
        t3 = time.perf_counter()
        if i >= warmup:
            t_full.append(t3 - t0); t_lg.append(t2 - t1)
    return B / statistics.median(t_full), B / statistics.median(t_lg), torch.get_num_threads()


def extra_workloads(dev, X, w):
    """BASELINE configs 4 and 5 as driver-visible numbers (N = 1): one rank's shard of config 4 (50-D, k = 6, 2^21 of the
    2^24 samples per step) and config 5 (wide MLP 4 x 256, 50-D, k = 4, per-layer tcgen05 GEMMs of the generic family)."""
    import numpy as np
    import torch
    from eigenpde_nn_b200.engine import StepEngine
    out = {}

    def timed(eng, B, steps, warm=3):
        g = torch.Generator(device=dev).manual_seed(5)
        idx = [torch.randint(0, K_DATA, (B,), generator=g, device=dev, dtype=torch.int64) for _ in range(2)]
        params = (0.5 * torch.randn(eng.n_params, generator=torch.Generator().manual_seed(9))).to(dev)
        for i in range(warm):
            eng.step_device(idx[i & 1], B, params, ALPHA)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            eng.step_device(idx[i & 1], B, params, ALPHA)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, float(eng.out[0])

    B4 = 1 << 21
    eng = StepEngine(X, w, ARCH, 6, "Tanh", np.ones(D), BETA, [1.0, 0.8, 0.6, 0.3, 0.2, 0.1], device=dev)
    ms, loss = timed(eng, B4, 10)
    flop4 = 2 * 6 * (6 * M_MACS - 50 * 20)
    out["cfg4_50d_k6_shard"] = {"workload": "50-D, k=6, one rank's 2^21-sample shard of the 2^24-sample step", "batch": B4,
                                "samples_per_s": B4 / (ms * 1e-3), "ms_per_step": ms, "flop_per_sample": flop4,
                                "frac_fp32_ffma": flop4 * B4 / (ms * 1e-3) / 1e12 / FFMA_PEAK_TFLOPS, "fused": bool(eng.fused),
                                "loss": loss}
    del eng
    torch.cuda.empty_cache()
    Bw = 1 << 18
    archw = [50, 256, 256, 256, 256, 1]
    eng = StepEngine(X, w, archw, 4, "Tanh", np.ones(D), BETA, [1.0, 0.8, 0.6, 0.3], device=dev)
    ms, loss = timed(eng, Bw, 4, warm=2)
    Mw = sum(archw[l] * archw[l + 1] for l in range(len(archw) - 1))
    flopw = 2 * 4 * (6 * Mw - archw[0] * archw[1])
    out["cfg5_wide_4x256_k4"] = {"workload": "50-D, 4x256 Tanh, k=4 (per-layer tcgen05 3xTF32 GEMMs)", "batch": Bw,
                                 "samples_per_s": Bw / (ms * 1e-3), "ms_per_step": ms, "flop_per_sample": flopw,
                                 "achieved_tflops": flopw * Bw / (ms * 1e-3) / 1e12,
                                 "frac_tf32x3_ceiling": flopw * Bw / (ms * 1e-3) / 1e12 / (TENSOR_PEAK_TFLOPS / 6.0),
                                 "fused": bool(eng.fused), "loss": loss}
    del eng
    torch.cuda.empty_cache()
    # BASELINE config 3 WITH the alignment layer: 30-D (10 atoms), k = 2, B = 10000 of K = 1e5 frames, NAMD-style
    # weights; epde_md_align once per data set, then the step with the per-row metric on the fused FFMA2 kernels + k_metric
    from eigenpde_nn_b200.engine import md_align
    rng = np.random.default_rng(0)
    Km, dm, Bm = 100000, 30, 10000
    Xm = (1.5 * rng.standard_normal((1, dm)) + np.sqrt(0.1) * rng.standard_normal((Km, dm))).astype(np.float32)
    wm = np.exp(2.0 * rng.standard_normal(Km)).astype(np.float32); wm /= wm.mean()
    ri = [0, 2, 5, 7, 8]
    ref = Xm[0].reshape(-1, 3)[ri].astype(np.float64); ref -= ref.mean(0, keepdims=True)
    beta_md = 1.0 / (0.001987191 * 300.0)
    diag = np.full(dm, 1e-5 * 1e7 * beta_md)
    Xa, Mx = md_align(Xm, ref, ri, diag, device=dev)
    eng = StepEngine(Xa, torch.from_numpy(wm), [dm, 20, 20, 20, 1], 2, "Tanh", diag, beta_md, [1.0, 0.8], device=dev, metric=Mx)
    idx = torch.randint(0, Km, (Bm,), generator=torch.Generator(device=dev).manual_seed(3), device=dev, dtype=torch.int64)
    params = (0.5 * torch.randn(eng.n_params, generator=torch.Generator().manual_seed(9))).to(dev)
    for _ in range(3):
        eng.step_device(idx, Bm, params, ALPHA)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.step_device(idx, Bm, params, ALPHA)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out["cfg3_md_30d_k2_aligned"] = {"workload": "30-D MD frames through the alignment layer (per-row metric), k=2, B=10000",
                                     "batch": Bm, "samples_per_s": Bm / (ms * 1e-3), "ms_per_step": ms, "fused": bool(eng.fused),
                                     "loss": float(eng.out[0])}
    del eng, Xa, Mx
    torch.cuda.empty_cache()
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path, timed on the host cores.
    /root/reference is Python and cannot travel to the GPU box; the oracle port restates it op by op
    (pinned against the reference's own outputs in tests/golden)."""
    if rank != 0:
        return
    B = 20000                      # the reference's own stage-2 batch size (params.cfg:39)
    t0 = time.perf_counter()
    full, lg, cores = cpu_port_rate(B, args.steps, args.warmup)
    line = {"impl": "reference", "metric": "samples/s, fused loss+grad step, 50-D k=3", "value": lg,
            "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * B / lg, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "ex1_50d_k3", "d": D, "k": KNETS, "arch": ARCH, "batch_per_step": B},
            "cpu_baseline": {"value": lg, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": "B=20000 per step, %d timed steps, loss+grad only (update_step as shipped incl. "
                                       "minibatch+Adam: %.0f samples/s)" % (args.steps, full)},
            "e2e": {"value": lg, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=1 << 20, help="samples per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the config-4 / config-5 side measurements (N = 1)")
    ap.add_argument("--knets", type=int, default=3, help="k (default 3 = the headline workload; 6 with --batch 2097152 on "
                                                         "8 GPUs = BASELINE config 4, an ad-hoc run, not the bench line)")
    args = ap.parse_args()
    if args.knets != 3:
        global KNETS, EIG_W, FLOP_ALG, FLOP_ALG_P1
        KNETS = args.knets
        EIG_W = [1.0, 0.8, 0.6, 0.3, 0.2, 0.1][:KNETS]
        FLOP_ALG = 2 * KNETS * (6 * M_MACS - 50 * 20)
        FLOP_ALG_P1 = 2 * KNETS * 2 * M_MACS
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    from eigenpde_nn_b200.engine import StepEngine

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    W = max(args.warmup, 3)
    B = args.batch
    X, w = synth_data(dev, seed=1234)                       # replicated data set (872 MB fp32)
    eng = StepEngine(X, w, ARCH, KNETS, "Tanh", np.ones(D), BETA, EIG_W, device=dev, process_group=pg)
    assert eng.fused
    params_h = init_params(7).pin_memory()
    params = params_h.to(dev)
    # minibatch indices: the reference's own stream (MT19937, random.seed on every rank -> the same GLOBAL stream), drawn
    # on the device; rank r keeps the r-th contiguous slice of each global batch
    import random
    from eigenpde_nn_b200.engine import DeviceMinibatch, DrawAhead
    random.seed(20251020)
    mb = DeviceMinibatch(K_DATA, dev).seed_from_python()
    nbatch = 4                                              # distinct index batches for the device-resident timing, cycled
    idx_dev = [mb.draw(world * B, lo=rank * B, n=B) for _ in range(nbatch)]

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    exchange = "none" if world == 1 else (("peer-folded" if eng._fold_exchange() else "peer") if eng._exchange is not None else "nccl")
    # ---- N > 1: the sharded step against a single-rank step on the union batch (same kernels, no exchange) -----------
    parity = None
    if world > 1:
        idx_union = mb.draw(world * B)                      # every rank holds the whole global batch of this draw
        o_n, g_n = eng.step_device(idx_union[rank * B:(rank + 1) * B].contiguous(), B, params, ALPHA)
        o_n, g_n = o_n.clone(), g_n.clone()
        single = StepEngine(X, w, ARCH, KNETS, "Tanh", np.ones(D), BETA, EIG_W, device=dev)
        o_1, g_1 = single.step_device(idx_union, world * B, params, ALPHA)
        torch.cuda.synchronize()
        rel_out = float(((o_n[:4 + KNETS] - o_1[:4 + KNETS]).abs() / o_1[:4 + KNETS].abs().clamp_min(1e-300)).max())
        rel_grad = float((g_n - g_1).norm() / g_1.norm())
        same_cvec = bool(torch.equal(o_n[4 + KNETS:4 + 2 * KNETS], o_1[4 + KNETS:4 + 2 * KNETS]))
        t = torch.tensor([rel_out, rel_grad, 0.0 if same_cvec else 1.0], dtype=torch.float64, device=dev)
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        parity = {"n_rank_vs_single_rank": {"scalars_max_rel": float(t[0]), "grad_norm_rel": float(t[1]),
                                            "cvec_equal": bool(t[2] == 0.0)},
                  "tolerance": 1e-5, "ok": bool(t[0] <= 1e-5 and t[1] <= 1e-5 and t[2] == 0.0), "exchange": exchange}
        del single, idx_union
        torch.cuda.empty_cache()

    # ---- device-resident timing ------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # every step announces the next step's (already drawn) index batch: its rows are gathered on a second stream while the
    # step computes (engine._step_ahead; EPDE_GATHER_AHEAD=0 switches the look-ahead off).  The loops run on the engine's
    # compute stream (the high-priority stream of the look-ahead scheme) so that no hand-over events are needed per step.
    cs = eng.compute_stream()
    if cs is not None:
        cs.wait_stream(torch.cuda.current_stream(dev))
        torch.cuda.set_stream(cs)
    for i in range(W):
        eng.step_device(idx_dev[i % nbatch], B, params, ALPHA, next_idx=idx_dev[(i + 1) % nbatch])
    barrier()
    sampler.mark_start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        eng.step_device(idx_dev[i % nbatch], B, params, ALPHA, next_idx=idx_dev[(i + 1) % nbatch])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    out_h = eng.out.cpu().numpy()
    # self-check of the look-ahead path: the last timed step (rows gathered one step ahead, SMs reserved) against the same step
    # with its own gather (the pending look-ahead does not match it, so the step gathers for itself)
    ahead_check = None
    if eng.compute_stream() is not None:
        last = idx_dev[(args.steps - 1) % nbatch]
        o_a, g_a = eng.out.clone(), eng.grad.clone()
        eng.step_device(last, B, params, ALPHA)
        rel_o = float(((eng.out[:4] - o_a[:4]).abs() / o_a[:4].abs().clamp_min(1e-300)).max())
        rel_g = float((eng.grad - g_a).norm() / g_a.norm())
        ahead_check = {"scalars_max_rel": rel_o, "grad_norm_rel": rel_g, "ok": bool(rel_o <= 1e-9 and rel_g <= 1e-5)}

    # ---- per-phase timing of the dominant kernel (k_pass2 inside epde_step_finish) ----------
    import ctypes as C
    from eigenpde_nn_b200 import _lib
    ws = eng._workspace(B)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    eigw = (C.c_double * KNETS)(*EIG_W)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_p1 = t_p2 = 0.0
    nphase = min(args.steps, 10)
    for i in range(nphase):
        ib = idx_dev[i % nbatch]
        a = (C.byref(eng.desc), X.data_ptr(), w.data_ptr(), ib.data_ptr(), B, params.data_ptr(), eng.diag.data_ptr())
        evs[0].record()
        _lib.check(eng.lib.epde_step_partial(*a, ws.data_ptr(), ws.numel(), eng.sums.data_ptr(), st), "partial")
        evs[1].record()
        _lib.check(eng.lib.epde_step_finish(*a, BETA, ALPHA, eigw, eng.sums.data_ptr(), ws.data_ptr(), ws.numel(),
                                            eng.out.data_ptr(), eng.grad.data_ptr(), st), "finish")
        evs[2].record()
        torch.cuda.synchronize()
        t_p1 += evs[0].elapsed_time(evs[1]); t_p2 += evs[1].elapsed_time(evs[2])
    t_p1 /= nphase; t_p2 /= nphase

    # ---- end to end through the host-facing API --------------------------------------------
    res_pin = torch.empty(eng.result.numel(), dtype=torch.uint8).pin_memory()      # scalars (fp64) + gradient (fp32), one copy
    out_pin = res_pin[:8 * eng.out.numel()].view(torch.float64)
    par_stage = torch.empty_like(params)
    from eigenpde_nn_b200.engine import LATER
    use_ahead = eng._ahead_ok()
    # the draws of the next TWO steps are kept in flight: step i + 1's indices are complete while step i runs (its rows are
    # gathered beside step i's phase 2), step i + 2's are being generated
    ahead = DrawAhead(mb, world * B, lo=rank * B, n=B, depth=2 if use_ahead else 1).start()

    def e2e_step(i):
        idx_stage = ahead.take(defer=True)                             # fresh draw from the MT19937 stream (device, side stream)
        par_stage.copy_(params_h, non_blocking=True)                   # H2D: current parameters
        eng.step_device(idx_stage, B, par_stage, ALPHA, next_idx=LATER if use_ahead else None)
        if use_ahead:
            nxt, drawn = ahead.peek(1)
            eng.gather_next(nxt, after=drawn)                          # the next step's rows are gathered beside this step's phase 2
        ahead.done_with(idx_stage)
        res_pin.copy_(eng.result, non_blocking=True)                   # D2H: loss / eigenvalues / penalty + gradient for the host optimiser
        ahead.enqueue_next()                                           # one more draw, behind this step's kernels
        torch.cuda.current_stream().synchronize()                      # update_step returns host values (:236)
        return float(out_pin[0])

    for i in range(W):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    barrier()
    e2e_s = time.perf_counter() - t0
    # clocks are sampled (20 ms period) over the whole measured span: device-timed steps, per-phase
    # timing and the end-to-end steps -- all of them back-to-back launches of the same kernels
    sampler.mark_stop()
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms, e2e_s, t_p1, t_p2], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s, t_p1, t_p2 = (float(v) for v in t.cpu())
    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    total = world * B * args.steps
    value = total / (ms * 1e-3)
    ach2 = (FLOP_ALG - FLOP_ALG_P1) * B / (t_p2 * 1e-3) / 1e12
    traffic = None
    prof = os.path.join(ROOT, "profiles", "pass2_dram_bytes.json")
    if os.path.exists(prof):
        pj = json.load(open(prof))
        traffic = pj.get("dram_bytes_per_launch") if pj.get("batch") == B else None
    line = {
        "metric": "samples/s, fused loss+grad step, 50-D k=%d" % KNETS, "value": value, "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ex1_50d_k%d" % KNETS, "d": D, "k": KNETS, "arch": ARCH, "activation": "Tanh",
                   "batch_per_gpu": B, "global_batch": world * B, "dataset_rows": K_DATA,
                   "parallelism": ("dp%d (batch sharded by contiguous slices; 2 small all-reduces per step: %s)" % (
                       world, {"peer": "one-shot NVLink peer-memory kernel epde_peer_allreduce",
                               "peer-folded": "NVLink peer-memory exchanges inside the step's own kernels (epde_step_sharded: "
                                              "k_reduce_coef_tc and the last CTA of k_fin2), no exchange launch"}.get(exchange, "NCCL")))
                   if world > 1 else "single GPU", "exchange": exchange,
                   "l2": "inputs larger than L2: 872 MB data set gathered by fresh random indices each step"},
        "loss": float(out_h[0]), "eig_vals": [float(v) for v in out_h[4:4 + KNETS]],
        # dominant kernel: k_pass2_tc (tcgen05: kind::tf32 3xTF32 mat-vecs, kind::f16 two-term-split gradient rounds).  `peak` is the driver-measured dense bf16
        # figure (sustained: the kernel is timed inside a long step); a TF32 operand pair runs at half that rate and
        # the fp32-accurate split needs three MMAs per product, so 1/6 of `peak` is the ceiling of this formulation
        # before any padding (N 20 -> 32, stacked gradient rounds).  `fp32_ffma` keeps the FP32-pipe comparison of the
        # FFMA2 kernel family (epde_set_fused_impl(0)): the same algorithmic FLOP against the measured FFMA peak.
        "roofline": {"bound": "tensor", "kernel": "k_pass2_tc (epde_step_finish)", "achieved": ach2,
                     "peak": TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ach2 / TENSOR_PEAK_TFLOPS,
                     "traffic": traffic, "ms_per_launch": t_p2,
                     "peak_source": TENSOR_PEAK_SOURCE,
                     "tf32x3_ceiling": {"peak": TENSOR_PEAK_TFLOPS / 6.0, "frac": ach2 / (TENSOR_PEAK_TFLOPS / 6.0)},
                     "fp32_ffma": {"peak": FFMA_PEAK_TFLOPS, "frac": ach2 / FFMA_PEAK_TFLOPS,
                                   "peak_source": "measured FFMA peak, tools/microbench/ffma_peak.cu (nominal 74.4)"},
                     "step": {"achieved": FLOP_ALG * B / ((t_p1 + t_p2) * 1e-3) / 1e12,
                              "frac": FLOP_ALG * B / ((t_p1 + t_p2) * 1e-3) / 1e12 / TENSOR_PEAK_TFLOPS,
                              "frac_fp32_ffma": FLOP_ALG * B / ((t_p1 + t_p2) * 1e-3) / 1e12 / FFMA_PEAK_TFLOPS,
                              "ms_phase1": t_p1, "ms_phase2": t_p2, "flop_per_sample": FLOP_ALG},
                     "hbm": {"achieved_gbs": BYTES_ALG * B / ((t_p1 + t_p2) * 1e-3) / 1e9, "peak_gbs": 6544.3}},
        "e2e": {"value": total / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": 4 * eng.n_params,
                "d2h_bytes_per_step": 8 * eng.out.numel() + 4 * eng.n_params,
                "minibatch": "fresh draw every step from the reference's MT19937 stream, generated on the device inside the "
                             "timed region (epde_minibatch_indices_device), one step ahead on a second stream; no index upload; "
                             + ("every rank jumps ahead to its slice of the %d-draw global batch (x^J mod the characteristic "
                                "polynomial, epde_mt_jump_device) and generates only its own %d draws" % (world * B, B)
                                if world > 1 else "%d draws per step" % B)},
        # kernels per step: k_pack_tc, k_gather128, k_pass1_tc, k_sums_y, k_reduce_coef_tc, k_pass2_tc, k_fin2 - also with N > 1 when the
        # exchanges are folded into k_reduce_coef_tc / k_fin2; otherwise k_reduce_tc + k_coef and two exchange launches (10)
        "gpu_launches": (7 if exchange in ("none", "peer-folded") else 10) * args.steps, "clocks": clocks,
    }
    if parity is not None:
        line["parity"] = parity
    if ahead_check is not None:
        line["gather_look_ahead"] = dict(ahead_check, reserve_sms=eng.RESERVE_SMS,
                                         note="value and e2e announce the next (already drawn) minibatch to every step: its rows are "
                                              "gathered on a second stream beside phase 2; EPDE_GATHER_AHEAD=0 disables it")
    if world == 1 and not args.no_extra and KNETS == 3:
        line["extra"] = extra_workloads(dev, X, w)
    if world == 1 and not args.no_cpu_baseline:
        full, lg, cores = cpu_port_rate(20000, 12, 3)
        line["cpu_baseline"] = {"value": lg, "unit": "samples/s", "cores": cores, "kind": "port",
                                "sample": "oracle/ref_port.py (torch fp64 autograd), B=20000, 12 timed steps, loss+grad "
                                          "only; update_step as shipped (minibatch+Adam): %.0f samples/s" % full}
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
