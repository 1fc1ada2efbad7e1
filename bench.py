"""Benchmark of the hot path: fused Jacobian+Hessian callback evaluations.

Metric (BASELINE.json): "Jacobian+Hessian callback evals/s at 10^5 mesh nodes".
Workload (``configs[1]``): cart-pole swing-up, Lobatto, mesh scaled to 33 333
sections x 4 nodes = 100 000 collocation nodes (num_x = 500 001,
num_c = 399 997, nnz_G = 3 899 963, nnz_H = 500 000), full-Hessian callbacks.
One *step* = one evaluation of (G values, H values) for one synthetic iterate
(x_tilde ~ U(-0.5, 0.5), lam ~ N(0, 1), sigma = 1) = ONE fused kernel launch.

* ``value``  : device-resident inputs/outputs, CUDA events on the launch stream.
  A ring of buffer sets larger than L2 is cycled so no step finds its inputs or
  last outputs in L2.
* ``e2e``    : same evaluation through the C-ABI call with HOST (pinned) buffers;
  H2D of x/lam/sigma and D2H of the values are inside the timed region.
* ``roofline``: algorithmic bytes 8*(num_x+nnz_G) + 8*(num_x+num_c+nnz_H)
   per launch / mean launch time, against the measured copy
  bandwidth of MEASURED_PEAKS.json.
* ``cpu_baseline``: the oracle port (``oracle/blockwise.py``, numpy, 1 thread)
  on a bounded sample of the same workload.  One thread is the reference's own
  degree of parallelism: its callbacks are CasADi SX virtual-machine evaluations,
  single-threaded by construction ("where time goes").
  ``cpu_baseline.all_cores`` adds, for transparency, the aggregate rate of one such
  single-threaded evaluation stream per host core (independent iterates).
* ``amortised``: informational -- the same kernel with 8 iterates per launch.
* ``--impl reference``: the reference's CPU implementation of this path.  The
  live reference (CasADi) cannot be installed here (no casadi/pyproprop wheel,
  no network), so this arm times the oracle port.

Multi-GPU (``torchrun``): weak scaling, one independent multi-start instance of
the same workload per rank, no data-path collective; time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "Jacobian+Hessian callback evals/s at 10^5 mesh nodes"
UNIT = "evals/s"
K_SECTIONS, N_K = 33333, 4


def workload_config(n_gpus):
    return {"workload": "cart_pole_swing_up explicit, Lobatto, 33333 sections x 4 nodes "
                        "= 100000 collocation nodes, derivative_level=2 (G+H fused)",
            "num_x": 500001, "num_c": 399997, "nnz_G": 3899963, "nnz_H": 500000,
            "inputs": "x~U(-0.5,0.5), lam~N(0,1), sigma=1, numpy default_rng(seed)",
            "l2_policy": "ring of 6 device buffer sets (255 MB > 126 MB L2), "
                         "one set per step",
            "launch": "one fused kernel per evaluation, back to back on one stream with "
                      "programmatic dependent launch (the next kernel's table-only prologue "
                      "overlaps the previous kernel's drain; it waits for that kernel's "
                      "completion before touching x, lam or any output)",
            "parallelism": f"{n_gpus} independent instance(s), one per GPU"}


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.rows = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_rate(seconds_budget=12.0, max_evals=200):
    """Time the oracle port on the same workload; returns (evals/s, n, seconds)."""
    from helpers import build_case
    from pycollo_b200 import examples
    low, B, _ = build_case(examples.cart_pole_swing_up(), "lobatto", K_SECTIONS, N_K,
                           seed=0, unit_scaling=True)
    rng = np.random.default_rng(0)
    x = rng.uniform(-0.5, 0.5, low.S.num_x)
    lam = rng.standard_normal(low.S.num_c)
    B.G_nonzeros(x)
    B.H_nonzeros(x, 1.0, lam)                       # warm-up: builds merge plans
    n, t0 = 0, time.perf_counter()
    while n < max_evals and (time.perf_counter() - t0) < seconds_budget:
        B.G_nonzeros(x)
        B.H_nonzeros(x, 1.0, lam)
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def batched_rate(low, scal, E, torch, dev, stream, alg_bytes, peak, batch=8, steps=20):
    """Same kernel, `batch` independent iterates per launch (multi-start sweep,
    grid.y = batch): what the fixed per-launch costs amortise to.  Informational;
    the headline `value` stays one evaluation per launch."""
    S = low.S
    eng = E.Engine(S, low.layouts, low.header, batch=batch, device=dev.index)
    eng.set_scaling(*scal)
    what = E.EVAL_JAC | E.EVAL_HESS
    g = torch.Generator(device=dev).manual_seed(7)
    R = 2                                             # 2 x 8 x 46 MB > L2
    xs = [torch.rand(batch, S.num_x, dtype=torch.float64, device=dev, generator=g) - 0.5 for _ in range(R)]
    ls = [torch.randn(batch, S.num_c, dtype=torch.float64, device=dev, generator=g) for _ in range(R)]
    js = [torch.empty(batch, S.nnz_g, dtype=torch.float64, device=dev) for _ in range(R)]
    hs = [torch.empty(batch, S.nnz_h, dtype=torch.float64, device=dev) for _ in range(R)]
    for i in range(3):
        eng.eval_ptr(what, xs[i % R], lam=ls[i % R], jac=js[i % R], hess=hs[i % R], stream=stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        eng.eval_ptr(what, xs[i % R], lam=ls[i % R], jac=js[i % R], hess=hs[i % R], stream=stream)
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / steps / batch
    gbs = alg_bytes / us / 1e3
    return {"batch_per_launch": batch, "us_per_eval": us, "evals_per_s": 1e6 / us,
            "achieved_GBs": gbs, "frac": gbs / peak,
            "note": "same kernel, 8 independent iterates per launch (grid.y); informational"}


def _cpu_worker(seconds, q):
    """One process of the all-cores CPU throughput leg: the oracle port on the same
    workload, evaluations counted over a fixed wall-clock window."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    from helpers import build_case
    from pycollo_b200 import examples
    low, B, _ = build_case(examples.cart_pole_swing_up(), "lobatto", K_SECTIONS, N_K,
                           seed=0, unit_scaling=True)
    rng = np.random.default_rng(os.getpid())
    x = rng.uniform(-0.5, 0.5, low.S.num_x)
    lam = rng.standard_normal(low.S.num_c)
    B.G_nonzeros(x)
    B.H_nonzeros(x, 1.0, lam)
    q.put(("ready", os.getpid()))
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        B.G_nonzeros(x)
        B.H_nonzeros(x, 1.0, lam)
        n += 1
    q.put(("done", n, time.perf_counter() - t0))


def cpu_oracle_rate_all_cores(seconds=8.0, max_procs=16):
    """Aggregate evals/s of `cores` independent evaluation streams, one process per
    host core (the port is single-threaded, as the reference's evaluation is; this
    is the multi-start throughput a host could reach with the same algorithm)."""
    import multiprocessing as mp
    cores = max(1, min(max_procs, os.cpu_count() or 1))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(seconds, q)) for _ in range(cores)]
    for p in procs:
        p.start()
    total, longest, done = 0, 0.0, 0
    deadline = time.perf_counter() + 120.0 + seconds
    while done < cores:
        try:
            msg = q.get(timeout=5)
        except Exception:
            if time.perf_counter() > deadline or not any(p.is_alive() for p in procs):
                for p in procs:
                    if p.is_alive():
                        p.terminate()
                raise RuntimeError("all-cores CPU leg: workers did not report")
            continue
        if msg[0] == "done":
            total += msg[1]
            longest = max(longest, msg[2])
            done += 1
    for p in procs:
        p.join(timeout=30)
    return {"value": total / longest, "unit": UNIT, "cores": cores,
            "how": f"{cores} processes, one independent evaluation stream each, "
                   f"{total} evaluations in {longest:.1f} s (windows overlap after a "
                   f"staggered start: an upper bound of the sustained rate)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warmup = args.steps, args.warmup
    from helpers import build_case
    from pycollo_b200 import examples
    low, B, _ = build_case(examples.cart_pole_swing_up(), "lobatto", K_SECTIONS, N_K,
                           seed=0, unit_scaling=True)
    rng = np.random.default_rng(0)
    xs = [rng.uniform(-0.5, 0.5, low.S.num_x) for _ in range(2)]
    lams = [rng.standard_normal(low.S.num_c) for _ in range(2)]
    for i in range(max(1, min(warmup, 3))):
        B.G_nonzeros(xs[i % 2])
        B.H_nonzeros(xs[i % 2], 1.0, lams[i % 2])
    steps = max(1, min(steps, 60))                  # bounded: ~0.25 s per eval
    t0 = time.perf_counter()
    for i in range(steps):
        B.G_nonzeros(xs[i % 2])
        B.H_nonzeros(xs[i % 2], 1.0, lams[i % 2])
    dt = time.perf_counter() - t0
    value = steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": f"{steps} fused G+H evaluations of the full "
                                       f"10^5-node workload by oracle/blockwise.py "
                                       f"(numpy, 1 thread); the live reference "
                                       f"(CasADi) is not installable here"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    try:
        line["cpu_baseline"]["all_cores"] = cpu_oracle_rate_all_cores()
    except Exception as exc:
        line["cpu_baseline"]["all_cores"] = {"error": repr(exc)[:200]}
    print(json.dumps(line))
    return 0


def run_cuda(args):
    import torch
    import torch.distributed as dist
    from helpers import build_case, make_engine
    from pycollo_b200 import engine as E
    from pycollo_b200 import examples

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    steps, warmup = args.steps, max(args.warmup, 3)

    low, _, scal = build_case(examples.cart_pole_swing_up(), "lobatto", K_SECTIONS, N_K,
                              seed=0, unit_scaling=True, oracle=False)
    S = low.S
    eng = E.Engine(S, low.layouts, low.header, batch=1, device=local_rank)
    eng.set_scaling(*scal)
    what = E.EVAL_JAC | E.EVAL_HESS
    alg_bytes = 8 * (S.num_x + S.nnz_g) + 8 * (S.num_x + S.num_c + S.nnz_h)

    # ---- device-resident ring (> L2) ------------------------------------
    R = 6
    rng = np.random.default_rng(1000 + rank)
    dev = torch.device("cuda", local_rank)
    xs = [torch.from_numpy(rng.uniform(-0.5, 0.5, S.num_x)).to(dev) for _ in range(R)]
    lams = [torch.from_numpy(rng.standard_normal(S.num_c)).to(dev) for _ in range(R)]
    jacs = [torch.empty(S.nnz_g, dtype=torch.float64, device=dev) for _ in range(R)]
    hess = [torch.empty(S.nnz_h, dtype=torch.float64, device=dev) for _ in range(R)]
    stream = torch.cuda.current_stream().cuda_stream

    # one pre-bound C-ABI call per ring slot: pcx_eval with its arguments converted
    # once, as a compiled host would hold them
    calls = [eng.bind(what, xs[k], lam=lams[k], jac=jacs[k], hess=hess[k], stream=stream)
             for k in range(R)]

    def step(i):
        calls[i % R]()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    l0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(steps):
        step(i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - l0

    # ---- end to end through the C ABI with pinned host buffers ------------
    hx = torch.from_numpy(rng.uniform(-0.5, 0.5, S.num_x)).pin_memory()
    hl = torch.from_numpy(rng.standard_normal(S.num_c)).pin_memory()
    hs = torch.ones(1, dtype=torch.float64).pin_memory()
    hj = torch.empty(S.nnz_g, dtype=torch.float64).pin_memory()
    hh = torch.empty(S.nnz_h, dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(steps, 50))
    for _ in range(3):
        eng.eval_ptr(what, hx, lam=hl, sigma=hs, jac=hj, hess=hh, space=E.PCX_HOST,
                     stream=stream)
    barrier()
    l1 = eng.launch_count
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.eval_ptr(what, hx, lam=hl, sigma=hs, jac=hj, hess=hh, space=E.PCX_HOST,
                     stream=stream)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    launches += eng.launch_count - l1
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms, 1e3 * e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = float(t[0]), float(t[1])
    if rank == 0:
        value = world * steps / (ms_max * 1e-3)
        e2e_value = world * e2e_steps / (e2e_ms_max * 1e-3)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
        else:
            peak, peak_src = 6650.0, "fallback"
        achieved = alg_bytes / (ms_max * 1e-3 / steps) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        amortised = None
        if world == 1:
            amortised = batched_rate(low, scal, E, torch, dev, stream, alg_bytes, peak)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, n, dt = cpu_oracle_rate()
            cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"{n} fused G+H evaluations of the full 10^5-node "
                             f"workload in {dt:.1f} s by oracle/blockwise.py "
                             f"(numpy, 1 thread)"}
            try:
                cpu["all_cores"] = cpu_oracle_rate_all_cores()
            except Exception as exc:                      # never fail the bench for this leg
                cpu["all_cores"] = {"error": repr(exc)[:200]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": ms_max / steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": workload_config(world),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": alg_bytes,
                             "kernel": "pcx_fill_12 (fused Jacobian+Hessian fill)"},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": UNIT,
                        "h2d_bytes_per_step": 8 * (S.num_x + S.num_c + 1),
                        "d2h_bytes_per_step": 8 * (S.nnz_g + S.nnz_h),
                        "steps": e2e_steps, "api": "pcx_eval(..., PCX_HOST) on pinned "
                                                   "host buffers"},
                "gpu_launches": int(launches), "clocks": clocks,
                "amortised": amortised}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
