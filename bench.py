"""Benchmark of the hot path: fused Jacobian+Hessian callback evaluations.

Metric (BASELINE.json): "Jacobian+Hessian callback evals/s at 10^5 mesh nodes".
Workload (``configs[1]``): cart-pole swing-up, Lobatto, mesh scaled to 33 333
sections x 4 nodes = 100 000 collocation nodes (num_x = 500 001,
num_c = 399 997, nnz_G = 3 899 963, nnz_H = 500 000), full-Hessian callbacks.
One *step* = one evaluation of (G values, H values) for one synthetic iterate
(x_tilde ~ U(-0.5, 0.5), lam ~ N(0, 1), sigma = 1) = ONE fused kernel launch.

* ``value``  : device-resident inputs/outputs.  The K timed launches are issued by
  ONE foreign call (``pcx_eval_many``): the stream is gated until all K are
  enqueued and the bracketing CUDA events are recorded on the launch stream inside
  that call, so the number does not depend on how fast Python enqueues.  A ring of
  buffer sets larger than L2 is cycled so no step finds its inputs or last outputs
  in L2.
  The ring's iterates are independent (a sweep / multi-start stream) and are
  declared so (``PCX_EVAL_INDEPENDENT``): a kernel does not wait for its predecessor's
  completion.  ``ordered`` reports the same K launches in stream order.
* ``latency_us``: one evaluation at a time, synchronised before and after, cold
  ring slot -- what a strictly sequential host (one IPOPT) sees per callback.
* ``e2e``    : same evaluation through the C-ABI call with HOST (pinned) buffers;
  H2D of x/lam/sigma and D2H of the values are inside the timed region.
* ``roofline``: algorithmic bytes 8*(num_x+nnz_G) + 8*(num_x+num_c+nnz_H)
   per launch / mean launch time, against the measured copy
  bandwidth of MEASURED_PEAKS.json.
* ``cpu_baseline``: the oracle port (``oracle/blockwise.py``, numpy, 1 thread) on a
  bounded sample of the same workload -- one thread is the reference's own degree
  of parallelism (CasADi SX virtual machine driven by one IPOPT).  Beside it:
  ``numba`` (same port, node functions in one ``numba.prange`` kernel over all host
  cores) and ``all_cores`` (one single-threaded stream per
  core on independent iterates).
* ``strong`` (N > 1 only): BASELINE config 4 -- ONE Delta III mesh (4 phases,
  ~10^6 nodes) split over the N ranks by tile ranges, border values exchanged
  inside the kernel over NVLink peer memory; checked in-process against the
  unsharded evaluation, timed as max over ranks.
* ``config4_1gpu`` (N = 1 only, informational): the same Delta III mesh on ONE GPU --
  the large-expression-body kernel (four phases sharing one body), ms per fused
  G+H evaluation and its fraction of the measured HBM rate (``--no-config4`` skips it).
* ``--impl reference``: the reference's CPU implementation of this path.  The
  live reference (CasADi) cannot be installed here (no casadi/pyproprop wheel,
  no network), so this arm times the oracle port with every host
  thread it can use -- inside one evaluation (numba node kernel) and as one
  independent stream per core; `value` is the better, the 1-thread figure beside it.

Multi-GPU (``torchrun``): the headline is weak scaling, one independent
multi-start instance of the same workload per rank, no data-path collective;
time = max over ranks.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Jacobian+Hessian callback evals/s at 10^5 mesh nodes"
UNIT = "evals/s"
K_SECTIONS, N_K = 33333, 4
STRONG_K = 83333                       # x 3 new nodes x 4 phases = 10^6 nodes (config 4)


def workload_config(n_gpus):
    return {"workload": "cart_pole_swing_up explicit, Lobatto, 33333 sections x 4 nodes "
                        "= 100000 collocation nodes, derivative_level=2 (G+H fused)",
            "num_x": 500001, "num_c": 399997, "nnz_G": 3899963, "nnz_H": 500000,
            "inputs": "x~U(-0.5,0.5), lam~N(0,1), sigma=1, numpy default_rng(seed)",
            "l2_policy": "ring of 6 device buffer sets (255 MB > 126 MB L2), "
                         "one set per step",
            "launch": "one fused kernel per evaluation; the K timed launches are enqueued "
                      "by one C call (pcx_eval_many) behind a stream gate, back to back on "
                      "one stream.  The ring's iterates are independent and declared so "
                      "(PCX_EVAL_INDEPENDENT): a kernel does not wait for the completion of "
                      "its predecessor (distinct buffers, rotating scratch), so consecutive "
                      "kernels overlap; results are bitwise those of ordered execution.  "
                      "`ordered` = the same launches in stream order, `latency_us` = one "
                      "synchronised evaluation (a sequential solver)",
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


# ------------------------------------------------------------------ CPU legs --
def _oracle_port(node_eval="numpy"):
    """The CPU restatement of the reference's algorithm on the bench workload
    (the one place besides tests/ and smoke() that may execute oracle/)."""
    from examples import problems
    from examples.cases import lower_case
    from oracle.blockwise import BlockwiseNLP
    ocp = problems.cart_pole_swing_up()
    low, meshes, scal = lower_case(ocp, "lobatto", K_SECTIONS, N_K, seed=0, unit_scaling=True)
    B = BlockwiseNLP(ocp, low.ir.full_bounds,
                     [dict(N=m.N, sI=m.sI_matrix, sA=m.sA_matrix, W=m.W_matrix) for m in meshes],
                     W_ocp=scal[2], w=scal[3], prune=low.S.prune,
                     scaling_method=ocp.settings.scaling_method, node_eval=node_eval)
    return low, B


def cpu_oracle_rate(node_eval="numpy", seconds_budget=10.0, max_evals=200):
    """Time the oracle port on the same workload; returns (evals/s, n, seconds)."""
    low, B = _oracle_port(node_eval)
    rng = np.random.default_rng(0)
    x = rng.uniform(-0.5, 0.5, low.S.num_x)
    lam = rng.standard_normal(low.S.num_c)
    B.G_nonzeros(x)
    B.H_nonzeros(x, 1.0, lam)                       # warm-up: merge plans, numba compile
    n, t0 = 0, time.perf_counter()
    while n < max_evals and (time.perf_counter() - t0) < seconds_budget:
        B.G_nonzeros(x)
        B.H_nonzeros(x, 1.0, lam)
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def _cpu_worker(seconds, q):
    """One process of the all-cores CPU throughput leg: the oracle port on the same
    workload, evaluations counted over a fixed wall-clock window."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    low, B = _oracle_port("numpy")
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


def cpu_baseline_block(single_budget=10.0):
    """The three CPU figures reported beside the GPU number."""
    rate, n, dt = cpu_oracle_rate("numpy", single_budget)
    cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
           "sample": f"{n} fused G+H evaluations of the full 10^5-node workload in "
                     f"{dt:.1f} s by oracle/blockwise.py (numpy, 1 thread)"}
    try:
        import numba
        r2, n2, dt2 = cpu_oracle_rate("numba", single_budget)
        cpu["numba"] = {"value": r2, "unit": UNIT, "cores": int(numba.get_num_threads()),
                        "how": f"same port, node functions in one numba.njit(parallel=True) "
                               f"prange kernel over the mesh nodes; {n2} evaluations in "
                               f"{dt2:.1f} s (sparse assembly stays single-threaded numpy)"}
    except Exception as exc:                          # never fail the bench for this leg
        cpu["numba"] = {"error": repr(exc)[:200]}
    try:
        cpu["all_cores"] = cpu_oracle_rate_all_cores()
    except Exception as exc:
        cpu["all_cores"] = {"error": repr(exc)[:200]}
    return cpu


def run_reference(args):
    """The reference arm: the CPU implementation of the path on the box's host cores,
    with all the host threads it can use.  Two ways exist to use them with the
    reference's single-threaded algorithm: inside one evaluation (node functions in a
    numba.prange kernel; the sparse assembly stays serial) or one independent
    evaluation stream per core (multi-start throughput).  Both are measured, each
    step a bounded sample; `value` is the better of the two, the other figures and
    the one-thread rate are reported beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warmup = args.steps, args.warmup
    kind_eval, cores = "numpy", 1
    try:
        import numba
        kind_eval, cores = "numba", int(numba.get_num_threads())
    except Exception:
        pass
    low, B = _oracle_port(kind_eval)
    rng = np.random.default_rng(0)
    xs = [rng.uniform(-0.5, 0.5, low.S.num_x) for _ in range(2)]
    lams = [rng.standard_normal(low.S.num_c) for _ in range(2)]
    for i in range(max(1, min(warmup, 3))):
        B.G_nonzeros(xs[i % 2])
        B.H_nonzeros(xs[i % 2], 1.0, lams[i % 2])
    steps = max(1, min(steps, 60))                  # bounded: ~0.1 s per eval
    t0 = time.perf_counter()
    for i in range(steps):
        B.G_nonzeros(xs[i % 2])
        B.H_nonzeros(xs[i % 2], 1.0, lams[i % 2])
    dt = time.perf_counter() - t0
    inside = steps / dt
    extra = {"inside_one_evaluation": {"value": inside, "unit": UNIT, "cores": cores,
                                       "how": f"node functions by {kind_eval} on {cores} thread(s), "
                                              f"{steps} evaluations in {dt:.1f} s"}}
    value, how, used = inside, f"{kind_eval} node kernel on {cores} thread(s) inside one evaluation", cores
    try:
        r1, n1, dt1 = cpu_oracle_rate("numpy", 8.0)
        extra["single_thread"] = {"value": r1, "unit": UNIT, "cores": 1,
                                  "sample": f"{n1} evaluations in {dt1:.1f} s"}
        allc = cpu_oracle_rate_all_cores()
        extra["all_cores"] = allc
        if allc["value"] > value:
            value, used = allc["value"], allc["cores"]
            how = (f"{allc['cores']} independent single-threaded evaluation streams, one per "
                   f"host core (the better of the two ways to use every core)")
    except Exception as exc:
        extra["all_cores"] = {"error": repr(exc)[:200]}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 / value, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": dict({"value": value, "unit": UNIT, "cores": used, "kind": "port",
                                  "sample": f"fused G+H evaluations of the full 10^5-node workload "
                                            f"by oracle/blockwise.py: {how}; the live reference "
                                            f"(CasADi) is not installable here"}, **extra),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------ GPU legs --
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
    sets = [dict(x=torch.rand(batch, S.num_x, dtype=torch.float64, device=dev, generator=g) - 0.5,
                 lam=torch.randn(batch, S.num_c, dtype=torch.float64, device=dev, generator=g),
                 jac=torch.empty(batch, S.nnz_g, dtype=torch.float64, device=dev),
                 hess=torch.empty(batch, S.nnz_h, dtype=torch.float64, device=dev))
            for _ in range(R)]
    args = eng.make_args(sets)
    eng.eval_many(what, args, 3, stream=stream, gate=False, timed=False)
    torch.cuda.synchronize()
    ms = eng.eval_many(what, args, steps, stream=stream, gate=True, timed=True)
    us = 1e3 * ms / steps / batch
    gbs = alg_bytes / us / 1e3
    return {"batch_per_launch": batch, "us_per_eval": us, "evals_per_s": 1e6 / us,
            "achieved_GBs": gbs, "frac": gbs / peak,
            "note": "same kernel, 8 independent iterates per launch (grid.y); informational"}


def config4_one_gpu(args, torch, E, dev, peak):
    """BASELINE config 4 on ONE GPU (informational, N = 1 only): Delta III, 4 phases,
    ~10^6 collocation nodes, fused G+H, device-resident, the launches enqueued by one
    C call behind a stream gate and timed by CUDA events on the launch stream.  Two
    output sets of 1.4 GB each alternate (>> L2)."""
    from examples import problems
    from examples.cases import lower_case
    K = args.strong_sections
    what = E.EVAL_JAC | E.EVAL_HESS
    low, _, scal = lower_case(problems.delta_iii_launch_vehicle(), "lobatto", K, 4, seed=0)
    S = low.S
    eng = E.Engine(S, low.layouts, low.header, device=dev.index or 0, structure=False)
    eng.set_scaling(*scal)
    g = torch.Generator(device=dev).manual_seed(0)
    x = 0.1 + 0.3 * torch.rand(S.num_x, dtype=torch.float64, device=dev, generator=g)
    lam = torch.randn(S.num_c, dtype=torch.float64, device=dev, generator=g)
    sets = [dict(x=x, lam=lam, jac=torch.empty(S.nnz_g, dtype=torch.float64, device=dev),
                 hess=torch.empty(S.nnz_h, dtype=torch.float64, device=dev)) for _ in range(2)]
    cargs = eng.make_args(sets)
    st = torch.cuda.current_stream().cuda_stream
    steps = 10
    eng.eval_many(what, cargs, 3, stream=st, gate=False, timed=False)
    torch.cuda.synchronize()
    ms = eng.eval_many(what, cargs, steps, stream=st, gate=True, timed=True) / steps
    alg = 8 * (S.num_x + S.nnz_g) + 8 * (S.num_x + S.num_c + S.nnz_h)
    info = eng.variant_info(what)
    out = {"workload": "delta_iii_launch_vehicle, 4 phases x %d sections x 4 nodes = %d collocation "
                       "nodes, one GPU" % (K, int(sum(t_.N for t_ in S.ph))),
           "num_x": int(S.num_x), "nnz_G": int(S.nnz_g), "nnz_H": int(S.nnz_h),
           "tiles": int(S.num_tiles), "threads": int(S.threads),
           "phases_sharing_one_body": [int(lay.leader) for lay in low.layouts],
           "ms_per_eval": ms, "evals_per_s": 1e3 / ms, "steps": steps,
           "algorithmic_bytes_per_launch": int(alg), "achieved_GBs": alg / ms / 1e6,
           "frac": alg / (ms * 1e-3) / 1e9 / peak, "kernel": info}
    del eng, sets, cargs, x, lam
    torch.cuda.empty_cache()
    return out


def strong_scaling(args, torch, dist, E, rank, world, local_rank):
    """BASELINE config 4 under the driver's own launch: ONE Delta III mesh (4 phases,
    ~10^6 nodes) split over the ranks by tile ranges, border values exchanged inside
    the kernel over NVLink peer memory (no NCCL call on the data path).  Asserts
    parity with the unsharded evaluation (each value slot has exactly one writer:
    the sum over ranks of the zero-initialised arrays must equal it) and times
    unsharded (every rank runs it; max over ranks) and sharded, max over ranks."""
    from examples import problems
    from examples.cases import lower_case
    from pycollo_b200.parallel import MeshSharder
    K = args.strong_sections
    dev = torch.device("cuda", local_rank)
    what = E.EVAL_JAC | E.EVAL_HESS
    ocp = problems.delta_iii_launch_vehicle()
    # the tiling is built for world x 148 SMs: the tile count is chosen for a rank's own
    # 148 SMs and its share of the mesh
    low, _, scal = lower_case(ocp, "lobatto", K, 4, seed=0, sm_count=148 * world)
    S = low.S
    eng = E.Engine(S, low.layouts, low.header, device=local_rank, structure=False)
    eng.set_scaling(*scal)
    g = torch.Generator(device=dev).manual_seed(0)              # same x / lam on every rank
    x = 0.1 + 0.3 * torch.rand(S.num_x, dtype=torch.float64, device=dev, generator=g)
    lam = torch.randn(S.num_c, dtype=torch.float64, device=dev, generator=g)
    R = 2
    jac = [torch.zeros(S.nnz_g, dtype=torch.float64, device=dev) for _ in range(R)]
    hes = [torch.zeros(S.nnz_h, dtype=torch.float64, device=dev) for _ in range(R)]
    st = torch.cuda.current_stream().cuda_stream
    steps = max(5, min(args.steps, 30))
    sets = [dict(x=x, lam=lam, jac=jac[k], hess=hes[k]) for k in range(R)]
    cargs = eng.make_args(sets)
    # ---- unsharded on this rank (the N = 1 time of the same engine) ------------
    eng.eval_many(what, cargs, 3, stream=st, gate=False, timed=False)
    torch.cuda.synchronize()
    ms1 = eng.eval_many(what, cargs, steps, stream=st, gate=True, timed=True) / steps
    ref_j, ref_h = jac[(steps - 1) % R].clone(), hes[(steps - 1) % R].clone()
    # the tiling a single GPU would choose for itself (rank 0 only; the others wait)
    ms1_own = 0.0
    if rank == 0 and world > 1:
        low1, _, _ = lower_case(problems.delta_iii_launch_vehicle(), "lobatto", K, 4, seed=0)
        eng1 = E.Engine(low1.S, low1.layouts, low1.header, device=local_rank, structure=False)
        eng1.set_scaling(*scal)
        a1 = eng1.make_args(sets)
        eng1.eval_many(what, a1, 3, stream=st, gate=False, timed=False)
        torch.cuda.synchronize()
        ms1_own = eng1.eval_many(what, a1, steps, stream=st, gate=True, timed=True) / steps
        tiles_own = int(low1.S.num_tiles)
        del eng1, a1, low1
    dist.barrier()
    # ---- sharded, fused exchange ----------------------------------------------------
    sh = MeshSharder(eng, world, rank, border_rank=0, fused=True)
    jac[0].zero_()
    hes[0].zero_()
    torch.cuda.synchronize()
    dist.barrier()
    sh.evaluate(what, x, lam=lam, jac=jac[0], hess=hes[0])
    dist.all_reduce(jac[0])
    dist.all_reduce(hes[0])                                     # check only: sum of disjoint slabs
    torch.cuda.synchronize()
    ej = float((jac[0] - ref_j).abs().max() / ref_j.abs().max())
    eh = float((hes[0] - ref_h).abs().max() / ref_h.abs().max())
    assert ej <= 1e-13 and eh <= 1e-13, f"sharded != unsharded: jac {ej:.2e} hess {eh:.2e}"
    for i in range(3):
        sh.evaluate(what, x, lam=lam, jac=jac[i % R], hess=hes[i % R])
    torch.cuda.synchronize()
    dist.barrier()
    # 6 untimed launches first, in the same call: their exchange brings the ranks into
    # step, so the timed ones do not contain the ranks' start-up skew (which, over a
    # ~2 ms timed region, would otherwise be charged to the border rank)
    msN = eng.eval_many(what, cargs, steps, stream=st, gate=True, timed=True, warm=6) / steps
    assert eng.status() == 0, "fused exchange timed out"
    t = torch.tensor([ms1, msN, ms1_own], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms1, msN, ms1_own = float(t[0]), float(t[1]), float(t[2])
    best1 = min(ms1, ms1_own) if ms1_own > 0 else ms1
    alg = 8 * (S.num_x + S.nnz_g) + 8 * (S.num_x + S.num_c + S.nnz_h)
    lo, hi = sh.lo, sh.hi
    del jac, hes, ref_j, ref_h
    torch.cuda.empty_cache()
    return {"workload": "delta_iii_launch_vehicle, 4 phases x %d sections x 4 nodes = %d "
                        "collocation nodes, ONE mesh over %d GPUs (tile ranges)"
                        % (K, int(sum(t_.N for t_ in S.ph)), world),
            "num_x": int(S.num_x), "nnz_G": int(S.nnz_g), "nnz_H": int(S.nnz_h),
            "tiles": int(S.num_tiles), "tiles_rank0": int(hi - lo),
            "scaling": "strong", "ms_per_eval_1gpu": best1, "ms_per_eval": msN,
            "speedup_vs_1gpu": best1 / msN, "evals_per_s": 1e3 / msN,
            "ms_per_eval_1gpu_detail": {"tiling_for_N_gpus": ms1, "tiling_for_1_gpu": ms1_own,
                                        "note": "speedup is quoted against the faster of the two "
                                                "unsharded timings"},
            "algorithmic_GBs": alg / msN / 1e6,
            "comm": "fused in-kernel exchange over NVLink peer memory (CUDA IPC mapping of "
                    "the border rank's buffer; st.release.sys / ld.acquire.sys epoch flags, "
                    "double-buffered shares); no NCCL call on the data path",
            "nvlink_bytes_per_eval": int(8 * (S.bv_size + 1) * (world - 1)),
            "parity": {"vs": "unsharded evaluation on the same engine", "rel_err_jac": ej,
                       "rel_err_hess": eh}, "steps": steps}


def run_cuda(args):
    import torch
    import torch.distributed as dist
    from examples import problems
    from examples.cases import lower_case
    from pycollo_b200 import engine as E

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    steps, warmup = args.steps, args.warmup
    if warmup < 3:
        print(f"bench.py: --warmup {warmup} raised to 3 (timing rule: W >= 3)", file=sys.stderr)
        warmup = 3

    low, _, scal = lower_case(problems.cart_pole_swing_up(), "lobatto", K_SECTIONS, N_K,
                              seed=0, unit_scaling=True)
    S = low.S
    eng = E.Engine(S, low.layouts, low.header, batch=1, device=local_rank)
    eng.set_scaling(*scal)
    what = E.EVAL_JAC | E.EVAL_HESS
    alg_bytes = 8 * (S.num_x + S.nnz_g) + 8 * (S.num_x + S.num_c + S.nnz_h)

    # ---- device-resident ring (> L2) ------------------------------------
    R = 6
    rng = np.random.default_rng(1000 + rank)
    dev = torch.device("cuda", local_rank)
    sets = [dict(x=torch.from_numpy(rng.uniform(-0.5, 0.5, S.num_x)).to(dev),
                 lam=torch.from_numpy(rng.standard_normal(S.num_c)).to(dev),
                 jac=torch.empty(S.nnz_g, dtype=torch.float64, device=dev),
                 hess=torch.empty(S.nnz_h, dtype=torch.float64, device=dev)) for _ in range(R)]
    stream = torch.cuda.current_stream().cuda_stream
    cargs = eng.make_args(sets)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the ring holds INDEPENDENT iterates (a sweep / multi-start stream): declared as
    # such, consecutive kernels do not wait for each other's completion and their
    # load / compute / store phases interleave.  The same K steps are also timed in
    # stream order (`ordered`: every kernel waits for its predecessor).
    sweep = what | E.EVAL_INDEPENDENT
    eng.eval_many(sweep, cargs, warmup, stream=stream, gate=False, timed=False)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    l0 = eng.launch_count
    barrier()
    ms = eng.eval_many(sweep, cargs, steps, stream=stream, gate=True, timed=True)
    barrier()
    launches = eng.launch_count - l0
    eng.eval_many(what, cargs, warmup, stream=stream, gate=False, timed=False)
    barrier()
    ms_ordered = eng.eval_many(what, cargs, steps, stream=stream, gate=True, timed=True)
    barrier()
    launches += steps + warmup

    # ---- single-evaluation latency (sequential host) --------------------------
    lat = []
    one = [eng.make_args([s]) for s in sets]
    for i in range(3 + 24):
        torch.cuda.synchronize()
        t_ms = eng.eval_many(what, one[i % R], 1, stream=stream, gate=False, timed=True)
        if i >= 3:
            lat.append(1e3 * t_ms)
    launches += 27

    # ---- end to end through the C ABI with pinned host buffers ------------
    hx = torch.from_numpy(rng.uniform(-0.5, 0.5, S.num_x)).pin_memory()
    hl = torch.from_numpy(rng.standard_normal(S.num_c)).pin_memory()
    hs = torch.ones(1, dtype=torch.float64).pin_memory()
    hj = torch.empty(S.nnz_g, dtype=torch.float64).pin_memory()
    hh = torch.empty(S.nnz_h, dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(steps, 50))
    # the host keeps its value arrays between evaluations (a solver's buffers), so the
    # Jacobian slots that do not depend on the iterate are fetched once, not every step
    # (PCX_EVAL_CONST_RESIDENT; the bytes actually copied are reported)
    call = eng.bind(what | E.EVAL_CONST_RESIDENT, hx, lam=hl, sigma=hs, jac=hj, hess=hh,
                    space=E.PCX_HOST, stream=stream)
    for _ in range(3):
        call()
    barrier()
    l1 = eng.launch_count
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        call()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    launches += eng.launch_count - l1
    e2e_d2h = eng.last_d2h_bytes
    # the same without the promise: every value crosses PCIe every step
    call_full = eng.bind(what, hx, lam=hl, sigma=hs, jac=hj, hess=hh, space=E.PCX_HOST, stream=stream)
    for _ in range(2):
        call_full()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        call_full()
    torch.cuda.synchronize()
    e2e_full_s = time.perf_counter() - t0
    launches += 2 * (e2e_steps + 2)
    # a sweep over iterates with a host consumer (multi-start): the same host-space
    # evaluations pipelined by pcx_sweep_host -- the upload of step i+1 and the kernel
    # of step i run under the download of step i-1 (PCIe is full duplex).  Every step
    # still uploads its inputs from pinned memory and downloads its values.
    hsets = [dict(x=hx, lam=hl, sigma=hs, jac=hj, hess=hh),
             dict(x=torch.from_numpy(rng.uniform(-0.5, 0.5, S.num_x)).pin_memory(),
                  lam=torch.from_numpy(rng.standard_normal(S.num_c)).pin_memory(), sigma=hs,
                  jac=torch.empty(S.nnz_g, dtype=torch.float64).pin_memory(),
                  hess=torch.empty(S.nnz_h, dtype=torch.float64).pin_memory())]
    hargs = eng.make_args(hsets)
    eng.sweep_host(what | E.EVAL_CONST_RESIDENT, hargs, 4, stream=stream)
    barrier()
    t0 = time.perf_counter()
    eng.sweep_host(what | E.EVAL_CONST_RESIDENT, hargs, e2e_steps, stream=stream)
    torch.cuda.synchronize()
    e2e_sweep_s = time.perf_counter() - t0
    launches += e2e_steps + 4
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms, 1e3 * e2e_s, float(np.median(lat)), ms_ordered, 1e3 * e2e_full_s,
                      1e3 * e2e_sweep_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max, lat_us, ms_ord, e2e_full_ms, e2e_sweep_ms = (float(v) for v in t)

    strong = None
    if world > 1 and not args.no_strong:
        del sets, cargs, one, hj, hh
        torch.cuda.empty_cache()
        try:
            strong = strong_scaling(args, torch, dist, E, rank, world, local_rank)
        except Exception as exc:                      # reported, never hides the headline
            strong = {"error": repr(exc)[:300]}

    if rank == 0:
        value = world * steps / (ms_max * 1e-3)
        e2e_seq_value = world * e2e_steps / (e2e_ms_max * 1e-3)
        e2e_value = world * e2e_steps / (e2e_sweep_ms * 1e-3)
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
        cpu = None
        config4 = None
        if world == 1:
            amortised = batched_rate(low, scal, E, torch, dev, stream, alg_bytes, peak)
            if not args.no_config4:
                try:
                    config4 = config4_one_gpu(args, torch, E, dev, peak)
                    launches += 13
                except Exception as exc:              # informational: never hides the headline
                    config4 = {"error": repr(exc)[:300]}
            if not args.no_cpu_baseline:
                cpu = cpu_baseline_block()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": ms_max / steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": workload_config(world),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": alg_bytes,
                             "kernel": "pcx_fill_12 (fused Jacobian+Hessian fill)"},
                "ordered": {"value": world * steps / (ms_ord * 1e-3), "unit": UNIT,
                            "us_per_eval": 1e3 * ms_ord / steps,
                            "frac": alg_bytes / (ms_ord * 1e-3 / steps) / 1e9 / peak,
                            "how": "the same K launches in stream order: every kernel waits for "
                                   "the completion of its predecessor before touching its data "
                                   "(programmatic dependent launch still overlaps the table-only "
                                   "prologue)"},
                "latency_us": {"value": lat_us, "frac": alg_bytes / lat_us / 1e3 / peak,
                               "how": "one evaluation per synchronised call, cold ring slot, "
                                      "CUDA events around the single launch; median of 24"},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": UNIT,
                        "h2d_bytes_per_step": 8 * (S.num_x + S.num_c + 1),
                        "d2h_bytes_per_step": int(e2e_d2h),
                        "steps": e2e_steps,
                        "api": "pcx_sweep_host(JAC|HESS|CONST_RESIDENT) over pinned host argument "
                               "sets: a stream of independent iterates with a host consumer, "
                               "pipelined (upload of step i+1 and kernel of step i under the "
                               "download of step i-1).  The host's value arrays persist between "
                               "evaluations, so the iterate-independent Jacobian slots (whole "
                               "variable blocks; listed by the structure builder) are fetched by "
                               "the first evaluation into an array only",
                        "sequential": {
                            "value": e2e_seq_value, "unit": UNIT,
                            "api": "one blocking pcx_eval(JAC|HESS|CONST_RESIDENT, PCX_HOST) per "
                                   "step: what a sequential solver sees"},
                        "all_values_every_step": {
                            "value": world * e2e_steps / (e2e_full_ms * 1e-3), "unit": UNIT,
                            "d2h_bytes_per_step": 8 * (S.nnz_g + S.nnz_h)}},
                "gpu_launches": int(launches), "clocks": clocks,
                "amortised": amortised}
        if strong is not None:
            line["strong"] = strong
        if config4 is not None:
            line["config4_1gpu"] = config4
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
    ap.add_argument("--no-config4", action="store_true",
                    help="skip the informational Delta III 10^6-node leg of a one-GPU run")
    ap.add_argument("--no-strong", action="store_true",
                    help="skip the mesh-sharded (config 4) leg of a multi-GPU run")
    ap.add_argument("--strong-sections", type=int, default=STRONG_K,
                    help="sections per phase of the sharded Delta III mesh")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
