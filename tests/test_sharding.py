"""One mesh over several ranks (BASELINE config 4).

CPU: world_size-2 gloo processes walk disjoint tile ranges of the device tables
with the kernel emulator, all_reduce the border-exchange buffer and apply the
border stage; the reassembled vectors equal the unsharded evaluation and the
oracle, and every value slot has exactly one writer.
GPU: the same two stages through the C ABI (pcx_set_shard / pcx_apply_border)
with two engines on one device standing in for two ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import build_case, max_err
from pycollo_b200 import engine as E
from examples import problems as examples
from pycollo_b200.parallel import shard_range

KEYS = ("c", "dy", "jac", "hess", "grad")


def _case():
    # free final time + integral + 3 phases with linkages: every kind of
    # reduction and border entry is present; small tiles -> several per phase
    return build_case(examples.multiphase_sliding_mass(), "lobatto", 6, [4, 3, 5, 4, 6, 3],
                      None, seed=3, max_tile_nodes=8)


def _inputs(S):
    rng = np.random.default_rng(11)
    return rng.uniform(-0.5, 0.5, S.num_x), rng.standard_normal(S.num_c)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from emulator import emulate
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        low, _, scal = _case()
        S = low.S
        tabs = E.build_tables(S, low.layouts)
        st = E.scaling_tables(S, low.layouts, *scal)
        x, lam = _inputs(S)
        lo, hi = shard_range(S.num_tiles, world, rank)
        out = emulate(low, tabs, st, x, lam, 0.7, flags=63, tile_range=(lo, hi), stage=1)
        xb = torch.from_numpy(out.pop("xbuf").copy())
        dist.all_reduce(xb)                                   # the only data-path collective
        fin = emulate(low, tabs, st, x, lam, 0.7, flags=63, stage=2, xbuf=xb.numpy())
        res = {}
        for k in KEYS:
            mine = out[k]
            if rank == 0:                                     # border_rank = 0
                mine = np.where(np.isnan(fin[k]), mine, fin[k])
            wrote = torch.from_numpy((~np.isnan(mine)).astype(np.float64))
            vals = torch.from_numpy(np.nan_to_num(mine, nan=0.0))
            dist.all_reduce(wrote)
            dist.all_reduce(vals)                             # sum of disjoint slabs
            res[k] = (vals.numpy(), wrote.numpy())
        if rank == 0:
            q.put((res, float(fin["f"])))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_mesh_sharding_matches_unsharded():
    from emulator import emulate
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res, f = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    low, B, scal = _case()
    S = low.S
    assert S.num_tiles >= 6
    x, lam = _inputs(S)
    full = emulate(low, E.build_tables(S, low.layouts), E.scaling_tables(S, low.layouts, *scal),
                   x, lam, 0.7, flags=63)
    ref = dict(c=B.c(x), dy=B.dy(x), jac=B.G_nonzeros(x), hess=B.H_nonzeros(x, 0.7, lam), grad=B.g(x))
    for k in KEYS:
        vals, wrote = res[k]
        if k == "grad":
            # tiles zero the dense gradient, the border stage then overwrites its
            # structural entries: a zero and a value may come from different ranks
            assert np.all(wrote >= 1.0)
        else:
            assert np.all(wrote == 1.0), (k, np.flatnonzero(wrote != 1.0)[:5])   # one writer per slot
        assert max_err(vals, full[k]) <= 1e-13, k
        assert max_err(vals, ref[k]) <= 1e-12, k
    assert abs(f - B.J(x)) <= 1e-12 * max(1.0, abs(B.J(x)))


@pytest.mark.gpu
@pytest.mark.parametrize("problem,K,nodes,kw", [
    ("multiphase_sliding_mass", 6, [4, 3, 5, 4, 6, 3], dict(max_tile_nodes=8)),
    ("delta_iii_launch_vehicle", 40, 4, dict(max_tile_nodes=24)),
    ("cart_pole_swing_up", 3000, 4, {}),
])
def test_gpu_two_stage_sharded_evaluation(problem, K, nodes, kw):
    low, B, scal = build_case(getattr(examples, problem)(), "lobatto", K, nodes, None,
                              seed=3, oracle=(K <= 40), **kw)
    S = low.S
    what = E.EVAL_C | E.EVAL_DY | E.EVAL_JAC | E.EVAL_HESS | E.EVAL_F | E.EVAL_GRAD
    rng = np.random.default_rng(2)
    x = torch.from_numpy(rng.uniform(-0.5, 0.5, S.num_x)).cuda()
    lam = torch.from_numpy(rng.standard_normal(S.num_c)).cuda()
    sig = torch.tensor([0.7], dtype=torch.float64, device="cuda")

    def bufs():
        z = lambda n: torch.zeros(n, dtype=torch.float64, device="cuda")
        return dict(f=z(1), grad=z(S.num_x), c=z(S.num_c), dy=z(S.num_dy), jac=z(S.nnz_g), hess=z(S.nnz_h))

    whole = E.Engine(S, low.layouts, low.header)
    whole.set_scaling(*scal)
    ref = bufs()
    whole.eval_ptr(what, x, lam=lam, sigma=sig, **ref)
    torch.cuda.synchronize()

    world = 3
    out = bufs()                                              # ONE set of full-size arrays
    engines, xbufs = [], []
    for r in range(world):
        eng = E.Engine(S, low.layouts, low.header)
        eng.set_scaling(*scal)
        eng.set_shard(*shard_range(S.num_tiles, world, r))
        ptr, n = eng.shard_buffer()
        from pycollo_b200.parallel import _DeviceArray
        xbufs.append(torch.as_tensor(_DeviceArray(ptr, n), device="cuda"))
        engines.append(eng)
        eng.eval_ptr(what, x, lam=lam, sigma=sig, **out)      # stage 1: disjoint slabs
    torch.cuda.synchronize()
    total = torch.stack(xbufs).sum(0)                         # what NCCL all_reduce would do
    xbufs[0].copy_(total)
    o2 = {k: v for k, v in out.items() if k != "dy"}
    engines[0].apply_border(what, x, lam=lam, sigma=sig, **o2)   # stage 2 on the border rank
    torch.cuda.synchronize()
    for k in ("c", "dy", "jac", "hess", "grad", "f"):
        a, b = out[k].cpu().numpy(), ref[k].cpu().numpy()
        assert max_err(a, b) <= 1e-13, k
    if B is not None:
        xn, ln = x.cpu().numpy(), lam.cpu().numpy()
        assert max_err(out["jac"].cpu().numpy(), B.G_nonzeros(xn)) <= 1e-12
        assert max_err(out["hess"].cpu().numpy(), B.H_nonzeros(xn, 0.7, ln)) <= 1e-12
        assert max_err(out["c"].cpu().numpy(), B.c(xn)) <= 1e-12


@pytest.mark.gpu
def test_gpu_fused_peer_exchange_same_device():
    """border_mode 3 (exchange fused into the kernel over peer memory) with three
    "ranks" on one device sharing one exchange buffer: the non-border ranks run
    first on the stream, the border rank last (it waits for their shares)."""
    low, B, scal = build_case(examples.multiphase_sliding_mass(), "lobatto", 6, [4, 3, 5, 4, 6, 3],
                              None, seed=3, max_tile_nodes=8)
    S = low.S
    what = E.EVAL_C | E.EVAL_JAC | E.EVAL_HESS | E.EVAL_F | E.EVAL_GRAD
    rng = np.random.default_rng(4)
    world = 3
    engines = []
    for r in range(world):
        eng = E.Engine(S, low.layouts, low.header)
        eng.set_scaling(*scal)
        eng.set_shard(*shard_range(S.num_tiles, world, r))
        engines.append(eng)
    engines[0].exchange_alloc(world)
    base = engines[0].exchange_buffer()
    for r, eng in enumerate(engines):
        eng.exchange_attach_ptr(r, world, 0, base)
    z = lambda n: torch.zeros(n, dtype=torch.float64, device="cuda")
    for trial in range(6):                                   # epochs pair the calls
        xn, ln = rng.uniform(-0.5, 0.5, S.num_x), rng.standard_normal(S.num_c)
        x, lam = torch.from_numpy(xn).cuda(), torch.from_numpy(ln).cuda()
        sig = torch.tensor([0.7], dtype=torch.float64, device="cuda")
        out = dict(f=z(1), grad=z(S.num_x), c=z(S.num_c), jac=z(S.nnz_g), hess=z(S.nnz_h))
        for r in (2, 1, 0):
            engines[r].eval_ptr(what, x, lam=lam, sigma=sig, **out)
        torch.cuda.synchronize()
        assert max_err(out["jac"].cpu().numpy(), B.G_nonzeros(xn)) <= 1e-12
        assert max_err(out["hess"].cpu().numpy(), B.H_nonzeros(xn, 0.7, ln)) <= 1e-12
        assert max_err(out["c"].cpu().numpy(), B.c(xn)) <= 1e-12
        assert max_err(out["grad"].cpu().numpy(), B.g(xn)) <= 1e-12
        assert max_err(out["f"].cpu().numpy(), [B.J(xn)]) <= 1e-12


@pytest.mark.gpu
def test_gpu_fused_exchange_across_processes():
    """The cross-process path (CUDA IPC + st.release.sys / ld.acquire.sys + slot
    back-pressure): two processes, one GPU each, 40 evaluations enqueued without
    host synchronisation while the border rank starts late.  Needs 2 GPUs."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    script = os.path.join(os.path.dirname(__file__), "dist_fused_exchange.py")
    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", "29533", script],
        capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "fused exchange across 2 processes" in res.stdout
