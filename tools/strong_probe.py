"""Where does the time of a mesh-sharded evaluation go?  Per-rank device times of
(a) this rank's tiles only (stage 1, no exchange), (b) the fused exchange, for the
Delta III mesh of bench.py's strong leg.
  torchrun --nproc-per-node N tools/strong_probe.py [K]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from examples import problems
from examples.cases import lower_case
from pycollo_b200 import engine as E
from pycollo_b200.parallel import MeshSharder, shard_range

K = int(sys.argv[1]) if len(sys.argv) > 1 else 83333
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
what = E.EVAL_JAC | E.EVAL_HESS
low, _, scal = lower_case(problems.delta_iii_launch_vehicle(), "lobatto", K, 4, seed=0, sm_count=148 * world)
S = low.S
g = torch.Generator(device=dev).manual_seed(0)
x = 0.1 + 0.3 * torch.rand(S.num_x, dtype=torch.float64, device=dev, generator=g)
lam = torch.randn(S.num_c, dtype=torch.float64, device=dev, generator=g)
sets = [dict(x=x, lam=lam, jac=torch.zeros(S.nnz_g, dtype=torch.float64, device=dev),
             hess=torch.zeros(S.nnz_h, dtype=torch.float64, device=dev)) for _ in range(2)]
st = torch.cuda.current_stream().cuda_stream
out = {}
# (a) stage 1 only
eng = E.Engine(S, low.layouts, low.header, device=local, structure=False)
eng.set_scaling(*scal)
args = eng.make_args(sets)
eng.set_shard(*shard_range(S.num_tiles, world, rank))
eng.eval_many(what, args, 5, stream=st, gate=False, timed=False)
torch.cuda.synchronize(); dist.barrier()
out["stage1_us"] = 1e3 * eng.eval_many(what, args, 40, stream=st, gate=True, timed=True) / 40
# is it the GPU or the tile range?  every rank times every rank's range, alone and together
out["each_range_us"] = []
for r in range(world):
    eng.set_shard(*shard_range(S.num_tiles, world, r))
    eng.eval_many(what, args, 3, stream=st, gate=False, timed=False)
    torch.cuda.synchronize(); dist.barrier()
    out["each_range_us"].append(round(1e3 * eng.eval_many(what, args, 20, stream=st, gate=True, timed=True) / 20, 1))
del eng
# (b) fused exchange, 20 / 200 timed launches after 6 / 20 warm ones
eng = E.Engine(S, low.layouts, low.header, device=local, structure=False)
eng.set_scaling(*scal)
args = eng.make_args(sets)
sh = MeshSharder(eng, world, rank, border_rank=0, fused=True)
for i in range(3):
    sh.evaluate(what, x, lam=lam, jac=sets[0]["jac"], hess=sets[0]["hess"])
torch.cuda.synchronize(); dist.barrier()
out["fused20_us"] = 1e3 * eng.eval_many(what, args, 20, stream=st, gate=True, timed=True, warm=6) / 20
torch.cuda.synchronize(); dist.barrier()
out["fused200_us"] = 1e3 * eng.eval_many(what, args, 200, stream=st, gate=True, timed=True, warm=20) / 200
torch.cuda.synchronize(); dist.barrier()
out["fused200_nogate_us"] = 1e3 * eng.eval_many(what, args, 200, stream=st, gate=False, timed=True, warm=20) / 200
allr = [None] * world
dist.all_gather_object(allr, {k: (round(v, 1) if isinstance(v, float) else v) for k, v in out.items()})
if rank == 0:
    print(json.dumps(dict(world=world, tiles=int(S.num_tiles), per_rank=allr)), flush=True)
dist.destroy_process_group()
