"""Config 4: one Delta III mesh (4 phases) sharded over the ranks by tile ranges.

torchrun --nproc-per-node N tools/shard_bench.py [sections_per_phase] [steps]
Checks the sharded evaluation (stage 1 + NCCL all_reduce of the border buffer +
stage 2) against the unsharded one on rank 0's device, then times it: per eval
every rank computes its slab of G and H (device-resident, ring of buffers > L2),
max over ranks."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import torch.distributed as dist
from helpers import build_case
from pycollo_b200 import engine as E
from examples import problems as examples
from pycollo_b200.parallel import MeshSharder

K = int(sys.argv[1]) if len(sys.argv) > 1 else 83333          # x3 nodes x4 phases ~ 10^6 nodes
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
fused = len(sys.argv) > 3 and sys.argv[3] == "fused"
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
low, _, scal = build_case(examples.delta_iii_launch_vehicle(), "lobatto", K, 4, seed=0, oracle=False)
S = low.S
what = E.EVAL_JAC | E.EVAL_HESS
g = torch.Generator(device="cuda").manual_seed(0)              # same x/lam on every rank
x = 0.1 + 0.3 * torch.rand(S.num_x, dtype=torch.float64, device="cuda", generator=g)   # away from r = 0
lam = torch.randn(S.num_c, dtype=torch.float64, device="cuda", generator=g)
eng = E.Engine(S, low.layouts, low.header, device=local)
eng.set_scaling(*scal)
R = 2
jac = [torch.zeros(S.nnz_g, dtype=torch.float64, device="cuda") for _ in range(R)]
hes = [torch.zeros(S.nnz_h, dtype=torch.float64, device="cuda") for _ in range(R)]
st = torch.cuda.current_stream().cuda_stream
# reference: unsharded evaluation on this rank
eng.eval_ptr(what, x, lam=lam, jac=jac[1], hess=hes[1], stream=st)
torch.cuda.synchronize()
ref_j, ref_h = jac[1].clone(), hes[1].clone()
sh = MeshSharder(eng, world, rank, border_rank=0, fused=fused)
jac[0].zero_(); hes[0].zero_()
sh.evaluate(what, x, lam=lam, jac=jac[0], hess=hes[0])
if world > 1:
    dist.all_reduce(jac[0]); dist.all_reduce(hes[0])           # test-only gather: sum of disjoint slabs
torch.cuda.synchronize()
ej = float((jac[0] - ref_j).abs().max() / ref_j.abs().max())
eh = float((hes[0] - ref_h).abs().max() / ref_h.abs().max())
# timing: stage 1 + all_reduce + stage 2, no gather
for i in range(3):
    sh.evaluate(what, x, lam=lam, jac=jac[i % R], hess=hes[i % R])
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    sh.evaluate(what, x, lam=lam, jac=jac[i % R], hess=hes[i % R])
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    alg = 8 * (S.num_x + S.nnz_g) + 8 * (S.num_x + S.num_c + S.nnz_h)
    print(json.dumps(dict(workload="delta_iii 4 phases", exchange="fused peer-memory" if fused and world > 1 else "nccl all_reduce + 2nd launch", nodes=int(sum(t.N for t in S.ph)), n_gpus=world,
                          num_x=S.num_x, nnz_G=S.nnz_g, nnz_H=S.nnz_h, tiles=S.num_tiles,
                          rel_err_jac=ej, rel_err_hess=eh, ms_per_eval=round(float(ms), 4),
                          evals_per_s=round(1e3 / float(ms), 1),
                          algorithmic_GBs=round(alg / float(ms) / 1e6, 1))), flush=True)
if world > 1:
    dist.destroy_process_group()
