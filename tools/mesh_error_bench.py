"""Config 3: mesh-refinement error re-evaluation on the device.

x_tilde -> dy (PCX_EVAL_DY) -> x_ph (pcx_refit_to_ph) -> errors (pcx_mesh_error), all
device-resident, against the CPU chain of oracle/mesh_error.py (the reference's
own algorithm: numpy polynomial fits per state and section + Python loops)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from pycollo_b200 import engine as E
from examples import problems as examples
from pycollo_b200.backend import lower_problem
from pycollo_b200.mesh import Mesh, PhaseMesh
from pycollo_b200.mesh_refinement import MeshErrorEvaluator, create_ph_mesh
from pycollo_b200.quadrature import Quadrature


def run(problem, K, nodes, cpu=True, steps=30):
    ocp = getattr(examples, problem)()
    ocp.settings.scaling_method = "none"
    mesh = Mesh(Quadrature("lobatto"), [PhaseMesh(K, None, nodes) for _ in ocp.phases], 2, 16)
    low = lower_problem(ocp, mesh.p)
    S = low.S
    base = E.Engine(S, low.layouts, low.header)
    base.set_scaling(np.ones(S.n_var_ocp), np.zeros(S.n_var_ocp), np.ones(S.n_con_ocp), 1.0)
    ev = MeshErrorEvaluator(ocp, mesh)
    ne, ns = ev.engine.mesh_error_sizes()
    g = torch.Generator(device="cuda").manual_seed(0)
    x = 0.2 + 0.4 * torch.rand(S.num_x, dtype=torch.float64, device="cuda", generator=g)
    dy = torch.empty(S.num_dy, dtype=torch.float64, device="cuda")
    xph = torch.empty(base.refit_size(), dtype=torch.float64, device="cuda")
    ab = torch.empty(ne, dtype=torch.float64, device="cuda")
    re = torch.empty(ne, dtype=torch.float64, device="cuda")
    mx = torch.empty(ns, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def step():
        base.eval_ptr(E.EVAL_DY, x, dy=dy, stream=st)
        base.refit_to_ph_ptr(x, dy, xph, stream=st)
        ev.engine.mesh_error_ptr(xph, ab, re, mx, stream=st)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / steps
    out = dict(problem=problem, sections=K, nodes=int(sum(m.N for m in mesh.p)),
               ph_nodes=int(sum(m.N for m in ev.ph_mesh.p)), gpu_us=round(us, 1))
    if cpu:
        from helpers import cpu_mesh_error_chain          # tests/: the only home of oracle imports
        xh, dyh = x.cpu().numpy(), dy.cpu().numpy()
        t0 = time.perf_counter()
        worst = cpu_mesh_error_chain(ocp, low, ev.low, mesh, ev.ph_mesh, xh, dyh)
        out["cpu_ms"] = round(1e3 * (time.perf_counter() - t0), 1)
        out["speedup"] = round(out["cpu_ms"] * 1e3 / us)
        out["max_rel_err_gpu_vs_cpu"] = float(abs(worst - float(mx.max())))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    run("free_flying_robot", 200, 5)
    run("space_shuttle_reentry", 200, 5)
    run("free_flying_robot", 20000, 5, cpu=False)
    run("space_shuttle_reentry", 20000, 5, cpu=False)
