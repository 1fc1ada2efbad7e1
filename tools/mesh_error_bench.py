"""Config 3: mesh-refinement error re-evaluation on the device.

x_tilde -> dy (PCX_EVAL_DY) -> x_ph (pcx_refit_to_ph) -> errors (pcx_mesh_error), all
device-resident, against the CPU chain of oracle/mesh_error.py (the reference's
own algorithm: numpy polynomial fits per state and section + Python loops)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from pycollo_b200 import engine as E, examples
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
        from oracle import mesh_error as OM
        from oracle.blockwise import BlockwiseNLP
        xh, dyh = x.cpu().numpy(), dy.cpu().numpy()
        ph = ev.ph_mesh
        B = BlockwiseNLP(ocp, ev.low.ir.full_bounds,
                         [dict(N=m.N, sI=m.sI_matrix, sA=m.sA_matrix, W=m.W_matrix) for m in ph.p],
                         scaling_method="none")
        sI = [m.sI_matrix for m in ph.p]
        t0 = time.perf_counter()
        xs = []
        for ip, (irp, t) in enumerate(zip(low.ir.phases, S.ph)):
            ny, nu, N = irp.n_y, irp.n_u, t.N
            y = xh[t.x_off:t.x_off + ny * N].reshape(ny, N)
            u = xh[t.x_off + ny * N:t.x_off + (ny + nu) * N].reshape(nu, N)
            d = dyh[t.dy_off:t.dy_off + ny * N].reshape(ny, N)
            tv = xh[t.q_col + irp.n_q:t.q_col + irp.n_q + irp.n_t]
            T = (tv[-1] if irp.t_needed[1] else float(irp.tF)) - (tv[0] if irp.t_needed[0] else float(irp.t0))
            bnd, bph = mesh.mesh_index_boundaries[ip], ph.mesh_index_boundaries[ip]
            yp, up = OM.fit_section_polys(mesh.tau[ip], y, d, u, T, bnd, mesh.N_K[ip])
            y_ph = OM.interpolate_to_ph(y, yp, bnd, bph, ph.tau[ip])
            u_ph = OM.interpolate_to_ph(u, up, bnd, bph, ph.tau[ip])
            xs += [y_ph.ravel(), u_ph.ravel(), xh[t.q_col:t.q_col + irp.n_q + irp.n_t]]
        xs.append(xh[S.s_off:])
        x_ph = np.concatenate(xs)
        dyp = B.dy(x_ph)
        o = 0
        worst = 0.0
        for ip, (irp, t) in enumerate(zip(low.ir.phases, ev.low.S.ph)):
            ny, Nph = irp.n_y, t.N
            y_ph = x_ph[t.x_off:t.x_off + ny * Nph].reshape(ny, Nph)
            a, r, m = OM.phase_mesh_error(dyp[o:o + ny * Nph], y_ph, sI[ip], 0.5 * T, ph.N_K[ip],
                                          ph.mesh_index_boundaries[ip])
            o += ny * Nph
            worst = max(worst, float(m.max()))
        out["cpu_ms"] = round(1e3 * (time.perf_counter() - t0), 1)
        out["speedup"] = round(out["cpu_ms"] * 1e3 / us)
        out["max_rel_err_gpu_vs_cpu"] = float(abs(worst - float(mx.max())))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    run("free_flying_robot", 200, 5)
    run("space_shuttle_reentry", 200, 5)
    run("free_flying_robot", 20000, 5, cpu=False)
    run("space_shuttle_reentry", 20000, 5, cpu=False)
