"""Generate golden fixtures from the *reference's own* modules.

TEST INFRASTRUCTURE ONLY -- run in the build container where ``/root/reference``
is mounted; the GPU box never sees the reference, only the committed ``.npz``
files this script writes under ``tests/golden/``.

What is produced
----------------
``quadrature_{lobatto,radau}.npz``
    points / weights / Butcher array / integration block ("A_matrix") for
    every order 2..20, by *executing* ``pycollo/quadrature.py`` (loaded by file
    path with a tiny ``pyproprop`` shim, see ``oracle/refshim``).
``mesh_*.npz``
    ``Mesh.generate_single_phase`` outputs (tau, h_K, N_K, boundaries, W,
    the integration CSR ``sI_matrix`` and the difference CSR ``sA_matrix``) of
    ``pycollo/mesh.py:236-356`` for a uniform and a ragged mesh under both
    quadrature schemes.
``iteration_scaling_{brachistochrone,double_pendulum}.npz``
    the reference's golden arrays ``EXPECT_V/R/V_INV/X/X_TILDE``
    (``tests/unit/iteration_scaling_test_data_*.py``) -- data, not code.

Usage:  python oracle/make_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_modules():
    sys.path.insert(0, os.path.join(HERE, "refshim"))
    quad = _load(os.path.join(REF, "pycollo", "quadrature.py"), "_ref_quadrature")
    mesh = _load(os.path.join(REF, "pycollo", "mesh.py"), "_ref_mesh")
    return quad, mesh


def make_backend(quad_mod, method, cmin=2, cmax=10):
    settings = types.SimpleNamespace(
        quadrature_method=method,
        collocation_points_min=cmin,
        collocation_points_max=cmax,
    )
    ocp = types.SimpleNamespace(settings=settings)
    backend = types.SimpleNamespace(ocp=ocp)
    backend.quadrature = quad_mod.Quadrature(backend)
    return backend


def dump_quadrature(quad_mod, method):
    backend = make_backend(quad_mod, method)
    q = backend.quadrature
    out = {}
    for order in range(2, 21):                      # Settings allow 2..20 (settings.py:234-251)
        out[f"points_{order}"] = np.asarray(q.quadrature_point(order), dtype=float)
        out[f"weights_{order}"] = np.asarray(q.quadrature_weight(order), dtype=float)
        out[f"butcher_{order}"] = np.asarray(q.butcher_array(order), dtype=float)
        out[f"A_{order}"] = np.asarray(q.A_matrix(order), dtype=float)
        out[f"D_{order}"] = np.asarray(q.D_matrix(order), dtype=float)
    np.savez(os.path.join(OUT, f"quadrature_{method}.npz"), **out)


def dump_mesh(quad_mod, mesh_mod, method, tag, sizes, nodes):
    backend = make_backend(quad_mod, method)
    phase = types.SimpleNamespace(
        optimal_control_problem=types.SimpleNamespace(settings=backend.ocp.settings))
    pm = mesh_mod.PhaseMesh(phase, number_mesh_sections=len(nodes),
                            mesh_section_sizes=sizes,
                            number_mesh_section_nodes=nodes)
    m = mesh_mod.Mesh(backend, [pm])
    sI = m.sI_matrix[0].tocsr()
    sA = m.sA_matrix[0].tocsr()
    np.savez(
        os.path.join(OUT, f"mesh_{method}_{tag}.npz"),
        section_sizes=np.asarray(pm.mesh_section_sizes, dtype=float),
        section_nodes=np.asarray(pm.number_mesh_section_nodes, dtype=np.int64),
        tau=m.tau[0], h_K=m.h_K[0], N=np.int64(m.N[0]), K=np.int64(m.K[0]),
        mesh_index_boundaries=np.asarray(m.mesh_index_boundaries[0], dtype=np.int64),
        W=m.W_matrix[0],
        sI_data=sI.data, sI_indices=sI.indices.astype(np.int64),
        sI_indptr=sI.indptr.astype(np.int64),
        sA_data=sA.data, sA_indices=sA.indices.astype(np.int64),
        sA_indptr=sA.indptr.astype(np.int64),
    )


def dump_iteration_scaling():
    base = os.path.join(REF, "tests", "unit")
    for long, short in (("brachistochrone", "BR"), ("double_pendulum", "DP")):
        mod = _load(os.path.join(base, f"iteration_scaling_test_data_{long}.py"),
                    f"_ref_data_{long}")
        np.savez(
            os.path.join(OUT, f"iteration_scaling_{long}.npz"),
            V=getattr(mod, f"EXPECT_V_{short}"),
            r=getattr(mod, f"EXPECT_R_{short}"),
            V_inv=getattr(mod, f"EXPECT_V_INV_{short}"),
            x=getattr(mod, f"EXPECT_X_{short}"),
            x_tilde=getattr(mod, f"EXPECT_X_TILDE_{short}"),
        )


class _Anything:
    """Import-time stand-in for the absent third-party ``casadi`` module: the
    reference's ``utils.py`` / ``mesh_refinement.py`` only *name* casadi symbols
    at import; ``phase_mesh_error`` itself is pure numpy."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def load_reference_mesh_refinement():
    """Import ``pycollo.mesh_refinement`` from the reference tree without running
    ``pycollo/__init__.py`` (which needs casadi/pyproprop for real)."""
    sys.path.insert(0, os.path.join(HERE, "refshim"))
    pkg = types.ModuleType("pycollo")
    pkg.__path__ = [os.path.join(REF, "pycollo")]
    sys.modules.setdefault("pycollo", pkg)
    ca = types.ModuleType("casadi")
    ca.__getattr__ = lambda name: _Anything()
    sys.modules.setdefault("casadi", ca)
    import importlib
    return importlib.import_module("pycollo.mesh_refinement"), \
        importlib.import_module("pycollo.mesh"), importlib.import_module("pycollo.quadrature")


def dump_mesh_error(tag, method, problem, sizes, nodes, seed):
    """Golden mesh errors: the reference's own ``phase_mesh_error`` (and its own
    ph ``Mesh``) executed on a synthetic trajectory; only the per-node dynamics
    ``dy_ph`` come from the oracle (the reference evaluates them with CasADi)."""
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle.blockwise import BlockwiseNLP
    from examples import problems as examples
    from pycollo_b200.backend import lower_problem
    from pycollo_b200.mesh import PhaseMesh as OurPhaseMesh, PhaseMeshData
    from pycollo_b200.quadrature import Quadrature as OurQuadrature
    mr, mesh_mod, quad_mod = load_reference_mesh_refinement()
    backend = make_backend(quad_mod, method, 2, 16)
    phase = types.SimpleNamespace(
        optimal_control_problem=types.SimpleNamespace(settings=backend.ocp.settings))
    nodes = np.asarray(nodes, dtype=np.int64)
    pm_ph = mesh_mod.PhaseMesh(phase, number_mesh_sections=len(nodes),
                               mesh_section_sizes=sizes,
                               number_mesh_section_nodes=nodes + 1)      # :75-86
    ocp = getattr(examples, problem)()
    ocp.settings.quadrature_method = method
    ocp.settings.scaling_method = "none"
    P = len(ocp.phases)
    ref_ph = mesh_mod.Mesh(backend, [pm_ph] * P)
    quad = OurQuadrature(method)
    ours_ph = [PhaseMeshData(quad, OurPhaseMesh(len(nodes), sizes, nodes + 1), 2, 17)
               for _ in range(P)]
    low = lower_problem(ocp, ours_ph)
    B = BlockwiseNLP(ocp, low.ir.full_bounds,
                     [dict(N=m.N, sI=m.sI_matrix, sA=m.sA_matrix, W=m.W_matrix) for m in ours_ph],
                     scaling_method="none")
    rng = np.random.default_rng(seed)
    # smooth synthetic trajectory inside the variable bounds + a rough component
    from pycollo_b200.backend import Bounds
    bnd = Bounds(low.ir)
    lo, hi = bnd.x_bnd_lower, bnd.x_bnd_upper
    lo = np.where(np.isfinite(lo), lo, -1.0)
    hi = np.where(np.isfinite(hi), hi, 1.0)
    x_ph = np.zeros(B.num_x)
    v = 0
    out = {}
    for ip, ph in enumerate(low.ir.phases):
        N = int(ref_ph.N[ip])
        tau = np.asarray(ref_ph.tau[ip])
        for a in range(ph.n_y + ph.n_u):
            mid, amp = 0.5 * (lo[v] + hi[v]), 0.2 * (hi[v] - lo[v])
            w1, w2, p1 = rng.uniform(0.5, 3.0), rng.uniform(4.0, 9.0), rng.uniform(0, 6.28)
            x_ph[B.x_off[ip] + a * N:B.x_off[ip] + (a + 1) * N] = \
                mid + amp * np.sin(w1 * tau + p1) + 0.05 * amp * np.cos(w2 * tau)
            v += 1
        nq, nt = ph.n_q, ph.n_t
        base = B.x_off[ip] + (ph.n_y + ph.n_u) * N
        for k in range(nq + nt):
            x_ph[base + k] = rng.uniform(lo[v] + 0.25 * (hi[v] - lo[v]), hi[v] - 0.25 * (hi[v] - lo[v]))
            v += 1
    for j in range(low.ir.n_s):
        x_ph[B.s_off + j] = 0.5 * (lo[v] + hi[v])
        v += 1
    dy_all = B.dy(x_ph)
    dyo = 0
    for ip, ph in enumerate(low.ir.phases):
        N = int(ref_ph.N[ip])
        y_ph = x_ph[B.x_off[ip]:B.x_off[ip] + ph.n_y * N].reshape(ph.n_y, N)
        u_ph = x_ph[B.x_off[ip] + ph.n_y * N:B.x_off[ip] + (ph.n_y + ph.n_u) * N].reshape(ph.n_u, N)
        base = B.x_off[ip] + (ph.n_y + ph.n_u) * N
        tvals = x_ph[base + ph.n_q:base + ph.n_q + ph.n_t]
        t0 = tvals[0] if ph.t_needed[0] else float(ph.t0)
        tF = tvals[-1] if ph.t_needed[1] else float(ph.tF)
        dy_p = dy_all[dyo:dyo + ph.n_y * N]
        dyo += ph.n_y * N
        fake = types.SimpleNamespace(
            dy_ph_callables=[None] * P, ph_mesh=ref_ph,
            it=types.SimpleNamespace(mesh=types.SimpleNamespace(K=ref_ph.K)),
            absolute_mesh_errors=[], relative_mesh_errors=[],
            maximum_relative_mesh_errors=[])
        fake.dy_ph_callables[ip] = lambda x, _d=dy_p: _d
        p_ns = types.SimpleNamespace(i=ip, num_y_var=ph.n_y)
        p_data = types.SimpleNamespace(stretch=0.5 * (tF - t0))
        mr.PattersonRaoMeshRefinement.phase_mesh_error(fake, p_ns, p_data, y_ph, u_ph, x_ph)
        out[f"abs_{ip}"] = fake.absolute_mesh_errors[0]
        out[f"rel_{ip}"] = fake.relative_mesh_errors[0]
        out[f"max_{ip}"] = fake.maximum_relative_mesh_errors[0]
        out[f"dy_{ip}"] = dy_p
        out[f"stretch_{ip}"] = np.float64(0.5 * (tF - t0))
        sI = ref_ph.sI_matrix[ip].tocsr()
        out[f"sI_data_{ip}"] = sI.data
        out[f"sI_indices_{ip}"] = sI.indices.astype(np.int64)
        out[f"sI_indptr_{ip}"] = sI.indptr.astype(np.int64)
    np.savez(os.path.join(OUT, f"mesh_error_{tag}.npz"), x_ph=x_ph,
             section_sizes=np.asarray(sizes, dtype=float), section_nodes=nodes,
             num_phases=np.int64(P), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    quad_mod, mesh_mod = load_reference_modules()
    ragged_sizes = [0.05, 0.2, 0.1, 0.15, 0.3, 0.2]
    ragged_nodes = [4, 7, 2, 10, 3, 5]
    for method in ("lobatto", "radau"):
        dump_quadrature(quad_mod, method)
        dump_mesh(quad_mod, mesh_mod, method, "uniform10x4", None, [4] * 10)
        dump_mesh(quad_mod, mesh_mod, method, "ragged6", ragged_sizes, ragged_nodes)
    dump_iteration_scaling()
    sizes = [0.1, 0.15, 0.05, 0.2, 0.1, 0.12, 0.08, 0.2]
    dump_mesh_error("robot_lobatto", "lobatto", "free_flying_robot", sizes, [4, 6, 3, 8, 5, 4, 9, 2], 1)
    dump_mesh_error("robot_radau", "radau", "free_flying_robot", sizes, [4, 6, 3, 8, 5, 4, 9, 3], 2)
    dump_mesh_error("shuttle_lobatto", "lobatto", "space_shuttle_reentry", sizes, [5, 4, 7, 3, 6, 4, 2, 8], 3)
    dump_mesh_error("multiphase_lobatto", "lobatto", "multiphase_sliding_mass", sizes, [3, 5, 4, 6, 2, 7, 4, 5], 4)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
