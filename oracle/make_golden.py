"""Generate golden fixtures from the *reference's own* modules.

TEST INFRASTRUCTURE ONLY -- run in the build container where ``/root/reference``
is mounted; the GPU box never sees the reference, only the committed ``.npz``
files this script writes under ``tests/golden/``.

What is produced
----------------
``quadrature_{lobatto,radau}.npz``
    points / weights / Butcher array / integration block ("A_matrix") for
    every order 2..10, by *executing* ``pycollo/quadrature.py`` (loaded by file
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
    for order in range(2, 11):
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
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
