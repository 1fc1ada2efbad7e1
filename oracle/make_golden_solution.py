"""Golden vectors for the solution post-processing / mesh-error chain (rows N2,
a11, a12) produced by EXECUTING THE UNMODIFIED REFERENCE (see
``oracle/make_golden_nlp.py`` for how it is imported).  TEST INFRASTRUCTURE ONLY.

For a synthetic smooth iterate x_tilde on the first mesh the reference's own
  ``Casadi.process_solution`` -> ``CasadiSolution``        (solution/casadi_solution.py:6-86:
       unscale, dy = dy_iter_callable(x), per-phase slices)
  ``SolutionABC.interpolate_solution_{lobatto,radau}``    (solution/solution_abc.py:60-142)
  ``PattersonRaoMeshRefinement(solution)``                 (mesh_refinement.py:61-240:
       ph mesh, dy_ph CasADi functions, construct_x_ph, phase_mesh_error)
are run and ``x``, ``dy``, ``x_ph`` and the absolute / relative / per-section maximum
errors are written to ``tests/golden/solution_<case>.npz``.

Usage:  python oracle/make_golden_solution.py [name ...]
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_nlp import OUT, ROOT, import_reference          # noqa: E402

CASES = ("brachistochrone_lobatto", "brachistochrone_lobatto_ragged", "cart_pole_radau",
         "free_flying_robot_lobatto", "multiphase_lobatto", "hypersensitive_radau")


def generate(name, pycollo, casadi):
    from examples.cases import build_golden_problem
    import pycollo.backend as ref_backend
    ocp = build_golden_problem(name, api=pycollo)
    ocp.settings.display_mesh_result_graph = False
    with contextlib.redirect_stdout(io.StringIO()):
        ocp.initialise()
    backend = ocp._backend
    it = backend.mesh_iterations[0]
    # a smooth iterate strictly inside the bounds: the scaled guess plus a smooth
    # perturbation of every mesh variable (same rule on every run: seeded)
    rng = np.random.default_rng(sum(map(ord, name)) + 7)
    x = np.array(it.guess_x, dtype=np.float64)
    for sl_list, N_list in ((it.y_slices, it.mesh.N), (it.u_slices, it.mesh.N)):
        for sl, N in zip(sl_list, N_list):
            nvar = (sl.stop - sl.start) // N
            for a in range(nvar):
                tau = np.linspace(-1, 1, N)
                w1, p1 = rng.uniform(1.0, 4.0), rng.uniform(0, 6.28)
                seg = slice(sl.start + a * N, sl.start + (a + 1) * N)
                x[seg] = np.clip(x[seg] + 0.08 * np.sin(w1 * tau + p1), -0.45, 0.45)
    nlp_result = ref_backend.NlpResult(
        solution={"x": casadi.DM(x), "f": casadi.DM(float(backend.evaluate_J(x)))},
        info=None, solve_time=0.0)
    with contextlib.redirect_stdout(io.StringIO()):
        sol = backend.process_solution(it, nlp_result)               # iteration.py:510
        from pycollo.mesh_refinement import PattersonRaoMeshRefinement
        mr = PattersonRaoMeshRefinement(sol)                         # solution_abc.py:147-151
    x_ph, y_ph, u_ph = mr.construct_x_ph()
    out = {"x": x, "objective": np.float64(sol.objective),
           "dy": np.concatenate([np.ravel(d) for d in sol._dy]),
           "x_ph": np.asarray(x_ph, dtype=np.float64),
           "num_phases": np.int64(len(backend.p))}
    for ip in range(len(backend.p)):
        out[f"abs_{ip}"] = mr.absolute_mesh_errors[ip]
        out[f"rel_{ip}"] = mr.relative_mesh_errors[ip]
        out[f"max_{ip}"] = mr.maximum_relative_mesh_errors[ip]
        out[f"tau_ph_{ip}"] = np.asarray(mr.ph_mesh.tau[ip], dtype=np.float64)
        nm = mr.next_iter_mesh
        out[f"next_N_K_{ip}"] = np.asarray(nm.N_K[ip], dtype=np.int64)
        out[f"next_h_K_{ip}"] = np.asarray(nm.h_K[ip], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, f"solution_{name}.npz"), **out)
    print(f"{name}: x_ph {len(x_ph)} entries, worst relative mesh error "
          f"{max(float(m.max()) for m in mr.maximum_relative_mesh_errors):.3e}", flush=True)


def main(argv):
    pycollo, casadi = import_reference()
    for name in (argv or CASES):
        generate(name, pycollo, casadi)


if __name__ == "__main__":
    main(sys.argv[1:])
