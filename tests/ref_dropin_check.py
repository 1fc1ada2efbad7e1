"""Run in a subprocess by tests/test_reference_dropin.py (build container only: needs /root/reference).

The drop-in boundary on the reference's OWN classes: the unmodified reference package is imported
(under the test-side casadi / pyproprop stand-ins of oracle/refshim), its backend option set gets
the entry INTEGRATION.md section 2 describes -- ``"cuda"`` -> ``pycollo_b200.backend.Cuda`` in
``BACKENDS`` (``pycollo/backend.py:1925-1927``) --, the user script selects
``problem.settings.backend = "cuda"`` and the REFERENCE's ``OptimalControlProblem.initialise()``
(``optimal_control_problem.py:316-337``) drives our backend through its own call sequence
(``_initialise_backend`` -> ``create_bounds`` -> ``create_scaling`` -> ``create_quadrature`` ->
``postprocess_problem_backend`` -> ``create_initial_mesh`` -> ``create_guess`` ->
``create_mesh_iterations``).  No GPU here, so the engine is deferred (``settings.defer_engine``);
everything up to the device -- layout, patterns, scaling, guess, bounds -- is compared with the
goldens the reference's Casadi backend produced.  Prints one JSON line per case.
"""
import contextlib
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np                                                   # noqa: E402
import make_golden_nlp as M                                          # noqa: E402

pycollo, _casadi = M.import_reference()
import pycollo.backend as ref_backend                                # noqa: E402
from examples.cases import build_golden_problem                      # noqa: E402
from pycollo_b200.backend import Cuda                                # noqa: E402

B = ref_backend.BACKENDS                                             # the patch of INTEGRATION.md section 2
B.options += ("cuda",)
B.handles += (Cuda,)
B.dispatcher["cuda"] = Cuda

for name in sys.argv[1:]:
    ocp = build_golden_problem(name, api=pycollo)
    assert type(ocp).__module__ == "pycollo.optimal_control_problem"
    ocp.settings.display_mesh_result_graph = False
    ocp.settings.backend = "cuda"
    ocp.settings.defer_engine = True
    with contextlib.redirect_stdout(io.StringIO()):
        ocp.initialise()
    backend = ocp._backend
    it = backend.mesh_iterations[0]
    g = np.load(os.path.join(ROOT, "tests", "golden", f"nlp_{name}.npz"))
    gr, gc = it.S.G_structure()
    hr, hc = it.S.H_structure()
    eq = np.array_equal
    print(json.dumps(dict(
        name=name, backend=type(backend).__module__ + "." + type(backend).__name__,
        initialised=bool(ocp._is_initialised),
        sizes=bool(it.S.num_x == int(g["num_x"]) and it.S.num_c == int(g["num_c"])),
        G_pattern=bool(eq(gr, g["G_row"]) and eq(gc, g["G_col"])),
        H_pattern=bool(eq(hr, g["H_row"]) and eq(hc, g["H_col"])),
        V_r=bool(eq(it.scaling.V, g["V"]) and eq(it.scaling.r, g["r"])),
        guess=bool(eq(it.guess_x_tilde, g["guess_x"])),
        x_bounds=bool(eq(it.x_bnd_l, g["x_bnd_l"]) and eq(it.x_bnd_u, g["x_bnd_u"])))), flush=True)
