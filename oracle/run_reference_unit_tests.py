"""Run the REFERENCE's own unit tests for the callback path under the test-side stand-ins.

TEST INFRASTRUCTURE ONLY (build container: needs ``/root/reference``).  The stand-ins of
``oracle/refshim`` (a sympy-backed ``casadi``, a minimal ``pyproprop``) are what lets the
unmodified reference package execute here and produce the golden vectors of
``tests/golden/nlp_*.npz`` (``oracle/make_golden_nlp.py``).  What validates the stand-ins
themselves is that the reference's own tests for this path pass on top of them:
``tests/unit/test_iteration.py`` (sizes, slices, J, g, c known answers),
``test_iteration_scaling.py`` and ``test_scaling.py`` (V, r, x <-> x_tilde golden arrays),
``test_quadrature.py`` (weights), ``test_initialisation.py``,
``test_optimal_control_problem.py``.

Usage:  python oracle/run_reference_unit_tests.py [extra pytest args]
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_nlp                                    # noqa: E402

FILES = ["test_iteration.py", "test_iteration_scaling.py", "test_scaling.py", "test_quadrature.py",
         "test_initialisation.py", "test_optimal_control_problem.py"]


def main(argv):
    make_golden_nlp.import_reference()                    # sys.path, casadi / pyproprop / matplotlib stand-ins
    import pytest
    unit = os.path.join(make_golden_nlp.REF, "tests", "unit")
    os.chdir("/tmp")                                      # the reference tree is read-only
    return pytest.main(["-q", "-p", "no:cacheprovider", "--rootdir", make_golden_nlp.REF]
                       + [os.path.join(unit, f) for f in FILES] + list(argv))


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
