"""Row (b): ``backend="cuda"`` selected on the REFERENCE's own ``OptimalControlProblem``.

Build container only (``/root/reference`` is not on the GPU box): see
``tests/ref_dropin_check.py`` for what runs.  A subprocess keeps the reference package and the
casadi stand-in out of this process's module table."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["brachistochrone_lobatto", "cart_pole_radau", "multiphase_radau",
         "double_pendulum_lobatto", "shuttle_radau", "free_flying_robot_lobatto"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/pycollo"),
                    reason="the reference tree is only present in the build container")
def test_reference_initialise_drives_the_cuda_backend():
    res = subprocess.run([sys.executable, os.path.join(HERE, "ref_dropin_check.py")] + CASES,
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]
    rows = [json.loads(line) for line in res.stdout.splitlines() if line.startswith("{")]
    assert [r["name"] for r in rows] == CASES
    for r in rows:
        assert r["backend"] == "pycollo_b200.backend.Cuda", r
        for key in ("initialised", "sizes", "G_pattern", "H_pattern", "V_r", "guess", "x_bounds"):
            assert r[key] is True, (r["name"], key)


@pytest.mark.skipif(not os.path.isdir("/root/reference/pycollo"),
                    reason="the reference tree is only present in the build container")
def test_reference_unit_tests_pass_under_the_stand_ins():
    """What validates ``oracle/refshim`` (and with it the reference-executed goldens): the
    reference's OWN unit tests for this path pass on top of it
    (``oracle/run_reference_unit_tests.py``: 30 passed in ~5 min).  The regular run leaves out
    the five double-pendulum cases that take a minute each under sympy;
    ``PCX_REFERENCE_UNIT_FULL=1`` runs everything."""
    args = [] if os.environ.get("PCX_REFERENCE_UNIT_FULL") else \
        ["-k", "not _dp and not create_iter_var_symbols and not generate_scaling_symbols"]
    res = subprocess.run([sys.executable, os.path.join(os.path.dirname(HERE), "oracle",
                                                       "run_reference_unit_tests.py")] + args,
                         capture_output=True, text=True, timeout=1500)
    tail = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else ""
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    passed = int(tail.split(" passed")[0].split()[-1])
    assert passed >= (30 if not args else 20), tail
    assert "failed" not in tail
