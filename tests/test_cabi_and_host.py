"""C-ABI library loads without a GPU and exports every symbol include/pcx.h
declares; evaluations fail loudly (no CPU fallback); host-side API behaviour."""
import ctypes
import os
import re

import numpy as np
import pytest
import sympy as sym

import pycollo_b200
from pycollo_b200 import OptimalControlProblem
from examples import problems as examples
from pycollo_b200 import engine as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = E.load_library()
    header = open(os.path.join(ROOT, "include", "pcx.h")).read()
    names = sorted(set(re.findall(r"\b(pcx_[a-z_0-9]+)\s*\(", header)))
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"libpcx.so does not export {name}"
    assert lib.pcx_version().startswith(b"pcx")
    n = lib.pcx_table_count()
    listed = {lib.pcx_table_name(i).decode(): lib.pcx_table_elem_size(i) for i in range(n)}
    assert set(listed) == set(E._TABLE_DTYPES)
    for k, size in listed.items():
        assert np.dtype(E._TABLE_DTYPES[k]).itemsize == size, k


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ocp = examples.brachistochrone()
    with pytest.raises(E.PcxError, match="no CPU fallback"):
        ocp.initialise()                      # compiles the engine: needs the device
    ocp = examples.brachistochrone()
    ocp.settings.defer_engine = True          # structure-only use stops before the device
    ocp.initialise()
    with pytest.raises(E.PcxError, match="no CPU fallback"):
        ocp._backend.evaluate_J(np.zeros(125))


def test_pcx_create_rejects_bad_spec():
    lib = E.load_library()
    h = ctypes.c_void_p()
    assert lib.pcx_create(None, ctypes.byref(h)) == -1
    spec = E._Spec(batch=0, num_tiles=1, threads=128, problem_header=b"x")
    assert lib.pcx_create(ctypes.byref(spec), ctypes.byref(h)) == -1
    assert b"bad batch" in lib.pcx_last_error(None)


def test_settings_validation_mirrors_reference():
    ocp = OptimalControlProblem("x")
    assert ocp.settings.backend == "cuda"
    for name in ("casadi", "hsad", "pycollo", "sympy"):       # test_initialisation.py:33-49
        with pytest.raises(ValueError):
            ocp.settings.backend = name
    with pytest.raises(ValueError):
        ocp.settings.quadrature_method = "gauss"
    with pytest.raises(ValueError):
        ocp.settings.scaling_method = "guess"
    with pytest.raises(ValueError):
        ocp.settings.derivative_level = 3
    ocp.settings.quadrature_method = "RADAU"
    assert ocp.settings.quadrature_method == "radau"


def test_phase_symbols_and_errors():
    x, v, u = sym.symbols("x v u")
    ocp = OptimalControlProblem("p")
    ph = ocp.new_phase("A", state_variables=[x, v], control_variables=u)
    assert str(ph.initial_time_variable) == "t0_P0" and str(ph.final_time_variable) == "tF_P0"
    assert str(ph.initial_state_variables.x) == "x_P0(t0)"
    assert str(ph.final_state_variables[1]) == "v_P0(tF)"
    ph.integrand_functions = [u ** 2]
    assert str(ph.integral_variables[0]) == "q0_P0"
    ph.state_equations = [v]
    with pytest.raises(ValueError, match="state equation"):
        ph._check_variables_and_equations()
    ph.state_equations = {v: u, x: v}                 # dict form is ordered by state
    assert ph.state_equations == (v, u)
    with pytest.raises(ValueError):
        ph.state_variables = [x, x]
    import torch
    if not torch.cuda.is_available():                     # solve() initialises: needs the device
        with pytest.raises((ValueError, E.PcxError)):
            ocp.solve()


def test_initialise_layout_and_scaling_against_golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "iteration_scaling_double_pendulum.npz"))
    ocp = examples.double_pendulum()
    ocp.settings.defer_engine = True
    ocp.initialise()
    it = ocp._backend.mesh_iterations[0]
    assert (it.num_x, it.num_c) == (190, 121)
    np.testing.assert_array_equal(it.scaling.V, g["V"])
    np.testing.assert_array_equal(it.scaling.r, g["r"])
    np.testing.assert_allclose(it.scaling.V_inv, g["V_inv"], rtol=1e-15)
    np.testing.assert_allclose(it.guess_x, g["x"], atol=1e-15)
    np.testing.assert_allclose(it.guess_x_tilde, g["x_tilde"], atol=1e-15)
    np.testing.assert_allclose(it.scaling.unscale_x(it.scaling.scale_x(g["x"])), g["x"],
                               atol=1e-13)
    # slices (tests/unit/test_iteration.py:210-234)
    assert it.y_slices == [slice(0, 124)] and it.u_slices == [slice(124, 186)]
    assert it.q_slices == [slice(186, 187)] and it.t_slices == [slice(187, 188)]
    assert it.s_slice == slice(188, 190)
    assert it.c_defect_slices == [slice(0, 120)] and it.c_integral_slices == [slice(120, 121)]
    # state endpoint constraints are variable bounds, not rows of c
    assert it.x_bnd_l[0] == it.x_bnd_u[0] == -0.25


def test_undefined_symbol_is_reported():
    x, u, k = sym.symbols("x u k")
    ocp = OptimalControlProblem("p")
    ph = ocp.new_phase("A", state_variables=[x], control_variables=[u])
    ph.state_equations = [k * u]
    ph.bounds.initial_time, ph.bounds.final_time = 0, 1
    ph.bounds.state_variables = [[0, 1]]
    ph.bounds.control_variables = [[0, 1]]
    ph.guess.time = [0, 1]
    ph.guess.state_variables = [[0, 1]]
    ph.guess.control_variables = [[0, 0]]
    ocp.objective_function = ph.final_state_variables[0]
    with pytest.raises(ValueError, match="'k' is not defined"):
        ocp.initialise()


def test_nlp_callback_orderings_are_permutations():
    from pycollo_b200.nlp import NlpCallbacks
    ocp = examples.cart_pole_swing_up()
    ocp.settings.defer_engine = True
    ocp.initialise()
    it = ocp._backend.mesh_iterations[0]
    cb = NlpCallbacks(it)
    gr, gc = cb.jacobianstructure()
    assert np.all(np.diff(gr * it.S.num_x + gc) > 0)          # row-major
    hr, hc = cb.hessianstructure()
    assert np.all(hr >= hc)                                     # lower triangle
    assert sorted(cb.g_perm) == list(range(it.S.nnz_g))
    assert sorted(cb.h_perm) == list(range(it.S.nnz_h))
    assert pycollo_b200.__version__


def test_backend_symbol_primitives():
    """``sym / const / substitute_pycollo_sym / expr_as_numeric`` of the backend surface
    (``pycollo/backend.py:81-124, 1343-1384``): the CUDA backend is sympy-native, so the
    substitution resolves auxiliary data (phase-level shadowing included) and folds
    constant variables -- the expressions the code generator sees."""
    import sympy as sym
    from examples import problems
    from pycollo_b200.backend import Cuda
    ocp = problems.double_pendulum()
    ocp.settings.defer_engine = True
    b = Cuda(ocp)
    assert b.sym("a") == sym.Symbol("a") and b.sym("M", 2, 3).shape == (2, 3)
    assert b.const(2.5) == sym.Float(2.5)
    assert b.expr_as_numeric(sym.Float(1.5) * 2) == 3.0 and b.expr_as_numeric(sym.Integer(2)).dtype == np.float64
    ph = ocp.phases[0]
    lowered = [b.substitute_pycollo_sym(e, 0) for e in ph.state_equations]
    roots = set(b.p[0].y) | set(b.p[0].u) | set(b.ir.s)
    assert all(e.free_symbols <= roots for e in lowered)            # aux data resolved
    assert tuple(lowered) == tuple(b.p[0].f)                        # = what codegen was given
    assert b.substitute_pycollo_sym(tuple(ph.state_equations), b.p[0]) == tuple(lowered)
    J = b.substitute_pycollo_sym(ocp.objective_function)
    assert J == b.ir.J and J.free_symbols <= set(b.ir.point_symbols)
    with pytest.raises(ValueError, match="not defined"):
        b.substitute_pycollo_sym(sym.Symbol("nobody_defined_me") + 1, 0)


# ---- a compiled host without Python in the process (pcx_create_from_file) ----------------

def _build_tnlp_host(tmp_path):
    import subprocess
    exe = tmp_path / "tnlp_host"
    libdir = os.path.join(ROOT, "pycollo_b200")
    res = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", f"-I{os.path.join(ROOT, 'include')}",
                          os.path.join(ROOT, "examples", "tnlp_host.cpp"), f"-L{libdir}", "-lpcx",
                          f"-Wl,-rpath,{libdir}", "-o", str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    return exe


def test_spec_file_is_parsed_and_bad_files_are_rejected(tmp_path):
    """``Engine.write_spec`` -> ``pcx_create_from_file``: a well-formed file gets as far as
    the device (PCX_OK with a GPU, PCX_ECUDA 'no CPU fallback' without one); unreadable,
    foreign, truncated and over-long files are PCX_EINVAL with a message."""
    from helpers import build_case
    low, _, scal = build_case(examples.brachistochrone(), "lobatto", 10, 4, oracle=False)
    path = str(tmp_path / "br.pcxspec")
    E.Engine.write_spec(path, low.S, low.layouts, low.header, scal)
    # the documented layout, read back in Python: nothing lost, nothing reordered
    back = E.Engine.read_spec(path)
    tables = E.build_tables(low.S, low.layouts)
    assert back["header"] == low.header
    assert back["fields"]["num_x"] == low.S.num_x and back["fields"]["nnz_g"] == low.S.nnz_g
    assert back["fields"]["num_tiles"] == low.S.num_tiles and back["fields"]["threads"] == low.S.threads
    for name, arr in tables.items():
        assert back["tables"][name] == arr.tobytes(), name
    assert back["tables"]["g_rows"] == low.S.G_structure()[0].astype(np.int64).tobytes()
    for got, want in zip(back["scaling"], E.scaling_tables(low.S, low.layouts, *scal)):
        assert np.array_equal(got, want)
    lib = E.load_library()
    h = ctypes.c_void_p()
    rc = lib.pcx_create_from_file(path.encode(), 0, ctypes.byref(h))
    if rc == 0:
        n = ctypes.c_int64()
        lib.pcx_sizes(h, ctypes.byref(n), None, None, None, None, None)
        assert n.value == low.S.num_x
        lib.pcx_destroy(h)
    else:
        assert rc == -2 and b"no CPU fallback" in lib.pcx_last_error(None)
    data = open(path, "rb").read()
    cases = {"missing": None, "magic": b"NOTASPEC" + data[8:], "short": data[:len(data) // 2 // 8 * 8],
             "odd": data[:-3], "long": data + b"\0" * 8, "empty": b""}
    for name, blob in cases.items():
        q = str(tmp_path / f"{name}.pcxspec")
        if blob is not None:
            open(q, "wb").write(blob)
        assert lib.pcx_create_from_file(q.encode(), 0, ctypes.byref(h)) == -1, name
        assert not h.value
        assert len(lib.pcx_last_error(None)) > 8
    assert lib.pcx_create_from_file(None, 0, ctypes.byref(h)) == -1


def test_tnlp_host_example_compiles_against_the_c_header(tmp_path):
    """include/pcx.h is consumable from C++ and the TNLP-shaped host links against libpcx.so."""
    exe = _build_tnlp_host(tmp_path)
    import subprocess
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 2 and "usage" in res.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["brachistochrone", "multiphase_sliding_mass"])
def test_compiled_host_evaluates_without_python(name, tmp_path):
    """The C++ TNLP-shaped host (examples/tnlp_host.cpp) creates the engine from a spec
    file and calls every callback; results equal the Python-created engine's bit for bit,
    in the row-major / lower-triangular order a cyipopt / IPOPT host is given."""
    import subprocess
    from helpers import build_case
    low, _, scal = build_case(getattr(examples, name)(), "lobatto", 10, 4, oracle=False)
    S = low.S
    eng = E.Engine(S, low.layouts, low.header)
    eng.set_scaling(*scal)
    spec = str(tmp_path / "p.pcxspec")
    eng.save_spec(spec)
    rng = np.random.default_rng(5)
    x = rng.uniform(0.1, 0.4, S.num_x)
    lam = rng.standard_normal(S.num_c)
    x.tofile(tmp_path / "x.bin")
    lam.tofile(tmp_path / "lam.bin")
    exe = _build_tnlp_host(tmp_path)
    res = subprocess.run([str(exe), spec, str(tmp_path / "x.bin"), str(tmp_path / "lam.bin"),
                          str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    raw = open(tmp_path / "out.bin", "rb").read()
    n, m, nj, nh = np.frombuffer(raw, dtype=np.int64, count=4)
    assert (n, m, nj, nh) == (S.num_x, S.num_c, S.nnz_g, S.nnz_h)
    off = 32

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=int(count), offset=off)
        off += a.nbytes
        return a
    f = take(np.float64, 1)[0]
    grad, g = take(np.float64, n), take(np.float64, m)
    jr, jc, jv = take(np.int32, nj), take(np.int32, nj), take(np.float64, nj)
    hr, hc, hv = take(np.int32, nh), take(np.int32, nh), take(np.float64, nh)
    assert off == len(raw)
    # the same three launches the host makes (same compiled variants => same bits)
    ref = eng.eval_host(E.EVAL_F | E.EVAL_GRAD | E.EVAL_C, x)
    ref.update(eng.eval_host(E.EVAL_JAC, x))
    ref.update(eng.eval_host(E.EVAL_HESS, x, lam=lam, sigma=0.75))
    rj, cj, pj = eng.structure_jac("row_major")
    rh, ch, ph = eng.structure_hess("tril_row_major")
    assert f == ref["f"][0]
    assert np.array_equal(grad, ref["grad"][0]) and np.array_equal(g, ref["c"][0])
    assert np.array_equal(jr, rj) and np.array_equal(jc, cj)
    assert np.array_equal(hr, rh) and np.array_equal(hc, ch)
    assert np.all(hr >= hc)                                           # lower triangle
    assert np.array_equal(jv, ref["jac"][0][pj])
    assert np.array_equal(hv, ref["hess"][0][ph])


def test_header_is_plain_c99(tmp_path):
    """The boundary is a C ABI: ``include/pcx.h`` compiles as C99 with -pedantic (no C++
    types, no torch types in the signatures) and a C program links against the library."""
    import subprocess
    src = tmp_path / "c99.c"
    src.write_text('#include "pcx.h"\n#include <stdio.h>\n'
                   'int main(void) { printf("%s %d\\n", pcx_version(), pcx_table_count()); return 0; }\n')
    libdir = os.path.join(ROOT, "pycollo_b200")
    exe = tmp_path / "c99"
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror",
                          f"-I{os.path.join(ROOT, 'include')}", str(src), f"-L{libdir}", "-lpcx",
                          f"-Wl,-rpath,{libdir}", "-o", str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("pcx ")


def test_spec_reader_survives_corrupted_files(tmp_path):
    """200 random corruptions of a valid spec file (flipped bytes, truncations, absurd length
    fields): ``pcx_create_from_file`` answers PCX_EINVAL -- or gets as far as the device --
    and never crashes.  Run in a subprocess so that a crash would fail this test only."""
    import subprocess
    import sys
    from helpers import build_case
    low, _, scal = build_case(examples.brachistochrone(), "lobatto", 10, 4, oracle=False)
    good = str(tmp_path / "good.pcxspec")
    E.Engine.write_spec(good, low.S, low.layouts, low.header, scal)
    script = tmp_path / "fuzz.py"
    script.write_text(f"""
import ctypes, random, sys
sys.path.insert(0, {ROOT!r})
from pycollo_b200 import engine as E
data = open({good!r}, 'rb').read()
lib = E.load_library(); h = ctypes.c_void_p()
random.seed(7)
for it in range(200):
    b = bytearray(data)
    mode = it % 4
    if mode == 0:
        for _ in range(random.randint(1, 8)):
            b[random.randrange(len(b))] = random.randrange(256)
    elif mode == 1:
        n = random.randrange(0, len(b)); b = b[:n - n % 8]
    elif mode == 2:
        off = random.randrange(0, len(b) // 8) * 8
        b[off:off + 8] = random.choice([b'\\xff' * 8, (2 ** 40).to_bytes(8, 'little'), bytes(8),
                                        len(b).to_bytes(8, 'little'), (2 ** 63 - 1).to_bytes(8, 'little')])
    else:
        b[random.randrange(8, 200)] = random.randrange(256)
    open({str(tmp_path / 'bad.pcxspec')!r}, 'wb').write(bytes(b))
    rc = lib.pcx_create_from_file({str(tmp_path / 'bad.pcxspec')!r}.encode(), 0, ctypes.byref(h))
    if rc == 0:
        lib.pcx_destroy(h); h = ctypes.c_void_p()
    assert rc in (0, -1, -2, -3, -4), rc
print('survived')
""")
    res = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "survived" in res.stdout, (res.returncode, res.stderr[-1500:])
