"""Named problem/mesh configurations shared by the golden-vector generator
(``oracle/make_golden_nlp.py``, which builds them on the reference's own
``pycollo`` package) and the parity tests (which build them on ``pycollo_b200``).
"""
from __future__ import annotations

from . import problems

RAGGED = dict(number_mesh_sections=5, mesh_section_sizes=[0.1, 0.3, 0.15, 0.25, 0.2],
              number_mesh_section_nodes=[4, 6, 3, 5, 4])
SMALL = dict(number_mesh_sections=4, mesh_section_sizes=None, number_mesh_section_nodes=4)

# name -> (builder, builder kwargs, mesh or None (= the default 10 sections x 4 nodes))
GOLDEN_CASES = {
    "brachistochrone_lobatto": ("brachistochrone", dict(quadrature_method="lobatto"), None),
    "brachistochrone_radau": ("brachistochrone", dict(quadrature_method="radau"), None),
    "brachistochrone_lobatto_ragged": ("brachistochrone", dict(quadrature_method="lobatto"), RAGGED),
    "hypersensitive_lobatto": ("hypersensitive", dict(quadrature_method="lobatto"), None),
    "hypersensitive_radau": ("hypersensitive", dict(quadrature_method="radau"), RAGGED),
    "cart_pole_lobatto": ("cart_pole_swing_up", dict(quadrature_method="lobatto"), None),
    "cart_pole_radau": ("cart_pole_swing_up", dict(quadrature_method="radau"), None),
    "free_flying_robot_lobatto": ("free_flying_robot", dict(quadrature_method="lobatto"), SMALL),
    "multiphase_lobatto": ("multiphase_sliding_mass", dict(num_phases=3), SMALL),
    "shuttle_lobatto": ("space_shuttle_reentry", dict(quadrature_method="lobatto"), SMALL),
    "free_flying_robot_radau": ("free_flying_robot", dict(quadrature_method="radau"), SMALL),
    "multiphase_radau": ("multiphase_sliding_mass", dict(num_phases=3, quadrature_method="radau"), RAGGED),
    "shuttle_radau": ("space_shuttle_reentry", dict(quadrature_method="radau"), SMALL),
    "double_pendulum_lobatto": ("double_pendulum", {}, None),
    "delta_iii_lobatto": ("delta_iii_launch_vehicle", {},
                          dict(number_mesh_sections=2, mesh_section_sizes=None,
                               number_mesh_section_nodes=4)),
}


def build_golden_problem(name, api=None):
    """The un-initialised problem of a golden case on either package."""
    builder, kwargs, mesh = GOLDEN_CASES[name]
    ocp = getattr(problems, builder)(api=api, **kwargs)
    if mesh is not None:
        ocp.settings.collocation_points_min = 2           # default 4 (quadrature.py:36)
        problems.set_mesh(ocp, **mesh)
    return ocp


def make_meshes(ocp, method, K, nodes, sizes=None):
    """K sections x `nodes` nodes for every phase (default reference-numerics tables)."""
    from pycollo_b200.mesh import PhaseMesh, PhaseMeshData
    from pycollo_b200.quadrature import Quadrature
    ocp.settings.quadrature_method = method
    quad = Quadrature(method)
    return [PhaseMeshData(quad, PhaseMesh(K, sizes, nodes), 2, 20) for _ in ocp.phases]


def lower_case(ocp, method, K, nodes, sizes=None, seed=0, unit_scaling=False,
               **structure_kwargs):
    """Lowered problem + (V, r, W_ocp, w) for a synthetic-iterate workload: the
    bounds-based variable scaling of ``pycollo/scaling.py:87-92`` and either unit
    or seeded random constraint / objective scaling.  No oracle involved."""
    import numpy as np
    from pycollo_b200.backend import Bounds, lower_problem
    meshes = make_meshes(ocp, method, K, nodes, sizes)
    low = lower_problem(ocp, meshes, **structure_kwargs)
    rng = np.random.default_rng(seed)
    W_ocp = np.ones(low.S.n_con_ocp) if unit_scaling else rng.uniform(0.5, 2.0, low.S.n_con_ocp)
    w = 1.0 if unit_scaling else 1.7
    if ocp.settings.scaling_method in (None, "none"):
        V, r = np.ones(low.S.n_var_ocp), np.zeros(low.S.n_var_ocp)
    else:
        bnd = Bounds(low.ir)
        V = bnd.x_bnd_upper - bnd.x_bnd_lower
        r = bnd.x_bnd_upper - V / 2
    return low, meshes, (V, r, W_ocp, w)
