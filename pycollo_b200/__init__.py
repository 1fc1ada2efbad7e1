"""pycollo_b200 -- B200-native evaluation engine for pycollo's NLP callbacks.

``backend="cuda"`` behind the ``OptimalControlProblem`` / ``Phase`` / ``Settings``
API: objective, gradient, constraints, constraint Jacobian and Lagrangian Hessian
of the direct-collocation NLP are evaluated by hand-written sm_100a CUDA kernels
(``csrc/``) reached through the C ABI of ``include/pcx.h``.
"""
from .problem import (EndpointBounds, EndpointGuess, OptimalControlProblem,  # noqa: F401
                      Phase, PhaseBounds, PhaseGuess, Settings)
from .mesh import PhaseMesh  # noqa: F401

__version__ = "0.1.0"
__all__ = ["OptimalControlProblem", "Settings", "PhaseBounds", "EndpointBounds",
           "PhaseGuess", "EndpointGuess", "Phase", "PhaseMesh"]
