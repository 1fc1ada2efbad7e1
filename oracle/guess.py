"""ORACLE -- guess interpolation to a new mesh.  TEST ONLY.

Restatement of ``pycollo/iteration.py:86-194`` (``interpolate_guess_to_mesh``):
every state / control row of the previous guess is passed through
``scipy.interpolate.interp1d(prev_tau, row, bounds_error=False,
fill_value="extrapolate")`` -- the very call the reference makes (``:128-134``) --
and q, t, s are carried over (``:166-168``); the result is chained in x order
(``:170-183``).  Only ``tests/`` may import this module.
"""
import numpy as np
from scipy import interpolate


def interpolate_guess_to_mesh(prev_tau, tau, prev_y, prev_u, prev_q, prev_t, prev_s):
    """Lists over phases (``prev_y[p]`` is (n_y, M_p) ...) -> guess_x on the new mesh."""
    parts = []
    for ptau, t, y, u, q, tt in zip(prev_tau, tau, prev_y, prev_u, prev_q, prev_t):
        for block in (y, u):
            new = np.empty((len(block), len(t)))
            for index, row in enumerate(block):                       # :127-135
                f = interpolate.interp1d(ptau, row, bounds_error=False, fill_value="extrapolate")
                new[index, :] = f(t)
            parts.append(new.ravel())
        parts += [np.ravel(q), np.ravel(tt)]
    parts.append(np.ravel(prev_s))
    return np.concatenate(parts)
