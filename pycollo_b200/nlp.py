"""cyipopt-style callback object over the CUDA engine.

The method set and argument order are those of the reference's (dead but
explicit) IPOPT plumbing: ``IPOPTProblem`` in ``pycollo/nlp.py:36-76`` --
``objective(x)``, ``gradient(x)``, ``constraints(x)``, ``jacobian(x)``,
``jacobianstructure()``, ``hessian(x, lagrange, obj_factor)``,
``hessianstructure()``, ``intermediate(...)``.  That path orders the Jacobian
row-major and uses the *lower* triangle of the Hessian
(``pycollo/iteration.py:930-933, 965-968, 1057``), whereas the engine stores the
live backend's CasADi order (CCS / upper triangle).  The two fixed permutations
are computed once here.
"""
from __future__ import annotations

import numpy as np

from . import engine as _engine


class NlpCallbacks:
    def __init__(self, iteration, ordering="cyipopt"):
        self.it = iteration
        S = iteration.S
        gr, gc = S.G_structure()
        hr, hc = S.H_structure()
        if ordering == "cyipopt":
            self.g_perm = np.lexsort((gc, gr))             # row-major
            self.h_perm = np.lexsort((hr, hc))             # tril, row-major
            self._g_struct = (gr[self.g_perm], gc[self.g_perm])
            # lower triangle: swap (row, col) of the stored upper triangle
            self._h_struct = (hc[self.h_perm], hr[self.h_perm])
        elif ordering == "casadi":
            self.g_perm = np.arange(len(gr))
            self.h_perm = np.arange(len(hr))
            self._g_struct = (gr, gc)
            self._h_struct = (hr, hc)
        else:
            raise ValueError("ordering must be 'cyipopt' or 'casadi'")
        self.ordering = ordering
        self.num_evals = dict(objective=0, gradient=0, constraints=0,
                              jacobian=0, hessian=0)
        self._fast = None

    # -- the hot callbacks: pinned staging, device-side permutation ------------
    def _fast_path(self):
        """Persistent buffers for ``jacobian`` / ``hessian``: pinned host staging for
        x, lam and the values (copies at PCIe speed instead of pageable-memory
        speed), and the cyipopt permutation applied on the device (``pcx_gather``)
        instead of a 4-million-element numpy fancy index per call."""
        if self._fast is None:
            import torch
            it = self.it
            S = it.S
            eng = it.create_engine()
            dev = torch.device("cuda", it.device)
            f64 = dict(dtype=torch.float64)
            b = dict(
                torch=torch, eng=eng, stream=torch.cuda.Stream(device=dev),
                hx=torch.empty(S.num_x, **f64).pin_memory(), dx=torch.empty(S.num_x, device=dev, **f64),
                hl=torch.empty(S.num_c, **f64).pin_memory(), dl=torch.empty(S.num_c, device=dev, **f64),
                ds=torch.ones(1, device=dev, **f64), hs=torch.ones(1, **f64).pin_memory(),
                dg=torch.empty(S.nnz_g, device=dev, **f64), dgp=torch.empty(S.nnz_g, device=dev, **f64),
                hg=torch.empty(S.nnz_g, **f64).pin_memory(),
                dh=torch.empty(S.nnz_h, device=dev, **f64), dhp=torch.empty(S.nnz_h, device=dev, **f64),
                hh=torch.empty(S.nnz_h, **f64).pin_memory(),
                gperm=torch.from_numpy(np.ascontiguousarray(self.g_perm, dtype=np.int64)).to(dev),
                hperm=torch.from_numpy(np.ascontiguousarray(self.h_perm, dtype=np.int64)).to(dev))
            self._fast = b
        return self._fast

    def _values(self, what, x, lagrange=None, obj_factor=1.0):
        b = self._fast_path()
        torch, eng, st = b["torch"], b["eng"], b["stream"]
        jac = what == _engine.EVAL_JAC
        b["hx"].numpy()[:] = x
        with torch.cuda.stream(st):
            b["dx"].copy_(b["hx"], non_blocking=True)
            if not jac:
                b["hl"].numpy()[:] = lagrange
                b["hs"][0] = float(obj_factor)
                b["dl"].copy_(b["hl"], non_blocking=True)
                b["ds"].copy_(b["hs"], non_blocking=True)
            raw, perm, out, host = (b["dg"], b["gperm"], b["dgp"], b["hg"]) if jac else \
                (b["dh"], b["hperm"], b["dhp"], b["hh"])
            if jac:
                eng.eval_ptr(what, b["dx"], jac=raw, stream=st.cuda_stream)
            else:
                eng.eval_ptr(what, b["dx"], lam=b["dl"], sigma=b["ds"], hess=raw, stream=st.cuda_stream)
            if self.ordering == "cyipopt":
                eng.gather(raw, perm, raw.numel(), out, stream=st.cuda_stream)
            else:
                out = raw
            host.copy_(out, non_blocking=True)
        st.synchronize()
        return host.numpy()          # pinned staging: valid until the next call of this kind

    def objective(self, x):
        self.num_evals["objective"] += 1
        return float(self.it.evaluate(_engine.EVAL_F, x)["f"][0])

    def gradient(self, x):
        self.num_evals["gradient"] += 1
        return self.it.evaluate(_engine.EVAL_GRAD, x)["grad"][0]

    def constraints(self, x):
        self.num_evals["constraints"] += 1
        return self.it.evaluate(_engine.EVAL_C, x)["c"][0]

    def jacobian(self, x):
        self.num_evals["jacobian"] += 1
        return self._values(_engine.EVAL_JAC, x)

    def jacobianstructure(self):
        return self._g_struct

    def hessian(self, x, lagrange, obj_factor):
        self.num_evals["hessian"] += 1
        return self._values(_engine.EVAL_HESS, x, lagrange, obj_factor)

    def hessianstructure(self):
        return self._h_struct

    def intermediate(self, alg_mod, iter_count, obj_value, inf_pr, inf_du, mu,
                     d_norm, regularization_size, alpha_du, alpha_pr, ls_trials):
        self.last_iterate = dict(iter_count=iter_count, obj_value=obj_value,
                                 inf_pr=inf_pr, inf_du=inf_du)
