"""cyipopt-style callback object over the CUDA engine.

The method set and argument order are those of the reference's (dead but
explicit) IPOPT plumbing: ``IPOPTProblem`` in ``pycollo/nlp.py:36-76`` --
``objective(x)``, ``gradient(x)``, ``constraints(x)``, ``jacobian(x)``,
``jacobianstructure()``, ``hessian(x, lagrange, obj_factor)``,
``hessianstructure()``, ``intermediate(...)``.  That path orders the Jacobian
row-major and uses the *lower* triangle of the Hessian
(``pycollo/iteration.py:930-933, 965-968, 1057``), whereas the engine stores the
live backend's CasADi order (CCS / upper triangle); the fixed permutation comes
from ``pcx_structure_jac(PCX_ORDER_ROW_MAJOR)`` and is applied on the device.

An NLP solver asks for f, grad f, c and the Jacobian at the SAME iterate and
then for the Hessian there: the object caches by iterate.  The first callback
that sees a new x ships it to the device once (pinned staging) and evaluates
f, grad f and c in one fused launch; ``jacobian`` and ``hessian`` reuse the
device-resident x (``hessian`` ships only the multipliers).  Nothing is
recomputed and x crosses PCIe once per iterate.

``register_inputs=True`` (default) page-locks the caller's own x / multiplier arrays the
first time their address is seen (``cudaHostRegister``; an NLP solver hands its callbacks
the same few vectors over and over) so the upload is one DMA from the caller's memory: the
copy into a pinned staging buffer -- 0.2-0.4 ms per 4 MB vector on the bench host, as long
as the PCIe transfer itself -- disappears.  It applies when the object does not need its
own copy of x to compare with (``new_x`` flag given, or ``x_check="sampled"``); an address
that cannot be registered simply takes the staging path.

With ``derivative_level=1`` the object has no ``hessian`` / ``hessianstructure``
members at all (``NlpCallbacksFirstOrder``), which is how a cyipopt host selects
its limited-memory quasi-Newton mode (``pycollo/nlp.py:61-62``).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import engine as _engine

_LIBC = ctypes.CDLL(None)
_LIBC.memcmp.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]


class NlpCallbacksFirstOrder:
    """objective / gradient / constraints / jacobian (+ structure)."""

    MAX_REGISTERED = 8

    def __init__(self, iteration, ordering="cyipopt", x_check="full", register_inputs=True):
        self.it = iteration
        self.register_inputs = bool(register_inputs)
        self._registered = {}          # address -> (nbytes, torch view of the caller's array)
        self._x_sample = None
        self.num_registered_uploads = 0
        S = iteration.S
        if ordering not in ("cyipopt", "casadi"):
            raise ValueError("ordering must be 'cyipopt' or 'casadi'")
        if x_check not in ("full", "sampled"):
            raise ValueError("x_check must be 'full' or 'sampled'")
        self.ordering, self.x_check = ordering, x_check
        # a deferred iteration (structure-only use, no device) has no engine yet
        eng = None if getattr(iteration, "deferred", False) else iteration.create_engine()
        order = "row_major" if ordering == "cyipopt" else "ccs"
        if eng is not None and eng.has_structure:      # from the C ABI (pcx_structure_*)
            gr, gc, self.g_perm = eng.structure_jac(order)
            hr, hc, self.h_perm = eng.structure_hess(
                "tril_row_major" if ordering == "cyipopt" else "triu_ccs")
        else:                                          # host tables (same rule)
            gr, gc = S.G_structure()
            hr, hc = S.H_structure()
            if ordering == "cyipopt":
                self.g_perm = np.lexsort((gc, gr))
                gr, gc = gr[self.g_perm], gc[self.g_perm]
                hr, hc = hc, hr
            else:
                self.g_perm = np.arange(len(gr))
            self.h_perm = np.arange(len(hr))
        self._g_struct, self._h_struct = (gr, gc), (hr, hc)
        self._identity_g = bool(np.array_equal(self.g_perm, np.arange(len(self.g_perm))))
        self.num_evals = dict(objective=0, gradient=0, constraints=0, jacobian=0, hessian=0)
        self.num_launches = dict(point=0, jacobian=0, hessian=0)
        self.num_x_uploads = 0
        self._b = None
        self._have_x = False
        self._fresh = set()

    # -- persistent buffers ---------------------------------------------------------
    def _buffers(self):
        """Pinned host staging for x, lam and every result (copies at PCIe speed,
        results handed out as views -- valid until the next call of the same kind),
        device buffers, the permutation on the device."""
        if self._b is None:
            import torch
            it = self.it
            S = it.S
            eng = it.create_engine()
            dev = torch.device("cuda", it.device)
            f64 = dict(dtype=torch.float64)
            pin = lambda n: torch.empty(n, **f64).pin_memory()
            devb = lambda n: torch.empty(n, device=dev, **f64)
            b = dict(torch=torch, eng=eng, stream=torch.cuda.Stream(device=dev),
                     hx=pin(S.num_x), dx=devb(S.num_x), hl=pin(S.num_c), dl=devb(S.num_c),
                     hs=torch.ones(1, **f64).pin_memory(), ds=torch.ones(1, device=dev, **f64),
                     hf=pin(1), df=devb(1), hgrad=pin(S.num_x), dgrad=devb(S.num_x),
                     hc=pin(S.num_c), dc=devb(S.num_c),
                     dg=devb(S.nnz_g), dgp=devb(S.nnz_g), hg=pin(S.nnz_g),
                     dh=devb(S.nnz_h), hh=pin(S.nnz_h),
                     gperm=torch.from_numpy(np.ascontiguousarray(self.g_perm, dtype=np.int64)).to(dev))
            b["hx_np"] = b["hx"].numpy()
            self._b = b
        return self._b

    # -- the caller's own arrays as DMA sources -----------------------------------------
    def _caller_view(self, arr, b):
        """A torch view of ``arr`` whose pages are locked (registered on first sight of
        the address), or None: then the staging copy is used."""
        if not self.register_inputs or not isinstance(arr, np.ndarray) or arr.nbytes < (1 << 16):
            return None
        ptr, nbytes = arr.ctypes.data, arr.nbytes
        hit = self._registered.get(ptr)
        if hit is not None and hit[0] == nbytes:
            return hit[1]
        torch = b["torch"]
        lib = _engine.load_library()       # pcx_host_register: leaves no stale CUDA error behind
        if hit is not None:                                   # same address, other length
            lib.pcx_host_unregister(ptr)
            del self._registered[ptr]
        while len(self._registered) >= self.MAX_REGISTERED:   # oldest first
            old = next(iter(self._registered))
            lib.pcx_host_unregister(old)
            del self._registered[old]
        # a host that hands over a fresh temporary every time would pay a page-lock
        # per call: give up on registering after a few dozen distinct addresses
        self._num_registrations = getattr(self, "_num_registrations", 0) + 1
        if self._num_registrations > 4 * self.MAX_REGISTERED:
            self.register_inputs = False
            return None
        if lib.pcx_host_register(ptr, nbytes) != 0:
            return None
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")                   # read-only arrays are only read
            view = torch.from_numpy(arr)
        # keep the view, not the caller's array object, alive: the registration is per
        # ADDRESS, the solver owns the memory
        self._registered[ptr] = (nbytes, view)
        return view

    def close(self):
        """Unlock the caller's arrays registered by this object."""
        if self._registered:
            lib = _engine.load_library()
            for ptr in list(self._registered):
                lib.pcx_host_unregister(ptr)
            self._registered.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- iterate cache ----------------------------------------------------------------
    def _same_x(self, x, b):
        if not self._have_x:
            return False
        hx = b["hx_np"]
        if x.shape != hx.shape:
            return False
        if self.x_check == "sampled":
            step = max(1, x.size // 4096)
            ref = self._x_sample if self._x_sample is not None else hx[::step]
            last = self._x_last if self._x_sample is not None else hx[-1]
            return bool(np.array_equal(x[::step], ref) and x[-1] == last)
        return _LIBC.memcmp(x.ctypes.data, hx.ctypes.data, hx.nbytes) == 0

    def _touch(self, x, new_x=None):
        """Make ``x`` the device-resident iterate.  ``new_x`` (the flag IPOPT's own
        TNLP interface carries) skips the comparison: True = x changed, False =
        same x as the previous callback."""
        b = self._buffers()
        x_in = x
        x = np.ascontiguousarray(x, dtype=np.float64)
        flagged = new_x is not None
        if new_x is None:
            new_x = not self._same_x(x, b)
        if new_x or not self._have_x:
            torch, st = b["torch"], b["stream"]
            # the exact compare needs this object's own copy of x; with the solver's
            # new_x flag or the sampled compare the caller's array is the DMA source
            src = None
            if (flagged or self.x_check == "sampled") and x is x_in \
                    and x.shape == b["hx_np"].shape:
                src = self._caller_view(x, b)
            if src is None:
                np.copyto(b["hx_np"], x)
                src = b["hx"]
                self._x_sample = None
            else:
                self.num_registered_uploads += 1
                if self.x_check == "sampled" and not flagged:
                    step = max(1, x.size // 4096)
                    self._x_sample, self._x_last = x[::step].copy(), x[-1]
            with torch.cuda.stream(st):
                b["dx"].copy_(src, non_blocking=True)
            # every callback synchronises the stream before it returns, so the caller
            # may reuse its array as soon as it has the result
            self._have_x = True
            self._fresh.clear()
            self.num_x_uploads += 1
        return b

    def _point(self, x, new_x):
        """f, grad f and c at the iterate: one fused launch."""
        b = self._touch(x, new_x)
        if "point" not in self._fresh:
            torch, eng, st = b["torch"], b["eng"], b["stream"]
            with torch.cuda.stream(st):
                eng.eval_ptr(_engine.EVAL_F | _engine.EVAL_GRAD | _engine.EVAL_C, b["dx"],
                             f=b["df"], grad=b["dgrad"], c=b["dc"], stream=st.cuda_stream)
                b["hf"].copy_(b["df"], non_blocking=True)
                b["hgrad"].copy_(b["dgrad"], non_blocking=True)
                b["hc"].copy_(b["dc"], non_blocking=True)
            st.synchronize()
            self._fresh.add("point")
            self.num_launches["point"] += 1
        return b

    def objective(self, x, new_x=None):
        self.num_evals["objective"] += 1
        return float(self._point(x, new_x)["hf"][0])

    def gradient(self, x, new_x=None):
        self.num_evals["gradient"] += 1
        return self._point(x, new_x)["hgrad"].numpy()

    def constraints(self, x, new_x=None):
        self.num_evals["constraints"] += 1
        return self._point(x, new_x)["hc"].numpy()

    def jacobian(self, x, new_x=None):
        self.num_evals["jacobian"] += 1
        b = self._touch(x, new_x)
        if "jac" not in self._fresh:
            torch, eng, st = b["torch"], b["eng"], b["stream"]
            with torch.cuda.stream(st):
                eng.eval_ptr(_engine.EVAL_JAC, b["dx"], jac=b["dg"], stream=st.cuda_stream)
                src = b["dg"]
                if not self._identity_g:
                    eng.gather(b["dg"], b["gperm"], b["dg"].numel(), b["dgp"],
                               stream=st.cuda_stream)
                    src = b["dgp"]
                b["hg"].copy_(src, non_blocking=True)
            st.synchronize()
            self._fresh.add("jac")
            self.num_launches["jacobian"] += 1
        return b["hg"].numpy()

    def jacobianstructure(self):
        return self._g_struct

    def intermediate(self, alg_mod, iter_count, obj_value, inf_pr, inf_du, mu,
                     d_norm, regularization_size, alpha_du, alpha_pr, ls_trials):
        self.last_iterate = dict(iter_count=iter_count, obj_value=obj_value,
                                 inf_pr=inf_pr, inf_du=inf_du)


class NlpCallbacks(NlpCallbacksFirstOrder):
    """+ ``hessian(x, lagrange, obj_factor)`` / ``hessianstructure()``
    (``derivative_level=2``).  The lower triangle in row-major order IS the
    engine's upper triangle in CCS order with rows and columns swapped, so the
    Hessian values need no permutation in either ordering."""

    def __init__(self, iteration, ordering="cyipopt", x_check="full", register_inputs=True):
        if int(iteration.ocp.settings.derivative_level) < 2:
            raise ValueError("derivative_level=1: use NlpCallbacksFirstOrder "
                             "(no Hessian callback exists)")
        super().__init__(iteration, ordering, x_check, register_inputs)

    def hessian(self, x, lagrange, obj_factor, new_x=None):
        self.num_evals["hessian"] += 1
        b = self._touch(x, new_x)
        torch, eng, st = b["torch"], b["eng"], b["stream"]
        lam_in = lagrange
        lagrange = np.ascontiguousarray(lagrange, dtype=np.float64)
        src = (self._caller_view(lagrange, b)
               if lagrange is lam_in and lagrange.shape == (b["hl"].numel(),) else None)
        if src is None:
            np.copyto(b["hl"].numpy(), lagrange)
            src = b["hl"]
        b["hs"][0] = float(obj_factor)
        with torch.cuda.stream(st):
            b["dl"].copy_(src, non_blocking=True)
            b["ds"].copy_(b["hs"], non_blocking=True)
            eng.eval_ptr(_engine.EVAL_HESS, b["dx"], lam=b["dl"], sigma=b["ds"], hess=b["dh"],
                         stream=st.cuda_stream)
            b["hh"].copy_(b["dh"], non_blocking=True)
        st.synchronize()
        self.num_launches["hessian"] += 1
        return b["hh"].numpy()

    def hessianstructure(self):
        return self._h_struct
