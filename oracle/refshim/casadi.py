"""sympy-backed stand-in for the third-party ``casadi`` module (absent here).

TEST INFRASTRUCTURE ONLY -- never imported by the product.  Its purpose is to
*execute the unmodified reference package* (``/root/reference/pycollo``) in the
build container so that golden vectors for the constraint Jacobian G and the
Lagrangian Hessian H can be emitted by the reference's own statements
(``pycollo/backend.py:1403-1679, 1681-1693, 1713-1771``) instead of by a
restatement.  ``oracle/make_golden.py`` is the only user.

What is modelled (the surface the reference touches):

``SX``         sparse matrix of scalar expressions; an entry is a sympy expression
               (sympy's automatic canonicalisation plays the role of SXElem's
               on-the-fly folding of ``0*x``, ``1*x``, ``x+0``, ``x-x``); entries
               absent from ``_d`` are structural zeros.  Element-wise operator
               patterns follow CasADi: union for +/-, intersection for *.
``DM``         numeric (possibly sparse) matrix; ``np.array(DM)`` densifies,
               ``nonzeros()`` lists the structural non-zeros column by column.
``substitute`` simultaneous replacement of symbols.
``jacobian``   pattern = *dependency* of row expression i on variable j (what
               ``Function::jac_sparsity`` propagates through the SX graph: a
               constant has no dependency, so coefficients that are exactly zero
               in the mesh matrices -- Radau's last integration column -- do not
               create entries); values = sympy derivatives.  ``jacobian_check``
               reports entries whose derivative is identically zero although the
               dependency exists (none for the fixtures: the two rules coincide).
``gradient``   sparse column, same rule.
``Function``   numeric evaluation through ``sympy.lambdify`` (double precision) or,
               with ``precise=True``, mpmath at 40 digits rounded once.
``nlpsol``     the five ``get_function`` members the reference and CasADi's
               nlpsol expose -- ``nlp_f``, ``nlp_grad_f``, ``nlp_g``, ``nlp_jac_g``
               and ``nlp_hess_l`` = upper triangle (``triu``) of
               ``jacobian(gradient(lam_f*f + lam_g.g, x), x)`` in CCS order, which
               is how CasADi's ``Nlpsol`` oracle builds it -- and a ``__call__``
               that solves the NLP with scipy's ``trust-constr`` (IPOPT is absent;
               used only for objective pins, never for iteration counts).

CCS order everywhere: non-zeros sorted by (column, row), as CasADi stores them
and as ``Casadi.evaluate_G_structure`` (``backend.py:1747-1761``) reads them.
"""
import math as _math
import numbers as _numbers

import numpy as _np
import scipy.sparse as _sparse
import sympy as _sp

__version__ = "refshim-sympy"

# Function objects compiled while this is True evaluate with mpmath at 40 digits
# and round once to double ("exact" golden values, free of the shim's own
# operation-order rounding); oracle/make_golden_nlp.py switches it on.
PRECISE = False


def _is_number(v):
    return isinstance(v, (_numbers.Number, _np.generic)) and not isinstance(v, bool) or isinstance(v, bool)


def _to_expr(v):
    if isinstance(v, _sp.Basic):
        return v
    if isinstance(v, (bool, _np.bool_)):
        return _sp.Integer(int(v))
    if isinstance(v, (int, _np.integer)):
        return _sp.Integer(int(v))
    if isinstance(v, (float, _np.floating)):
        f = float(v)
        if f == int(f) and abs(f) < 2 ** 53:
            # integral doubles are exact integers: lets sympy fold 0*x, 1*x, x**2.0
            # the way SXElem folds against the numeric values
            return _sp.Integer(int(f))
        return _sp.Float(f)                  # 53-bit mantissa: the exact double
    raise TypeError(f"cannot convert {type(v)} to an SX entry")


class SX:
    __array_priority__ = 1000
    __array_ufunc__ = None

    def __init__(self, *args):
        self.shape = (0, 0)
        self._d = {}
        if len(args) == 0:
            return
        if len(args) == 2 and all(isinstance(a, (int, _np.integer)) for a in args):
            self.shape = (int(args[0]), int(args[1]))
            return
        (v,) = args
        if isinstance(v, SX):
            self.shape = v.shape
            self._d = dict(v._d)
        elif isinstance(v, DM):
            self.shape = v.shape
            self._d = {k: _to_expr(val) for k, val in v._items()}
        elif _sparse.issparse(v):
            m = v.tocoo()
            self.shape = m.shape
            self._d = {(int(i), int(j)): _to_expr(val) for i, j, val in zip(m.row, m.col, m.data)}
        elif isinstance(v, _np.ndarray) or isinstance(v, (list, tuple)):
            arr = _np.array(v, dtype=object)
            if arr.ndim == 0:
                arr = arr.reshape(1, 1)
            elif arr.ndim == 1:
                arr = arr.reshape(-1, 1)
            self.shape = arr.shape
            for (i, j), e in _np.ndenumerate(arr):
                if isinstance(e, SX):
                    if e.shape != (1, 1):
                        raise ValueError("nested non-scalar SX")
                    if (0, 0) in e._d:
                        self._d[(i, j)] = e._d[(0, 0)]
                else:
                    self._d[(i, j)] = _to_expr(e)
        else:
            self.shape = (1, 1)
            self._d = {(0, 0): _to_expr(v)}

    # ---- construction ---------------------------------------------------
    @staticmethod
    def sym(name, rows=1, cols=1):
        out = SX(int(rows), int(cols))
        n = rows * cols
        if n == 1:
            out._d[(0, 0)] = _sp.Symbol(name, real=True)
        else:
            k = 0
            for j in range(cols):
                for i in range(rows):
                    out._d[(i, j)] = _sp.Symbol(f"{name}_{k}", real=True)
                    k += 1
        return out

    @staticmethod
    def zeros(rows=1, cols=1):
        out = SX(rows, cols)
        for i in range(rows):
            for j in range(cols):
                out._d[(i, j)] = _sp.Integer(0)
        return out

    @staticmethod
    def _from_items(shape, items):
        out = SX(shape[0], shape[1])
        out._d = dict(items)
        return out

    # ---- inspection -----------------------------------------------------
    def size(self):
        return self.shape

    def size1(self):
        return self.shape[0]

    def size2(self):
        return self.shape[1]

    def numel(self):
        return self.shape[0] * self.shape[1]

    def nnz(self):
        return len(self._d)

    def _keys_ccs(self):
        return sorted(self._d, key=lambda k: (k[1], k[0]))

    def colind(self):
        counts = [0] * (self.shape[1] + 1)
        for (_, j) in self._d:
            counts[j + 1] += 1
        for j in range(self.shape[1]):
            counts[j + 1] += counts[j]
        return counts

    def row(self):
        return [k[0] for k in self._keys_ccs()]

    def sparsity(self):
        return self

    def nonzeros(self):
        return [SX(self._d[k]) for k in self._keys_ccs()]

    def is_scalar(self):
        return self.shape == (1, 1)

    def _scalar(self):
        if self.shape != (1, 1):
            raise ValueError(f"expected a scalar SX, got shape {self.shape}")
        return self._d.get((0, 0), _sp.Integer(0))

    def is_symbolic(self):
        return self.shape == (1, 1) and isinstance(self._d.get((0, 0)), _sp.Symbol)

    def is_constant(self):
        return all(not e.free_symbols for e in self._d.values())

    def name(self):
        e = self._scalar()
        if not isinstance(e, _sp.Symbol):
            raise RuntimeError("name() of a non-symbolic SX")
        return e.name

    def __hash__(self):
        if self.shape == (1, 1):
            return hash(self._scalar())
        return hash((self.shape, tuple(sorted((k, hash(v)) for k, v in self._d.items()))))

    def __eq__(self, other):
        if not isinstance(other, SX):
            try:
                other = SX(other)
            except TypeError:
                return False
        return self.shape == other.shape and self._d == other._d

    def __ne__(self, other):
        return not self.__eq__(other)

    def __bool__(self):
        e = self._scalar()
        if e.free_symbols:
            raise RuntimeError("truth value of a symbolic SX")
        return bool(e != 0)

    def __float__(self):
        return float(self._scalar())

    def __int__(self):
        return int(self._scalar())

    def __repr__(self):
        if self.shape == (1, 1):
            return f"SX({self._scalar()})"
        return f"SX({self.shape[0]}x{self.shape[1]}, {len(self._d)}nz)"

    __str__ = __repr__

    @property
    def T(self):
        return SX._from_items((self.shape[1], self.shape[0]),
                              {(j, i): e for (i, j), e in self._d.items()})

    def __len__(self):
        return self.shape[0] if self.shape[1] == 1 else self.shape[0] * self.shape[1]

    def __iter__(self):
        # column-vector iteration (CasADi would raise; the reference never relies on it)
        for i in range(len(self)):
            yield self[i]

    def __getitem__(self, key):
        r, c = self.shape
        if isinstance(key, tuple):
            i, j = key
            rows = range(r)[i] if isinstance(i, slice) else [i % r if i < 0 else i]
            cols = range(c)[j] if isinstance(j, slice) else [j % c if j < 0 else j]
            out = SX(len(rows), len(cols))
            for a, ii in enumerate(rows):
                for b, jj in enumerate(cols):
                    if (ii, jj) in self._d:
                        out._d[(a, b)] = self._d[(ii, jj)]
            return out
        n = r * c
        if isinstance(key, slice):
            idx = list(range(n)[key])
        else:
            k = int(key)
            if k < 0:
                k += n
            if not 0 <= k < n:
                raise IndexError("SX index out of range")
            idx = [k]
        out = SX(len(idx), 1)
        for a, k in enumerate(idx):
            ij = (k % r, k // r)                  # column-major linear index
            if ij in self._d:
                out._d[(a, 0)] = self._d[ij]
        return out

    # ---- element-wise algebra --------------------------------------------
    def _binary(self, other, fn, kind):
        """kind: 'union' (f(0,0)=0), 'inter' (f(x,0)=f(0,y)=0), 'left' (f(0,y)=0), 'dense'."""
        if not isinstance(other, SX):
            other = SX(other)
        a, b = self, other
        if a.shape == (1, 1) and b.shape != (1, 1):
            shape = b.shape
            ea = a._d.get((0, 0))
            get_a = lambda k: ea
            keys_a = None if ea is not None else set()
            get_b = b._d.get
            keys_b = set(b._d)
        elif b.shape == (1, 1) and a.shape != (1, 1):
            shape = a.shape
            eb = b._d.get((0, 0))
            get_b = lambda k: eb
            keys_b = None if eb is not None else set()
            get_a = a._d.get
            keys_a = set(a._d)
        else:
            if a.shape != b.shape:
                raise ValueError(f"shape mismatch {a.shape} vs {b.shape}")
            shape = a.shape
            get_a, get_b = a._d.get, b._d.get
            keys_a, keys_b = set(a._d), set(b._d)
        allk = {(i, j) for i in range(shape[0]) for j in range(shape[1])}
        ka = allk if keys_a is None else keys_a
        kb = allk if keys_b is None else keys_b
        if kind == "union":
            keys = ka | kb
        elif kind == "inter":
            keys = ka & kb
        elif kind == "left":
            keys = ka
        else:
            keys = allk
        zero = _sp.Integer(0)
        out = SX(shape[0], shape[1])
        for k in keys:
            x = get_a(k)
            y = get_b(k)
            out._d[k] = fn(zero if x is None else x, zero if y is None else y)
        return out

    def __add__(self, o):
        return self._binary(o, lambda x, y: x + y, "union")

    def __radd__(self, o):
        return SX(o)._binary(self, lambda x, y: x + y, "union")

    def __sub__(self, o):
        return self._binary(o, lambda x, y: x - y, "union")

    def __rsub__(self, o):
        return SX(o)._binary(self, lambda x, y: x - y, "union")

    def __mul__(self, o):
        return self._binary(o, lambda x, y: x * y, "inter")

    def __rmul__(self, o):
        return SX(o)._binary(self, lambda x, y: x * y, "inter")

    def __truediv__(self, o):
        return self._binary(o, lambda x, y: x / y, "left")

    def __rtruediv__(self, o):
        return SX(o)._binary(self, lambda x, y: x / y, "left")

    def __pow__(self, o):
        return self._binary(o, lambda x, y: x ** y, "dense")

    def __rpow__(self, o):
        return SX(o)._binary(self, lambda x, y: x ** y, "dense")

    def __neg__(self):
        return SX._from_items(self.shape, {k: -e for k, e in self._d.items()})

    def __pos__(self):
        return self

    def __abs__(self):
        return fabs(self)

    def _map(self, fn, zero_preserving):
        if zero_preserving:
            return SX._from_items(self.shape, {k: fn(e) for k, e in self._d.items()})
        out = SX(self.shape[0], self.shape[1])
        zero = _sp.Integer(0)
        for i in range(self.shape[0]):
            for j in range(self.shape[1]):
                out._d[(i, j)] = fn(self._d.get((i, j), zero))
        return out


class DM:
    """Numeric matrix; ``_pattern`` (CCS-ordered keys) is kept when it came from a
    sparse symbolic matrix so that ``nonzeros()`` follows the symbolic pattern."""
    __array_priority__ = 1000

    def __init__(self, v=None, pattern=None, shape=None):
        if isinstance(v, DM):
            self._a, self._pattern = v._a.copy(), v._pattern
            return
        if isinstance(v, SX):
            if not v.is_constant():
                raise RuntimeError("DM(SX) of a non-constant expression")
            a = _np.zeros(v.shape)
            for (i, j), e in v._d.items():
                a[i, j] = float(e)
            self._a, self._pattern = a, v._keys_ccs()
            return
        if _sparse.issparse(v):
            m = v.tocoo()
            self._a = _np.asarray(v.todense(), dtype=float)
            self._pattern = sorted({(int(i), int(j)) for i, j in zip(m.row, m.col)},
                                   key=lambda k: (k[1], k[0]))
            return
        if v is None:
            a = _np.zeros((0, 0)) if shape is None else _np.zeros(shape)
        else:
            a = _np.array(v, dtype=float)
        if a.ndim == 0:
            a = a.reshape(1, 1)
        elif a.ndim == 1:
            a = a.reshape(-1, 1)
        self._a = a
        self._pattern = pattern

    @property
    def shape(self):
        return self._a.shape

    def _items(self):
        if self._pattern is not None:
            return [((i, j), self._a[i, j]) for (i, j) in self._pattern]
        return [((i, j), v) for (i, j), v in _np.ndenumerate(self._a)]

    def __array__(self, dtype=None, copy=None):
        return self._a.astype(dtype) if dtype is not None else self._a

    def full(self):
        return self._a

    def nonzeros(self):
        return [float(v) for _, v in sorted(self._items(), key=lambda kv: (kv[0][1], kv[0][0]))]

    def nnz(self):
        return len(self._items())

    def size1(self):
        return self._a.shape[0]

    def size2(self):
        return self._a.shape[1]

    def __float__(self):
        return float(self._a.reshape(-1)[0])

    def __getitem__(self, k):
        return DM(self._a.reshape(-1, order="F")[k] if not isinstance(k, tuple) else self._a[k])

    def __len__(self):
        return self._a.shape[0]

    def __repr__(self):
        return f"DM({self._a.tolist()})"

    def _sx(self):
        return SX(self)

    def __add__(self, o): return self._sx() + o
    def __radd__(self, o): return o + self._sx()
    def __sub__(self, o): return self._sx() - o
    def __rsub__(self, o): return o - self._sx()
    def __mul__(self, o): return self._sx() * o
    def __rmul__(self, o): return o * self._sx()
    def __truediv__(self, o): return self._sx() / o
    def __rtruediv__(self, o): return o / self._sx()
    def __neg__(self): return -self._sx()


def _as_sx(v):
    return v if isinstance(v, SX) else SX(v)


# ---- structural helpers ------------------------------------------------------
def vertcat(*args):
    parts = [_as_sx(a) for a in args]
    parts = [p for p in parts if p.shape[0] * max(p.shape[1], 1) > 0 or p.shape[0] > 0]
    if not parts:
        return SX(0, 1)
    cols = max(p.shape[1] for p in parts)
    out_d, r0 = {}, 0
    for p in parts:
        if p.shape[1] != cols and p.shape[0] > 0:
            raise ValueError("vertcat: column mismatch")
        for (i, j), e in p._d.items():
            out_d[(r0 + i, j)] = e
        r0 += p.shape[0]
    return SX._from_items((r0, cols), out_d)


def horzcat(*args):
    return vertcat(*[_as_sx(a).T for a in args]).T


def vertsplit(v, *unused):
    v = _as_sx(v)
    return [v[i, :] for i in range(v.shape[0])]


def blockcat(*args):
    rows = args[0] if len(args) == 1 else args
    return vertcat(*[horzcat(*r) for r in rows])


def symvar(v):
    v = _as_sx(v)
    syms = set()
    for e in v._d.values():
        syms |= e.free_symbols
    return [SX(s) for s in sorted(syms, key=lambda s: s.name)]


def substitute(expr, remove, add):
    expr, remove, add = _as_sx(expr), _as_sx(remove), _as_sx(add)
    if remove.numel() != add.numel():
        raise ValueError("substitute: size mismatch")
    if remove.numel() == 0:
        return SX(expr)
    mapping = {}
    zero = _sp.Integer(0)
    for i in range(remove.shape[0]):
        for j in range(remove.shape[1]):
            k = remove._d.get((i, j))
            if not isinstance(k, _sp.Symbol):
                raise RuntimeError("substitute: can only replace symbols")
            val = add._d.get((i, j), zero)
            if k != val:
                mapping[k] = val
    if not mapping:
        return SX(expr)
    keys = set(mapping)
    out = {}
    for ij, e in expr._d.items():
        out[ij] = e.xreplace(mapping) if (e.free_symbols & keys) else e
    return SX._from_items(expr.shape, out)


def mtimes(a, b):
    a, b = _as_sx(a), _as_sx(b)
    if a.shape == (1, 1) or b.shape == (1, 1):
        return a * b
    if a.shape[1] != b.shape[0]:
        raise ValueError(f"mtimes: {a.shape} x {b.shape}")
    brow = {}
    for (k, j), e in b._d.items():
        brow.setdefault(k, []).append((j, e))
    acc = {}
    for (i, k), ea in a._d.items():
        for j, eb in brow.get(k, ()):
            acc.setdefault((i, j), []).append(ea * eb)
    return SX._from_items((a.shape[0], b.shape[1]), {ij: _sp.Add(*terms) for ij, terms in acc.items()})


def dot(a, b):
    a, b = _as_sx(a), _as_sx(b)
    if a.shape != b.shape:
        if a.shape == b.shape[::-1]:
            b = b.T
        else:
            raise ValueError("dot: shape mismatch")
    terms = [e * b._d[k] for k, e in a._d.items() if k in b._d]
    return SX(_sp.Add(*terms)) if terms else SX(1, 1)


def sum1(a):
    a = _as_sx(a)
    acc = {}
    for (i, j), e in a._d.items():
        acc.setdefault((0, j), []).append(e)
    return SX._from_items((1, a.shape[1]), {k: _sp.Add(*t) for k, t in acc.items()})


def _depends(expr_syms, var_index):
    return sorted((var_index[s] for s in expr_syms if s in var_index))


def _var_index(x):
    x = _as_sx(x)
    if x.shape[1] != 1 and x.shape[0] == 1:
        x = x.T
    idx = {}
    for (i, j), e in x._d.items():
        if not isinstance(e, _sp.Symbol):
            raise RuntimeError("differentiation variable is not purely symbolic")
        idx[e] = i
    return x, idx


def jacobian(f, x):
    f = _as_sx(f)
    x, idx = _var_index(x)
    if f.shape[1] != 1:
        raise NotImplementedError("jacobian of a non-vector")
    out = {}
    for (i, _), e in f._d.items():
        for s in e.free_symbols:
            j = idx.get(s)
            if j is not None:
                out[(i, j)] = _sp.diff(e, s)
    return SX._from_items((f.shape[0], x.shape[0]), out)


def jacobian_check(jac):
    """Entries that are structurally present (dependency) but identically zero."""
    return [k for k, e in jac._d.items() if e == 0]


def gradient(f, x):
    return jacobian(_as_sx(f), x).T


def hessian(f, x):
    g = gradient(f, x)
    return jacobian(g, x), g


def triu(a):
    a = _as_sx(a)
    return SX._from_items(a.shape, {(i, j): e for (i, j), e in a._d.items() if i <= j})


def tril(a):
    a = _as_sx(a)
    return SX._from_items(a.shape, {(i, j): e for (i, j), e in a._d.items() if i >= j})


# ---- element-wise functions (names as sympy's lambdify printer emits them) ----
def _unary(fn, zero_preserving):
    def apply(v):
        if isinstance(v, (SX, DM)):
            return _as_sx(v)._map(fn, zero_preserving)
        return _as_sx(v)._map(fn, zero_preserving)
    return apply


sin = _unary(_sp.sin, True)
cos = _unary(_sp.cos, False)
tan = _unary(_sp.tan, True)
asin = arcsin = _unary(_sp.asin, True)
acos = arccos = _unary(_sp.acos, False)
atan = arctan = _unary(_sp.atan, True)
sinh = _unary(_sp.sinh, True)
cosh = _unary(_sp.cosh, False)
tanh = _unary(_sp.tanh, True)
asinh = _unary(_sp.asinh, True)
acosh = _unary(_sp.acosh, False)
atanh = _unary(_sp.atanh, True)
exp = _unary(_sp.exp, False)
log = _unary(_sp.log, False)
sqrt = _unary(_sp.sqrt, True)
fabs = _unary(_sp.Abs, True)
sign = _unary(_sp.sign, True)
floor = _unary(_sp.floor, True)
ceil = _unary(_sp.ceiling, True)


def atan2(y, x):
    return _as_sx(y)._binary(x, lambda a, b: _sp.atan2(a, b), "dense")


arctan2 = atan2


def fmin(a, b):
    return _as_sx(a)._binary(b, lambda x, y: _sp.Min(x, y), "dense")


def fmax(a, b):
    return _as_sx(a)._binary(b, lambda x, y: _sp.Max(x, y), "dense")


def power(a, b):
    return _as_sx(a) ** b


pi = _math.pi
inf = _math.inf
e = _math.e


# ---- numeric evaluation -------------------------------------------------------
class Function:
    def __init__(self, name, args, outs, *unused, precise=False):
        self._name = name
        self._args = [_as_sx(a) for a in args]
        self._outs = [_as_sx(o) for o in outs]
        self._fn = None
        self._precise = precise or PRECISE

    def name(self):
        return self._name

    def sx_in(self, i=None):
        return self._args if i is None else self._args[i]

    def sx_out(self, i=None):
        return self._outs if i is None else self._outs[i]

    def n_in(self):
        return len(self._args)

    def _compile(self):
        syms = []
        for a in self._args:
            col = a if a.shape[1] == 1 or a.shape[0] != 1 else a.T
            zero = _sp.Integer(0)
            for j in range(col.shape[1]):
                for i in range(col.shape[0]):
                    s = col._d.get((i, j), zero)
                    if not isinstance(s, _sp.Symbol):
                        s = _sp.Dummy()            # structural zero / constant input slot
                    syms.append(s)
        self._keys = [o._keys_ccs() for o in self._outs]
        exprs = [o._d[k] for o, ks in zip(self._outs, self._keys) for k in ks]
        self._syms = syms
        self._exprs = exprs
        if self._precise:
            import mpmath
            self._fn = _sp.lambdify(syms, exprs, modules="mpmath", cse=True)
            self._mp = mpmath
        else:
            self._fn = _sp.lambdify(syms, exprs, modules=["math"], cse=True)

    def __call__(self, *vals):
        if self._fn is None:
            self._compile()
        flat = []
        for a, v in zip(self._args, vals):
            n = a.numel()
            if n == 0:
                continue
            if isinstance(v, DM):
                arr = _np.asarray(v._a, dtype=float).reshape(-1, order="F")
            else:
                arr = _np.asarray(v, dtype=float).reshape(-1)
            if arr.size != n:
                raise ValueError(f"{self._name}: argument has {arr.size} entries, expected {n}")
            flat.extend(arr.tolist())
        if self._precise:
            mp = self._mp
            with mp.workdps(40):
                res = self._fn(*[mp.mpf(v) for v in flat])
                res = [float(r) for r in res]
        else:
            res = self._fn(*flat)
        outs, pos = [], 0
        for o, ks in zip(self._outs, self._keys):
            a = _np.zeros(o.shape)
            for k in ks:
                a[k] = res[pos]
                pos += 1
            outs.append(DM(a, pattern=list(ks)))
        return outs[0] if len(outs) == 1 else outs


def _running_error(expr, env, cache, u=2.0 ** -53):
    """(value, bound) of a first-order running error analysis of evaluating
    ``expr`` in fp64 with exact inputs: every operation contributes ``u*|result|``
    and propagates the bounds of its operands through its partial derivatives."""
    hit = cache.get(expr)
    if hit is not None:
        return hit
    if expr.is_Symbol:
        res = (env[expr], 0.0)
    elif expr.is_Number:
        v = float(expr)
        res = (v, 0.0 if (expr.is_Integer or v == 0.0) else 0.0)
    elif expr.is_Add:
        v = e = 0.0
        for a in expr.args:
            va, ea = _running_error(a, env, cache, u)
            v += va
            e += ea + u * abs(v)
        res = (v, e)
    elif expr.is_Mul:
        v, e = 1.0, 0.0
        for a in expr.args:
            va, ea = _running_error(a, env, cache, u)
            e = abs(va) * e + abs(v) * ea
            v *= va
            e += u * abs(v)
        res = (v, e)
    elif expr.is_Pow:
        b, p_ = expr.args
        vb, eb = _running_error(b, env, cache, u)
        vp, ep = _running_error(p_, env, cache, u)
        v = vb ** vp
        dvb = abs(vp * vb ** (vp - 1.0)) if vb != 0.0 else 0.0
        e = dvb * eb + u * abs(v) * max(1.0, abs(vp))
        if ep:
            e += abs(v * _math.log(abs(vb))) * ep if vb else 0.0
        res = (v, e)
    else:
        vals, errs = zip(*[_running_error(a, env, cache, u) for a in expr.args])
        fn = _sp.lambdify([], expr.func(*[_sp.Float(v) for v in vals]), modules="math")
        v = float(fn())
        e = u * abs(v) * 2.0
        for i, (va, ea) in enumerate(zip(vals, errs)):
            if ea:
                d = _sp.Dummy()
                args = [_sp.Float(x) for x in vals]
                args[i] = d
                dv = float(_sp.diff(expr.func(*args), d).subs(d, va))
                e += abs(dv) * ea
        res = (v, e)
    cache[expr] = res
    return res


def error_bounds(function, *vals):
    """Per structural non-zero of each output of ``function``: a bound of the
    absolute error of evaluating the reference's expression in fp64 at ``vals``
    (same order as ``nonzeros()``).  Tells an ill-conditioned entry (cancellation
    after ``x = V*x_tilde + r``) from a real discrepancy."""
    if function._fn is None:
        function._compile()
    flat = []
    for a, v in zip(function._args, vals):
        if a.numel():
            flat.extend(_np.asarray(v, dtype=float).reshape(-1).tolist())
    env = {s: float(v) for s, v in zip(function._syms, flat)}
    cache = {}
    out = []
    for o, ks in zip(function._outs, function._keys):
        out.append(_np.array([_running_error(o._d[k], env, cache)[1] for k in ks]))
    return out


class _NlpSolver:
    """The callable object ``ca.nlpsol`` returns: CasADi's Nlpsol oracle functions
    + a host solve (scipy trust-constr instead of IPOPT)."""

    def __init__(self, name, plugin, nlp, opts=None):
        self._x = _as_sx(nlp["x"])
        self._f = _as_sx(nlp["f"])
        self._g = _as_sx(nlp.get("g", SX(0, 1)))
        self._p = _as_sx(nlp.get("p", SX(0, 1)))
        self._opts = opts or {}
        self._fns = {}
        self.stats_ = {}

    def get_function(self, name):
        if name in self._fns:
            return self._fns[name]
        x, p, f, g = self._x, self._p, self._f, self._g
        if name == "nlp_f":
            fn = Function(name, [x, p], [f])
        elif name == "nlp_g":
            fn = Function(name, [x, p], [g])
        elif name == "nlp_grad_f":
            fn = Function(name, [x, p], [f, gradient(f, x)])
        elif name == "nlp_jac_g":
            fn = Function(name, [x, p], [g, jacobian(g, x)])
        elif name == "nlp_hess_l":
            # CasADi's Nlpsol oracle: triu(jacobian(gradient(lam_f*f + lam_g.g, x), x)).
            # lam_f and the lam_g[i] are independent symbols, so that matrix is
            # exactly  lam_f*hess(f) + sum_i lam_g[i]*hess(g_i)  with the union
            # pattern; forming it row by row keeps sympy's work proportional to the
            # few variables each constraint row depends on.
            lam_f = SX.sym("lam_f")
            lam_g = SX.sym("lam_g", g.shape[0])
            _, idx = _var_index(x)
            acc = {}
            rows = [(lam_f._scalar(), f._scalar())]
            rows += [(lam_g._d[(i, 0)], e) for (i, _), e in sorted(g._d.items())]
            for mult, e in rows:
                syms = [s_ for s_ in e.free_symbols if s_ in idx]
                for sj in syms:
                    dj = _sp.diff(e, sj)
                    for sk in dj.free_symbols:
                        if sk in idx and idx[sj] <= idx[sk]:
                            acc.setdefault((idx[sj], idx[sk]), []).append(mult * _sp.diff(dj, sk))
                        elif sk in idx and sk not in e.free_symbols:
                            raise RuntimeError("derivative depends on a new symbol")
            n = x.shape[0] if x.shape[1] == 1 else x.shape[1]
            hess = SX._from_items((n, n), {k: _sp.Add(*t) for k, t in acc.items()})
            fn = Function(name, [x, p, lam_f, lam_g], [hess])
        else:
            raise KeyError(name)
        self._fns[name] = fn
        return fn

    def stats(self):
        return self.stats_

    def __call__(self, x0=None, lbx=None, ubx=None, lbg=None, ubg=None, p=None, **unused):
        from scipy.optimize import Bounds, NonlinearConstraint, minimize
        n, m = self._x.shape[0], self._g.shape[0]
        f_fn, gf_fn = self.get_function("nlp_f"), self.get_function("nlp_grad_f")
        g_fn, jg_fn = self.get_function("nlp_g"), self.get_function("nlp_jac_g")
        h_fn = self.get_function("nlp_hess_l")
        jkeys = jg_fn._outs[1]._keys_ccs()
        jr = _np.array([k[0] for k in jkeys]); jc = _np.array([k[1] for k in jkeys])
        hkeys = h_fn._outs[0]._keys_ccs()
        hr = _np.array([k[0] for k in hkeys]); hc = _np.array([k[1] for k in hkeys])

        def fun(x):
            return float(f_fn(x, []))

        def grad(x):
            return _np.array(gf_fn(x, [])[1]).reshape(-1)

        def con(x):
            return _np.array(g_fn(x, [])).reshape(-1)

        def jac(x):
            vals = _np.array(jg_fn(x, [])[1].nonzeros())
            return _sparse.csr_matrix((vals, (jr, jc)), shape=(m, n))

        def hess_of(sig, lam):
            def h(x, *a):
                vals = _np.array(h_fn(x, [], sig, lam).nonzeros())
                up = _sparse.coo_matrix((vals, (hr, hc)), shape=(n, n))
                return (up + _sparse.triu(up, 1).T).tocsr()
            return h

        lbx = _np.full(n, -_np.inf) if lbx is None else _np.asarray(lbx, float).reshape(-1)
        ubx = _np.full(n, _np.inf) if ubx is None else _np.asarray(ubx, float).reshape(-1)
        lbg = _np.asarray(lbg, float).reshape(-1) if m else _np.zeros(0)
        ubg = _np.asarray(ubg, float).reshape(-1) if m else _np.zeros(0)
        cons = []
        if m:
            cons = [NonlinearConstraint(con, lbg, ubg, jac=jac,
                                        hess=lambda x, v: hess_of(0.0, v)(x))]
        ip = self._opts.get("ipopt", {})
        res = minimize(fun, _np.asarray(x0, float).reshape(-1), jac=grad,
                       hess=lambda x: hess_of(1.0, _np.zeros(m))(x),
                       bounds=Bounds(lbx, ubx, keep_feasible=False), constraints=cons,
                       method="trust-constr",
                       options={"gtol": max(float(ip.get("tol", 1e-8)), 1e-10), "xtol": 1e-12,
                                "maxiter": int(ip.get("max_iter", 2000)), "verbose": 0})
        self.stats_ = {"success": bool(res.success), "iter_count": int(res.nit),
                       "return_status": str(res.message)}
        lam_g = _np.asarray(res.v[0]) if m else _np.zeros(0)
        return {"x": DM(res.x), "f": DM(res.fun), "g": DM(con(res.x)) if m else DM(),
                "lam_x": DM(_np.zeros(n)), "lam_g": DM(lam_g), "lam_p": DM()}


def nlpsol(name, plugin, nlp, opts=None):
    return _NlpSolver(name, plugin, nlp, opts)
