"""Minimal stand-in for the third-party ``pyproprop`` package (absent here).

TEST INFRASTRUCTURE ONLY.  The reference modules ``pycollo/quadrature.py`` and
``pycollo/mesh.py`` only need ``pyproprop.Options`` to be importable; nothing
else of pyproprop is exercised when those two modules are loaded in isolation
by ``oracle/make_golden.py``.  This shim is never imported by the product.
"""


class Options:
    def __init__(self, options, default=None, unsupported=(), handles=None):
        self.options = tuple(options)
        self.default = default
        if not isinstance(unsupported, (tuple, list)):
            unsupported = (unsupported,)
        self.unsupported = tuple(unsupported)
        self.handles = handles


def processed_property(name, **kwargs):
    private = "_" + name

    def getter(self):
        return getattr(self, private)

    def setter(self, value):
        setattr(self, private, value)

    return property(getter, setter)
