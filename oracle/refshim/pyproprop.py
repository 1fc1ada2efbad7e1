"""Stand-in for the third-party ``pyproprop`` package (absent here, no wheel).

TEST INFRASTRUCTURE ONLY -- never imported by the product.  It lets
``oracle/make_golden.py`` import and *execute* the unmodified reference package
from ``/root/reference`` in the build container.  Only the behaviour the
reference relies on is provided (``pycollo/settings.py:171-372``,
``guess.py:67-107``, ``backend.py:1925-1927`` ...): validated descriptors with
type casting, option sets with a dispatcher.  Error *messages* are not
reproduced -- nothing in the golden generation depends on them.
"""
import numpy as np


class Options:
    """Option set: ``.options``, ``.default``, ``.unsupported``, ``.handles``,
    ``.dispatcher`` (option -> handle)."""

    def __init__(self, options, default=None, unsupported=(), handles=None):
        if not isinstance(options, (tuple, list)):
            options = (options,)
        self.options = tuple(options)
        self.default = default if default is not None else self.options[0]
        if not isinstance(unsupported, (tuple, list)):
            unsupported = (unsupported,)
        self.unsupported = tuple(unsupported)
        self.handles = tuple(handles) if handles is not None else None
        self.dispatcher = (dict(zip(self.options, self.handles))
                           if self.handles is not None else {})

    def __iter__(self):
        return iter(self.options)

    def __contains__(self, item):
        return item in self.options


def processed_property(name, description=None, type=None, optional=False, default=None,
                       iterable_allowed=False, cast=False, min=None, max=None,
                       exclusive=False, at_least=None, at_most=None, options=None,
                       unsupported_options=(), method=None, read_only=False, len=None,
                       **unused):
    private = "_" + name
    desc = description or name
    kind = type
    import builtins

    def check_one(self, value):
        if value is None:
            if optional:
                return None
            if default is not None:
                return default
            # the reference stores None for several non-optional attributes at
            # construction time (e.g. Iteration.solution); accept it
            return None
        if cast and kind is not None and not isinstance(value, kind):
            if kind is np.ndarray:
                value = np.array(value)
            elif kind is bool:
                value = bool(value)
            else:
                value = kind(value)
        elif kind is not None and not cast and not isinstance(value, kind):
            if kind is float and isinstance(value, (int, np.integer, np.floating)):
                value = float(value)
            elif not isinstance(value, kind):
                raise TypeError(f"{desc} must be of type {kind}, got {builtins.type(value)}")
        if options is not None:
            opts = options.options if isinstance(options, Options) else tuple(options)
            unsup = (tuple(options.unsupported) if isinstance(options, Options) else ())
            if not isinstance(unsupported_options, (tuple, list)):
                extra = (unsupported_options,)
            else:
                extra = tuple(unsupported_options)
            unsup = unsup + extra
            if value not in opts:
                raise ValueError(f"{value!r} is not a valid option of {desc}")
            if value in unsup:
                raise ValueError(f"{value!r} is not currently supported as {desc}")
        if min is not None:
            if (value <= min) if exclusive else (value < min):
                raise ValueError(f"{desc} must be greater than {min}")
        if max is not None:
            if (value >= max) if exclusive else (value > max):
                raise ValueError(f"{desc} must be less than {max}")
        if at_least is not None and getattr(self, "_" + at_least, None) is not None:
            if value < getattr(self, "_" + at_least):
                raise ValueError(f"{desc} must be at least {at_least}")
        if at_most is not None and getattr(self, "_" + at_most, None) is not None:
            if value > getattr(self, "_" + at_most):
                raise ValueError(f"{desc} must be at most {at_most}")
        if method is not None:
            value = method(value)
        return value

    def getter(self):
        return getattr(self, private)

    def setter(self, value):
        if read_only and hasattr(self, private):
            raise AttributeError(f"{desc} is read-only")
        if iterable_allowed and isinstance(value, (tuple, list)):
            value = builtins.type(value)(check_one(self, v) for v in value)
        elif read_only:
            pass
        else:
            value = check_one(self, value)
        setattr(self, private, value)

    return property(getter, setter)
