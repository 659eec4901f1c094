"""
Optional link to the reference package.

``pints_b200`` is standalone: it never needs ``pints`` to be installed.  When
the reference *is* importable, the mirror classes in this package additionally
inherit from the reference's abstract base classes, so they pass the
``isinstance`` checks inside the reference's controllers
(``pints/_optimisers/__init__.py:392-417``, ``pints/_log_pdfs.py:182-190``,
``pints/_error_measures.py:176-181``) and can be mixed freely with reference
objects.
"""
import importlib

_pints = None
_tried = False


def pints_or_none():
    global _pints, _tried
    if not _tried:
        _tried = True
        try:
            _pints = importlib.import_module('pints')
        except Exception:       # not installed / not on the path
            _pints = None
    return _pints


def base(name):
    """The reference's abstract base class ``pints.<name>`` or ``object``."""
    p = pints_or_none()
    cls = getattr(p, name, None) if p is not None else None
    return cls if isinstance(cls, type) else object


def is_instance(obj, local_class, name):
    """``isinstance`` against this package's class or the reference's
    ``pints.<name>`` -- the latter only where the reference is importable
    (``base(name)`` is ``object`` otherwise, which everything is)."""
    ref = base(name)
    return isinstance(obj, local_class) or \
        (ref is not object and isinstance(obj, ref))
