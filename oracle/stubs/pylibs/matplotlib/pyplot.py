"""No-op matplotlib.pyplot (see matplotlib/__init__.py)."""
from . import _nothing


def subplots(*a, **k):
    return _nothing, _nothing


def __getattr__(name):
    return _nothing
