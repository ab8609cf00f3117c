"""No-op seaborn (test infrastructure; see ../matplotlib/__init__.py)."""
from matplotlib import _nothing


def __getattr__(name):
    return _nothing
