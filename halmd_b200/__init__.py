"""halmd_b200 — Blackwell-native MD step loop behind HALMD's module API.

``from halmd_b200 import mdsim, observables`` gives the mirror of the Lua modules of the same names
(see halmd_b200/api.py); ``halmd_b200._lib`` is the raw ctypes binding of the C ABI in include/*.h.
Importing the package loads halmd_b200/libhalmd_b200.so and fails loudly when it has not been built.
"""
from . import _lib  # noqa: F401  (loads the shared library)
from .api import HxError, mdsim, observables  # noqa: F401

__all__ = ["mdsim", "observables", "HxError"]
