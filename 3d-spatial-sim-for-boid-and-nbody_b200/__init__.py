"""B200-native Barnes-Hut n-body step and boids neighbour-rule update.

The directory name is the project's (`3d-spatial-sim-for-boid-and-nbody_b200`), which is
not a Python identifier: import it as ``b200sim`` through the alias module at the repo root.

Host code is Python over a C-ABI CUDA library (include/b200sim.h) through ctypes.  There is
no CPU path in this package.
"""
from . import _lib  # noqa: F401
from ._lib import B200Error, pinned_empty  # noqa: F401

__all__ = ["B200Error", "pinned_empty", "nbody", "boids", "config", "presets"]
