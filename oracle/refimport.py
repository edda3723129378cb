"""Import the UNMODIFIED reference kernels headless (authoring container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so nothing that
runs there (-m gpu tests, smoke(), bench.py) may call this; it is used by
tests/golden/make_golden.py to generate fixtures and by CPU tests that cross-check the
oracle against the live reference when the tree is present (skipped otherwise).

The reference hard-imports PyOpenGL at module top (nbody/simulation.py:16-17,
boids/flock.py:6-7), which is not installed: empty stub modules are registered first.
Its kernels use numba cache=True and the tree is read-only, so NUMBA_CACHE_DIR is
pointed at a writable directory.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("B200SIM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "nbody")):
        return False
    try:
        import numba  # noqa: F401
    except Exception:
        return False
    return True


def _stub_opengl() -> None:
    if "OpenGL" in sys.modules:
        return
    ogl = types.ModuleType("OpenGL")
    gl = types.ModuleType("OpenGL.GL")
    gl.__all__ = []
    arrays = types.ModuleType("OpenGL.arrays")
    vbo = types.ModuleType("OpenGL.arrays.vbo")
    ogl.GL, ogl.arrays, arrays.vbo = gl, arrays, vbo
    sys.modules.update({"OpenGL": ogl, "OpenGL.GL": gl, "OpenGL.arrays": arrays, "OpenGL.arrays.vbo": vbo})


def _stub_zstandard() -> None:
    """tools/record.py:228 hard-imports `zstandard` (not installed): a minimal shim over the system
    libzstd (the same binding the product's codec uses) with the two classes the recorder touches."""
    if "zstandard" in sys.modules:
        return
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import b200sim  # noqa: F401
    from b200sim import codec

    class ZstdCompressor:
        def __init__(self, level=3, threads=0, **_kw):
            self.level = level

        def compress(self, data):
            return codec.zstd_compress(bytes(data), self.level, 0)

    class ZstdDecompressor:
        def decompress(self, data, max_output_size=0):
            return codec.zstd_decompress(bytes(data))

    m = types.ModuleType("zstandard")
    m.ZstdCompressor, m.ZstdDecompressor = ZstdCompressor, ZstdDecompressor
    m.__version__ = "shim-libzstd-" + codec.zstd_version()
    sys.modules["zstandard"] = m


def load_recorder():
    """The reference's frame codec functions, unmodified (tools/record.py:88-326)."""
    if not available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    _stub_opengl()
    _stub_zstandard()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    rec = importlib.import_module("tools.record")
    return types.SimpleNamespace(compress_frame=rec.compress_frame, decompress_frame=rec.decompress_frame,
                                 load_frame=rec.load_frame, save_frame=rec.save_frame, module=rec)


def load():
    """Returns a namespace with the reference's hot-path functions."""
    if not available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/b200sim_numba_cache")
    _stub_opengl()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    sim = importlib.import_module("nbody.simulation")
    flock = importlib.import_module("boids.flock")
    presets = importlib.import_module("tools.presets")
    ns = types.SimpleNamespace(
        simulation=sim, flock=flock, presets=presets,
        build_octree=sim.build_octree, compute_forces_barnes_hut=sim.compute_forces_barnes_hut,
        update_positions_velocities=sim.update_positions_velocities, compute_bounds=sim.compute_bounds,
        compute_colors_by_velocity=sim.compute_colors_by_velocity,
        generate_distribution=presets.generate_distribution, get_preset_config=presets.get_preset_config,
        Flock=flock.Flock)
    return ns
