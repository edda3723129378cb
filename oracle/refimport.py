"""Import the UNMODIFIED reference kernels headless.

TEST INFRASTRUCTURE ONLY.  The reference tree is looked up at B200SIM_REFERENCE_ROOT, then
/root/reference (authoring container), then oracle/_ref (the git-ignored verbatim copy made by
oracle/make_ref.py, which travels to the GPU box).  Used by tests/golden/make_golden.py to generate
fixtures, by tests that cross-check the oracle / the CUDA path against the live reference (skipped when
no tree is present) and by `bench.py --impl reference` / its `cpu_baseline` leg.  The product package
never imports this module.

The reference hard-imports PyOpenGL at module top (nbody/simulation.py:16-17,
boids/flock.py:6-7), which is not installed: empty stub modules are registered first.
Its kernels use numba cache=True and the tree is read-only, so NUMBA_CACHE_DIR is
pointed at a writable directory.
"""
from __future__ import annotations

import os
import sys
import types

def _find_root() -> str:
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in (os.environ.get("B200SIM_REFERENCE_ROOT"), "/root/reference", os.path.join(here, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "nbody")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


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
    libzstd with the two classes the recorder touches.  The binding is this module's own (ctypes on
    libzstd.so.1), independent of the product's codec, so frames written through it are not circular
    evidence for the product's codec."""
    if "zstandard" in sys.modules:
        return
    import ctypes as C
    import ctypes.util
    z = C.CDLL(ctypes.util.find_library("zstd") or "libzstd.so.1")
    z.ZSTD_compressBound.restype = C.c_size_t
    z.ZSTD_compressBound.argtypes = [C.c_size_t]
    z.ZSTD_compress.restype = C.c_size_t
    z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int]
    z.ZSTD_decompress.restype = C.c_size_t
    z.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]
    z.ZSTD_getFrameContentSize.restype = C.c_ulonglong
    z.ZSTD_getFrameContentSize.argtypes = [C.c_char_p, C.c_size_t]
    z.ZSTD_isError.restype = C.c_uint
    z.ZSTD_isError.argtypes = [C.c_size_t]
    z.ZSTD_versionString.restype = C.c_char_p

    class ZstdCompressor:
        def __init__(self, level=3, threads=0, **_kw):
            self.level = level

        def compress(self, data):
            data = bytes(data)
            cap = z.ZSTD_compressBound(len(data))
            buf = C.create_string_buffer(cap)
            n = z.ZSTD_compress(buf, cap, data, len(data), self.level)
            if z.ZSTD_isError(n):
                raise RuntimeError("ZSTD_compress failed")
            return buf.raw[:n]

    class ZstdDecompressor:
        def decompress(self, data, max_output_size=0):
            data = bytes(data)
            size = z.ZSTD_getFrameContentSize(data, len(data))
            if size >= (1 << 62):
                size = max_output_size or 64 * len(data)
            buf = C.create_string_buffer(int(size) or 1)
            n = z.ZSTD_decompress(buf, int(size), data, len(data))
            if z.ZSTD_isError(n):
                raise RuntimeError("ZSTD_decompress failed")
            return buf.raw[:n]

    m = types.ModuleType("zstandard")
    m.ZstdCompressor, m.ZstdDecompressor = ZstdCompressor, ZstdDecompressor
    m.__version__ = "shim-libzstd-" + z.ZSTD_versionString().decode()
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
