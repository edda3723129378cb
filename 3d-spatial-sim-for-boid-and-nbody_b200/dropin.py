"""Run the reference's own tools against the B200 backend, unchanged.

    python -m b200sim.dropin --reference-root /path/to/reference record --preset quick_galaxy
    python -m b200sim.dropin --reference-root /path/to/reference --seed 0 record --preset tiny_galaxy --frames 6
    python -m b200sim.dropin --reference-root /path/to/reference nbody_main

``--seed S`` calls ``numpy.random.seed(S)`` before the tool starts: the reference's generators draw from
numpy's global RandomState (tools/presets.py:91-1390) and are otherwise unseeded.

``install()`` puts the reference checkout on ``sys.path`` and registers this package's
``nbody.gpu_backend`` under the reference's module name, so ``from nbody.gpu_backend import
get_backend, Backend, create_gpu_simulation`` (tools/record.py:760, nbody/simulation.py:511)
resolves to the B200 Barnes-Hut backend.  No reference file is modified.

Two launcher-side conveniences for headless boxes (NOT part of the hot path): empty ``OpenGL``
stubs when PyOpenGL is absent (nbody/simulation.py:16-17 imports it at module top although the
recorder never draws) and a ctypes ``zstandard`` shim over the system libzstd when the wheel is
absent (tools/record.py:228).
"""
from __future__ import annotations

import ctypes as C
import ctypes.util
import importlib
import os
import sys
import types


def _stub_opengl():
    try:
        import OpenGL  # noqa: F401
        return
    except Exception:
        pass
    ogl, gl = types.ModuleType("OpenGL"), types.ModuleType("OpenGL.GL")
    gl.__all__ = []
    arrays, vbo = types.ModuleType("OpenGL.arrays"), types.ModuleType("OpenGL.arrays.vbo")
    ogl.GL, ogl.arrays, arrays.vbo = gl, arrays, vbo
    sys.modules.update({"OpenGL": ogl, "OpenGL.GL": gl, "OpenGL.arrays": arrays, "OpenGL.arrays.vbo": vbo})


def _shim_zstandard():
    try:
        import zstandard  # noqa: F401
        return
    except Exception:
        pass
    path = ctypes.util.find_library("zstd") or "libzstd.so.1"
    z = C.CDLL(path)
    z.ZSTD_compressBound.restype = C.c_size_t
    z.ZSTD_compressBound.argtypes = [C.c_size_t]
    z.ZSTD_compress.restype = C.c_size_t
    z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    z.ZSTD_decompress.restype = C.c_size_t
    z.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    z.ZSTD_getFrameContentSize.restype = C.c_ulonglong
    z.ZSTD_getFrameContentSize.argtypes = [C.c_void_p, C.c_size_t]
    z.ZSTD_isError.restype = C.c_uint
    z.ZSTD_isError.argtypes = [C.c_size_t]

    class ZstdCompressor:
        def __init__(self, level=3, threads=0, **_kw):
            self.level = level

        def compress(self, data: bytes) -> bytes:
            data = bytes(data)
            cap = z.ZSTD_compressBound(len(data))
            buf = C.create_string_buffer(cap)
            n = z.ZSTD_compress(buf, cap, data, len(data), self.level)
            if z.ZSTD_isError(n):
                raise RuntimeError("ZSTD_compress failed")
            return buf.raw[:n]

    class ZstdDecompressor:
        def decompress(self, data: bytes, max_output_size: int = 0) -> bytes:
            data = bytes(data)
            size = z.ZSTD_getFrameContentSize(data, len(data))
            if size >= (1 << 62):
                size = max_output_size or 64 * len(data)
            buf = C.create_string_buffer(int(size) or 1)
            n = z.ZSTD_decompress(buf, int(size), data, len(data))
            if z.ZSTD_isError(n):
                raise RuntimeError("ZSTD_decompress failed")
            return buf.raw[:n]

    mod = types.ModuleType("zstandard")
    mod.ZstdCompressor, mod.ZstdDecompressor = ZstdCompressor, ZstdDecompressor
    mod.__b200_shim__ = True
    sys.modules["zstandard"] = mod


def install(reference_root: str, headless: bool = True):
    """Make the reference importable with the B200 backend in place of nbody/gpu_backend.py."""
    reference_root = os.path.abspath(reference_root)
    if not os.path.isdir(os.path.join(reference_root, "nbody")):
        raise FileNotFoundError(f"no reference checkout at {reference_root}")
    if headless:
        _stub_opengl()
        _shim_zstandard()
        os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/b200sim_numba_cache")
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    from .nbody import gpu_backend as ours
    ref_nbody = importlib.import_module("nbody")            # the reference's package (its __init__ runs)
    sys.modules["nbody.gpu_backend"] = ours
    setattr(ref_nbody, "gpu_backend", ours)
    return ours


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    root = os.environ.get("B200SIM_REFERENCE_ROOT", "/root/reference")
    seed = None
    while argv and argv[0] in ("--reference-root", "--seed"):
        if argv[0] == "--reference-root":
            root = argv[1]
        else:
            seed = int(argv[1])
        argv = argv[2:]
    if not argv:
        print(__doc__)
        return 2
    tool, rest = argv[0], argv[1:]
    install(root)
    if seed is not None:
        import numpy as np
        np.random.seed(seed)
    sys.argv = [tool] + rest
    if tool == "record":
        return importlib.import_module("tools.record").main()
    if tool == "nbody_main":
        return importlib.import_module("nbody_main").main()
    raise SystemExit(f"unknown tool {tool!r} (record | nbody_main)")


if __name__ == "__main__":
    sys.exit(main())
