"""Frame codec of the reference's recorder: the `frame_NNNN.zstd` files that
tools.playback / tools.export read.

Format (tools/record.py:231-279): 1 byte tag (1 = absolute float32, 2 = int16 delta x 1000 against the
previous frame) + u32 length + zstd(positions payload) + u32 length + zstd(colours payload).  The
reference computes the format-2 payload on the host from two float32 frames; here it can come straight
from the device (`B200BarnesHutSimulation.frame_delta_begin`, 12 instead of 24 B/body over PCIe) and
the two zstd streams of a frame are compressed concurrently (ctypes releases the GIL), optionally with
libzstd's own worker threads.  Same function names and argument meaning as the reference's
`compress_frame` / `decompress_frame` / `load_frame` so a maintainer can swap the import.

zstd itself is the system libzstd (bound with ctypes; the image has no `zstandard` package).  The
compressed bytes are standard zstd frames: any zstd (the reference's `zstandard` included) reads them.
"""
from __future__ import annotations

import ctypes as C
import ctypes.util
import os
import struct
import threading
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

DEFAULT_LEVEL = 19   # the reference's level (tools/record.py:252)
_lock = threading.Lock()
_zstd = None


def _lib():
    global _zstd
    with _lock:
        if _zstd is None:
            name = ctypes.util.find_library("zstd") or "libzstd.so.1"
            z = C.CDLL(name)
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
            z.ZSTD_getErrorName.restype = C.c_char_p
            z.ZSTD_getErrorName.argtypes = [C.c_size_t]
            z.ZSTD_versionString.restype = C.c_char_p
            z.ZSTD_createCCtx.restype = C.c_void_p
            z.ZSTD_freeCCtx.argtypes = [C.c_void_p]
            z.ZSTD_CCtx_setParameter.restype = C.c_size_t
            z.ZSTD_CCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
            z.ZSTD_compress2.restype = C.c_size_t
            z.ZSTD_compress2.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
            _zstd = z
    return _zstd


def zstd_version() -> str:
    return _lib().ZSTD_versionString().decode()


def _check(code: int) -> int:
    z = _lib()
    if z.ZSTD_isError(code):
        raise RuntimeError("zstd: " + z.ZSTD_getErrorName(code).decode())
    return code


_ZSTD_c_compressionLevel, _ZSTD_c_nbWorkers = 100, 400


def zstd_compress(data, level: int = DEFAULT_LEVEL, workers: int = 0) -> bytes:
    """One zstd frame.  workers > 0 uses libzstd's multi-threaded compressor (different bytes, same
    content); workers = 0 is what `zstandard.ZstdCompressor(level, threads=1).compress` produces."""
    z = _lib()
    buf = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else data.reshape(-1).view(np.uint8)
    n = buf.size
    cap = z.ZSTD_compressBound(n)
    dst = np.empty(cap, np.uint8)
    src = buf.ctypes.data if n else None
    if workers <= 0:
        size = _check(z.ZSTD_compress(dst.ctypes.data, cap, src, n, level))
    else:
        ctx = z.ZSTD_createCCtx()
        try:
            _check(z.ZSTD_CCtx_setParameter(ctx, _ZSTD_c_compressionLevel, level))
            if z.ZSTD_isError(z.ZSTD_CCtx_setParameter(ctx, _ZSTD_c_nbWorkers, workers)):
                pass   # single-threaded libzstd build: compress on this thread
            size = _check(z.ZSTD_compress2(ctx, dst.ctypes.data, cap, src, n))
        finally:
            z.ZSTD_freeCCtx(ctx)
    return dst[:size].tobytes()


def zstd_decompress(data: bytes) -> bytes:
    z = _lib()
    src = np.frombuffer(data, np.uint8)
    size = z.ZSTD_getFrameContentSize(src.ctypes.data, src.size)
    if size in (2 ** 64 - 1, 2 ** 64 - 2):
        raise RuntimeError("zstd: frame without a content size")
    dst = np.empty(int(size), np.uint8)
    got = _check(z.ZSTD_decompress(dst.ctypes.data if size else None, int(size), src.ctypes.data, src.size))
    return dst[:got].tobytes()


# ----------------------------------------------------------------------------- payloads
def delta_payload(frame: np.ndarray, prev: np.ndarray) -> np.ndarray:
    """int16((frame - prev) * 1000) on float32 arrays: tools/record.py:256-262 (what the device kernel
    `frame_delta_kernel` reproduces bit for bit for in-range deltas)."""
    with np.errstate(invalid="ignore"):
        return ((np.asarray(frame, np.float32) - np.asarray(prev, np.float32)) * 1000).astype(np.int16)


def pack_frame(comp_format: int, pos_payload, col_payload, level: int = DEFAULT_LEVEL, workers: int = 0,
               pool: ThreadPoolExecutor | None = None) -> bytes:
    """tag + u32 + zstd(positions) + u32 + zstd(colours) (tools/record.py:270-277).  With a pool the two
    streams are compressed concurrently."""
    if pool is not None:
        fp = pool.submit(zstd_compress, pos_payload, level, workers)
        fc = pool.submit(zstd_compress, col_payload, level, workers)
        pc, cc = fp.result(), fc.result()
    else:
        pc, cc = zstd_compress(pos_payload, level, workers), zstd_compress(col_payload, level, workers)
    return struct.pack("B", comp_format) + struct.pack("I", len(pc)) + pc + struct.pack("I", len(cc)) + cc


def compress_frame(positions: np.ndarray, colors: np.ndarray, prev_positions: np.ndarray = None,
                   prev_colors: np.ndarray = None, level: int = DEFAULT_LEVEL, workers: int = 0, pool=None) -> bytes:
    """Same signature and bytes layout as the reference's compress_frame (tools/record.py:234-279)."""
    if prev_positions is not None and prev_colors is not None:
        return pack_frame(2, delta_payload(positions, prev_positions), delta_payload(colors, prev_colors), level, workers, pool)
    return pack_frame(1, np.ascontiguousarray(positions, np.float32), np.ascontiguousarray(colors, np.float32), level, workers, pool)


def compress_delta_frame(pos_delta_i16: np.ndarray, col_delta_i16: np.ndarray, level: int = DEFAULT_LEVEL, workers: int = 0,
                         pool=None) -> bytes:
    """Format-2 frame from device-produced int16 deltas (frame_delta_begin)."""
    if pos_delta_i16.dtype != np.int16 or col_delta_i16.dtype != np.int16:
        raise ValueError("delta payloads must be int16")
    return pack_frame(2, np.ascontiguousarray(pos_delta_i16), np.ascontiguousarray(col_delta_i16), level, workers, pool)


def decompress_frame(data: bytes, prev_positions: np.ndarray = None, prev_colors: np.ndarray = None):
    """tools/record.py:282-326: -> (positions, colors) float32 (n,3)."""
    if len(data) < 1:
        raise ValueError("Invalid compressed data")
    comp_format = data[0]
    off = 1
    (ps,) = struct.unpack("I", data[off:off + 4]); off += 4
    pos_c = data[off:off + ps]; off += ps
    (cs,) = struct.unpack("I", data[off:off + 4]); off += 4
    col_c = data[off:off + cs]
    pos_data, col_data = zstd_decompress(pos_c), zstd_decompress(col_c)
    if comp_format == 1:
        return (np.frombuffer(pos_data, np.float32).reshape(-1, 3), np.frombuffer(col_data, np.float32).reshape(-1, 3))
    if comp_format == 2:
        if prev_positions is None or prev_colors is None:
            raise ValueError("Delta compression requires previous frame")
        pd = np.frombuffer(pos_data, np.int16).reshape(-1, 3).astype(np.float32) / 1000.0
        cd = np.frombuffer(col_data, np.int16).reshape(-1, 3).astype(np.float32) / 1000.0
        return prev_positions + pd, prev_colors + cd
    raise ValueError(f"Unknown compression format: {comp_format}")


def frame_path(rec_dir, frame_idx: int) -> Path:
    return Path(rec_dir) / f"frame_{frame_idx:04d}.zstd"


def load_frame(rec_dir, frame_idx: int, prev_positions=None, prev_colors=None):
    """Reads frame_NNNN.zstd (or the recorder's uncompressed frame_NNNN.npz).  A delta frame asked for
    without its predecessor walks back to the last frame that stands alone and decodes forward, as the
    reference does iteratively (tools/record.py:99-210)."""
    rec_dir = Path(rec_dir)
    z, npz = frame_path(rec_dir, frame_idx), rec_dir / f"frame_{frame_idx:04d}.npz"
    if z.exists():
        data = z.read_bytes()
        if len(data) > 0 and data[0] == 2 and (prev_positions is None or prev_colors is None):
            if frame_idx == 0:
                raise ValueError(f"Frame {frame_idx:04d} appears to be delta-compressed but is the first frame")
            chain, base, i = [], None, frame_idx - 1
            while i >= 0:
                pz, pn = frame_path(rec_dir, i), rec_dir / f"frame_{i:04d}.npz"
                if pz.exists():
                    d = pz.read_bytes()
                    if not d:
                        break
                    chain.append(d)
                    if d[0] == 1:
                        break
                    i -= 1
                elif pn.exists():
                    with np.load(pn) as f:
                        base = (f["positions"].copy(), f["colors"].copy())
                    break
                else:
                    raise FileNotFoundError(f"Frame {i:04d} not found (needed for delta decompression)")
            if base is None:
                if not chain or chain[-1][0] != 1:
                    raise ValueError(f"Frame {frame_idx:04d} appears to be delta-compressed but no base frame (format 1) found")
                base = decompress_frame(chain.pop(), None, None)
            p, c = base
            for d in reversed(chain):
                p, c = decompress_frame(d, p, c)
            prev_positions, prev_colors = p, c
        return decompress_frame(data, prev_positions, prev_colors)
    if npz.exists():
        with np.load(npz) as f:
            return f["positions"].copy(), f["colors"].copy()
    raise FileNotFoundError(f"Frame {frame_idx:04d} not found")


class FrameWriter:
    """Writes a recording's frames in the reference's compressed format while the simulation keeps
    stepping: `submit` returns at once, compression (two streams per frame in parallel) and file IO run
    on worker threads -- the role of the reference's BackgroundCompressor (tools/record.py:329-490)
    without the uncompressed round trip through the disk."""

    def __init__(self, rec_dir, level: int = DEFAULT_LEVEL, threads: int = 4, zstd_workers: int = 0):
        self.rec_dir = Path(rec_dir)
        self.rec_dir.mkdir(parents=True, exist_ok=True)
        self.level, self.zstd_workers = level, zstd_workers
        self._streams = ThreadPoolExecutor(max(2, threads))
        self._frames = ThreadPoolExecutor(max(1, threads // 2))
        self._pending = []
        self.bytes_in = self.bytes_out = 0

    def _write(self, frame_idx: int, comp_format: int, pos_payload, col_payload):
        data = pack_frame(comp_format, pos_payload, col_payload, self.level, self.zstd_workers, self._streams)
        tmp = self.rec_dir / f".frame_{frame_idx:04d}.zstd.tmp"
        tmp.write_bytes(data)
        os.replace(tmp, frame_path(self.rec_dir, frame_idx))
        return pos_payload.nbytes + col_payload.nbytes, len(data)

    def submit_absolute(self, frame_idx: int, positions: np.ndarray, colors: np.ndarray):
        """Format 1.  The arrays are copied: the caller may reuse its (pinned) buffers at once."""
        p, c = np.array(positions, np.float32, copy=True), np.array(colors, np.float32, copy=True)
        self._pending.append(self._frames.submit(self._write, frame_idx, 1, p, c))

    def submit_delta(self, frame_idx: int, pos_delta_i16: np.ndarray, col_delta_i16: np.ndarray):
        """Format 2 from the device-produced deltas."""
        p, c = np.array(pos_delta_i16, np.int16, copy=True), np.array(col_delta_i16, np.int16, copy=True)
        self._pending.append(self._frames.submit(self._write, frame_idx, 2, p, c))

    def flush(self):
        for f in self._pending:
            i, o = f.result()
            self.bytes_in += i
            self.bytes_out += o
        self._pending = []

    def close(self):
        self.flush()
        self._frames.shutdown()
        self._streams.shutdown()
