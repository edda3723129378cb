"""ctypes binding of libb200sim.so (C ABI: include/b200sim.h).

There is no CPU fallback: if the library is missing or fails to load, importing a symbol
raises.  The library is built in-tree by build.py (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

N_PHASES = 8
PHASE_NAMES = ("keygen", "sort", "gather", "build", "extract", "traverse", "exchange", "integrate")


class NBodyStats(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("steps", C.c_int64), ("records", C.c_int64), ("interactions", C.c_int64),
        ("bounds", C.c_double), ("error_flags", C.c_uint32), ("sm_count", C.c_int32),
        ("bytes_allocated", C.c_int64), ("timed_steps", C.c_int64), ("phase_ms", C.c_double * N_PHASES),
        ("pair_records", C.c_int64), ("trav_pair_slots", C.c_int64), ("trav_lane_pairs", C.c_int64),
        ("trav_batches", C.c_int64), ("trav_stack_max", C.c_int64), ("trav_shared_pairs", C.c_int64),
        ("trav_kernel", C.c_int32), ("reserved0", C.c_int32), ("trav_sure_pairs", C.c_int64),
    ]


N_BOIDS_PHASES = 5
BOIDS_PHASE_NAMES = ("cells", "sort", "gather", "table", "rules")
BOIDS_PARAM_FIELDS = ("bounds", "max_speed", "max_force", "wall_margin", "wall_weight", "perception_radius",
                      "separation_radius", "separation_weight", "alignment_weight", "cohesion_weight",
                      "color_blend_rate")


class BoidsParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in BOIDS_PARAM_FIELDS]


class BoidsStats(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("steps", C.c_int64), ("num_cells", C.c_int64), ("grid_dim", C.c_int32),
        ("key_bits", C.c_int32), ("cell_size", C.c_double), ("grid_offset", C.c_double),
        ("neighbor_pairs", C.c_int64), ("bytes_allocated", C.c_int64), ("launches", C.c_int64),
        ("timed_steps", C.c_int64), ("phase_ms", C.c_double * N_BOIDS_PHASES),
    ]


class B200Error(RuntimeError):
    """Raised for every non-zero status of the C ABI (callers of the reference backend catch
    Exception around construction: nbody/simulation.py:533-540, tools/record.py:781-784)."""


_dp, _fp = C.POINTER(C.c_double), C.POINTER(C.c_float)
_h = C.c_void_p

# name -> (restype, argtypes); every entry must be declared in include/b200sim.h
SIGNATURES = {
    "b200_last_error": (C.c_char_p, []),
    "b200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "b200_device_info": (C.c_int, [C.c_int, C.c_char_p, C.c_int]),
    "b200_nbody_create": (C.c_int, [C.c_int64, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_double,
                                    C.c_int, C.POINTER(_h)]),
    "b200_nbody_create_multi": (C.c_int, [C.c_int64, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_double,
                                          C.c_uint32, C.POINTER(_h)]),
    "b200_generate_distribution": (C.c_int, [C.c_char_p, C.c_int64, C.c_double, C.c_double, C.c_uint64, C.c_int, _dp, _dp, _dp]),
    "b200_nbody_create_generated": (C.c_int, [C.c_char_p, C.c_int64, C.c_double, C.c_double, C.c_uint64, C.c_double, C.c_double,
                                              C.c_double, C.c_double, C.c_int, C.POINTER(_h)]),
    "b200_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "b200_nbody_comm_init": (C.c_int, [_h, C.c_void_p, C.c_int, C.c_int]),
    "b200_nbody_world": (C.c_int, [_h, C.POINTER(C.c_int)]),
    "b200_cost_weighted_split": (C.c_int, [C.POINTER(C.c_uint64), C.c_int, C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_int64)]),
    "b200_nbody_get_shard": (C.c_int, [_h, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "b200_nbody_destroy": (C.c_int, [_h]),
    "b200_nbody_step": (C.c_int, [_h, C.c_double]),
    "b200_nbody_step_n": (C.c_int, [_h, C.c_double, C.c_int]),
    "b200_nbody_compute_accelerations": (C.c_int, [_h, _fp]),
    "b200_nbody_compute_colors": (C.c_int, [_h, C.c_double]),
    "b200_nbody_get_positions": (C.c_int, [_h, _fp]),
    "b200_nbody_get_positions_f64": (C.c_int, [_h, _dp]),
    "b200_nbody_get_velocities": (C.c_int, [_h, _dp]),
    "b200_nbody_get_colors": (C.c_int, [_h, _fp]),
    "b200_nbody_sync": (C.c_int, [_h]),
    "b200_nbody_set_state": (C.c_int, [_h, _dp, _dp]),
    "b200_nbody_set_params": (C.c_int, [_h, C.c_double, C.c_double, C.c_double, C.c_double]),
    "b200_nbody_get_keys": (C.c_int, [_h, C.POINTER(C.c_uint64)]),
    "b200_nbody_get_perm": (C.c_int, [_h, C.POINTER(C.c_uint32)]),
    "b200_nbody_get_stats": (C.c_int, [_h, C.POINTER(NBodyStats)]),
    "b200_nbody_reset_stats": (C.c_int, [_h]),
    "b200_nbody_set_profiling": (C.c_int, [_h, C.c_int]),
    "b200_nbody_set_counting": (C.c_int, [_h, C.c_int]),
    "b200_nbody_count_interactions": (C.c_int, [_h, C.POINTER(C.c_int64)]),
    "b200_nbody_state_checksum": (C.c_int, [_h, C.POINTER(C.c_uint64)]),
    "b200_nbody_timed_steps": (C.c_int, [_h, C.c_double, C.c_int, C.POINTER(C.c_float)]),
    "b200_nbody_launch_count": (C.c_int, [_h, C.POINTER(C.c_int64)]),
    "b200_nbody_frame_begin": (C.c_int, [_h, C.c_double, _fp, _fp]),
    "b200_nbody_frame_wait": (C.c_int, [_h]),
    "b200_nbody_frame_delta_begin": (C.c_int, [_h, C.c_double, C.POINTER(C.c_int16), C.POINTER(C.c_int16)]),
    "b200_nbody_visible_frame": (C.c_int, [_h, C.c_double, _dp, _fp, _fp, C.POINTER(C.c_int64)]),
    "b200_nbody_visible_frame_device": (C.c_int, [_h, C.c_double, _dp, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]),
    "b200_nbody_set_state_begin": (C.c_int, [_h, _dp, _dp]),
    "b200_nbody_set_state_commit": (C.c_int, [_h]),
    "b200_nbody_set_state_begin_rows": (C.c_int, [_h, _dp, _dp, C.c_int64, C.c_int64]),
    "b200_nbody_upload_staging": (C.c_int, [_h, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "b200_nbody_upload_wait": (C.c_int, [_h]),
    "b200_nbody_frame_begin_rows": (C.c_int, [_h, C.c_double, _fp, _fp, C.c_int64, C.c_int64]),
    "b200_nbody_set_stream": (C.c_int, [_h, C.c_void_p, C.c_int]),
    "b200_nbody_set_shard": (C.c_int, [_h, C.c_int64, C.c_int64]),
    "b200_nbody_sharded_sort_setup": (C.c_int, [_h, C.c_int64, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "b200_nbody_sort_local": (C.c_int, [_h, C.c_int]),
    "b200_nbody_step_begin_sorted": (C.c_int, [_h]),
    "b200_nbody_step_begin": (C.c_int, [_h]),
    "b200_nbody_step_end": (C.c_int, [_h, C.c_double]),
    "b200_nbody_acc_buffer": (C.c_int, [_h, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "b200_fp32_peak_tflops": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "b200_host_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    "b200_host_free": (C.c_int, [C.c_void_p]),
    "b200_boids_create": (C.c_int, [C.c_int64, _dp, _dp, _dp, C.POINTER(BoidsParams), C.c_int, C.POINTER(_h)]),
    "b200_boids_destroy": (C.c_int, [_h]),
    "b200_boids_step": (C.c_int, [_h, C.c_double]),
    "b200_boids_get_state": (C.c_int, [_h, _dp, _dp, _dp]),
    "b200_boids_set_state": (C.c_int, [_h, _dp, _dp, _dp]),
    "b200_boids_get_cell_indices": (C.c_int, [_h, C.POINTER(C.c_int32)]),
    "b200_boids_get_stats": (C.c_int, [_h, C.POINTER(BoidsStats)]),
    "b200_boids_reset_stats": (C.c_int, [_h]),
    "b200_boids_set_profiling": (C.c_int, [_h, C.c_int]),
    "b200_boids_timed_steps": (C.c_int, [_h, C.c_double, C.c_int, C.POINTER(C.c_float)]),
    "b200_boids_sync": (C.c_int, [_h]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building first if the sources are newer) and type every entry point."""
    global _lib
    if _lib is None:
        path = os.environ.get("B200SIM_LIB") or _build.LIB_PATH     # B200SIM_LIB: an A/B build of the library (development)
        if path == _build.LIB_PATH and _build.is_stale():
            try:
                _build.build()
            except Exception as e:  # no nvcc on this box and no prebuilt library
                if not os.path.exists(path):
                    raise B200Error(f"libb200sim.so is not built and cannot be built here: {e}") from e
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)   # AttributeError if the library lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(status: int) -> None:
    if status != 0:
        msg = load().b200_last_error()
        raise B200Error(f"libb200sim status {status}: {msg.decode() if msg else 'unknown error'}")


def device_count() -> int:
    n = C.c_int(0)
    st = load().b200_device_count(C.byref(n))
    return int(n.value) if st == 0 else 0


def device_info(device: int = 0) -> str:
    buf = C.create_string_buffer(256)
    check(load().b200_device_info(device, buf, 256))
    return buf.value.decode()


def fp32_peak_tflops(device: int = 0) -> float:
    v = C.c_double(0.0)
    check(load().b200_fp32_peak_tflops(device, C.byref(v)))
    return float(v.value)


class _PinnedBlock:
    """Owner of one b200_host_alloc block; numpy arrays made by pinned_empty keep it alive through .base."""

    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        check(load().b200_host_alloc(int(nbytes), C.byref(self.ptr)))
        self.nbytes = int(nbytes)

    def __del__(self):
        ptr, self.ptr = getattr(self, "ptr", None), None
        if ptr and _lib is not None:
            _lib.b200_host_free(ptr)


def pinned_empty(shape, dtype):
    """numpy.empty in page-locked host memory (b200_host_alloc): the buffers to hand to frame_begin /
    frame_delta_begin / set_state_begin so that the copies really overlap the next step.  The memory is
    released when the last array viewing it is garbage-collected."""
    import numpy as np
    dt = np.dtype(dtype)
    shape = (int(shape),) if np.isscalar(shape) else tuple(int(x) for x in shape)
    count = 1
    for x in shape:
        count *= x
    block = _PinnedBlock(max(count * dt.itemsize, 1))
    raw = (C.c_char * block.nbytes).from_address(block.ptr.value)
    raw._b200_block = block            # the ctypes array is the numpy array's base object: ties the lifetimes
    return np.frombuffer(raw, dtype=dt, count=count).reshape(shape)
