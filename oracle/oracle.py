"""ctypes front end of the CPU oracle (oracle/bh_oracle.c, oracle/boids_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Parity pinning: checked against fixtures generated from the reference itself
(tests/golden/make_golden.py -> tests/golden/*.npz, tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

REFERENCE_MAX_TREE_NODES = 8_000_000  # nbody/simulation.py:35


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc, OpenMP)."""
    srcs = [os.path.join(_HERE, f) for f in ("bh_oracle.c", "boids_oracle.c", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip, i64p, u8p, fp = (C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_uint8), C.POINTER(C.c_float))
        L.orc_compute_bounds.restype = C.c_double
        L.orc_compute_bounds.argtypes = [dp, C.c_int64]
        L.orc_build_octree.restype = C.c_int64
        L.orc_build_octree.argtypes = [dp, dp, C.c_int64, C.c_double, C.c_int64, C.c_int,
                                       dp, dp, dp, dp, ip, ip, u8p]
        L.orc_compute_forces.restype = None
        L.orc_compute_forces.argtypes = [dp, dp, dp, dp, dp, ip, ip, u8p, C.c_int64, i64p, C.c_int64,
                                         C.c_double, C.c_double, C.c_double, C.c_int,
                                         i64p, i64p, ip, i64p]
        L.orc_update.restype = None
        L.orc_update.argtypes = [dp, dp, dp, C.c_double, C.c_double, C.c_int64]
        L.orc_colors.restype = None
        L.orc_colors.argtypes = [dp, fp, C.c_int64, C.c_double]
        L.orc_direct_sum.restype = None
        L.orc_direct_sum.argtypes = [dp, dp, C.c_int64, i64p, C.c_int64, C.c_double, C.c_double, dp]
        L.orc_morton_keys.restype = None
        L.orc_morton_keys.argtypes = [dp, C.c_int64, C.c_double, C.POINTER(C.c_uint64)]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_boids_assign_cells.restype = None
        L.orc_boids_assign_cells.argtypes = [dp, ip, C.c_double, C.c_int32, C.c_double, C.c_int64]
        L.orc_boids_sort_and_lists.restype = None
        L.orc_boids_sort_and_lists.argtypes = [ip, ip, ip, ip, C.c_int64, C.c_int64]
        L.orc_boids_flocking.restype = None
        L.orc_boids_flocking.argtypes = [dp, dp, dp, ip, ip, ip, dp, dp, dp, dp,
                                         C.c_double, C.c_int32, C.c_double, C.c_double, C.c_double,
                                         C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                         C.c_int64, ip]
        L.orc_boids_physics.restype = None
        L.orc_boids_physics.argtypes = [dp, dp, dp, dp, dp, dp, dp, C.c_double, C.c_double, C.c_double,
                                        C.c_double, C.c_double, C.c_double, C.c_int64]
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(t: int) -> None:
    lib().orc_set_num_threads(int(t))


def compute_bounds(pos) -> float:
    pos = _f64(pos)
    return float(lib().orc_compute_bounds(_p(pos, C.c_double), len(pos)))


class Octree:
    """Node arrays with the reference's names and dtypes (nbody/simulation.py:69-76)."""

    def __init__(self, max_nodes: int):
        self.max_nodes = int(max_nodes)
        self.node_centers = np.zeros((max_nodes, 3), np.float64)
        self.node_half_sizes = np.zeros(max_nodes, np.float64)
        self.node_masses = np.zeros(max_nodes, np.float64)
        self.node_com = np.zeros((max_nodes, 3), np.float64)
        self.node_children = np.full((max_nodes, 8), -1, np.int32)
        self.node_body_idx = np.full(max_nodes, -1, np.int32)
        self.node_is_leaf = np.ones(max_nodes, np.uint8)
        self.num_nodes = 0
        self.bounds = 0.0


def build_octree(pos, mass, bounds: float | None = None, max_nodes: int | None = None,
                 reference_cap: bool = False, max_depth: int = 0) -> Octree:
    """Sequential-insertion octree.  reference_cap=True reproduces the reference callers'
    allocation min(8M, 4N) and the MAX_TREE_NODES truncation; otherwise the pool is sized so
    the tree is never truncated (grown and rebuilt if needed)."""
    pos, mass = _f64(pos), _f64(mass)
    n = len(pos)
    if bounds is None:
        bounds = compute_bounds(pos)
    if reference_cap:
        alloc = min(REFERENCE_MAX_TREE_NODES, 4 * n) if max_nodes is None else max_nodes
        cap = REFERENCE_MAX_TREE_NODES
    else:
        alloc = max(64, 3 * n) if max_nodes is None else max_nodes
        cap = alloc
    while True:
        t = Octree(alloc)
        t.bounds = float(bounds)
        t.num_nodes = int(lib().orc_build_octree(
            _p(pos, C.c_double), _p(mass, C.c_double), n, bounds, min(cap, alloc), max_depth,
            _p(t.node_centers, C.c_double), _p(t.node_half_sizes, C.c_double),
            _p(t.node_masses, C.c_double), _p(t.node_com, C.c_double),
            _p(t.node_children, C.c_int32), _p(t.node_body_idx, C.c_int32),
            _p(t.node_is_leaf, C.c_uint8)))
        if reference_cap or t.num_nodes < alloc - 1:
            return t
        alloc *= 2
        cap = alloc


def compute_forces(pos, tree: Octree, theta: float, G: float, softening: float,
                   targets=None, stack_cap: int = 0, stats: dict | None = None):
    """Barnes-Hut accelerations (fp64).  stack_cap=64 reproduces the reference's fixed stack."""
    pos = _f64(pos)
    if targets is None:
        nt, tp = len(pos), None
    else:
        targets = np.ascontiguousarray(targets, np.int64)
        nt, tp = len(targets), _p(targets, C.c_int64)
    acc = np.zeros((nt, 3), np.float64)
    inter, visits, drops = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    peak = C.c_int32(0)
    lib().orc_compute_forces(
        _p(pos, C.c_double), _p(acc, C.c_double), _p(tree.node_half_sizes, C.c_double),
        _p(tree.node_masses, C.c_double), _p(tree.node_com, C.c_double),
        _p(tree.node_children, C.c_int32), _p(tree.node_body_idx, C.c_int32),
        _p(tree.node_is_leaf, C.c_uint8), tree.num_nodes, tp, nt, theta, G, softening, stack_cap,
        C.byref(inter), C.byref(visits), C.byref(peak), C.byref(drops))
    if stats is not None:
        stats.update(interactions=inter.value, visits=visits.value, peak_stack=peak.value,
                     drops=drops.value)
    return acc


def update(pos, vel, acc, damping: float, dt: float) -> None:
    """In-place kick-drift on contiguous fp64 arrays."""
    for a in (pos, vel, acc):
        assert a.dtype == np.float64 and a.flags.c_contiguous
    lib().orc_update(_p(pos, C.c_double), _p(vel, C.c_double), _p(acc, C.c_double), damping, dt, len(pos))


def colors(vel, max_speed: float):
    vel = _f64(vel)
    out = np.zeros((len(vel), 3), np.float32)
    lib().orc_colors(_p(vel, C.c_double), _p(out, C.c_float), len(vel), max_speed)
    return out


def visibility_mask(positions, cam_pos, cam_forward, cam_right, cam_up, tan_h: float, tan_v: float, far_dist: float):
    """compute_visibility_points (nbody/simulation.py:403-434), operation by operation in fp64 (numpy does not
    contract): positions are what the viewer holds, i.e. the backend's float32 positions widened to float64 (:816)."""
    p = np.asarray(positions, np.float64)
    cp, f, r, u = (np.asarray(a, np.float64) for a in (cam_pos, cam_forward, cam_right, cam_up))
    dx, dy, dz = p[:, 0] - cp[0], p[:, 1] - cp[1], p[:, 2] - cp[2]
    z = dx * f[0] + dy * f[1] + dz * f[2]
    x = dx * r[0] + dy * r[1] + dz * r[2]
    y = dx * u[0] + dy * u[1] + dz * u[2]
    hw, hh = z * tan_h * 1.2, z * tan_v * 1.2
    return ~((z < 0.1) | (z > far_dist)) & (np.abs(x) < hw) & (np.abs(y) < hh)


def direct_sum(pos, mass, G: float, softening: float, targets=None):
    pos, mass = _f64(pos), _f64(mass)
    if targets is None:
        nt, tp = len(pos), None
    else:
        targets = np.ascontiguousarray(targets, np.int64)
        nt, tp = len(targets), _p(targets, C.c_int64)
    acc = np.zeros((nt, 3), np.float64)
    lib().orc_direct_sum(_p(pos, C.c_double), _p(mass, C.c_double), len(pos), tp, nt, G, softening,
                         _p(acc, C.c_double))
    return acc


def morton_keys(pos, bounds: float | None = None):
    pos = _f64(pos)
    if bounds is None:
        bounds = compute_bounds(pos)
    keys = np.zeros(len(pos), np.uint64)
    lib().orc_morton_keys(_p(pos, C.c_double), len(pos), bounds, _p(keys, C.c_uint64))
    return keys


def sort_permutation(keys):
    """Stable sort: ties broken by original body index."""
    return np.argsort(keys, kind="stable").astype(np.uint32)


def nbody_step(pos, vel, mass, theta, G, softening, damping, dt, reference_cap=False, stack_cap=0,
               stats: dict | None = None):
    """One substep exactly as tools/record.py:835-858 sequences it.  In-place on pos/vel."""
    bounds = compute_bounds(pos)
    tree = build_octree(pos, mass, bounds, reference_cap=reference_cap)
    acc = compute_forces(pos, tree, theta, G, softening, stack_cap=stack_cap, stats=stats)
    update(pos, vel, acc, damping, dt)
    if stats is not None:
        stats.update(num_nodes=tree.num_nodes, bounds=bounds)
    return acc


# ----------------------------------------------------------------------------- boids

BOIDS_DEFAULTS = dict(  # config/boids.py:30-46
    bounds=500.0, max_speed=25.0, max_force=60.0, wall_margin=3.0, wall_weight=10.0,
    perception_radius=5.0, separation_radius=3.0, separation_weight=2.5,
    alignment_weight=1.0, cohesion_weight=1.0, color_blend_rate=1.0)


def boids_grid(params: dict):
    """boids/flock.py:478-481"""
    cell = float(params["perception_radius"])
    dim = int(np.ceil(params["bounds"] * 2 / cell)) + 2
    return cell, dim, float(params["bounds"] + cell)


def boids_step(pos, vel, col, dt: float, params: dict | None = None, neighbor_counts=None) -> None:
    """Flock.update (boids/flock.py:627-678) on contiguous fp64 (n,3) arrays, in place."""
    p = dict(BOIDS_DEFAULTS)
    if params:
        p.update(params)
    for a in (pos, vel, col):
        assert a.dtype == np.float64 and a.flags.c_contiguous
    n = len(pos)
    cell, dim, offset = boids_grid(p)
    ncell = dim ** 3
    L = lib()
    cell_idx = np.zeros(n, np.int32)
    sorted_idx = np.zeros(n, np.int32)
    starts = np.zeros(ncell, np.int32)
    counts = np.zeros(ncell, np.int32)
    L.orc_boids_assign_cells(_p(pos, C.c_double), _p(cell_idx, C.c_int32), cell, dim, offset, n)
    L.orc_boids_sort_and_lists(_p(cell_idx, C.c_int32), _p(sorted_idx, C.c_int32),
                               _p(starts, C.c_int32), _p(counts, C.c_int32), n, ncell)
    sepf = np.zeros((n, 3)); alif = np.zeros((n, 3)); cohf = np.zeros((n, 3))
    avgc = col.copy()
    ncp = None
    if neighbor_counts is not None:
        assert neighbor_counts.dtype == np.int32 and len(neighbor_counts) == n
        ncp = _p(neighbor_counts, C.c_int32)
    L.orc_boids_flocking(_p(pos, C.c_double), _p(vel, C.c_double), _p(col, C.c_double),
                         _p(sorted_idx, C.c_int32), _p(starts, C.c_int32), _p(counts, C.c_int32),
                         _p(sepf, C.c_double), _p(alif, C.c_double), _p(cohf, C.c_double), _p(avgc, C.c_double),
                         cell, dim, offset, p["perception_radius"], p["separation_radius"],
                         p["separation_weight"], p["alignment_weight"], p["cohesion_weight"],
                         p["max_speed"], p["max_force"], n, ncp)
    blend = min(1.0, float(p["color_blend_rate"]) * float(dt))
    L.orc_boids_physics(_p(pos, C.c_double), _p(vel, C.c_double), _p(col, C.c_double),
                        _p(sepf, C.c_double), _p(alif, C.c_double), _p(cohf, C.c_double), _p(avgc, C.c_double),
                        p["bounds"], p["wall_margin"], p["max_force"] * p["wall_weight"], p["max_speed"],
                        blend, float(dt), n)
