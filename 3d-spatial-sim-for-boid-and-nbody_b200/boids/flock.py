"""Device twin of the reference's ``Flock`` update path (boids/flock.py:454-678).

The reference has no backend layer for boids: the seam is ``Flock.update(dt)`` mutating
``positions / velocities / colors`` (n,3) float64 in place (:627-678), read by ``draw`` (:716-726).
``B200Flock`` keeps those attributes and that method; the state lives on the GPU and is copied
back to the host arrays after each update when ``mirror_host`` is on (the viewer needs it every
frame; benchmarks turn it off and call ``get_state`` when they want it).

``attach(flock)`` patches a live reference ``Flock`` instance so its ``update`` runs on the device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from ..config import boids as _cfg


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError(f"expected (n,3) array, got {a.shape}")
    return a


def make_params(overrides: dict | None = None) -> _lib.BoidsParams:
    d = {k: float(_cfg.BOIDS[k]) for k in _lib.BOIDS_PARAM_FIELDS}   # config/boids.py:30-46
    if overrides:
        for k, v in overrides.items():
            if k in d:
                d[k] = float(v)
    return _lib.BoidsParams(**d)


class B200Flock:
    def __init__(self, positions, velocities, colors, params: dict | None = None, device: int = 0,
                 mirror_host: bool = False):
        L = _lib.load()
        self.positions, self.velocities, self.colors = _f64(positions).copy(), _f64(velocities).copy(), _f64(colors).copy()
        if not (len(self.positions) == len(self.velocities) == len(self.colors)):
            raise ValueError("positions, velocities and colors disagree on n")
        self.num_boids = len(self.positions)
        self.mirror_host = mirror_host
        self._params = make_params(params)
        for k in _lib.BOIDS_PARAM_FIELDS:
            setattr(self, k, getattr(self._params, k))
        # boids/flock.py:478-481
        self.cell_size = float(self.perception_radius)
        self.grid_dim = int(np.ceil(self.bounds * 2 / self.cell_size)) + 2
        self.num_cells = self.grid_dim ** 3
        self.grid_offset = float(self.bounds + self.cell_size)
        self._L, self._h = L, C.c_void_p()
        dp = C.POINTER(C.c_double)
        _lib.check(L.b200_boids_create(self.num_boids, self.positions.ctypes.data_as(dp),
                                       self.velocities.ctypes.data_as(dp), self.colors.ctypes.data_as(dp),
                                       C.byref(self._params), int(device), C.byref(self._h)))

    @classmethod
    def random(cls, num_boids: int, seed: int = 0, params: dict | None = None, **kw):
        """Initial state with the laws of Flock.__init__ (boids/flock.py:488-490, :587-608),
        seeded (the reference draws from the unseeded global RandomState)."""
        p = make_params(params)
        rng = np.random.default_rng(seed)
        pos = (rng.random((num_boids, 3)) - 0.5) * 2 * p.bounds
        vel = (rng.random((num_boids, 3)) - 0.5) * p.max_speed
        hues = np.linspace(0, 1, num_boids, endpoint=False)
        rng.shuffle(hues)
        h6 = hues * 6.0
        i = h6.astype(np.int64) % 6
        f = h6 - np.floor(h6)
        s, v = 0.9, 1.0
        pp, q, t = v * (1 - s), v * (1 - s * f), v * (1 - s * (1 - f))
        vv = np.full(num_boids, v)
        pv = np.full(num_boids, pp)
        table = [(vv, t, pv), (q, vv, pv), (pv, vv, t), (pv, q, vv), (t, pv, vv), (vv, pv, q)]
        col = np.zeros((num_boids, 3))
        for idx, (r, g, b) in enumerate(table):
            m = i == idx
            col[m, 0], col[m, 1], col[m, 2] = r[m], g[m], b[m]
        return cls(pos, vel, col, params=params, **kw)

    # ---- the reference's method ---------------------------------------------------------------
    def update(self, dt: float):
        """Flock.update (boids/flock.py:627-678) on the device."""
        _lib.check(self._L.b200_boids_step(self._handle(), float(dt)))
        if self.mirror_host:
            self.get_state(out=(self.positions, self.velocities, self.colors))

    # ---- additions ----------------------------------------------------------------------------
    def get_state(self, out=None):
        if out is None:
            out = tuple(np.empty((self.num_boids, 3), np.float64) for _ in range(3))
        dp = C.POINTER(C.c_double)
        _lib.check(self._L.b200_boids_get_state(self._handle(), *(a.ctypes.data_as(dp) for a in out)))
        return out

    def set_state(self, positions, velocities, colors):
        p, v, c = _f64(positions), _f64(velocities), _f64(colors)
        if not (len(p) == len(v) == len(c) == self.num_boids):
            raise ValueError("set_state: n differs")
        dp = C.POINTER(C.c_double)
        _lib.check(self._L.b200_boids_set_state(self._handle(), p.ctypes.data_as(dp), v.ctypes.data_as(dp),
                                                c.ctypes.data_as(dp)))

    def get_cell_indices(self) -> np.ndarray:
        out = np.empty(self.num_boids, np.int32)
        _lib.check(self._L.b200_boids_get_cell_indices(self._handle(), out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def timed_steps(self, dt: float, nsteps: int) -> float:
        ms = C.c_float(0.0)
        _lib.check(self._L.b200_boids_timed_steps(self._handle(), float(dt), int(nsteps), C.byref(ms)))
        return float(ms.value)

    def set_profiling(self, enabled: bool):
        _lib.check(self._L.b200_boids_set_profiling(self._handle(), int(bool(enabled))))

    def reset_stats(self):
        _lib.check(self._L.b200_boids_reset_stats(self._handle()))

    def get_stats(self) -> dict:
        st = _lib.BoidsStats()
        _lib.check(self._L.b200_boids_get_stats(self._handle(), C.byref(st)))
        d = {k: getattr(st, k) for k, _ in st._fields_ if k != "phase_ms"}
        d["phase_ms"] = {name: st.phase_ms[i] for i, name in enumerate(_lib.BOIDS_PHASE_NAMES)}
        return d

    def sync(self):
        _lib.check(self._L.b200_boids_sync(self._handle()))

    def _handle(self):
        if not self._h:
            raise _lib.B200Error("flock is closed")
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            self._L.b200_boids_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def attach(flock, device: int = 0) -> B200Flock:
    """Patch a live reference ``Flock`` so ``flock.update(dt)`` runs on the GPU and its
    ``positions/velocities/colors`` arrays are refreshed in place (what draw() reads)."""
    params = {k: float(getattr(flock, k)) for k in _lib.BOIDS_PARAM_FIELDS if hasattr(flock, k)}
    dev = B200Flock(flock.positions, flock.velocities, flock.colors, params=params, device=device)

    def update(dt):
        _lib.check(dev._L.b200_boids_step(dev._handle(), float(dt)))
        dev.get_state(out=(flock.positions, flock.velocities, flock.colors))

    flock.update = update
    flock._b200 = dev
    return dev
