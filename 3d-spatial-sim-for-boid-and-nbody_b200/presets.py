"""Seeded synthetic initial conditions and the preset parameters the benchmarks are quoted on.

The reference's generators (tools/presets.py:91-1390) draw from numpy's unseeded global
RandomState and some loop per body in Python (e.g. cluster :380-393), which is unusable at
50 M bodies.  These are vectorised, seeded restatements of the three distributions named in
BASELINE.json -- the same density and velocity laws, not the same random stream -- used as
*synthetic inputs* by bench.py and the GPU tests.  The reference's own generators remain the
ones its tools call; this module replaces nothing on the drop-in path.

Preset parameter dicts mirror the reference's PRESETS entries (tools/presets.py:1397-2642):
same keys, same values.
"""
from __future__ import annotations

import numpy as np

# tools/presets.py:1552-1568, :1774-1790, :2478-2494, :2590-2642, :1868-1884
PRESETS = {
    "tiny_galaxy": dict(num_bodies=10_000, theta=0.95, G=0.2, softening=5.0, damping=1.0, spawn_radius=200.0,
                        distribution="galaxy", dt_per_frame=0.3, substeps=1),
    "tiny_collision": dict(num_bodies=15_000, theta=0.95, G=0.25, softening=5.0, damping=1.0, spawn_radius=250.0,
                           distribution="collision", dt_per_frame=0.3, substeps=1),
    "demo_cluster": dict(num_bodies=20_000, theta=0.95, G=0.15, softening=3.0, damping=1.0, spawn_radius=150.0,
                         distribution="cluster", dt_per_frame=0.2, substeps=1),
    "quick_galaxy": dict(num_bodies=100_000, theta=0.95, G=0.15, softening=3.0, damping=1.0, spawn_radius=500.0,
                         distribution="galaxy", dt_per_frame=0.2, substeps=1),
    "accurate_cluster_100k": dict(num_bodies=100_000, theta=0.5, G=0.05, softening=1.0, damping=1.0,
                                  spawn_radius=300.0, distribution="cluster", dt_per_frame=0.05, substeps=1),
    "4k_collision_1m": dict(num_bodies=1_000_000, theta=0.5, G=0.08, softening=1.5, damping=1.0, spawn_radius=900.0,
                            distribution="collision", dt_per_frame=0.06, substeps=5),
    # BASELINE.json config 5: extreme_50m_galaxy with the theta override 0.7
    "extreme_50m_galaxy_t07": dict(num_bodies=50_000_000, theta=0.7, G=0.04, softening=10.0, damping=1.0,
                                   spawn_radius=3000.0, distribution="galaxy", dt_per_frame=0.35, substeps=1),
}


def get_preset_config(key: str) -> dict:
    p = dict(PRESETS[key])
    p["session_name"] = key
    p["dt"] = p["dt_per_frame"] / p["substeps"]   # tools/record.py:749
    return p


def _rotation_curve(r, G, softening):
    """Softened enclosed-mass circular speed for unit masses (law of tools/presets.py:52-88)."""
    order = np.argsort(r, kind="stable")
    rs = r[order]
    m_enc = np.arange(1, len(r) + 1, dtype=np.float64)
    eps = 2.0 * softening
    r2 = rs * rs
    v = np.sqrt(G * m_enc * r2 / (r2 + eps * eps) ** 1.5)
    v *= np.maximum(r2 / (r2 + eps * eps), 0.3)
    out = np.empty_like(v)
    out[order] = v
    return out


def _disk(rng, n, R, G, scale_frac, cap_frac, height_frac, disp, spin):
    """Exponential disk with soft truncation in the XZ plane (laws of tools/presets.py:104-146)."""
    soft = R * (0.03 if cap_frac >= 1.0 else 0.025)
    r = rng.exponential(R * scale_frac, n)
    r = r * (1.0 - np.exp(-(R * cap_frac) / (r + 0.01)))
    r = np.maximum(r, R * 0.001)
    th = rng.uniform(0.0, 2.0 * np.pi, n)
    h = R * height_frac * (1.0 + np.sqrt(r / R) * 0.3)
    pos = np.empty((n, 3))
    pos[:, 0] = r * np.cos(th)
    pos[:, 1] = rng.normal(0.0, 1.0, n) * h
    pos[:, 2] = r * np.sin(th)
    vc = _rotation_curve(r, G, soft)
    vel = np.empty((n, 3))
    vel[:, 0] = -spin * vc * np.sin(th)
    vel[:, 2] = spin * vc * np.cos(th)
    sigma = vc * disp * (r / (r + 2.0 * soft)) + np.sqrt(G * n * 0.00005)
    vel[:, 0] += rng.normal(0.0, 1.0, n) * sigma
    vel[:, 2] += rng.normal(0.0, 1.0, n) * sigma
    vel[:, 1] = rng.normal(0.0, 1.0, n) * sigma * 0.25
    return pos, vel


def generate(distribution: str, n: int, R: float, G: float, seed: int = 0):
    """-> positions (n,3) f64, velocities (n,3) f64, masses (n) f64 (all 1.0, as the reference)."""
    rng = np.random.default_rng(seed)
    mass = np.ones(n, np.float64)
    if distribution == "galaxy":
        pos, vel = _disk(rng, n, R, G, 0.3, 1.0, 0.012, 0.12, +1.0)
        vel -= vel.mean(axis=0)
    elif distribution == "collision":       # laws of tools/presets.py:148-232
        half = n // 2
        sep = R * 0.5 * 3.5
        p1, v1 = _disk(rng, half, R, G, 0.25, 0.5, 0.01, 0.10, +1.0)
        p2, v2 = _disk(rng, n - half, R, G, 0.25, 0.5, 0.01, 0.10, -1.0)
        p1[:, 0] -= sep / 2
        p2[:, 0] += sep / 2
        p2[:, 1] += R * 0.15
        speed = np.sqrt(2.0 * G * (n * 0.001) / sep) * 0.6
        v1[:, 0] += speed
        v2[:, 0] -= speed
        pos, vel = np.concatenate([p1, p2]), np.concatenate([v1, v2])
    elif distribution == "cluster":         # Plummer sphere, laws of tools/presets.py:350-397
        a = R * 0.3
        u = rng.uniform(0.0, 1.0, n)
        r = np.clip(a / np.sqrt(u ** (-2.0 / 3.0) - 1.0), 0.0, R * 1.5)
        ph = rng.uniform(0.0, 2.0 * np.pi, n)
        ct = rng.uniform(-1.0, 1.0, n)
        st = np.sqrt(1.0 - ct * ct)
        pos = np.stack([r * st * np.cos(ph), r * ct, r * st * np.sin(ph)], axis=1)
        s2 = G * (n * 0.001) / (6.0 * a)
        sigma = np.sqrt(np.maximum(s2 / np.sqrt(1.0 + (r / a) ** 2), s2 * 0.01))
        vm = np.abs(rng.normal(0.0, 1.0, n) * sigma * np.sqrt(3.0))
        vph = rng.uniform(0.0, 2.0 * np.pi, n)
        vct = rng.uniform(-1.0, 1.0, n)
        vst = np.sqrt(1.0 - vct * vct)
        vel = np.stack([vm * vst * np.cos(vph), vm * vct, vm * vst * np.sin(vph)], axis=1)
        vel -= vel.mean(axis=0)
    elif distribution == "sphere":          # uniform ball, zero velocity
        d = rng.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        pos = d * (R * rng.uniform(0.0, 1.0, n) ** (1.0 / 3.0))[:, None]
        vel = np.zeros((n, 3))
    else:
        raise ValueError(f"unknown synthetic distribution {distribution!r}")
    return np.ascontiguousarray(pos), np.ascontiguousarray(vel), mass


DISTRIBUTIONS = (
    "galaxy", "collision", "spiral", "sphere", "ring", "shell", "cluster", "binary", "elliptical", "bar", "stream",
    "filament", "explosion", "disc", "vortex", "cube", "pleiades", "double_helix", "accretion_disk", "torus",
    "hourglass", "fibonacci", "triple", "rosette", "dyson")   # tools/presets.py:23-49


def generate_distribution(distribution: str, n: int, R: float, G: float, seed: int = 0, device: int = 0):
    """Drop-in for the reference's generate_distribution(distribution, n, R, G) (tools/presets.py:91-1390): all
    25 laws, generated ON THE GPU by libb200sim.so (csrc/generate.cu) from a seed -- reproducible, one thread
    per body, no per-body Python loops -- and copied back as the reference's three host arrays
    positions (n,3), velocities (n,3), masses (n) [fp64].  An unknown name gives the sphere, like the
    reference's final else.  Raises if the library or a GPU is missing (no CPU fallback);
    B200BarnesHutSimulation.from_distribution keeps the state on the device instead."""
    import ctypes as C
    from . import _lib
    L = _lib.load()
    pos, vel, mass = np.empty((n, 3)), np.empty((n, 3)), np.empty(n)
    dp = C.POINTER(C.c_double)
    _lib.check(L.b200_generate_distribution(str(distribution).encode(), int(n), float(R), float(G), int(seed), int(device),
                                            pos.ctypes.data_as(dp), vel.ctypes.data_as(dp), mass.ctypes.data_as(dp)))
    return pos, vel, mass


def generate_preset(key: str, seed: int = 0, num_bodies: int | None = None):
    cfg = get_preset_config(key)
    n = cfg["num_bodies"] if num_bodies is None else int(num_bodies)
    pos, vel, mass = generate(cfg["distribution"], n, cfg["spawn_radius"], cfg["G"], seed)
    return cfg, pos, vel, mass
