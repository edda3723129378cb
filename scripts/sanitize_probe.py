"""Small workload that touches every kernel family of libb200sim.so, meant to be run under compute-sanitizer
(scripts/sanitize.sh): memcheck / racecheck / synccheck / initcheck slow kernels down 10-100x, so the sizes are a
few thousand bodies.  No oracle, no torch: plain ctypes calls through the package, like a user of the drop-in.

    compute-sanitizer --tool memcheck python scripts/sanitize_probe.py [--graphs]

Without --graphs the step runs as plain launches (B200_NO_GRAPH=1), which gives the sanitizer per-kernel
attribution; with it the captured-graph replay path is exercised.
"""
from __future__ import annotations

import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if "--graphs" not in sys.argv:
    os.environ["B200_NO_GRAPH"] = "1"

import numpy as np  # noqa: E402

import b200sim  # noqa: E402,F401
from b200sim import presets  # noqa: E402
from b200sim.boids.flock import B200Flock  # noqa: E402
from b200sim.nbody.gpu_backend import B200BarnesHutSimulation  # noqa: E402


def nbody(n: int, theta: float, walk: str | None):
    if walk is None:
        os.environ.pop("B200_TRAV", None)
    else:
        os.environ["B200_TRAV"] = walk
    cfg, pos, vel, mass = presets.generate_preset("tiny_galaxy", seed=1, num_bodies=n)
    sim = B200BarnesHutSimulation(pos, vel, mass, cfg["G"], cfg["softening"], cfg["damping"], theta, device=0)
    sim.compute_accelerations()
    sim.count_interactions()
    for _ in range(3):
        sim.step(cfg["dt"])
    sim.compute_colors(15.0)
    p, v, c = sim.get_positions(), sim.get_velocities(), sim.get_colors()
    assert np.isfinite(p).all() and np.isfinite(v).all() and np.isfinite(c).all()
    # frame egress: plain, delta, visible
    fp, fc = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
    sim.frame_begin(15.0, fp, fc)
    sim.frame_wait()
    sim.step(cfg["dt"])
    dp, dc = np.empty((n, 3), np.int16), np.empty((n, 3), np.int16)
    sim.frame_delta_begin(15.0, dp, dc)
    sim.frame_wait()
    vp, vc = sim.visible_frame([0.0, 0.0, 3.0 * cfg["spawn_radius"]], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0],
                               math.radians(60.0), 16.0 / 9.0, 1.0e4)
    assert len(vp) == len(vc) <= n
    # state replacement: blocking and prefetched
    sim.set_state(pos, vel)
    sim.step(cfg["dt"])
    sim.set_state_begin(pos, vel)
    sim.set_state_commit()
    sim.step(cfg["dt"])
    sim.sync()
    st = sim.get_stats()
    sim.close()
    return st


def generators(n: int):
    for law in presets.DISTRIBUTIONS:
        pos, vel, mass = presets.generate_distribution(law, n, 100.0, 1.0, seed=3)
        assert np.isfinite(pos).all() and np.isfinite(vel).all() and (mass > 0).all(), law
    sim = B200BarnesHutSimulation.from_distribution("galaxy", n, 100.0, 1.0, 1.0, 1.0, 1.0, 0.7, seed=5)
    sim.step(0.01)
    sim.sync()
    sim.close()


def boids(n: int):
    rng = np.random.default_rng(0)
    flock = B200Flock((rng.random((n, 3)) - 0.5) * 80.0, (rng.random((n, 3)) - 0.5) * 25.0, rng.random((n, 3)),
                      params=dict(bounds=40.0), device=0)
    for _ in range(3):
        flock.update(1.0 / 60.0)
    p, v, c = flock.get_state()
    assert np.isfinite(p).all() and np.isfinite(v).all() and np.isfinite(c).all()
    flock.get_cell_indices()
    flock.close()
    flock = B200Flock.random(3000, seed=2)          # the default 202^3 grid
    flock.update(1.0 / 60.0)
    flock.sync()
    flock.close()


def main():
    print("[sanitize] nbody 3000 bodies, 32-body walk", nbody(3000, 0.7, "32").get("records"))
    print("[sanitize] nbody 5000 bodies, 64-body classed walk", nbody(5000, 0.5, "64").get("records"))
    print("[sanitize] nbody 1 body / 2 bodies / ragged tile")
    for n in (1, 2, 65):
        nbody(n, 0.5, None)
    generators(2000)
    print("[sanitize] generators ok")
    boids(4000)
    print("[sanitize] boids ok")


if __name__ == "__main__":
    main()
