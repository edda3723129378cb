"""Generate golden fixtures from the UNMODIFIED reference (authoring container only).

Run:  python tests/golden/make_golden.py
Needs /root/reference and numba; writes tests/golden/*.npz.  The fixtures are what pins
the CPU oracle (oracle/) -- the reference itself ships no golden vectors.

Inputs are seeded through numpy's legacy global RandomState (np.random.seed), which is
what the reference's generators draw from (tools/presets.py:91-1390, boids/flock.py:488-490).
numpy version recorded in each file.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refimport  # noqa: E402


def ref_tree(ref, pos, mass):
    """Tree exactly as tools/record.py:795-846 builds it."""
    n = len(pos)
    max_nodes = min(8_000_000, n * 4)
    t = dict(
        node_centers=np.zeros((max_nodes, 3)), node_half_sizes=np.zeros(max_nodes),
        node_masses=np.zeros(max_nodes), node_com=np.zeros((max_nodes, 3)),
        node_children=np.full((max_nodes, 8), -1, np.int32), node_body_idx=np.full(max_nodes, -1, np.int32),
        node_is_leaf=np.ones(max_nodes, np.bool_))
    bounds = ref.compute_bounds(pos, n)
    num_nodes = ref.build_octree(pos, mass, n, bounds, t["node_centers"], t["node_half_sizes"], t["node_masses"],
                                 t["node_com"], t["node_children"], t["node_body_idx"], t["node_is_leaf"])
    return bounds, num_nodes, t


def ref_forces(ref, pos, mass, t, num_nodes, theta, G, eps):
    n = len(pos)
    acc = np.zeros((n, 3))
    ref.compute_forces_barnes_hut(pos, mass, acc, t["node_centers"], t["node_half_sizes"], t["node_masses"],
                                  t["node_com"], t["node_children"], t["node_body_idx"], t["node_is_leaf"],
                                  num_nodes, n, theta, G, eps)
    return acc


def nbody_case(ref, name, dist, n, R, G, eps, thetas, dt, seed, mass_mode="ones"):
    np.random.seed(seed)
    pos, vel, mass = ref.generate_distribution(dist, n, R, G)
    if mass_mode == "varied":
        mass = np.random.uniform(0.5, 2.0, n)
    bounds, num_nodes, t = ref_tree(ref, pos, mass)
    out = dict(pos=pos, vel=vel, mass=mass, bounds=bounds, num_nodes=num_nodes, G=G, softening=eps,
               thetas=np.array(thetas), dt=dt, damping=0.999, seed=seed, numpy_version=np.__version__,
               distribution=dist, R=R)
    for k, v in t.items():
        out[k] = v[:num_nodes]
    for th in thetas:
        out[f"acc_theta_{th}"] = ref_forces(ref, pos, mass, t, num_nodes, th, G, eps)
    # two full substeps sequenced as tools/record.py:835-858, at thetas[0]
    p, v = pos.copy(), vel.copy()
    for _ in range(2):
        b, nn, tt = ref_tree(ref, p, mass)
        a = ref_forces(ref, p, mass, tt, nn, thetas[0], G, eps)
        ref.update_positions_velocities(p, v, a, 0.999, dt, n)
    out["pos_after2"], out["vel_after2"] = p, v
    col = np.zeros((n, 3), np.float32)
    ref.compute_colors_by_velocity(v, col, n, 15.0)
    out["colors_after2"] = col
    np.savez_compressed(os.path.join(HERE, f"nbody_{name}.npz"), **out)
    print(f"nbody_{name}: n={n} nodes={num_nodes} bounds={bounds:.4f}")


def colors_case(ref):
    """Sweep |v|/max_speed over every branch of the colour map (nbody/simulation.py:349-400)."""
    t = np.concatenate([np.linspace(0, 1.2, 1201), [0.15, 0.30, 0.45, 0.55, 0.90, 0.95, 0.99, 1.0]])
    rng = np.random.RandomState(7)
    d = rng.normal(size=(len(t), 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    vel = d * (t * 15.0)[:, None]
    col = np.zeros((len(t), 3), np.float32)
    ref.compute_colors_by_velocity(vel, col, len(t), 15.0)
    np.savez_compressed(os.path.join(HERE, "colors_sweep.npz"), vel=vel, colors=col, max_speed=15.0)
    print("colors_sweep:", len(t))


def boids_case(ref, name, n, bounds, seed, steps=2, dt=1.0 / 60.0):
    """Flock.update from a seeded state in a small box so every rule and the walls fire."""
    import config.boids as cb  # the reference's config module
    saved = dict(cb.BOIDS)
    cb.BOIDS["bounds"] = bounds
    try:
        np.random.seed(seed)
        f = ref.Flock(n)
        out = dict(pos0=f.positions.copy(), vel0=f.velocities.copy(), col0=f.colors.copy(), dt=dt,
                   params_keys=np.array(sorted(cb.BOIDS.keys())),
                   params_vals=np.array([float(cb.BOIDS[k]) for k in sorted(cb.BOIDS.keys())]),
                   grid_dim=f.grid_dim, cell_size=f.cell_size, grid_offset=f.grid_offset,
                   seed=seed, numpy_version=np.__version__)
        for s in range(1, steps + 1):
            f.update(dt)
            out[f"pos{s}"], out[f"vel{s}"], out[f"col{s}"] = f.positions.copy(), f.velocities.copy(), f.colors.copy()
            if s == 1:
                out["sep_forces1"] = f._sep_forces.copy()
                out["align_forces1"] = f._align_forces.copy()
                out["coh_forces1"] = f._coh_forces.copy()
                out["cell_indices1"] = f._cell_indices.copy()
    finally:
        cb.BOIDS.clear()
        cb.BOIDS.update(saved)
    np.savez_compressed(os.path.join(HERE, f"boids_{name}.npz"), **out)
    nb = int((np.abs(out["align_forces1"]).sum(1) > 0).sum())
    print(f"boids_{name}: n={n} grid_dim={out['grid_dim']} boids_with_neighbours={nb}")


def main():
    ref = refimport.load()
    nbody_case(ref, "galaxy_2k", "galaxy", 2000, 200.0, 0.2, 5.0, [0.5, 0.95], 0.3, seed=0)
    nbody_case(ref, "collision_3k", "collision", 3000, 250.0, 0.25, 5.0, [0.5, 0.95], 0.3, seed=1)
    nbody_case(ref, "cluster_2k", "cluster", 2000, 150.0, 0.15, 1.0, [0.3, 0.7, 1.5], 0.2, seed=2, mass_mode="varied")
    colors_case(ref)
    boids_case(ref, "dense_3k", 3000, bounds=30.0, seed=3)
    boids_case(ref, "sparse_2k", 2000, bounds=120.0, seed=4)


if __name__ == "__main__":
    main()
