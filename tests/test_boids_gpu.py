"""Parity of the CUDA boids update (through the C ABI) against the CPU oracle and the
reference-generated fixtures.  -m gpu."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as orc  # noqa: E402

RTOL = 1e-9   # fp64 on both sides; only the neighbour summation order differs


def _flock(pos, vel, col, params):
    from b200sim.boids.flock import B200Flock
    return B200Flock(pos, vel, col, params=params)


@pytest.mark.parametrize("case", ["dense_3k", "sparse_2k"])
def test_two_updates_match_reference_golden(golden_dir, case):
    g = np.load(os.path.join(golden_dir, f"boids_{case}.npz"))
    params = dict(zip([str(k) for k in g["params_keys"]], [float(x) for x in g["params_vals"]]))
    f = _flock(g["pos0"], g["vel0"], g["col0"], params)
    st = f.get_stats()
    assert (st["grid_dim"], st["cell_size"], st["grid_offset"]) == (int(g["grid_dim"]), float(g["cell_size"]), float(g["grid_offset"]))
    assert np.array_equal(f.get_cell_indices(), g["cell_indices1"])       # bit-exact integer work
    for s in (1, 2):
        f.update(float(g["dt"]))
        p, v, c = f.get_state()
        np.testing.assert_allclose(p, g[f"pos{s}"], rtol=RTOL, atol=1e-9)
        np.testing.assert_allclose(v, g[f"vel{s}"], rtol=RTOL, atol=1e-9)
        np.testing.assert_allclose(c, g[f"col{s}"], rtol=RTOL, atol=1e-12)


@pytest.mark.parametrize("n,bounds,seed", [(20_000, 60.0, 0), (50_000, 500.0, 1), (1, 20.0, 2), (2, 20.0, 3), (0, 20.0, 4),
                                           (777, 12.0, 5)])
def test_updates_match_oracle_synthetic(n, bounds, seed):
    rng = np.random.default_rng(seed)
    params = dict(bounds=bounds)
    pos = (rng.random((n, 3)) - 0.5) * 2 * bounds * 1.02   # some boids beyond the walls / outside the grid
    vel = (rng.random((n, 3)) - 0.5) * 25.0
    col = rng.random((n, 3))
    f = _flock(pos, vel, col, params)
    p, v, c = pos.copy(), vel.copy(), col.copy()
    pairs = 0
    for _ in range(3):
        nc = np.zeros(n, np.int32)
        orc.boids_step(p, v, c, 1.0 / 60.0, params, neighbor_counts=nc)
        pairs += int(nc.sum())
        f.update(1.0 / 60.0)
    gp, gv, gc = f.get_state()
    np.testing.assert_allclose(gp, p, rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(gv, v, rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(gc, c, rtol=RTOL, atol=1e-12)
    assert f.get_stats()["neighbor_pairs"] == pairs                          # identical neighbour sets


def test_default_config_grid_and_many_steps():
    """config/boids.py defaults: 202^3 = 8 242 408 cells, 23-bit keys; state keeps creation order."""
    from b200sim.boids.flock import B200Flock
    n = 30_000
    f = B200Flock.random(n, seed=7, params=dict(bounds=500.0))
    st = f.get_stats()
    assert (st["grid_dim"], st["num_cells"], st["key_bits"]) == (202, 8_242_408, 23)
    p0, v0, c0 = (a.copy() for a in (f.positions, f.velocities, f.colors))
    gp, gv, gc = f.get_state()
    assert np.array_equal(gp, p0) and np.array_equal(gv, v0) and np.array_equal(gc, c0)
    p, v, c = p0.copy(), v0.copy(), c0.copy()
    for _ in range(5):
        f.update(1 / 60)
        orc.boids_step(p, v, c, 1 / 60, dict(bounds=500.0))
    gp, gv, gc = f.get_state()
    np.testing.assert_allclose(gp, p, rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(gv, v, rtol=RTOL, atol=1e-9)
    speed = np.linalg.norm(gv, axis=1)
    assert speed.max() <= 25.0 * (1 + 1e-12)


def test_coincident_boids_and_set_state():
    rng = np.random.default_rng(3)
    base = (rng.random((400, 3)) - 0.5) * 30
    pos = np.concatenate([base, base[:100], np.zeros((50, 3))])   # exact duplicates: d2 = 0 is not a neighbour
    n = len(pos)
    vel = (rng.random((n, 3)) - 0.5) * 20
    col = rng.random((n, 3))
    params = dict(bounds=20.0)
    f = _flock(pos, vel, col, params)
    f.update(0.02)
    p, v, c = pos.copy(), vel.copy(), col.copy()
    orc.boids_step(p, v, c, 0.02, params)
    gp, gv, gc = f.get_state()
    np.testing.assert_allclose(gv, v, rtol=RTOL, atol=1e-9)
    f.set_state(pos, vel, col)
    f.update(0.02)
    gp2, gv2, _ = f.get_state()
    assert np.array_equal(gp, gp2) and np.array_equal(gv, gv2)                # deterministic
