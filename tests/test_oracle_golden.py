"""Pins the CPU oracle (oracle/) against fixtures generated from the reference itself
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

NBODY_CASES = ["galaxy_2k", "collision_3k", "cluster_2k"]


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


@pytest.mark.parametrize("case", NBODY_CASES)
def test_bounds_and_tree_identical_to_reference(golden_dir, case):
    g = _load(golden_dir, "nbody_" + case)
    pos, mass = g["pos"], g["mass"]
    assert orc.compute_bounds(pos) == float(g["bounds"])
    t = orc.build_octree(pos, mass, reference_cap=True)
    nn = int(g["num_nodes"])
    assert t.num_nodes == nn
    # same insertion order => node-for-node identical arrays
    assert np.array_equal(t.node_children[:nn], g["node_children"])
    assert np.array_equal(t.node_body_idx[:nn], g["node_body_idx"])
    assert np.array_equal(t.node_is_leaf[:nn].astype(bool), g["node_is_leaf"])
    assert np.array_equal(t.node_centers[:nn], g["node_centers"])
    assert np.array_equal(t.node_half_sizes[:nn], g["node_half_sizes"])
    np.testing.assert_allclose(t.node_masses[:nn], g["node_masses"], rtol=1e-14)
    np.testing.assert_allclose(t.node_com[:nn], g["node_com"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("case", NBODY_CASES)
def test_forces_match_reference(golden_dir, case):
    g = _load(golden_dir, "nbody_" + case)
    pos, mass = g["pos"], g["mass"]
    t = orc.build_octree(pos, mass, reference_cap=True)
    for th in g["thetas"]:
        ref = g[f"acc_theta_{th}"]
        st = {}
        acc = orc.compute_forces(pos, t, float(th), float(g["G"]), float(g["softening"]), stack_cap=64, stats=st)
        assert st["drops"] == 0
        rms = np.sqrt(((acc - ref) ** 2).sum() / (ref ** 2).sum())
        assert rms < 1e-13, (case, th, rms)
        # uncapped stack gives the same answer when nothing was dropped
        acc2 = orc.compute_forces(pos, t, float(th), float(g["G"]), float(g["softening"]))
        assert np.array_equal(acc, acc2)


@pytest.mark.parametrize("case", NBODY_CASES)
def test_two_steps_match_reference(golden_dir, case):
    g = _load(golden_dir, "nbody_" + case)
    p, v, m = g["pos"].copy(), g["vel"].copy(), g["mass"]
    for _ in range(2):
        orc.nbody_step(p, v, m, float(g["thetas"][0]), float(g["G"]), float(g["softening"]),
                       float(g["damping"]), float(g["dt"]), reference_cap=True, stack_cap=64)
    np.testing.assert_allclose(p, g["pos_after2"], rtol=1e-12, atol=1e-10)
    np.testing.assert_allclose(v, g["vel_after2"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(orc.colors(v, 15.0), g["colors_after2"], atol=2e-7)


def test_colour_map_every_branch(golden_dir):
    g = _load(golden_dir, "colors_sweep")
    got = orc.colors(g["vel"], float(g["max_speed"]))
    # branch boundaries can flip on fp64 rounding of |v|/max_speed; tolerate none in practice
    np.testing.assert_allclose(got, g["colors"], atol=2e-6)


@pytest.mark.parametrize("case", ["dense_3k", "sparse_2k"])
def test_boids_steps_match_reference(golden_dir, case):
    g = _load(golden_dir, "boids_" + case)
    params = dict(zip([str(k) for k in g["params_keys"]], [float(x) for x in g["params_vals"]]))
    cell, dim, off = orc.boids_grid(params)
    assert (cell, dim, off) == (float(g["cell_size"]), int(g["grid_dim"]), float(g["grid_offset"]))
    p, v, c = g["pos0"].copy(), g["vel0"].copy(), g["col0"].copy()
    for s in (1, 2):
        orc.boids_step(p, v, c, float(g["dt"]), params)
        np.testing.assert_allclose(p, g[f"pos{s}"], rtol=1e-12, atol=1e-10)
        np.testing.assert_allclose(v, g[f"vel{s}"], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(c, g[f"col{s}"], rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("case", NBODY_CASES)
def test_morton_implied_octree_equals_reference_tree(golden_dir, case):
    """the octree implied by 21-level Morton prefixes over the reference's
    root cube has exactly the reference's cells (internal cells = prefixes shared by >= 2
    bodies; leaves = single bodies under such a cell)."""
    g = _load(golden_dir, "nbody_" + case)
    pos = g["pos"]
    bounds = float(g["bounds"])
    keys = orc.morton_keys(pos, bounds)
    nn = int(g["num_nodes"])
    half = g["node_half_sizes"]
    level = np.rint(np.log2(bounds / half)).astype(int)
    assert level.max() <= 21
    # reference cell -> (level, prefix): prefix of any body below it; use the cell centre
    ckeys = orc.morton_keys(g["node_centers"], bounds)
    ref_internal = {(int(l), int(k) >> (3 * (21 - int(l)))) for l, k, leaf in
                    zip(level, ckeys, g["node_is_leaf"]) if not leaf}
    ref_leaves = {(int(l), int(k) >> (3 * (21 - int(l)))) for l, k, leaf in
                  zip(level, ckeys, g["node_is_leaf"]) if leaf}
    mine_internal, mine_leaves = set(), set()
    from collections import Counter
    counts = [Counter((keys >> np.uint64(3 * (21 - l))).tolist()) for l in range(22)]
    for l in range(22):
        for pref, cnt in counts[l].items():
            if cnt >= 2:
                mine_internal.add((l, pref))
            elif l > 0 and counts[l - 1][pref >> 3] >= 2:
                mine_leaves.add((l, pref))
    assert mine_internal == ref_internal
    assert mine_leaves == ref_leaves
    assert len(ref_internal) + len(ref_leaves) == nn


def test_morton_key_layout():
    """bit0 = x, bit1 = y, bit2 = z of each 3-bit group; level-1 octant in bits 62..60."""
    b = 100.0
    pos = np.array([[50.0, -50.0, -50.0], [-50.0, 50.0, -50.0], [-50.0, -50.0, 50.0], [-1e-9, -1e-9, -1e-9],
                    [0.0, 0.0, 0.0]])
    k = orc.morton_keys(pos, b)
    assert (int(k[0]) >> 60, int(k[1]) >> 60, int(k[2]) >> 60) == (1, 2, 4)
    assert int(k[3]) == int("000" + "111" * 20, 2)
    assert int(k[4]) == int("111" + "000" * 20, 2)
    assert int(k.max()) < 2 ** 63


def test_visibility_mask_restatement_matches_the_reference_frustum_test():
    """oracle.visibility_mask (the checker of the live-viewer path) against the reference's own
    compute_visibility_points (nbody/simulation.py:403-434) where the reference tree is present."""
    import math
    from oracle import refimport
    if not refimport.available():
        pytest.skip("no reference tree")
    ref = refimport.load()
    rng = np.random.default_rng(5)
    n = 50_000
    pos = (rng.normal(size=(n, 3)) * [300.0, 20.0, 300.0]).astype(np.float32).astype(np.float64)   # float32 positions widened (:816)
    eye = np.array([40.0, 120.0, 380.0])
    f = -eye / np.linalg.norm(eye)
    r = np.cross(f, [0.0, 1.0, 0.0]); r /= np.linalg.norm(r)
    u = np.cross(r, f)
    fov_v, aspect, far = math.radians(75), 16 / 9, 600.0
    tan_h, tan_v = math.tan(math.atan(math.tan(fov_v / 2) * aspect)), math.tan(fov_v / 2)
    mask = np.zeros(n, np.bool_)
    ref.simulation.compute_visibility_points(pos, eye, f, r, u, tan_h, tan_v, far, mask, n)
    mine = orc.visibility_mask(pos, eye, f, r, u, tan_h, tan_v, far)
    assert 0 < mask.sum() < n
    # (numba fastmath may contract the dot products: a body within an ulp of a frustum plane may flip)
    assert int((mask != mine).sum()) <= 2
