"""The UNMODIFIED reference driven through the B200 backend on a GPU (BASELINE config 1), and
the CUDA step against the reference's live Numba kernels.  -m gpu.

Needs the reference tree: oracle/_ref (the git-ignored verbatim copy made by oracle/make_ref.py, which
travels to the GPU box) or B200SIM_REFERENCE_ROOT; skipped when absent.  The recorder is run as a user
would run it -- `python -m b200sim.dropin ... record --preset tiny_galaxy` -- and the frames it WRITES
(its own save_frame / BackgroundCompressor / zstd+delta files, read back with its own load_frame) are
compared with the oracle's trajectory from the same seeded initial conditions.
"""
import os
import shutil
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as orc  # noqa: E402
from oracle import refimport  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _writable_reference_root():
    """The recorder writes PROJECT_ROOT/recordings (tools/record.py:34,43-47): it needs a writable tree."""
    for cand in (os.environ.get("B200SIM_REFERENCE_ROOT"), os.path.join(ROOT, "oracle", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "tools")) and os.access(cand, os.W_OK):
            return cand
    return None


needs_ref = pytest.mark.skipif(_writable_reference_root() is None or not refimport.available(),
                               reason="no writable reference tree (run oracle/make_ref.py in the authoring container)")


def _rms_rel(a, ref):
    return float(np.sqrt(((a - ref) ** 2).sum() / (ref ** 2).sum()))


@needs_ref
@pytest.mark.parametrize("preset,frames", [("tiny_galaxy", 6), ("quick_galaxy", 4)])
def test_reference_recorder_end_to_end_on_the_b200_backend(preset, frames):
    """tools.record --preset <preset> (tools/record.py:702-935, presets tools/presets.py:1774-1790,
    :2590-2606) with nbody.gpu_backend replaced by this package: the recorder must pick the CUDA backend
    (tools/record.py:759-784), run its frame loop (:818-832) on it and write frames that follow the oracle."""
    root = _writable_reference_root()
    rec_dir = Path(root) / "recordings" / preset
    shutil.rmtree(rec_dir, ignore_errors=True)
    env = dict(os.environ, B200SIM_REFERENCE_ROOT=root, NUMBA_CACHE_DIR="/tmp/b200sim_numba_cache", PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "b200sim.dropin", "--reference-root", root, "--seed", "0", "record", "--preset", preset,
           "--frames", str(frames)]
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    try:
        assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
        assert "[Record] GPU acceleration: cuda" in res.stdout, res.stdout[-3000:]
        assert "[CUDA] Using B200 Barnes-Hut kernel" in res.stdout
        assert "Using CPU backend" not in res.stdout
        rec = refimport.load_recorder()
        ref = refimport.load()
        cfg = ref.get_preset_config(preset)
        np.random.seed(0)                       # the same draw the launcher's --seed 0 made
        pos, vel, mass = rec.module._generate_initial_conditions(cfg)
        assert len(pos) == cfg["num_bodies"]
        dt = cfg["dt_per_frame"] / cfg["substeps"]
        p, v = pos.copy(), vel.copy()
        for f in range(frames):
            for _ in range(cfg["substeps"]):
                orc.nbody_step(p, v, mass, cfg["theta"], cfg["G"], cfg["softening"], cfg["damping"], dt)
            fp, fc = rec.load_frame(rec_dir, f)
            assert fp.shape == (len(pos), 3) and fc.shape == (len(pos), 3)
            # frame 0 of a batch is stored absolute (float32), later frames as int16 deltas * 1e-3
            # (tools/record.py:254-262): each adds at most 1e-3 of truncation per component
            tol = 1e-4 * np.abs(p - pos).max() + 4e-7 * np.abs(p).max() + 1.001e-3 * f
            assert np.abs(fp - p).max() <= tol, (f, np.abs(fp - p).max(), tol)
            assert np.abs(fc - orc.colors(v, 15.0)).max() <= 1e-3 + 1.001e-3 * f
    finally:
        shutil.rmtree(rec_dir, ignore_errors=True)


@needs_ref
def test_cuda_step_matches_the_live_numba_kernels():
    """The reference's own @njit kernels (nbody/simulation.py:63-317), called exactly as
    tools/record.py:835-858 sequences them, against the CUDA path on identical inputs: accelerations
    within the stated 1e-4 RMS, two substeps, colours."""
    from b200sim import presets
    from b200sim.nbody.gpu_backend import B200BarnesHutSimulation
    ref = refimport.load()
    n, G, eps, theta, damping, dt = 30_000, 0.1, 2.0, 0.6, 0.999, 0.05
    pos, vel, mass = presets.generate("collision", n, 400.0, G, 7)
    mass = np.random.default_rng(7).uniform(0.5, 2.0, n)
    sim = B200BarnesHutSimulation(pos, vel, mass, G, eps, damping, theta)
    acc_gpu = sim.compute_accelerations().astype(np.float64)

    max_nodes = min(8_000_000, n * 4)                   # tools/record.py:795
    nc, nh = np.zeros((max_nodes, 3)), np.zeros(max_nodes)
    nm, ncom = np.zeros(max_nodes), np.zeros((max_nodes, 3))
    nch, nb = np.full((max_nodes, 8), -1, np.int32), np.full(max_nodes, -1, np.int32)
    leaf = np.ones(max_nodes, np.bool_)
    acc = np.zeros((n, 3))
    p, v = pos.copy(), vel.copy()

    def substep():
        bounds = ref.compute_bounds(p, n)
        nch.fill(-1); nb.fill(-1); leaf.fill(True)
        nn = ref.build_octree(p, mass, n, bounds, nc, nh, nm, ncom, nch, nb, leaf)
        ref.compute_forces_barnes_hut(p, mass, acc, nc, nh, nm, ncom, nch, nb, leaf, nn, n, theta, G, eps)
        return bounds

    bounds = substep()
    assert _rms_rel(acc_gpu, acc) <= 1e-4, _rms_rel(acc_gpu, acc)
    assert sim.get_stats()["bounds"] == bounds
    ref.update_positions_velocities(p, v, acc, damping, dt, n)
    substep()
    ref.update_positions_velocities(p, v, acc, damping, dt, n)
    sim.step(dt); sim.step(dt)
    # the stated tolerance is RMS <= 1e-4 on the change; single bodies with a borderline MAC flip stay within 1e-3
    gp, gv = sim.get_positions_f64(), sim.get_velocities()
    assert _rms_rel(gp - pos, p - pos) <= 1e-4 and _rms_rel(gv - vel, v - vel) <= 1e-4
    assert np.abs(gp - p).max() <= 1e-3 * np.abs(p - pos).max()
    assert np.abs(gv - v).max() <= 1e-3 * np.abs(v - vel).max()
    col = np.zeros((n, 3), np.float32)
    ref.compute_colors_by_velocity(v, col, n, 15.0)
    sim.compute_colors(15.0)
    assert np.abs(sim.get_colors() - col).max() < 1e-3


@needs_ref
def test_reference_flock_update_matches_the_cuda_flock():
    """boids/flock.py:627-678 Flock.update (Numba) against B200Flock.update from the same state."""
    from b200sim.boids.flock import B200Flock
    ref = refimport.load()
    np.random.seed(3)
    cfg = ref.flock.config.BOIDS            # the reference's module-level config dict (config/boids.py:30-46)
    old = cfg["bounds"]
    cfg["bounds"] = 70.0                    # dense enough for every rule to act (~6 neighbours per boid)
    try:
        f = ref.Flock(20_000)
    finally:
        cfg["bounds"] = old
    pos, vel, col = f.positions.copy(), f.velocities.copy(), f.colors.copy()
    g = B200Flock(pos, vel, col, params=dict(bounds=70.0))
    for _ in range(3):
        f.update(1.0 / 60.0)
        g.update(1.0 / 60.0)
    gp, gv, gc = g.get_state()
    np.testing.assert_allclose(gp, f.positions, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(gv, f.velocities, rtol=1e-9, atol=1e-8)
    np.testing.assert_allclose(gc, f.colors, rtol=1e-9, atol=1e-9)


@needs_ref
def test_attach_patches_a_live_reference_flock():
    """b200sim.boids.flock.attach: a reference Flock object keeps its interface (update(dt) mutating
    positions / velocities / colors in place, what draw() reads: boids/flock.py:716-726) while the update runs
    on the device; compared with an untouched reference Flock stepping the same state."""
    from b200sim.boids.flock import attach
    ref = refimport.load()
    cfg = ref.flock.config.BOIDS
    old = cfg["bounds"]
    cfg["bounds"] = 50.0
    try:
        np.random.seed(5)
        a = ref.Flock(8_000)
        np.random.seed(5)
        b = ref.Flock(8_000)
    finally:
        cfg["bounds"] = old
    assert np.array_equal(a.positions, b.positions)
    pa = a.positions                       # the arrays must be refreshed IN PLACE
    dev = attach(a)
    for _ in range(2):
        a.update(1.0 / 60.0)
        b.update(1.0 / 60.0)
    assert a.positions is pa
    np.testing.assert_allclose(a.positions, b.positions, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(a.velocities, b.velocities, rtol=1e-9, atol=1e-8)
    np.testing.assert_allclose(a.colors, b.colors, rtol=1e-9, atol=1e-9)
    dev.close()
