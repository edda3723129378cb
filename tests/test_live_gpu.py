"""Live-viewer path: frustum test + creation-order compaction on the device against the oracle's
restatement of compute_visibility_points (nbody/simulation.py:403-434) and, when the reference tree is present
(oracle/_ref), against the reference's own NBodySimulation.update() + _compute_visibility() + draw()-style mask
gathers running on this backend.  -m gpu."""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as orc  # noqa: E402
from oracle import refimport  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _camera(eye, target):
    eye, target = np.asarray(eye, np.float64), np.asarray(target, np.float64)
    f = target - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, [0.0, 1.0, 0.0])
    r /= np.linalg.norm(r)
    u = np.cross(r, f)
    return eye, f, r, u


def _expected(sim, cam, fov_v, aspect, far, max_speed):
    sim.compute_colors(max_speed)
    pos32, col = sim.get_positions(), sim.get_colors()
    tan_h = math.tan(math.atan(math.tan(fov_v / 2) * aspect))
    mask = orc.visibility_mask(pos32.astype(np.float64), *cam, tan_h, math.tan(fov_v / 2), far)
    return pos32[mask], col[mask], mask


@pytest.mark.parametrize("n,eye,far", [(200_003, (0.0, 150.0, 600.0), 900.0), (200_003, (40.0, 5.0, 30.0), 10000.0),
                                       (1025, (0.0, 0.0, 900.0), 50.0), (5_000_000, (300.0, 200.0, 2500.0), 4000.0), (1, (0, 0, 5.0), 100.0)])
def test_visible_frame_equals_the_cpu_frustum_test_and_mask_gather(n, eye, far):
    from b200sim import presets
    from b200sim.nbody.gpu_backend import B200BarnesHutSimulation
    cfg, pos, vel, mass = presets.generate_preset("quick_galaxy" if n < 1_000_000 else "extreme_50m_galaxy_t07", 0, n)
    sim = B200BarnesHutSimulation(pos, vel, mass, cfg["G"], cfg["softening"], cfg["damping"], cfg["theta"])
    cam = _camera(eye, (0.0, 0.0, 0.0))
    fov_v, aspect = math.radians(75), 16 / 9
    for step in range(2):
        sim.step(cfg["dt"])
        vp, vc = sim.visible_frame(*cam, fov_v, aspect, far, 15.0)
        ep, ec, mask = _expected(sim, cam, fov_v, aspect, far, 15.0)
        assert len(vp) == int(mask.sum())
        assert np.array_equal(vp, ep) and np.array_equal(vc, ec)      # same bodies, creation order, bit-equal payload
    if n > 100_000 and far < 5000:
        assert 0 < len(vp) < n                                           # the case really culls
    if far == 50.0:
        assert len(vp) == 0                                              # everything beyond the far plane
    sim.close()


def test_visible_frame_into_device_buffers():
    """The interop variant: the compacted frame lands in caller-provided device memory (stand-in for mapped VBOs)."""
    import torch
    from b200sim import presets
    from b200sim.nbody.gpu_backend import B200BarnesHutSimulation
    n = 300_000
    cfg, pos, vel, mass = presets.generate_preset("quick_galaxy", 1, n)
    sim = B200BarnesHutSimulation(pos, vel, mass, cfg["G"], cfg["softening"], cfg["damping"], cfg["theta"])
    sim.step(cfg["dt"])
    cam = _camera((0.0, 100.0, 400.0), (0.0, 0.0, 0.0))
    fov_v, aspect, far = math.radians(60), 4 / 3, 700.0
    dpos = torch.zeros((n, 3), dtype=torch.float32, device="cuda")
    dcol = torch.zeros((n, 3), dtype=torch.float32, device="cuda")
    cam15 = np.concatenate([*cam, [math.tan(math.atan(math.tan(fov_v / 2) * aspect)), math.tan(fov_v / 2), far]])
    k = sim.visible_frame_device(cam15, dpos.data_ptr(), dcol.data_ptr(), 15.0)
    ep, ec, _ = _expected(sim, cam, fov_v, aspect, far, 15.0)
    assert k == len(ep)
    assert np.array_equal(dpos[:k].cpu().numpy(), ep) and np.array_equal(dcol[:k].cpu().numpy(), ec)
    sim.close()


@pytest.mark.skipif(not refimport.available(), reason="no reference tree (oracle/_ref)")
def test_reference_viewer_object_with_the_live_path_attached():
    """The reference's NBodySimulation (nbody/simulation.py:440-960), constructed through the drop-in so that it runs on
    this backend, once as it is (full copies + its own CPU frustum test + mask gathers) and once with attach_live():
    the arrays draw() would upload must be identical."""
    from b200sim import dropin
    from b200sim.nbody.live import attach_live
    dropin.install(refimport.REFERENCE_ROOT)
    import importlib
    refsim = importlib.import_module("nbody.simulation")
    sims = []
    for _ in range(2):
        np.random.seed(3)
        sims.append(refsim.NBodySimulation(num_bodies=120_000))
    plain, live = sims
    assert plain._use_gpu and live._use_gpu, "the reference object did not pick the B200 backend"
    attach_live(live)
    cam = _camera((0.0, 120.0, 380.0), (0.0, 0.0, 0.0))
    fov_v, aspect = math.radians(75), 16 / 9
    for _ in range(3):
        plain.update(0.016)
        live.update(0.016)
        plain._compute_visibility(*cam, fov_v, aspect)
        live._compute_visibility(*cam, fov_v, aspect)
        want_pos = plain.positions[plain._visible_mask].astype(np.float32)     # draw(), nbody/simulation.py:926-927
        want_col = plain.colors[plain._visible_mask]
        got_pos = live.positions[live._visible_mask].astype(np.float32)
        got_col = live.colors[live._visible_mask]
        assert plain._visible_count == live._visible_count and 0 < live._visible_count < 120_000
        assert np.array_equal(got_pos, want_pos) and np.array_equal(got_col, want_col)
