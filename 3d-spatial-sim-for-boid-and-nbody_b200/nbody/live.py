"""Live-viewer path on the B200 backend.

The reference's viewer (`NBodySimulation`, nbody/simulation.py:440-960) does, per displayed frame and with a GPU
backend: `step` + `compute_colors` + `get_positions().astype(float64)` + `get_colors()` (`_update_gpu`, :809-817),
then on the CPU a frustum test of every body (`_compute_visibility` -> `compute_visibility_points`, :403-434,
:880-904) and two boolean-mask gathers (`draw`, :926-927) whose results are uploaded to the VBOs (:936-937).
At 50 M bodies that is 1.2 GB over PCIe, a 1.2 GB float64 conversion and two 600 MB gathers per frame.

`attach_live(sim)` patches ONE reference `NBodySimulation` object (no reference file is modified) so that
  * `_update_gpu(dt)` only advances the device state, and
  * `_compute_visibility(...)` asks the device for the visible bodies (frustum test + creation-order compaction on
    the GPU, `B200BarnesHutSimulation.visible_frame`) and leaves them where the reference's own `draw()` picks its
    data up: `positions` / `colors` hold exactly the visible rows and `_visible_mask` is all-true, so
    `positions[mask].astype(float32)` and `colors[mask]` in `draw()` are the arrays the reference would have drawn.
Only the visible bodies cross PCIe.  With PyOpenGL + CUDA-GL interop the same device call can write into the mapped
VBOs instead (`visible_frame_device`); there is no display on the B200 box, so that variant is exercised with plain
device buffers.
"""
from __future__ import annotations

import types

import numpy as np


class LiveView:
    """The same data path without a reference object: update(dt) + visible(camera) -> the two VBO payloads."""

    def __init__(self, gpu_sim, max_speed_color: float = 15.0, fog_end: float = 10000.0):
        self.sim = gpu_sim
        self.max_speed_color = float(max_speed_color)   # config/nbody.py:73
        self.fog_end = float(fog_end)                   # config.CAMERA["far_clip"] (nbody/simulation.py:497)

    def update(self, dt: float):
        self.sim.step(min(dt, 0.02))                    # nbody/simulation.py:801-802

    def visible(self, cam_pos, cam_forward, cam_right, cam_up, fov=None, aspect=None):
        import math
        fov_rad = math.radians(fov) if fov else math.radians(75)      # nbody/simulation.py:913-914
        aspect = aspect if aspect else (16 / 9)
        return self.sim.visible_frame(cam_pos, cam_forward, cam_right, cam_up, fov_rad, aspect, self.fog_end,
                                      self.max_speed_color)


def attach_live(ref_sim):
    """Patch a reference NBodySimulation whose `_gpu_sim` is a B200BarnesHutSimulation (see module docstring)."""
    gpu = getattr(ref_sim, "_gpu_sim", None)
    if gpu is None or not hasattr(gpu, "visible_frame"):
        raise TypeError("attach_live needs a reference NBodySimulation running on the B200 backend")

    def _update_gpu(self, dt):
        self._gpu_sim.step(dt)

    def _compute_visibility(self, cam_pos, cam_forward, cam_right, cam_up, fov_v, aspect):
        vp, vc = self._gpu_sim.visible_frame(cam_pos, cam_forward, cam_right, cam_up, fov_v, aspect, self.fog_end,
                                             self.max_speed_color)
        self.positions, self.colors = vp, vc            # draw(): positions[mask].astype(float32), colors[mask]
        self._visible_count = len(vp)
        self._visible_mask = np.ones(len(vp), dtype=np.bool_)

    ref_sim._update_gpu = types.MethodType(_update_gpu, ref_sim)
    ref_sim._compute_visibility = types.MethodType(_compute_visibility, ref_sim)
    return ref_sim
