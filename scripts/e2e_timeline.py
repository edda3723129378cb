"""Host timestamps around every call of the pipelined frame loop (development aid, torch-free).
    python scripts/e2e_timeline.py [frames: 0|1]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import b200sim  # noqa
from b200sim import pinned_empty, presets
from b200sim.nbody.gpu_backend import B200BarnesHutSimulation

frames = len(sys.argv) > 1 and sys.argv[1] == "1"
cfg = presets.get_preset_config("extreme_50m_galaxy_t07")
n, dt = cfg["num_bodies"], cfg["dt"]
pos, vel, mass = presets.generate_distribution(cfg["distribution"], n, cfg["spawn_radius"], cfg["G"], seed=0)
hp, hv = pinned_empty((n, 3), np.float64), pinned_empty((n, 3), np.float64)
hp[:], hv[:] = pos, vel
out_p = [pinned_empty((n, 3), np.float32) for _ in range(2)]
out_c = [pinned_empty((n, 3), np.float32) for _ in range(2)]
sim = B200BarnesHutSimulation(pos, vel, mass, cfg["G"], cfg["softening"], cfg["damping"], cfg["theta"])
for _ in range(3):
    sim.step(dt)
sim.sync()
rows = []
T0 = time.perf_counter()


def call(name, i, fn):
    t = 1e3 * (time.perf_counter() - T0)
    fn()
    rows.append((i, name, t, 1e3 * (time.perf_counter() - T0)))


sim.set_state_begin(hp, hv)
sim.set_state_commit()
sim.set_state_begin(hp, hv)
sim.step(dt)
if frames:
    sim.frame_begin(15.0, out_p[0], out_c[0])
T0 = time.perf_counter()
for i in range(1, 6):
    call("commit", i, sim.set_state_commit)
    call("begin", i, lambda: sim.set_state_begin(hp, hv))
    call("step", i, lambda: sim.step(dt))
    if frames:
        call("frame_wait", i, sim.frame_wait)
        call("frame_begin", i, lambda: sim.frame_begin(15.0, out_p[i & 1], out_c[i & 1]))
if frames:
    call("frame_wait", 6, sim.frame_wait)
call("sync", 6, sim.sync)
for r in rows:
    print("%d %-12s enter %8.2f  leave %8.2f  (%6.2f)" % (r[0], r[1], r[2], r[3], r[3] - r[2]))
sim.set_state_commit()
sim.close()
