"""Timeline of the pipelined end-to-end frame loop (host timestamps around every call + device time of
the step), to find where the loop loses time.  Development aid."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import b200sim  # noqa
from b200sim import presets
from b200sim.nbody.gpu_backend import B200BarnesHutSimulation

key = sys.argv[1] if len(sys.argv) > 1 else "extreme_50m_galaxy_t07"
cfg, pos, vel, mass = presets.generate_preset(key, 0, None)
n, dt = len(pos), cfg["dt"]
sim = B200BarnesHutSimulation(pos, vel, mass, cfg["G"], cfg["softening"], cfg["damping"], cfg["theta"])
hp = torch.from_numpy(pos).pin_memory().numpy()
hv = torch.from_numpy(vel).pin_memory().numpy()
out_p = [torch.empty((n, 3), dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
out_c = [torch.empty((n, 3), dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
for _ in range(3):
    sim.step(dt)
sim.sync()
rows = []
T0 = time.perf_counter()


def stamp(name, i):
    rows.append((i, name, 1e3 * (time.perf_counter() - T0)))


k = 6
sim.set_state_begin(hp, hv)
for i in range(k):
    stamp("commit>", i); sim.set_state_commit(); stamp("commit<", i)
    if i + 1 < k:
        sim.set_state_begin(hp, hv); stamp("begin<", i)
    sim.step(dt); stamp("step<", i)
    sim.frame_wait(); stamp("frame_wait<", i)
    sim.frame_begin(15.0, out_p[i & 1], out_c[i & 1]); stamp("frame_begin<", i)
sim.frame_wait(); stamp("last_frame<", k)
sim.sync(); stamp("sync<", k)
for r in rows:
    print("%d %-14s %9.2f" % r)
