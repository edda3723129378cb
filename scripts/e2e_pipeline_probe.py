"""Where the pipelined end-to-end frame goes: phases of the from-scratch steps, copy bandwidths alone and
concurrent, wall time per frame.  Development aid."""
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


def wall(fn, reps=1):
    sim.sync(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    sim.sync(); torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps


for _ in range(3):
    sim.step(dt)
print("step (Morton-ordered state)      %8.2f ms" % wall(lambda: sim.step(dt), 3))
print("set_state_begin+commit (blocking) %8.2f ms" % wall(lambda: (sim.set_state_begin(hp, hv), sim.set_state_commit()), 2))
print("step right after an upload       %8.2f ms" % wall(lambda: sim.step(dt)))
sim.set_state(hp, hv)
sim.set_profiling(True); sim.reset_stats()
sim.step(dt); sim.sync()
st = sim.get_stats()
print("  phases of that step:", {k: round(v, 2) for k, v in st["phase_ms"].items()})
sim.set_profiling(False)
print("frame_begin+wait (blocking)      %8.2f ms" % wall(lambda: (sim.frame_begin(15.0, out_p[0], out_c[0]), sim.frame_wait()), 2))


def both():
    sim.set_state_begin(hp, hv)
    sim.frame_begin(15.0, out_p[0], out_c[0])
    sim.frame_wait()
    sim.set_state_commit()


print("upload + frame concurrently      %8.2f ms" % wall(both, 2))


def pipelined(k):
    sim.set_state_begin(hp, hv)
    for i in range(k):
        sim.set_state_commit()
        if i + 1 < k:
            sim.set_state_begin(hp, hv)
        sim.step(dt)
        sim.frame_wait()
        sim.frame_begin(15.0, out_p[i & 1], out_c[i & 1])
    sim.frame_wait()


pipelined(2)
print("pipelined frame                  %8.2f ms" % (wall(lambda: pipelined(6)) / 6))


def upload_only(k):
    sim.set_state_begin(hp, hv)
    for i in range(k):
        sim.set_state_commit()
        if i + 1 < k:
            sim.set_state_begin(hp, hv)
        sim.step(dt)


def frame_only(k):
    for i in range(k):
        sim.step(dt)
        sim.frame_wait()
        sim.frame_begin(15.0, out_p[i & 1], out_c[i & 1])
    sim.frame_wait()


def commit_only(k):
    for i in range(k):
        sim.set_state_begin(hp, hv)
        sim.set_state_commit()
    

upload_only(2)
print("pipelined upload + step          %8.2f ms" % (wall(lambda: upload_only(6)) / 6))
frame_only(2)
print("pipelined step + frame           %8.2f ms" % (wall(lambda: frame_only(6)) / 6))
sim.set_profiling(True); sim.reset_stats()
upload_only(4); sim.sync()
st = sim.get_stats()
print("  phases with concurrent upload:", {k: round(v / st["timed_steps"], 2) for k, v in st["phase_ms"].items()})
sim.set_profiling(False)
