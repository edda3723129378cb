"""Single-GPU emulation of the sharded sort at full size: times the local sorts and the merge."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import b200sim  # noqa
from b200sim import presets
from b200sim.nbody.gpu_backend import B200BarnesHutSimulation
from b200sim.nbody.sharded import slice_size

key = sys.argv[1] if len(sys.argv) > 1 else "extreme_50m_galaxy_t07"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg, pos, vel, mass = presets.generate_preset(key, 0, None)
n = len(pos)
sim = B200BarnesHutSimulation(pos, vel, mass, cfg["G"], cfg["softening"], cfg["damping"], cfg["theta"])
S = slice_size(n, world)
sim.sharded_sort_setup(S, world)
for _ in range(3):
    sim.step(cfg["dt"])
sim.sync()
def t(fn, name, reps=3):
    ts = []
    for _ in range(reps):
        sim.sync(); t0 = time.perf_counter(); fn(); sim.sync(); ts.append(1e3 * (time.perf_counter() - t0))
    print(f"{name:40s} {min(ts):8.3f} ms")
t(lambda: sim.sort_local(0), "sort_local(0): keygen + local sort")
def all_local():
    for r in range(world):
        sim.sort_local(r)
t(all_local, f"all {world} local sorts")
sim.set_profiling(True); sim.reset_stats()
for _ in range(3):
    all_local(); sim.step_begin_sorted(); sim.step_end(cfg["dt"])
st = sim.get_stats()
print({k: round(v / st["timed_steps"], 3) for k, v in st["phase_ms"].items()})
