"""Traversal timing of ONE build of the library on a device-generated workload (development aid; torch-free, a
run costs a few seconds of box time).  Use with B200SIM_LIB=<variant .so> (scripts/build_variants.sh).
    python scripts/trav_probe.py [preset] [bodies|-] [steps] [--count]
Prints per-phase ms (profiled plain launches), the captured-step wall time, the state checksum after the
run (equal checksums = bit-identical trajectories across builds) and, with --count, the walk statistics."""
import os
import sys
import time

sys.path.insert(0, ".")
import b200sim  # noqa
from b200sim import presets
from b200sim.nbody.gpu_backend import B200BarnesHutSimulation

key = sys.argv[1] if len(sys.argv) > 1 else "extreme_50m_galaxy_t07"
cfg = presets.get_preset_config(key)
n = cfg["num_bodies"] if len(sys.argv) <= 2 or sys.argv[2] in ("-", "None") else int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("-") else 8
sim = B200BarnesHutSimulation.from_distribution(cfg["distribution"], n, cfg["spawn_radius"], cfg["G"], cfg["G"], cfg["softening"],
                                                cfg["damping"], cfg["theta"], seed=0)
dt = cfg["dt"]
for _ in range(3):
    sim.step(dt)
sim.sync()
sim.reset_stats(); sim.set_profiling(True)
for _ in range(steps):
    sim.step(dt)
sim.sync()
st = sim.get_stats()
sim.set_profiling(False)
ph = {k: v / st["timed_steps"] for k, v in st["phase_ms"].items()}
t0 = time.perf_counter()
for _ in range(steps):
    sim.step(dt)
sim.sync()
wall = 1e3 * (time.perf_counter() - t0) / steps
cs = sim.state_checksum()
line = (f"{os.environ.get('B200SIM_LIB', 'shipped'):32s} traverse {ph['traverse']:7.3f} ms | step {sum(ph.values()):7.3f} (captured {wall:7.3f}) | "
        f"checksum {cs[0] & 0xffffffff:08x}{cs[1] & 0xffffffff:08x}")
if "--count" in sys.argv:
    sim.reset_stats()
    inter = sim.count_interactions()
    c = sim.get_stats()
    slots = max(c["trav_pair_slots"], 1)
    line += (f" | inter/body {inter / n:6.1f} {20 * inter / (ph['traverse'] * 1e-3) / 1e12:5.2f} TFLOP/s slots/batch {slots / max(c['trav_batches'], 1):5.1f} "
             f"stack max {c['trav_stack_max']} sure {c['trav_sure_pairs'] / slots:.3f} lane-util {c['trav_lane_pairs'] / (32 * slots):.3f}")
print(line, flush=True)
sim.close()
