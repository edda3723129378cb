"""Quick per-phase timing probe of the n-body step (development aid)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import b200sim  # noqa
from b200sim import presets
from b200sim.nbody.gpu_backend import B200BarnesHutSimulation

key = sys.argv[1] if len(sys.argv) > 1 else "4k_collision_1m"
n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] not in ("-", "None") else None
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
t0 = time.time()
cfg, pos, vel, mass = presets.generate_preset(key, 0, n)
print(f"generated {len(pos)} bodies in {time.time()-t0:.1f}s", flush=True)
sim = B200BarnesHutSimulation(pos, vel, mass, cfg["G"], cfg["softening"], cfg["damping"], cfg["theta"])
for _ in range(3):
    sim.step(cfg["dt"])
sim.sync()
def run(counting):
    sim.reset_stats()
    sim.set_profiling(True)
    sim.set_counting(counting)
    t0 = time.time()
    for _ in range(steps):
        sim.step(cfg["dt"])
    sim.sync()
    wall = time.time() - t0
    st = sim.get_stats()
    sim.set_counting(False)
    sim.set_profiling(False)
    return st, wall

st, wall = run(False)
print("wall ms/step", 1e3 * wall / steps)
tot = 0
for k, v in st["phase_ms"].items():
    print(f"  {k:10s} {v/st['timed_steps']:9.3f} ms")
    tot += v / st["timed_steps"]
print("  total", tot, "ms  -> body-updates/s", len(pos) / (tot * 1e-3))
tr = st["phase_ms"]["traverse"] / st["timed_steps"] * 1e-3
if "--nocount" not in sys.argv:
    st, _ = run(True)
    ips = st["interactions"] / st["timed_steps"]
    print("interactions/body", ips / len(pos), "records", st["records"], "bounds", st["bounds"])
    print("traversal (counting) ms", st["phase_ms"]["traverse"] / st["timed_steps"])
    if st["trav_pair_slots"]:
        print("traversal: pair evals/body", st["trav_pair_slots"] / st["timed_steps"] / len(pos) * 32,
              "lane utilisation", st["trav_lane_pairs"] / (32 * st["trav_pair_slots"]),
              "pair evals/batch", st["trav_pair_slots"] / st["trav_batches"], "stack max", st["trav_stack_max"],
              "shared fraction of evals", 2 * st.get("trav_shared_pairs", 0) / st["trav_pair_slots"])
    print("traversal TFLOP/s (20 flop/interaction)", 20 * ips / tr / 1e12)
for _ in range(3):
    sim.step(cfg["dt"])
sim.sync()
t0 = time.time()
for _ in range(steps):
    sim.step(cfg["dt"])
sim.sync()
print("unprofiled wall ms/step", 1e3 * (time.time() - t0) / steps, "(captured step)")
