"""A/B timing of the traversal variants on one workload (development aid): B200_TRAV is read when a
simulation is created, so every variant gets its own handle on the same initial state.
    python scripts/trav_ab.py <preset> <bodies|-> <steps> <variant> [<variant> ...]      variants: 32 64 64o"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import b200sim  # noqa
from b200sim import presets
from b200sim.nbody.gpu_backend import B200BarnesHutSimulation

key = sys.argv[1]
n = None if sys.argv[2] in ("-", "None") else int(sys.argv[2])
steps = int(sys.argv[3])
variants = sys.argv[4:]
t0 = time.time()
cfg, pos, vel, mass = presets.generate_preset(key, 0, n)
if os.environ.get("AB_THETA"):
    cfg["theta"] = float(os.environ["AB_THETA"])
print(f"{key}: {len(pos)} bodies theta {cfg['theta']} generated in {time.time() - t0:.1f}s", flush=True)
ref_acc = None
for v in variants:
    os.environ["B200_TRAV"] = v
    sim = B200BarnesHutSimulation(pos, vel, mass, cfg["G"], cfg["softening"], cfg["damping"], cfg["theta"])
    acc = sim.compute_accelerations()
    st0 = sim.get_stats()
    if ref_acc is None:
        ref_acc = acc.astype(np.float64)
        rms = 0.0
    else:
        rms = float(np.sqrt(((acc - ref_acc) ** 2).sum() / (ref_acc ** 2).sum()))
    for _ in range(3):
        sim.step(cfg["dt"])
    sim.sync()
    sim.reset_stats(); sim.set_profiling(True)
    for _ in range(steps):
        sim.step(cfg["dt"])
    sim.sync()
    st = sim.get_stats()
    sim.set_profiling(False)
    ph = {k: v2 / st["timed_steps"] for k, v2 in st["phase_ms"].items()}
    sim.reset_stats()
    inter = sim.count_interactions()
    cst = sim.get_stats()
    tr = ph["traverse"]
    slots = max(cst["trav_pair_slots"], 1)
    print(f"  B200_TRAV={v:4s} traverse {tr:8.3f} ms  step {sum(ph.values()):8.3f} ms  inter/body {inter / len(pos):7.1f}  "
          f"{20 * inter / (tr * 1e-3) / 1e12:6.2f} TFLOP/s  lane-util {cst['trav_lane_pairs'] / (32 * slots):.3f}  "
          f"sure {cst['trav_sure_pairs'] / slots:.3f}  both {2 * cst['trav_shared_pairs'] / slots:.3f}  slots/batch {slots / max(cst['trav_batches'], 1):.1f}  "
          f"rms vs first {rms:.2e}  int/body first pass {st0['interactions'] / len(pos):.1f}", flush=True)
    sim.close()
