"""Per-phase timing of the boids update at the BASELINE config (1 M boids, config/boids.py defaults)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import b200sim  # noqa
from b200sim.boids.flock import B200Flock

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
pre = int(sys.argv[3]) if len(sys.argv) > 3 else 0
f = B200Flock.random(n, seed=0)
dt = 1.0 / 60.0
for _ in range(3 + pre):
    f.update(dt)
f.sync()
ms = f.timed_steps(dt, steps)
print(f"n={n} after {3 + pre} steps: {ms / steps:.3f} ms/step -> {n / (ms / steps * 1e-3):.3e} boid-updates/s")
f.reset_stats()
f.set_profiling(True)
for _ in range(steps):
    f.update(dt)
f.sync()
st = f.get_stats()
for k, v in st["phase_ms"].items():
    print(f"  {k:8s} {v / st['timed_steps']:8.3f} ms")
print("neighbour pairs/boid/step", st["neighbor_pairs"] / st["timed_steps"] / n, "grid", st["grid_dim"], "key bits", st["key_bits"])
t0 = time.perf_counter()
out = f.get_state()
print(f"get_state (72 B/boid D2H, pageable): {1e3 * (time.perf_counter() - t0):.2f} ms")
