"""Times the pieces of the end-to-end frame (set_state / step / colours / getters) separately."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import b200sim  # noqa
from b200sim import presets
from b200sim.nbody.gpu_backend import B200BarnesHutSimulation

key = sys.argv[1] if len(sys.argv) > 1 else "4k_collision_1m"
n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] not in ("-", "None") else None
cfg, pos, vel, mass = presets.generate_preset(key, 0, n)
n = len(pos)
sim = B200BarnesHutSimulation(pos, vel, mass, cfg["G"], cfg["softening"], cfg["damping"], cfg["theta"])
hp = torch.from_numpy(pos).pin_memory().numpy()
hv = torch.from_numpy(vel).pin_memory().numpy()
out_p = torch.empty((n, 3), dtype=torch.float32).pin_memory().numpy()
out_c = torch.empty((n, 3), dtype=torch.float32).pin_memory().numpy()


def t(name, fn, reps=3):
    fn(); sim.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    sim.sync()
    ms = 1e3 * (time.perf_counter() - t0) / reps
    print(f"{name:28s} {ms:9.3f} ms")
    return ms


t("set_state pinned (48 B/body)", lambda: sim.set_state(hp, hv))
t("set_state pageable", lambda: sim.set_state(pos, vel))
t("step", lambda: sim.step(cfg["dt"]))
t("compute_colors", lambda: sim.compute_colors(15.0))
t("get_positions pinned (12 B)", lambda: sim.get_positions(out=out_p))
t("get_colors pinned (12 B)", lambda: sim.get_colors(out=out_c))
t("get_positions fresh", lambda: sim.get_positions())
x = torch.empty(n * 6, dtype=torch.float64, device="cuda")
hx = torch.from_numpy(np.concatenate([hp.ravel(), hv.ravel()])).pin_memory()
torch.cuda.synchronize()
t0 = time.perf_counter(); x.copy_(hx, non_blocking=True); torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"torch pinned H2D {x.numel()*8/1e9:.2f} GB: {1e3*dt:.2f} ms = {x.numel()*8/dt/1e9:.1f} GB/s")
