"""One process, several GPUs through the C ABI alone (b200_nbody_create_multi, no torch): the device-mask
handle is compared step by step with a single-GPU twin.  python scripts/mgpu_mask_check.py [bodies] [mask]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import b200sim  # noqa
from b200sim import _lib, presets
from b200sim.nbody.gpu_backend import B200BarnesHutSimulation

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_003
mask = int(sys.argv[2], 0) if len(sys.argv) > 2 else (1 << _lib.device_count()) - 1
if bin(mask).count("1") < 2:
    print("needs at least two devices in the mask")
    sys.exit(2)
pos, vel, mass = presets.generate("collision", n, 500.0, 0.1, 2)
mass = np.random.default_rng(0).uniform(0.5, 2.0, n)
multi = B200BarnesHutSimulation(pos, vel, mass, 0.1, 2.0, 0.9995, 0.6, device_mask=mask)
twin = B200BarnesHutSimulation(pos, vel, mass, 0.1, 2.0, 0.9995, 0.6, device=0)
ok = multi.world() == bin(mask).count("1")
print(f"world {multi.world()} (mask {mask:#x})", flush=True)
shards = set()
for step in range(18):   # crosses the cost-weighted rebalancing of the shards (after steps 1, 8 and 16)
    multi.step(0.1); twin.step(0.1)
    shards.add(multi.get_shard())
    if step < 3 or step in (8, 9, 17):
        same = np.array_equal(multi.get_positions_f64(), twin.get_positions_f64()) and np.array_equal(multi.get_velocities(), twin.get_velocities())
        ok &= same
        print(f"step {step}: shard of device 0 {multi.get_shard()}: identical to the single-GPU twin: {same}", flush=True)
ok &= all(b % 4096 == 0 and (e % 4096 == 0 or e == n) for b, e in shards)   # cost-weighted boundaries (4096-body chunks), not the equal-count 64-body-tile split
print(f"shard ranges of device 0 seen: {sorted(shards)}", flush=True)
multi.set_state(pos * 1.01, vel); twin.set_state(pos * 1.01, vel)
for step in range(3):
    multi.step(0.05); twin.step(0.05)
same = np.array_equal(multi.get_positions_f64(), twin.get_positions_f64())
ok &= same
print(f"after set_state + 3 steps: identical: {same}", flush=True)
multi.compute_colors(15.0); twin.compute_colors(15.0)
ok &= np.array_equal(multi.get_colors(), twin.get_colors())
a, b = multi.compute_accelerations(), twin.compute_accelerations()
ok &= np.array_equal(a, b)
ok &= multi.state_checksum() == twin.state_checksum()
print("MASK CHECK", "PASSED" if ok else "FAILED", flush=True)
multi.close(); twin.close()
sys.exit(0 if ok else 1)
