"""The sharded sort's merge on ONE GPU (development aid): the slices of `world` ranks are sorted one after the other
into this device's exchange buffers and merged, so merge_runs_kernel can be timed / profiled without N GPUs.
    python scripts/merge_probe.py [world] [bodies]"""
import sys

sys.path.insert(0, ".")
import b200sim  # noqa
from b200sim import presets
from b200sim.nbody.gpu_backend import B200BarnesHutSimulation
from b200sim.nbody.sharded import slice_size

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = presets.get_preset_config("extreme_50m_galaxy_t07")
n = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["num_bodies"]
sim = B200BarnesHutSimulation.from_distribution("galaxy", n, cfg["spawn_radius"], cfg["G"], cfg["G"], cfg["softening"], cfg["damping"], cfg["theta"])
S = slice_size(n, world)
sim.sharded_sort_setup(S, world)
sim.set_shard(0, (n // world) // 64 * 64)      # a 1 / world shard keeps the traversal short
sim.step_begin(); sim.step_end(cfg["dt"])
for _ in range(3):
    for r in range(world):
        sim.sort_local(r)
    sim.step_begin_sorted()
    sim.step_end(cfg["dt"])
sim.sync()
print("done")
