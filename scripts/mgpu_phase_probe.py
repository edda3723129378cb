"""Per-rank phase times of the sharded step on N GPUs (development aid): the workload is generated ON THE DEVICE
(from_distribution), so a run costs seconds of box time instead of the bench's host generation.
    torchrun --nproc-per-node N scripts/mgpu_phase_probe.py [preset] [bodies|-] [steps]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import b200sim  # noqa
from b200sim import presets
from b200sim.nbody.gpu_backend import B200BarnesHutSimulation
from b200sim.nbody.sharded import ShardedSimulation

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_stream(torch.cuda.Stream())
key = sys.argv[1] if len(sys.argv) > 1 else "extreme_50m_galaxy_t07"
cfg = presets.get_preset_config(key)
n = cfg["num_bodies"] if len(sys.argv) <= 2 or sys.argv[2] in ("-", "None") else int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 16
sim = B200BarnesHutSimulation.from_distribution(cfg["distribution"], n, cfg["spawn_radius"], cfg["G"], cfg["G"], cfg["softening"],
                                                cfg["damping"], cfg["theta"], seed=0, device=local)
sh = ShardedSimulation(sim, rank, world)
dt = cfg["dt"]
for _ in range(10):      # warm-up: crosses the first two rebalancing points of the cost-weighted split
    sh.step(dt)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for _ in range(steps):
    sh.step(dt)
torch.cuda.synchronize()
el = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(el, op=dist.ReduceOp.MAX)
sim.reset_stats(); sim.set_profiling(True)
for _ in range(4):
    sh.step(dt)
sim.sync()
st = sim.get_stats()
sim.set_profiling(False)
ph = {k: v / st["timed_steps"] for k, v in st["phase_ms"].items()}
line = f"[rank {rank}] shard {sim.get_shard()} " + " ".join(f"{k} {v:.3f}" for k, v in ph.items()) + f" | sum {sum(ph.values()):.3f}"
if world > 1:
    lines = [None] * world
    dist.all_gather_object(lines, line)
else:
    lines = [line]
if rank == 0:
    print(f"{key}: {n} bodies on {world} GPU(s): {1e3 * el.item() / steps:.3f} ms/step (wall, max over ranks, {steps} steps)", flush=True)
    for ln in lines:
        print(ln, flush=True)
if world > 1 and "--nccl" in sys.argv:
    # what NCCL itself needs for the sharded sort's exchange at this size: 8 B keys + 4 B positions per body
    S = -(-n // world)
    for dtype, name in ((torch.int64, "keys 8 B"), (torch.int32, "vals 4 B")):
        buf = torch.zeros(S * world, dtype=dtype, device="cuda")
        for _ in range(3):
            dist.all_gather_into_tensor(buf, buf[rank * S:(rank + 1) * S])
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            dist.all_gather_into_tensor(buf, buf[rank * S:(rank + 1) * S])
        torch.cuda.synchronize()
        if rank == 0:
            print(f"NCCL all-gather of the {name}/body slices ({S * buf.element_size() / 1e6:.0f} MB per rank): {(time.perf_counter() - t0) * 100:.3f} ms", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
