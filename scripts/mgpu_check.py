"""torchrun check of the multi-GPU plumbing (sharded traversal, sharded sort, sharded host traffic):
every rank compares its replica, step by step, with an unsharded twin running on the same GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import b200sim  # noqa
from b200sim import presets
from b200sim.nbody.gpu_backend import B200BarnesHutSimulation
from b200sim.nbody.sharded import ShardedSimulation

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_003
pos, vel, mass = presets.generate("collision", n, 500.0, 0.1, 2)
mass = np.random.default_rng(0).uniform(0.5, 2.0, n)
mk = lambda: B200BarnesHutSimulation(pos, vel, mass, 0.1, 2.0, 1.0, 0.6, device=local)
twin = mk()
sh = ShardedSimulation(mk(), rank, world)
ok = True
shards = set()
for step in range(18):   # crosses the cost-weighted rebalancing of the shards (after steps 1, 8 and 16)
    sh.step(0.1); twin.step(0.1)
    shards.add(sh.get_shard())
    if step < 3 or step in (8, 9, 17):
        same = np.array_equal(sh.get_positions_f64(), twin.get_positions_f64()) and np.array_equal(sh.get_velocities(), twin.get_velocities())
        ok &= same
        print(f"[rank {rank}] step {step}: shard {sh.get_shard()} identical to the unsharded twin: {same}", flush=True)
print(f"[rank {rank}] shard ranges seen: {sorted(shards)}", flush=True)
if world > 1:
    ok &= all(b % 4096 == 0 and (e % 4096 == 0 or e == n) for b, e in shards)   # cost-weighted boundaries (4096-body chunks), not the equal-count 64-body-tile split
# sharded host traffic
rng = np.random.default_rng(5)
p2 = torch.from_numpy(pos + rng.normal(size=pos.shape)).pin_memory().numpy()
v2 = torch.from_numpy(vel[::-1].copy()).pin_memory().numpy()
sh.set_state_begin(p2, v2); sh.set_state_commit()
twin.set_state(p2, v2)
same = np.array_equal(sh.get_positions_f64(), twin.get_positions_f64()) and np.array_equal(sh.get_velocities(), twin.get_velocities())
ok &= same
print(f"[rank {rank}] sharded upload identical: {same}", flush=True)
for step in range(3):
    sh.step(0.1); twin.step(0.1)
ok &= np.array_equal(sh.get_positions_f64(), twin.get_positions_f64())
fp = torch.zeros((n, 3), dtype=torch.float32).pin_memory().numpy()
fc = torch.zeros((n, 3), dtype=torch.float32).pin_memory().numpy()
sh.frame_begin(15.0, fp, fc); sh.frame_wait()
twin.compute_colors(15.0)
b, e = sh.host_rows()
same = np.array_equal(fp[b:e], twin.get_positions()[b:e]) and np.array_equal(fc[b:e], twin.get_colors()[b:e]) \
    and not fp[:b].any() and not fp[e:].any()
ok &= same
print(f"[rank {rank}] frame rows [{b},{e}) identical: {same}", flush=True)
t = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MGPU CHECK", "PASSED" if t.item() == 1.0 else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
