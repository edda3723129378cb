"""Morton-range sharding of the Barnes-Hut step over the GPUs of one box.

The sharded step itself lives in the library (csrc/multi.cu, include/b200sim.h "several GPUs behind the C
ABI"): every rank holds the full replicated fp64 state, radix-sorts one slice of it (NCCL all-gather of the
sorted runs, merge by counting), builds the whole tree (deterministic, so identical everywhere), and traverses
and integrates only its contiguous Morton range; the traversal kernel stores the new positions and velocities
of those bodies straight into the next-state buffers of every rank over NVLink peer mappings, and an 8-byte
all-reduce ends the step.  Replicas stay bit-identical to each other and to a single-GPU run.

This module is the one-process-per-GPU (torchrun) front end: ``torch.distributed`` carries the 128-byte NCCL
id to the ranks and provides the stream the library works on; the collectives of the step are issued by the
library.  (One process driving several GPUs needs no torch at all: ``B200BarnesHutSimulation(...,
device_mask=0xff)``.)  The reference has no multi-GPU code; this is new design.
"""
from __future__ import annotations

from typing import List, Tuple

TILE = 64  # a warp owns 64 consecutive sorted bodies (two per lane); slices are whole tiles, so a rank's
           # tiles are tiles of the single-GPU pass and the accelerations are bit-identical


def slice_size(n: int, world: int) -> int:
    """Bodies per rank: equal whole-tile slices covering n (the last may be short or empty)."""
    tiles = (n + TILE - 1) // TILE
    return ((tiles + world - 1) // world) * TILE


def partition_equal(n: int, world: int) -> List[Tuple[int, int]]:
    """[begin, end) of every rank in sorted-body order."""
    s = slice_size(n, world)
    return [(min(r * s, n), min((r + 1) * s, n)) for r in range(world)]


def all_gather_slices(buf, rank: int, world: int, group=None):
    """In-place all-gather of equal slices of a (world * S, C) tensor: rank r contributes rows
    [r*S, (r+1)*S).  Works on CUDA tensors over NCCL and on CPU tensors over gloo (tests)."""
    import torch.distributed as dist
    s = buf.shape[0] // world
    assert s * world == buf.shape[0]
    mine = buf[rank * s:(rank + 1) * s]
    if buf.is_cuda:
        dist.all_gather_into_tensor(buf, mine, group=group)
    else:   # gloo has no all_gather_into_tensor on every build: gather into views
        chunks = [buf[r * s:(r + 1) * s] for r in range(world)]
        dist.all_gather(chunks, mine.clone(), group=group)
    return buf


def broadcast_unique_id(rank: int, group=None) -> bytes:
    """Rank 0 draws the NCCL id through the library, torch.distributed carries it to the others."""
    import torch
    import torch.distributed as dist
    from .gpu_backend import B200BarnesHutSimulation
    uid = B200BarnesHutSimulation.nccl_unique_id() if rank == 0 else bytes(128)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor(list(uid), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0, group=group)
    return bytes(t.cpu().tolist())


class _DeviceArray:
    """Minimal __cuda_array_interface__ carrier so torch can view library-owned memory."""

    def __init__(self, ptr: int, shape, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False),
                                         "version": 3, "strides": None}


class ShardedSimulation:
    """Rank `rank` of `world` (one process per GPU): wraps this rank's B200BarnesHutSimulation replica; same
    duck type (step / compute_colors / get_* / sync) as the single-GPU object.  Every call that changes the
    state must be made by all ranks alike."""

    def __init__(self, sim, rank: int, world: int, group=None):
        import torch
        self.sim, self.rank, self.world, self.group = sim, rank, world, group
        self.n = sim.n
        self.S = slice_size(sim.n, world)
        self._torch = torch
        dev = torch.device("cuda", sim.device)
        # all library work (kernels and the library's own NCCL calls) on torch's current stream
        sim.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        self._stage = None
        if world > 1:
            sim.comm_init(broadcast_unique_id(rank, group), rank, world)

    def step(self, dt: float):
        self.sim.step(dt)            # world > 1: the sharded fused step inside the library

    def shard_bodies(self) -> int:
        """Bodies this rank sorts, traverses and integrates."""
        b, e = partition_equal(self.n, self.world)[self.rank]
        return e - b

    def set_state(self, positions, velocities):
        self.sim.set_state(positions, velocities)

    # sharded host traffic: every rank moves 1/world of the rows over its own PCIe link; the upload is
    # completed over NVLink (all-gather of the staging slices), the frame stays split across the ranks
    def host_rows(self):
        R = -(-self.n // self.world)
        return min(self.rank * R, self.n), min((self.rank + 1) * R, self.n)

    def set_state_begin(self, positions, velocities):
        self.sim.set_state_begin(positions, velocities, rows=self.host_rows() if self.world > 1 else None)

    def set_state_commit(self):
        if self.world > 1:
            torch = self._torch
            if self._stage is None:
                R = -(-self.n // self.world)
                pp, vp = self.sim.upload_staging()
                dev = torch.device("cuda", self.sim.device)
                self._stage = tuple(torch.as_tensor(_DeviceArray(p, (R * self.world, 3), "<f8"), device=dev) for p in (pp, vp))
            self.sim.upload_wait()
            for t in self._stage:
                all_gather_slices(t, self.rank, self.world, self.group)
        self.sim.set_state_commit()

    def frame_begin(self, max_speed, out_positions, out_colors):
        self.sim.frame_begin(max_speed, out_positions, out_colors, rows=self.host_rows() if self.world > 1 else None)

    def __getattr__(self, name):   # compute_colors, get_positions, ... are replica-local
        return getattr(self.sim, name)
