"""Morton-range sharding of the Barnes-Hut step over the GPUs of one box.

One process per GPU (torchrun); ``torch.distributed`` (NCCL over NVLink/NVSwitch) is the
plumbing.  Every rank holds the full replicated fp64 state, sorts and builds the whole tree
(redundant, cheap), traverses only its contiguous slice of the Morton-sorted bodies, then one
all-gather of the accelerations (16 B/body) makes every rank's buffer complete and every rank
integrates all bodies.  Integration is elementwise and the sort is deterministic, so the
replicas stay bit-identical without exchanging positions, velocities or tree data.

The reference has no multi-GPU code; this is new design.
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


class _DeviceArray:
    """Minimal __cuda_array_interface__ carrier so torch can view library-owned memory."""

    def __init__(self, ptr: int, shape, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False),
                                         "version": 3, "strides": None}


class ShardedSimulation:
    """Wraps a B200BarnesHutSimulation replica on this rank's GPU; same duck type
    (step / compute_colors / get_* / sync) as the single-GPU object."""

    def __init__(self, sim, rank: int, world: int, group=None, sharded_sort: bool = True):
        import torch
        self.sim, self.rank, self.world, self.group = sim, rank, world, group
        self.n = sim.n
        self.S = slice_size(sim.n, world)
        begin, end = partition_equal(sim.n, world)[rank]
        sim.set_shard(begin, end)
        ptr, cap = sim.acc_buffer()
        if cap < self.S * world:
            raise RuntimeError("accelerations buffer too small for the padded slices")
        self._torch = torch
        dev = torch.device("cuda", sim.device)
        self.acc_all = torch.as_tensor(_DeviceArray(ptr, (self.S * world, 4)), device=dev)
        # all library work on torch's current stream: the collective is ordered with the kernels
        sim.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        # sharded sort: every rank sorts one slice of the (Morton-ordered) state, the sorted slices are
        # all-gathered (12 B/body) and merged by counting on every rank.  Only once the state is in Morton
        # order (from the second step after an upload); the first step sorts everything on every rank.
        self.sharded_sort = sharded_sort and world > 1
        self._in_morton_order = False
        self._stage = None
        if self.sharded_sort:
            kp, vp = sim.sharded_sort_setup(self.S, world)
            self.keys_all = torch.as_tensor(_DeviceArray(kp, (self.S * world, 1), "<i8"), device=dev)
            self.vals_all = torch.as_tensor(_DeviceArray(vp, (self.S * world, 1), "<i4"), device=dev)

    def step(self, dt: float):
        if self.world == 1:      # nothing to exchange: the library's captured step (CUDA graph)
            self.sim.step(dt)
            self._in_morton_order = True
            return
        if self.sharded_sort and self._in_morton_order:
            self.sim.sort_local(self.rank)
            all_gather_slices(self.keys_all, self.rank, self.world, self.group)
            all_gather_slices(self.vals_all, self.rank, self.world, self.group)
            self.sim.step_begin_sorted()
        else:
            self.sim.step_begin()
        if self.world > 1:
            all_gather_slices(self.acc_all, self.rank, self.world, self.group)
        self.sim.step_end(dt)
        self._in_morton_order = True

    def shard_bodies(self) -> int:
        """Bodies this rank traverses (and sorts, with the sharded sort)."""
        b, e = partition_equal(self.n, self.world)[self.rank]
        return e - b

    # a new state arrives in creation order: the next step must sort everything
    def set_state(self, positions, velocities):
        self._in_morton_order = False
        self.sim.set_state(positions, velocities)

    # sharded host traffic: every rank moves 1/world of the rows over its own PCIe link; the upload is
    # completed over NVLink (all-gather of the staging slices), the frame stays split across the ranks
    def host_rows(self):
        R = -(-self.n // self.world)
        return min(self.rank * R, self.n), min((self.rank + 1) * R, self.n)

    def set_state_begin(self, positions, velocities):
        self.sim.set_state_begin(positions, velocities, rows=self.host_rows() if self.world > 1 else None)

    def set_state_commit(self):
        self._in_morton_order = False
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
