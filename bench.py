#!/usr/bin/env python
"""bench.py -- Barnes-Hut body-updates/s (BASELINE.json metric) on 1/2/4/8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload KEY] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (Morton keys -> radix sort -> LBVH/octree -> theta-MAC
traversal -> integrate) over all bodies of the workload.  Default workload: the 50 M-body
EXTREME galaxy at theta 0.7 (BASELINE.json configs[4]; it fits one GPU, and is the config
quoted for 1/2/4/8 GPUs => strong scaling); the 1 M-body `4k_collision_1m` line (configs[2])
is measured in the same run at N=1 and reported under "also".  Inputs are synthetic, seeded,
generated on the host and resident in HBM before the timed region; per-step touched data is
far larger than the 126 MB L2 at both sizes (no L2 flush needed).

Prints ONE JSON line on rank 0.  `--impl reference` times the reference's own CPU path on the same
config: its unmodified Numba kernels (nbody/simulation.py:63-317) imported from oracle/_ref and sequenced
exactly as tools/record.py:835-858, on all host cores, over a bounded sample (the first 1 M bodies of the
workload, the same sample `cpu_baseline` in the GPU arm uses); the C/OpenMP port of those kernels
(oracle/) is the fallback when the reference tree did not travel.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_INTERACTION = 20          # GPU-Gems-3 convention
CPU_SAMPLE_BODIES = 1_000_000      # bounded CPU sample (full oracle step on this many bodies)
# algorithmic HBM bytes per body per phase (DESIGN.md "Kernels"), fp64 master state
# gather = physical reorder of positions / masses / ids (4 + 36 R, 36 + 16 W; velocities are fetched through the
# permutation by the traversal's integration epilogue, never reordered) fused with the radix-tree topology (8 R, 25 W);
# build = prefix sums (32 R, 32 W) + children lists / allocation (21 R, ~24 W);
# extract = pair records (kids 16 + meta 16 + prefix sums 31 + leaves 16 R, ~50 W)
PHASE_BYTES = {"keygen": 24 + 8, "sort": 8 + 8 * (12 + 12), "gather": 4 + 36 + 36 + 16 + 8 + 25,
               "build": 64 + 45, "extract": 79 + 50, "integrate": 16 + 48 + 48}


def _traffic(workload: str, phase: str):
    """Measured DRAM bytes per launch/step of a phase from the committed ncu captures (profiles/)."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        with open(path) as f:
            e = json.load(f)[workload][phase]
        return e["dram_bytes_read"] + e["dram_bytes_write"]
    except Exception:
        return None


def _dispatch(workload: str):
    """Dispatch-slot accounting of the traversal from the committed ncu per-pipe capture (profiles/)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            return json.load(f)[workload]["traverse_dispatch"]
    except Exception:
        return None


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _dist_env():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, local, world


def _workload(key: str, bodies: int | None):
    import b200sim  # noqa: F401
    from b200sim import presets
    t0 = time.time()
    cfg, pos, vel, mass = presets.generate_preset(key, seed=0, num_bodies=bodies)
    return cfg, pos, vel, mass, time.time() - t0


# ----------------------------------------------------------------------------- CPU baseline
def _host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    import platform
    return platform.processor() or "unknown"


def _pin_host_threads():
    """torchrun exports OMP_NUM_THREADS=1: the CPU arm must not inherit it.  Must run before numba is imported."""
    t = str(_host_threads())
    os.environ["NUMBA_NUM_THREADS"] = t
    os.environ["OMP_NUM_THREADS"] = t
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/b200sim_numba_cache")


class CpuPath:
    """The reference's CPU substep on a bounded sample: kind 'reference' = the unmodified Numba kernels from
    oracle/_ref called as tools/record.py:835-858 does (arrays sized as :795-803); kind 'port' = the oracle's C/OpenMP
    restatement.  step() returns seconds spent in (bounds + pool reset, build, forces, integrate)."""

    def __init__(self, cfg, pos, vel, mass):
        _pin_host_threads()
        self.cfg = cfg
        self.n = len(pos)
        self.p, self.v, self.m = pos.copy(), vel.copy(), mass.copy()
        self.kind, self.info = "port", {}
        self.ref = None
        try:
            from oracle import refimport
            if refimport.available():
                self.ref = refimport.load()
                import numba
                numba.set_num_threads(min(_host_threads(), numba.config.NUMBA_NUM_THREADS))
                self.kind = "reference"
                self.info = {"numba": numba.__version__, "numba_threads": numba.get_num_threads(),
                             "reference_root": os.path.relpath(refimport.REFERENCE_ROOT, ROOT)}
        except Exception as e:   # numba missing / tree unreadable: the port still runs
            self.ref, self.kind, self.info = None, "port", {"reference_unavailable": repr(e)[:200]}
        from oracle import oracle as orc
        self.orc = orc
        if self.ref is None:
            orc.build()
            orc.set_num_threads(_host_threads())
            self.cores = orc.num_threads()
        else:
            self.cores = self.info["numba_threads"]
            n = self.n
            mx = min(8_000_000, n * 4)                              # tools/record.py:795
            self.nc, self.nh = np.zeros((mx, 3)), np.zeros(mx)
            self.nm, self.ncom = np.zeros(mx), np.zeros((mx, 3))
            self.nch, self.nb = np.full((mx, 8), -1, np.int32), np.full(mx, -1, np.int32)
            self.leaf = np.ones(mx, np.bool_)
            self.acc = np.zeros((n, 3))

    def step(self):
        c, n = self.cfg, self.n
        t = [time.perf_counter()]
        if self.ref is not None:
            r = self.ref
            bounds = r.compute_bounds(self.p, n)
            self.nch.fill(-1); self.nb.fill(-1); self.leaf.fill(True)
            t.append(time.perf_counter())
            nn = r.build_octree(self.p, self.m, n, bounds, self.nc, self.nh, self.nm, self.ncom, self.nch, self.nb, self.leaf)
            t.append(time.perf_counter())
            r.compute_forces_barnes_hut(self.p, self.m, self.acc, self.nc, self.nh, self.nm, self.ncom, self.nch, self.nb,
                                        self.leaf, nn, n, c["theta"], c["G"], c["softening"])
            t.append(time.perf_counter())
            r.update_positions_velocities(self.p, self.v, self.acc, c["damping"], c["dt"], n)
            t.append(time.perf_counter())
        else:
            o = self.orc
            bounds = o.compute_bounds(self.p)
            t.append(time.perf_counter())
            tree = o.build_octree(self.p, self.m, bounds)
            t.append(time.perf_counter())
            acc = o.compute_forces(self.p, tree, c["theta"], c["G"], c["softening"])
            t.append(time.perf_counter())
            o.update(self.p, self.v, acc, c["damping"], c["dt"])
            t.append(time.perf_counter())
        return [t[i + 1] - t[i] for i in range(4)]

    def threading_layer(self):
        if self.ref is None:
            return "openmp"
        try:
            import numba
            return numba.threading_layer()
        except Exception:
            return "unknown"


def _cpu_sample(pos, vel, mass):
    """The bounded CPU sample: the first CPU_SAMPLE_BODIES bodies of the workload (creation order is i.i.d.,
    so this is the same distribution at 1/50 of the density).  Both arms use exactly these rows."""
    n = min(CPU_SAMPLE_BODIES, len(pos))
    return pos[:n], vel[:n], mass[:n]


def _cpu_line(path: "CpuPath", times, full_bodies):
    """times: list of [prep, build, forces, integrate] seconds per step."""
    tot = [sum(t) for t in times]
    n = path.n
    best = min(tot)
    value = n / best
    split = {k: float(np.median([t[i] for t in times])) for i, k in enumerate(("bounds_and_reset_s", "build_s", "forces_s", "integrate_s"))}
    import math
    out = {"value": value, "unit": "body-updates/s", "cores": path.cores, "kind": path.kind,
           "sample": f"one full substep of the {'reference Numba kernels (oracle/_ref, tools/record.py:835-858 sequence)' if path.kind == 'reference' else 'oracle C/OpenMP port of the reference kernels'}"
                     f" on the first {n:,} bodies of the workload; best of {len(times)}: {best:.2f} s",
           "split_median": split, "cpu_model": _cpu_model(), "host_threads": _host_threads(), "threading_layer": path.threading_layer(),
           "note": "build_octree is sequential by construction (nbody/simulation.py:63-198); forces run on all cores (prange :225)"}
    out.update(path.info)
    if full_bodies > n:
        out["extrapolated_full_size_value"] = value * math.log(n) / math.log(full_bodies)
        out["extrapolation"] = (f"N log N scaling from {n:,} to {full_bodies:,} bodies (value x ln n / ln N); the reference itself cannot run "
                                "this size correctly (8 M-node cap drops bodies above ~5.4 M, 64-entry stack: SURVEY section 0)")
    return out


def cpu_baseline(cfg, pos, vel, mass, reps: int = 2):
    sp, sv, sm = _cpu_sample(pos, vel, mass)
    path = CpuPath(cfg, sp, sv, sm)
    if path.kind == "reference":
        path.step()                                    # JIT compile + first-touch (untimed)
    times = [path.step() for _ in range(reps)]
    return _cpu_line(path, times, len(pos))


def run_reference(args):
    rank, _local, world = _dist_env()
    if rank != 0:
        return 0
    _pin_host_threads()
    key = args.workload or "extreme_50m_galaxy_t07"
    cfg, pos, vel, mass, _ = _workload(key, args.bodies)
    full = len(pos)
    sp, sv, sm = _cpu_sample(pos, vel, mass)
    del pos, vel, mass
    path = CpuPath(cfg, sp, sv, sm)
    for _ in range(max(1, args.warmup)):
        path.step()                                    # Numba JIT + first touch, then W untimed substeps
    steps = max(1, args.steps)                         # exactly K timed substeps (~0.9 s each on the bounded sample)
    times = [path.step() for _ in range(steps)]
    total = sum(sum(t) for t in times)
    n = path.n
    value = n * steps / total
    cb = _cpu_line(path, times, full)
    cb["value"] = value
    line = {
        "impl": "reference", "metric": "body_updates_per_sec", "value": value, "unit": "body-updates/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": key, "bodies": full, "theta": cfg["theta"], "G": cfg["G"],
                   "softening": cfg["softening"], "dt": cfg["dt"], "distribution": cfg["distribution"], "seed": 0,
                   "sample_bodies": n,
                   "same_config": "same workload, parameters and seed as the GPU arm; each step is one full CPU substep on a bounded "
                                  f"sample (the first {n:,} of the {full:,} bodies -- the rows cpu_baseline in the GPU arm uses): the "
                                  "reference cannot hold the full size (8 M-node cap) and a full-size CPU step would take minutes"},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "body-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- boids (BASELINE configs[3])
BOIDS_PHASE_BYTES = {"cells": 24 + 4 + 4, "sort": 4 + 3 * 16, "gather": 4 + 72 + 72, "rules": 72 + 72}


def _measure_boids(args, torch, peaks):
    """1 M boids, config/boids.py defaults, U(-500,500)^3 positions (boids/flock.py:488-489), dt 1/60:
    device-timed Flock.update at step 3 (uniform) and after 500 steps (clustered), phases, e2e
    (update + get_state to pinned host memory), CPU oracle step on the same state."""
    from b200sim.boids.flock import B200Flock
    from oracle import oracle as orc
    n, dt = 1_000_000, 1.0 / 60.0
    f = B200Flock.random(n, seed=0)
    pos0, vel0, col0 = f.positions.copy(), f.velocities.copy(), f.colors.copy()
    for _ in range(3):
        f.update(dt)
    f.sync()
    steps = max(args.steps, 10)
    ms_uniform = f.timed_steps(dt, steps) / steps
    f.reset_stats(); f.set_profiling(True)
    for _ in range(steps):
        f.update(dt)
    f.sync()
    st = f.get_stats()
    f.set_profiling(False)
    phases = {}
    for k, v in st["phase_ms"].items():
        ms = v / max(st["timed_steps"], 1)
        e = {"ms": ms}
        if k in BOIDS_PHASE_BYTES and ms > 0:
            gbs = BOIDS_PHASE_BYTES[k] * n / (ms * 1e-3) / 1e9
            e.update(algorithmic_bytes_per_boid=BOIDS_PHASE_BYTES[k], achieved_gbs=gbs, frac_of_hbm_peak=gbs / peaks["hbm_gbs"])
        phases[k] = e
    pairs = st["neighbor_pairs"] / max(st["timed_steps"], 1) / n
    for _ in range(500 - 3 - 2 * steps):
        f.update(dt)
    f.sync()
    ms_clustered = f.timed_steps(dt, steps) / steps
    # e2e: update + full state to pinned host buffers (what Flock.draw reads: boids/flock.py:716-726)
    outs = tuple(torch.empty((n, 3), dtype=torch.float64).pin_memory().numpy() for _ in range(3))
    f.update(dt); f.get_state(out=outs)
    t0 = time.perf_counter()
    e2e_steps = 5
    for _ in range(e2e_steps):
        f.update(dt)
        f.get_state(out=outs)
    el = time.perf_counter() - t0
    f.close()
    _pin_host_threads()
    kind, cores, cpu_s = "port", None, None
    try:
        from oracle import refimport
        if refimport.available():
            ref = refimport.load()
            import numba
            rf = ref.Flock(n)                                       # boids/flock.py:459 (config defaults)
            rf.positions[:], rf.velocities[:], rf.colors[:] = pos0, vel0, col0
            rf.update(dt)                                           # JIT + first touch
            ts = []
            for _ in range(3):
                t0 = time.perf_counter(); rf.update(dt); ts.append(time.perf_counter() - t0)
            cpu_s, kind, cores = float(np.median(ts)), "reference", numba.get_num_threads()
    except Exception:
        kind = "port"
    if kind == "port":
        orc.set_num_threads(_host_threads())
        p, v, c = pos0, vel0, col0
        orc.boids_step(p, v, c, dt)                    # touch
        t0 = time.perf_counter()
        orc.boids_step(p, v, c, dt)
        cpu_s, cores = time.perf_counter() - t0, orc.num_threads()
    return {"workload": "boids_flock_1m", "boids": n, "dt": dt, "value": n / (ms_uniform * 1e-3), "unit": "boid-updates/s",
            "ms_per_step": ms_uniform, "ms_per_step_after_500_steps": ms_clustered,
            "value_after_500_steps": n / (ms_clustered * 1e-3), "neighbour_pairs_per_boid": pairs, "phases": phases,
            "dtype": "f64", "grid": {"dim": st["grid_dim"], "cells": st["num_cells"], "key_bits": st["key_bits"]},
            "e2e": {"value": n * e2e_steps / el, "unit": "boid-updates/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 72 * n, "what": "B200Flock.update + get_state(pos, vel, col) into pinned host buffers"},
            "cpu_baseline": {"value": n / cpu_s, "unit": "boid-updates/s", "cores": cores, "kind": kind, "cpu_model": _cpu_model(),
                             "sample": f"one Flock.update ({'reference boids/flock.py:627-678, Numba' if kind == 'reference' else 'oracle C/OpenMP port'}) "
                                       f"on the same 1,000,000-boid initial state; {cpu_s:.3f} s"}}


# ----------------------------------------------------------------------------- GPU arm
def _make_sim(cfg, pos, vel, mass, device, rank, world, torch):
    from b200sim.nbody.gpu_backend import B200BarnesHutSimulation
    from b200sim.nbody.sharded import ShardedSimulation
    sim = B200BarnesHutSimulation(pos, vel, mass, cfg["G"], cfg["softening"], cfg["damping"], cfg["theta"],
                                  device=device)
    return ShardedSimulation(sim, rank, world)


def _timed(sh, dt, steps, torch, dist, world):
    """K steps bracketed by barrier + device synchronize; device time by CUDA events on the
    stream the kernels run on (torch's current stream); max over ranks."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sh.step(dt)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def _measure(key, bodies, args, rank, local, world, torch, dist, with_e2e=True, quiet=False):
    cfg, pos, vel, mass, gen_s = _workload(key, bodies)
    n = len(pos)
    dt = cfg["dt"]
    sh = _make_sim(cfg, pos, vel, mass, local, rank, world, torch)
    sim = sh.sim
    for _ in range(max(args.warmup, 3)):
        sh.step(dt)
    torch.cuda.synchronize()

    # ---- headline: device-timed steps, inputs resident
    launches0 = sim.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ms_total = _timed(sh, dt, args.steps, torch, dist, world)
    clocks = sampler.stop() if rank == 0 else None
    launches = sim.launch_count() - launches0
    ms_per_step = ms_total / args.steps
    value = n / (ms_per_step * 1e-3)

    # ---- per-phase device times (CUDA events inside the library, same stream), separate pass.  Before every
    # profiled step a counting force pass runs on the SAME state (nothing integrated), so the interactions the
    # roofline divides by are exactly those of the traversals that were timed.
    sim.reset_stats()
    sim.set_profiling(True)
    inter_local = 0
    for _ in range(args.steps):
        sim.set_profiling(False)
        inter_local += sim.count_interactions()
        sim.set_profiling(True)
        sh.step(dt)
    torch.cuda.synchronize()
    st = sim.get_stats()
    sim.set_profiling(False)
    phase = {k: v / max(st["timed_steps"], 1) for k, v in st["phase_ms"].items()}
    inter = torch.tensor([float(inter_local) / args.steps], device="cuda", dtype=torch.float64)
    trav = torch.tensor([phase["traverse"]], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(inter, op=dist.ReduceOp.SUM)
        dist.all_reduce(trav, op=dist.ReduceOp.MAX)
    inter_step, trav_ms = float(inter.item()), float(trav.item())

    # ---- replicas (N > 1): every rank's state against rank 0's, and against an unsharded replay of the same
    # number of steps from the same initial state on rank 0's GPU
    replicas = None
    if world > 1:
        steps_done = st["steps"]
        cs = torch.tensor([c & 0x7fffffffffffffff for c in sim.state_checksum()], device="cuda", dtype=torch.int64)
        allcs = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(allcs, cs)
        same = all(bool(torch.equal(allcs[0], c)) for c in allcs)
        replay = None
        if rank == 0:
            from b200sim.nbody.gpu_backend import B200BarnesHutSimulation
            twin = B200BarnesHutSimulation(pos, vel, mass, cfg["G"], cfg["softening"], cfg["damping"], cfg["theta"], device=local)
            twin.step_n(dt, int(steps_done))
            tc = [c & 0x7fffffffffffffff for c in twin.state_checksum()]
            replay = tc == [int(x) for x in allcs[0].tolist()]
            twin.close()
        replicas = {"replicas_identical": same, "matches_single_gpu_replay": replay, "steps_compared": int(steps_done),
                    "how": "64-bit checksums of the fp64 positions and velocities keyed by creation index (b200_nbody_state_checksum), "
                           "all-gathered; rank 0 replays the same steps unsharded on its own GPU"}

    out = dict(cfg=cfg, n=n, gen_s=gen_s, ms_per_step=ms_per_step, value=value, launches=launches, clocks=clocks,
               phase=phase, inter_step=inter_step, trav_ms=trav_ms, stats=st, pos=pos, vel=vel, mass=mass, replicas=replicas,
               shard_bodies=sh.shard_bodies())

    # ---- end to end through the public API with HOST buffers, inside the timed region every step:
    # H2D of the step's inputs (48 B/body, pinned) + step + colours + D2H of positions and colours
    # (24 B/body, pinned).  Two variants: the reference-style blocking getters, and the asynchronous
    # frame egress / state prefetch API (copies overlap the next step's kernels; PCIe both ways).
    if with_e2e:
        pin_pos = torch.from_numpy(pos).pin_memory()
        pin_vel = torch.from_numpy(vel).pin_memory()
        hp, hv = pin_pos.numpy(), pin_vel.numpy()
        out_p = [torch.empty((n, 3), dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
        out_c = [torch.empty((n, 3), dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
        e2e_steps = max(2, min(args.steps, 50))      # ~50 ms each

        def frame_blocking():
            sh.set_state(hp, hv)                   # H2D of the step's inputs
            sh.step(dt)
            sim.compute_colors(15.0)
            sim.get_positions(out=out_p[0])        # D2H, as tools/record.py:828
            sim.get_colors(out=out_c[0])           # D2H, as tools/record.py:829

        def run_blocking(k):
            for _ in range(k):
                frame_blocking()

        def run_pipelined(k):
            # (N > 1: every rank moves 1/N of the rows over its own PCIe link; the upload is completed by
            # an all-gather over NVLink, each rank holds its rows of the frame)
            sh.set_state_begin(hp, hv)             # inputs of step 0
            for i in range(k):
                sh.set_state_commit()
                if i + 1 < k:
                    sh.set_state_begin(hp, hv)     # next step's inputs upload while this step computes
                sh.step(dt)
                sh.frame_wait()                    # host buffers of frame i-1 are complete
                sh.frame_begin(15.0, out_p[i & 1], out_c[i & 1])    # D2H overlaps the next step
            sh.frame_wait()

        def run_pipelined_steady(k):
            # the same loop with the pipeline already full: one untimed prologue iteration, then k iterations each of
            # which commits one upload, starts the next one (except the last: nothing dangles), runs one step and
            # completes one frame.  The upload committed by the first timed iteration is started right before the
            # clock (it runs inside the timed region), so exactly k uploads, k steps and k frames -- copies included --
            # complete between t0 and the closing synchronize.  (scripts/e2e_timeline.py shows the steady cycle.)
            sh.set_state_begin(hp, hv)
            sh.set_state_commit()
            sh.step(dt)
            sh.frame_begin(15.0, out_p[0], out_c[0])
            sh.frame_wait()
            torch.cuda.synchronize()               # prologue drained: no backlog of untimed work inside the timed region
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            sh.set_state_begin(hp, hv)
            for i in range(1, k + 1):
                sh.set_state_commit()
                if i < k:
                    sh.set_state_begin(hp, hv)
                sh.step(dt)
                sh.frame_wait()
                sh.frame_begin(15.0, out_p[i & 1], out_c[i & 1])
            sh.frame_wait()
            torch.cuda.synchronize()
            return time.perf_counter() - t0

        out_dp = [torch.empty((n, 3), dtype=torch.int16).pin_memory().numpy() for _ in range(2)]
        out_dc = [torch.empty((n, 3), dtype=torch.int16).pin_memory().numpy() for _ in range(2)]

        def run_pipelined_delta(k):
            # the same loop with the frame as the recorder's format-2 payload (int16 deltas produced on the
            # device): 12 instead of 24 B/body device-to-host.  Single GPU only.
            sim.frame_begin(15.0, out_p[0], out_c[0]); sim.frame_wait()      # the chain's absolute frame
            sh.set_state_begin(hp, hv)
            for i in range(k):
                sh.set_state_commit()
                if i + 1 < k:
                    sh.set_state_begin(hp, hv)
                sh.step(dt)
                sim.frame_wait()
                sim.frame_delta_begin(15.0, out_dp[i & 1], out_dc[i & 1])
            sim.frame_wait()

        def run_recorder_loop(k):
            # what tools/record.py:818-832 does per frame with substeps = 1: step, colours, positions and colours to
            # the host; the state stays on the device (no upload), the read-back overlaps the next step
            for i in range(k):
                sh.step(dt)
                sh.frame_wait()
                sh.frame_begin(15.0, out_p[i & 1], out_c[i & 1])
            sh.frame_wait()

        def timed(fn):
            fn(1)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn(e2e_steps)
            torch.cuda.synchronize()
            el = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(el, op=dist.ReduceOp.MAX)
            return n * e2e_steps / float(el.item())

        def pcie_probe():
            # what the link gives this process: the 48 B/body upload alone, the 24 B/body read-back alone, and both at
            # once (the steady-state e2e loop keeps both directions busy) -- the PCIe roofline of the e2e number
            dpos = torch.empty((n, 3), dtype=torch.float64, device="cuda")
            dvel = torch.empty((n, 3), dtype=torch.float64, device="cuda")
            dfp = torch.empty((n, 3), dtype=torch.float32, device="cuda")
            dfc = torch.empty((n, 3), dtype=torch.float32, device="cuda")
            op_t, oc_t = torch.from_numpy(out_p[0]), torch.from_numpy(out_c[0])
            s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

            def up():
                with torch.cuda.stream(s_up):
                    dpos.copy_(pin_pos, non_blocking=True); dvel.copy_(pin_vel, non_blocking=True)

            def dn():
                with torch.cuda.stream(s_dn):
                    op_t.copy_(dfp, non_blocking=True); oc_t.copy_(dfc, non_blocking=True)

            def t(fns):
                for f in fns:
                    f()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(2):
                    for f in fns:
                        f()
                torch.cuda.synchronize()
                return (time.perf_counter() - t0) / 2
            t_up, t_dn, t_both = t([up]), t([dn]), t([up, dn])
            return {"h2d_alone_gbs": 48 * n / t_up / 1e9, "d2h_alone_gbs": 24 * n / t_dn / 1e9,
                    "both_directions_ms": 1e3 * t_both, "h2d_alone_ms": 1e3 * t_up, "d2h_alone_ms": 1e3 * t_dn,
                    "bound_value": n / t_both, "bound_what": "bodies / time to move one step's 48 + 24 B/body over PCIe with both "
                    "directions active and NOTHING else running: the e2e loop cannot beat this"}

        pcie = pcie_probe() if world == 1 else None
        v_block = timed(run_blocking)
        v_pipe = timed(run_pipelined)
        steady_steps = e2e_steps
        run_pipelined_steady(2)
        el = torch.tensor([run_pipelined_steady(steady_steps)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        v_steady = n * steady_steps / float(el.item())
        v_delta = timed(run_pipelined_delta) if world == 1 else None
        v_rec = timed(run_recorder_loop)
        out["e2e"] = {"value": v_steady, "unit": "body-updates/s", "h2d_bytes_per_step": 48 * n, "d2h_bytes_per_step": 24 * n,
                      "steps": steady_steps,
                      "pipeline": "k iterations after one untimed (and drained) prologue iteration; exactly k 48 B/body uploads, k steps "
                                  "and k 24 B/body frames start and complete inside the timed region (the closing synchronize "
                                  "covers the last frame); uploads and read-backs overlap the neighbouring steps' kernels",
                      "pcie": pcie,
                      "with_fill_and_drain_value": v_pipe, "with_fill_and_drain_steps": e2e_steps,
                      "what": "per step, through the ctypes C-ABI with pinned HOST buffers: set_state_begin/commit (H2D of "
                              "positions + velocities) + step + frame_begin/wait (colours, D2H of positions + colours); the "
                              "copies run on side streams and overlap the neighbouring steps' kernels",
                      "delta_frames_value": v_delta,
                      "delta_frames_what": "same loop with frame_delta_begin (the recorder's int16 delta payload, computed on the "
                                           "device): 12 B/body device-to-host per step instead of 24",
                      "recorder_loop_value": v_rec,
                      "recorder_loop_what": "the reference recorder's own frame loop (tools/record.py:818-832, substeps 1): step + frame "
                                            "egress (24 B/body device-to-host) every step, no upload -- the state lives on the device",
                      "blocking_value": v_block,
                      "blocking_what": "same bytes with the reference-style blocking calls: set_state + step + compute_colors + "
                                       "get_positions + get_colors"}
    sim.close()
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank, local, world = _dist_env()
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
        local = 0
    # everything (kernels, collectives, timing events) on one non-default stream: the library captures its
    # step into a CUDA graph, which the legacy default stream does not allow
    torch.cuda.set_stream(torch.cuda.Stream())
    import b200sim
    from b200sim import _lib
    _lib.load()   # fails loudly if the CUDA library is missing

    key = args.workload or "extreme_50m_galaxy_t07"
    m = _measure(key, args.bodies, args, rank, local, world, torch, dist)
    cfg, n = m["cfg"], m["n"]

    peaks, peak_src = _peaks()
    fp32_peak = _lib.fp32_peak_tflops(local)
    achieved = FLOP_PER_INTERACTION * m["inter_step"] / (m["trav_ms"] * 1e-3) / 1e12 / world   # per GPU
    kernel = "traverse64c_kernel" if m["stats"].get("trav_kernel") == 64 else "traverse_kernel"
    roofline = {
        "kernel": kernel, "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": achieved / fp32_peak if fp32_peak else None,
        "traffic": (_traffic(key, "traverse") if world == 1 and n == cfg["num_bodies"] else None),
        "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full, profiles/r02_traffic.json): 116 B/body "
                        "of it is the fused integration epilogue (permutation, velocities, positions in; next fp64 state out); "
                        "the kernel is instruction-issue / FP32 bound (DRAM 3 % busy), its records are served from L2",
        "peak_source": "measured in this run: FFMA-chain microbenchmark (b200_fp32_peak_tflops); "
                       "MEASURED_PEAKS.json has no FP32 entry; nominal 148 SM x 128 x 2 x 1.965 GHz = 74.4",
        "algorithmic_flop_per_launch": FLOP_PER_INTERACTION * m["inter_step"] / world,
        "interactions_per_body": m["inter_step"] / n, "launch_ms": m["trav_ms"],
        "interactions_counted": "on the same states as the timed traversals (a counting pass before every profiled step)",
        "ceiling": "the kernel takes all of its SMSPs' dispatch slots (a packed fp32x2 instruction holds the port for 2 cycles: "
                   "2 x packed + other instructions = 1.006 x the active cycles, see 'dispatch'), and more resident warps do not "
                   "help (16 / 24 / 27 warps per SM: 28.2 / 26.0 / 26.05 ms): it is at the dispatch roofline of its instruction "
                   "mix; with every non-arithmetic instruction free it would reach 0.59 of the FFMA peak (DESIGN.md section 4: "
                   "0.63 useful interactions per evaluated lane-slot)",
        "dispatch": (_dispatch(key) if world == 1 and n == cfg["num_bodies"] else None),
        "note": "per GPU; not tensor-core work (no dense contraction); HBM phases under 'phases'",
    }
    # HBM phases: bytes one rank moves / that rank's time.  With N > 1 the key generation and the radix sort run over the
    # rank's slice only (sharded sort); gather, build, extract and integrate are replicated over all bodies.
    shard = m["shard_bodies"]
    phase_bodies = {"keygen": shard, "sort": shard}
    phases = {}
    for k, v in m["phase"].items():
        e = {"ms": v}
        if k == "integrate" and v < 0.02:
            e["note"] = "fused into the traversal kernel's epilogue (traverse.cuh finish_body): its 116 B/body move inside the traverse phase"
        elif k in PHASE_BYTES and v > 0:
            nb = phase_bodies.get(k, n) if world > 1 else n
            gbs = PHASE_BYTES[k] * nb / (v * 1e-3) / 1e9
            e.update(algorithmic_bytes_per_body=PHASE_BYTES[k], bodies_this_rank=nb, achieved_gbs=gbs, frac_of_hbm_peak=gbs / peaks["hbm_gbs"])
            if world > 1 and k == "gather":
                e["note"] = "includes the key/value all-gathers of the sharded sort and the merge of the sorted runs"
            if world > 1 and k == "extract":
                # locally essential tree: this rank writes only the records its shard may open, so the full tree's bytes
                # over its time would overstate the bandwidth
                for kk in ("achieved_gbs", "frac_of_hbm_peak"):
                    e.pop(kk, None)
                e["note"] = "locally essential tree: only the records this rank's shard may open are written (B200_LET=0: all)"
        phases[k] = e

    # the largest HBM-bound phase (radix sort) against the measured copy bandwidth
    sort_bodies = shard if world > 1 else n
    sort_bytes = PHASE_BYTES["sort"] * sort_bodies
    sort_ms = m["phase"].get("sort", 0.0)
    roofline_hbm = None
    if sort_ms > 0:
        gbs = sort_bytes / (sort_ms * 1e-3) / 1e9
        roofline_hbm = {"kernel": "hist_kernel + 8 x onesweep_kernel (radix sort of 63-bit keys + 32-bit payload)", "bound": "hbm",
                        "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                        "traffic": (_traffic(key, "sort") if n == cfg["num_bodies"] and world == 1 else None),
                        "algorithmic_bytes_per_launch": sort_bytes, "bodies_this_rank": sort_bodies, "launch_ms": sort_ms,
                        "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)"}
    line = {
        "metric": "body_updates_per_sec", "value": m["value"], "unit": "body-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": m["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": key, "bodies": n, "theta": cfg["theta"], "G": cfg["G"], "softening": cfg["softening"],
                   "dt": cfg["dt"], "distribution": cfg["distribution"], "seed": 0,
                   "parallelism": (f"morton-range x{world}: sharded sort (NCCL all-gather of the sorted runs), replicated tree, each rank traverses + "
                                   "integrates its range and stores the next state into every rank over NVLink peer mappings"
                                   if world > 1 else "single GPU"),
                   "l2": "per-step working set >> 126 MB L2 (no flush needed)",
                   "state_dtype": "f64 positions/velocities, f32 forces"},
        "steps_per_sec": 1e3 / m["ms_per_step"],
        "roofline": roofline, "roofline_hbm": roofline_hbm, "phases": phases, "hbm_peak_gbs": peaks["hbm_gbs"], "hbm_peak_source": peak_src,
        "e2e": m.get("e2e"), "gpu_launches": m["launches"], "clocks": m["clocks"],
        "host_generate_s": m["gen_s"],
    }
    if m.get("replicas") is not None:
        line["replicas"] = m["replicas"]
        line["replicas_identical"] = bool(m["replicas"]["replicas_identical"] and m["replicas"]["matches_single_gpu_replay"] is not False)
    if rank == 0 and world == 1:
        line["cpu_baseline"] = cpu_baseline(cfg, m["pos"], m["vel"], m["mass"])
    del m
    if world == 1 and key != "4k_collision_1m" and not args.no_also:
        a = _measure("4k_collision_1m", None, args, rank, local, world, torch, dist, with_e2e=True)
        ach = FLOP_PER_INTERACTION * a["inter_step"] / (a["trav_ms"] * 1e-3) / 1e12
        line["also"] = {"workload": "4k_collision_1m", "bodies": a["n"], "theta": a["cfg"]["theta"],
                        "value": a["value"], "unit": "body-updates/s", "ms_per_step": a["ms_per_step"],
                        "phases_ms": a["phase"], "interactions_per_body": a["inter_step"] / a["n"],
                        "roofline": {"kernel": "traverse64c_kernel" if a["stats"].get("trav_kernel") == 64 else "traverse_kernel",
                                     "bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                                     "frac": ach / fp32_peak if fp32_peak else None},
                        "e2e": a.get("e2e"),
                        "cpu_baseline": cpu_baseline(a["cfg"], a["pos"], a["vel"], a["mass"])}
    if world == 1 and not args.no_also:
        line["also_boids"] = _measure_boids(args, torch, peaks)
    if world == 1:
        # the same law generated ON THE DEVICE straight into a handle (csrc/generate.cu), next to host_generate_s
        import time as _t
        from b200sim.nbody.gpu_backend import B200BarnesHutSimulation
        t0 = _t.time()
        g = B200BarnesHutSimulation.from_distribution(cfg["distribution"], n, cfg["spawn_radius"], cfg["G"], cfg["G"], cfg["softening"],
                                                      cfg["damping"], cfg["theta"], seed=0, device=local)
        g.sync()
        line["device_generate_s"] = _t.time() - t0
        line["device_generate_what"] = ("B200BarnesHutSimulation.from_distribution: allocation of the handle + Philox generation + radius "
                                        "ranking (radix sort) + centre-of-mass shift on the device, no host arrays; the timed workload "
                                        "itself is still the host-generated instance (host_generate_s) so the numbers stay comparable "
                                        "across rounds")
        g.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, help="preset key (b200sim.presets.PRESETS)")
    ap.add_argument("--bodies", type=int, default=None, help="override the preset's body count")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary 1 M-body line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
