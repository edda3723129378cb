"""Where the end-to-end frame loop (bench.py's e2e: upload + step + frame every iteration) spends its time.
Torch-free and device-generated, so a run costs seconds of box time: every component is timed alone
(serialised, host clock around a synchronising call), then the pipelined loop and two reduced loops.
    python scripts/e2e_breakdown.py [preset] [bodies|-] [iterations]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import b200sim  # noqa
from b200sim import pinned_empty, presets
from b200sim.nbody.gpu_backend import B200BarnesHutSimulation

key = sys.argv[1] if len(sys.argv) > 1 else "extreme_50m_galaxy_t07"
cfg = presets.get_preset_config(key)
n = cfg["num_bodies"] if len(sys.argv) <= 2 or sys.argv[2] in ("-", "None") else int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 12
dt = cfg["dt"]
t0 = time.perf_counter()
pos, vel, mass = presets.generate_distribution(cfg["distribution"], n, cfg["spawn_radius"], cfg["G"], seed=0)
hp, hv = pinned_empty((n, 3), np.float64), pinned_empty((n, 3), np.float64)
hp[:], hv[:] = pos, vel
out_p = [pinned_empty((n, 3), np.float32) for _ in range(2)]
out_c = [pinned_empty((n, 3), np.float32) for _ in range(2)]
sim = B200BarnesHutSimulation(pos, vel, mass, cfg["G"], cfg["softening"], cfg["damping"], cfg["theta"])
del pos, vel
print(f"{key}: {n} bodies, setup {time.perf_counter() - t0:.1f} s", flush=True)
for _ in range(3):
    sim.step(dt)
sim.sync()


def ms(fn):
    t = time.perf_counter()
    fn()
    return 1e3 * (time.perf_counter() - t)


def both(*fns):
    def run():
        for f in fns:
            f()
    return run


print("-- components alone (ms)")
for rep in range(2):
    t_sorted = ms(both(lambda: sim.step(dt), sim.sync))
    t_up = ms(both(lambda: sim.set_state_begin(hp, hv), sim.set_state_commit))     # commit returns when the copy is done
    t_commit = ms(sim.sync)                                                       # staging -> state, id, bounds
    t_unsorted = ms(both(lambda: sim.step(dt), sim.sync))
    t_fk = ms(both(lambda: sim.frame_begin(15.0, out_p[0], out_c[0]), sim.sync))  # colours + un-permute kernels
    t_d2h = ms(sim.frame_wait)
    print(f"  upload 48 B/body {t_up:7.2f} | commit (device part) {t_commit:6.2f} | step after commit {t_unsorted:6.2f} | "
          f"step in Morton order {t_sorted:6.2f} | frame kernels {t_fk:6.2f} | frame D2H (rest) {t_d2h:6.2f}", flush=True)


def drained(fn):
    """fn(k) -> ms per iteration, started on an idle GPU (no backlog of untimed work inside the timed region)."""
    def run(k):
        sim.frame_wait()
        sim.sync()
        t = time.perf_counter()
        fn(k)
        sim.sync()
        return 1e3 * (time.perf_counter() - t) / k
    return run


@drained
def loop_full(k):          # bench.py's e2e: exactly k uploads, k steps, k frames
    sim.set_state_begin(hp, hv)
    for i in range(1, k + 1):
        sim.set_state_commit()
        if i < k:
            sim.set_state_begin(hp, hv)
        sim.step(dt)
        sim.frame_wait()
        sim.frame_begin(15.0, out_p[i & 1], out_c[i & 1])
    sim.frame_wait()


@drained
def loop_no_frame(k):
    sim.set_state_begin(hp, hv)
    for i in range(1, k + 1):
        sim.set_state_commit()
        if i < k:
            sim.set_state_begin(hp, hv)
        sim.step(dt)


@drained
def loop_no_upload(k):
    for i in range(1, k + 1):
        sim.step(dt)
        sim.frame_wait()
        sim.frame_begin(15.0, out_p[i & 1], out_c[i & 1])
    sim.frame_wait()


@drained
def loop_upload_only(k):
    for i in range(k):
        sim.set_state_begin(hp, hv)
        sim.set_state_commit()


print("-- loops (ms per iteration)")
for name, fn in (("upload + commit only", loop_upload_only), ("upload + step (no frame)", loop_no_frame),
                 ("step + frame (no upload: the recorder's loop)", loop_no_upload), ("upload + step + frame (bench e2e)", loop_full)):
    fn(2)
    v = fn(iters)
    print(f"  {name:48s} {v:7.2f} ms  -> {n / v * 1e3:.3e} body-updates/s", flush=True)
sim.close()
