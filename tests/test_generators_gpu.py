"""The device initial-condition generators (csrc/generate.cu through the C ABI) against their numpy twin
(oracle/generators.py: same laws, same Philox streams), which tests/test_generators_cpu.py pins to the
reference's generators (tools/presets.py:91-1390).  -m gpu."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import generators as gen  # noqa: E402

R, G = 300.0, 0.1
# fp64 on both sides; what differs is libm vs CUDA's log / cos / sin / pow (<= 2 ulp each), amplified where a law
# divides by a small cylindrical radius: relative to the scale of the array
TOL = 1e-9


def _close(a, b, what):
    scale = max(float(np.abs(b).max()), 1e-30)
    err = float(np.abs(a - b).max()) / scale
    assert err <= TOL, f"{what}: max |device - twin| / scale = {err:.3e}"


@pytest.mark.parametrize("dist", gen.DISTRIBUTIONS)
@pytest.mark.parametrize("n", [20_011])
def test_device_generator_equals_the_numpy_twin(dist, n):
    from b200sim import presets
    pos, vel, mass = presets.generate_distribution(dist, n, R, G, seed=11)
    tp, tv, tm = gen.generate(dist, n, R, G, seed=11)
    assert pos.shape == (n, 3) and vel.shape == (n, 3) and mass.shape == (n,)
    assert np.array_equal(mass, tm)
    _close(pos, tp, f"{dist} positions")
    _close(vel, tv, f"{dist} velocities")


@pytest.mark.parametrize("dist,n", [("galaxy", 1), ("collision", 2), ("triple", 7), ("bar", 5), ("dyson", 3), ("hourglass", 4),
                                    ("accretion_disk", 9), ("rosette", 4), ("cube", 27), ("cube", 28), ("filament", 33),
                                    ("double_helix", 1), ("sphere", 0)])
def test_ragged_and_tiny_sizes(dist, n):
    from b200sim import presets
    pos, vel, mass = presets.generate_distribution(dist, n, R, G, seed=3)
    tp, tv, tm = gen.generate(dist, n, R, G, seed=3)
    assert np.array_equal(mass, tm)
    if n:
        _close(pos, tp, f"{dist} positions")
        _close(vel, tv, f"{dist} velocities")


def test_seed_and_prefix_properties():
    """Counter-based streams: another seed gives another realisation, and a law without a global step (no rank, no
    centre-of-mass shift) gives body i the same state whatever n is."""
    from b200sim import presets
    a = presets.generate_distribution("shell", 5000, R, G, seed=1)
    b = presets.generate_distribution("shell", 5000, R, G, seed=1)
    c = presets.generate_distribution("shell", 5000, R, G, seed=2)
    d = presets.generate_distribution("shell", 1000, R, G, seed=1)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert not np.array_equal(a[0], c[0])
    assert np.array_equal(a[0][:1000], d[0])
    s = presets.generate_distribution("no_such_law", 1000, R, G, seed=1)     # tools/presets.py:1379: falls to the sphere
    t = presets.generate_distribution("sphere", 1000, R, G, seed=1)
    assert all(np.array_equal(x, y) for x, y in zip(s, t))


def test_created_from_distribution_on_device_and_full_size_speed():
    """from_distribution: the state is drawn straight into the handle (no host arrays); equal to generate + create.
    At full size (50 M bodies, BASELINE config 5's law) the device generator must take well under a second of
    device work (the reference's vectorised numpy galaxy takes ~19 s on the host, its looped laws hours)."""
    from b200sim import presets
    from b200sim.nbody.gpu_backend import B200BarnesHutSimulation
    n = 100_003
    sim = B200BarnesHutSimulation.from_distribution("collision", n, 900.0, 0.08, G=0.08, softening=1.5, damping=1.0,
                                                    theta=0.5, seed=5)
    pos, vel, mass = presets.generate_distribution("collision", n, 900.0, 0.08, seed=5)
    assert np.array_equal(sim.get_positions_f64(), pos)
    assert np.array_equal(sim.get_velocities(), vel)
    twin = B200BarnesHutSimulation(pos, vel, mass, 0.08, 1.5, 1.0, 0.5)
    sim.step(0.012)
    twin.step(0.012)
    assert np.array_equal(sim.get_positions_f64(), twin.get_positions_f64())
    sim.close()
    twin.close()
    n = 50_000_000
    t0 = time.time()
    big = B200BarnesHutSimulation.from_distribution("galaxy", n, 3000.0, 0.04, G=0.04, softening=10.0, damping=1.0,
                                                    theta=0.7, seed=0)
    big.sync()
    dt = time.time() - t0
    print(f"50 M-body galaxy created on the device in {dt:.2f} s (allocation included)")
    p = big.get_positions()
    r = np.sqrt(p[:, 0].astype(np.float64) ** 2 + p[:, 2] ** 2)
    # exponential disk, scale 0.3 R soft-capped at R (tools/presets.py:104-116): median radius of the law
    u = np.linspace(0.0005, 0.9995, 1000)
    rr = -np.log(u) * 900.0
    rr = np.maximum(rr * (1.0 - np.exp(-3000.0 / (rr + 0.01))), 3.0)
    assert abs(np.median(r) - np.median(rr)) <= 0.01 * np.median(rr)
    assert dt < 5.0
    big.close()
