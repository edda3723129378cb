"""The seeded restatement of the 25 initial-condition laws (oracle/generators.py, the CPU twin of the
device generators in csrc/generate.cu) against the UNMODIFIED reference generators
(tools/presets.py:91-1390) through the golden quantile tables of tests/golden/generators_ref.npz."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle import generators as gen  # noqa: E402
from make_golden_generators import REPS, summaries  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "generators_ref.npz"))
BAND = 3   # +-0.03 in probability (two-sample KS noise at n = 20 000 is ~0.014 at the 95 % level)


def test_golden_covers_every_reference_distribution():
    assert sorted(gen.DISTRIBUTIONS) == sorted(str(d) for d in GOLD["distributions"])
    assert len(gen.DISTRIBUTIONS) == 25


def test_philox_known_answers():
    """Philox4x32-10 known-answer vectors of the Random123 distribution (kat_vectors)."""
    out = gen.philox4x32(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = gen.philox4x32(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(x) for x in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = gen.philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(x) for x in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.parametrize("dist", gen.DISTRIBUTIONS)
def test_law_matches_reference_quantiles(dist):
    n, R, G = int(GOLD["n"]), float(GOLD["R"]), float(GOLD["G"])
    reps = REPS.get(dist, 1)
    band = BAND if reps == 1 else 5   # pooled realisations of a random node table: node-count noise on top
    pooled = []
    for rep in range(reps):
        pos, vel, mass = gen.generate(dist, n, R, G, seed=7 + rep)
        assert pos.shape == (n, 3) and vel.shape == (n, 3) and mass.shape == (n,)
        assert np.isfinite(pos).all() and np.isfinite(vel).all()
        assert abs(mass.sum() - float(GOLD[f"{dist}.mass_sum"])) <= 1e-9 * float(GOLD[f"{dist}.mass_sum"])
        pooled.append(summaries(pos, vel, mass))
    probs = GOLD["probs"]
    for key in pooled[0]:
        v = np.concatenate([s[key] for s in pooled])
        qref = GOLD[f"{dist}.{key}"]
        q = np.quantile(v, probs)
        eps = 1e-9 * max(np.abs(qref).max(), 1e-30) + 1e-12
        if key in ("vy", "vtan", "speed"):
            # the laws end with a centre-of-mass velocity shift (a sample mean): a common offset of ~std / sqrt(n)
            eps += 4.0 * float(np.std(v)) / np.sqrt(n)
        lo, hi = qref[:-2 * band] - eps, qref[2 * band:] + eps
        mid = q[band:-band]
        bad = (mid < lo) | (mid > hi)
        assert not bad.any(), (f"{dist}.{key}: quantile function leaves the +-0.0{band} band at p = "
                               f"{probs[band:-band][bad][:5]}: {mid[bad][:5]} vs [{lo[bad][:5]}, {hi[bad][:5]}]")


def test_streams_are_counter_based():
    """body i's draws do not depend on n or on the other bodies (what lets the device generate any slice)."""
    a, _, _ = gen.generate("shell", 1000, 300.0, 0.1, seed=3)
    b, _, _ = gen.generate("shell", 5000, 300.0, 0.1, seed=3)
    assert np.array_equal(a, b[:1000])
    c, _, _ = gen.generate("shell", 1000, 300.0, 0.1, seed=4)
    assert not np.array_equal(a, c)
