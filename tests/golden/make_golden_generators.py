"""Golden quantile tables of the UNMODIFIED reference generators (authoring container only).

Run:  python tests/golden/make_golden_generators.py
Needs /root/reference; writes tests/golden/generators_ref.npz.  For each of the 25 distributions of
tools/presets.py:91-1390 (n = 20 000, R = 300, G = 0.1, numpy's legacy global RandomState seeded with
1234) it stores the quantile functions (p = 0.01 .. 0.99) of seven scalar summaries of the bodies --
|pos|, cylindrical radius in the XZ plane, y, |vel|, tangential speed in the XZ plane, v_y and mass --
which pin the radial profile, thickness, rotation curve and velocity dispersion of every law.
tests/test_generators_cpu.py requires the seeded restatement (oracle/generators.py, the twin of the
device generators) to reproduce each quantile function inside a +-0.03 probability band.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

N, R, G, SEED = 20_000, 300.0, 0.1, 1234
PROBS = np.linspace(0.01, 0.99, 99)
# the cosmic web is one draw of ~180 active nodes with power-law weights: one realisation is not representative
# of the law, so both sides pool several realisations (and the test uses a wider band)
REPS = {"filament": 16}


def summaries(pos, vel, mass):
    """The seven scalar summaries (also used by the tests on the restatement's output)."""
    rcyl = np.sqrt(pos[:, 0] ** 2 + pos[:, 2] ** 2)
    vtan = (pos[:, 0] * vel[:, 2] - pos[:, 2] * vel[:, 0]) / np.maximum(rcyl, 1e-12)
    return dict(rad=np.sqrt((pos ** 2).sum(axis=1)), rcyl=rcyl, y=pos[:, 1], speed=np.sqrt((vel ** 2).sum(axis=1)),
                vtan=vtan, vy=vel[:, 1], mass=mass)


def main():
    from oracle import refimport
    ref = refimport.load()
    out = dict(n=N, R=R, G=G, seed=SEED, probs=PROBS, numpy_version=np.__version__,
               distributions=np.array(sorted(ref.presets.DISTRIBUTIONS)))
    for dist in sorted(ref.presets.DISTRIBUTIONS):
        pooled = []
        for rep in range(REPS.get(dist, 1)):
            np.random.seed(SEED + rep)
            pos, vel, mass = ref.generate_distribution(dist, N, R, G)
            pooled.append(summaries(pos, vel, mass))
        for k in pooled[0]:
            out[f"{dist}.{k}"] = np.quantile(np.concatenate([s[k] for s in pooled]), PROBS)
        out[f"{dist}.mass_sum"] = mass.sum()
        print(f"{dist:15s} |pos| median {np.median(np.sqrt((pos ** 2).sum(1))):9.3f}  |vel| median {np.median(np.sqrt((vel ** 2).sum(1))):8.4f}")
    np.savez_compressed(os.path.join(HERE, "generators_ref.npz"), **out)


if __name__ == "__main__":
    main()
