"""Parity of the headline configuration at FULL size (extreme_50m_galaxy_t07: 50 M bodies, theta 0.7), on
the GPU box: Morton keys and sort permutation bit-exact against the oracle quantiser + stable sort over
all 50 M bodies; accelerations of a random target sample against the oracle's uncapped Barnes-Hut (the
reference itself drops bodies above ~5.4 M: SURVEY section 0) and against an fp64 direct sum.
Writes gpurun_out/parity_50m.json (committed as profiles/r01_parity_50m.json).

    python scripts/parity_50m.py [bodies] [bh_targets] [direct_targets]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import b200sim  # noqa
from b200sim import presets
from b200sim.nbody.gpu_backend import B200BarnesHutSimulation
from oracle import oracle as orc

bodies = int(sys.argv[1]) if len(sys.argv) > 1 else None
n_bh = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000
n_ds = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000
key = "extreme_50m_galaxy_t07"
t0 = time.time()
cfg, pos, vel, mass = presets.generate_preset(key, 0, bodies)
n = len(pos)
G, eps, theta = cfg["G"], cfg["softening"], cfg["theta"]
out = {"workload": key, "bodies": n, "theta": theta, "G": G, "softening": eps, "host_threads": orc.num_threads()}
print(f"generated {n} bodies in {time.time() - t0:.1f} s", flush=True)


def rms_rel(a, ref):
    return float(np.sqrt(((a - ref) ** 2).sum() / (ref ** 2).sum()))


acc = {}
for walk in ("64", "32"):
    os.environ["B200_TRAV"] = walk
    sim = B200BarnesHutSimulation(pos, vel, mass, G, eps, cfg["damping"], theta)
    if walk == "64":
        t0 = time.time()
        gk, gp = sim.get_morton_keys(), sim.get_sort_permutation()
        keys = orc.morton_keys(pos)
        perm = orc.sort_permutation(keys)
        out["keys_bit_exact"] = bool(np.array_equal(gk, keys[perm]))
        out["permutation_bit_exact"] = bool(np.array_equal(gp, perm))
        out["keys_sorted"] = bool(np.all(gk[1:] >= gk[:-1]))
        out["duplicate_keys"] = int((gk[1:] == gk[:-1]).sum())
        print(f"keys/permutation checked in {time.time() - t0:.1f} s: {out['keys_bit_exact']} {out['permutation_bit_exact']}", flush=True)
        del gk, gp, keys, perm
    sim.reset_stats()
    acc[walk] = sim.compute_accelerations().astype(np.float64)
    st = sim.get_stats()
    out[f"walk{walk}"] = {"interactions_per_body": st["interactions"] / n, "stack_max": st["trav_stack_max"],
                          "records": st["records"], "error_flags": st["error_flags"]}
    sim.close()
os.environ.pop("B200_TRAV")
out["walks_agree_rms"] = rms_rel(acc["64"], acc["32"])

rng = np.random.default_rng(1)
tgt = np.sort(rng.choice(n, size=min(n_bh, n), replace=False))
t0 = time.time()
tree = orc.build_octree(pos, mass)
out["oracle_build_s"] = time.time() - t0
out["oracle_nodes"] = int(tree.num_nodes)
print(f"oracle tree: {tree.num_nodes} nodes in {out['oracle_build_s']:.1f} s", flush=True)
t0 = time.time()
st = {}
ref = orc.compute_forces(pos, tree, theta, G, eps, targets=tgt, stats=st)
out["oracle_forces_s"] = time.time() - t0
out["oracle_sample"] = {"targets": int(len(tgt)), "interactions_per_target": st["interactions"] / len(tgt), "peak_stack": st["peak_stack"]}
for walk in ("64", "32"):
    a = acc[walk][tgt]
    rel = np.linalg.norm(a - ref, axis=1) / np.linalg.norm(ref, axis=1)
    out[f"walk{walk}"].update(acc_rms_rel_vs_oracle_bh=rms_rel(a, ref), max_body_rel=float(rel.max()),
                              frac_bodies_rel_gt_1e3=float((rel > 1e-3).mean()))
del tree
ds_t = tgt[rng.choice(len(tgt), size=min(n_ds, len(tgt)), replace=False)]
t0 = time.time()
direct = orc.direct_sum(pos, mass, G, eps, targets=ds_t)
out["direct_sum_s"] = time.time() - t0
sel = np.searchsorted(tgt, ds_t)
out["bh_error_vs_direct_sum"] = {"targets": int(len(ds_t)), "oracle_bh": rms_rel(ref[sel], direct),
                                 "walk64": rms_rel(acc["64"][ds_t], direct), "walk32": rms_rel(acc["32"][ds_t], direct)}
out["tolerance"] = {"acc_rms_rel_vs_reference_bh": 1e-4, "bh_error_ratio_vs_reference": 1.05}
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/parity_50m.json", "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
ok = out["keys_bit_exact"] and out["permutation_bit_exact"] and all(
    out[f"walk{w}"]["acc_rms_rel_vs_oracle_bh"] <= 1e-4 and out[f"walk{w}"]["error_flags"] == 0 for w in ("64", "32")) and \
    out["bh_error_vs_direct_sum"]["walk64"] <= 1.05 * out["bh_error_vs_direct_sum"]["oracle_bh"] + 1e-6
print("PARITY 50M", "PASSED" if ok else "FAILED")
sys.exit(0 if ok else 1)
