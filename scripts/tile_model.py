"""CPU model of the traversal's tiles (development aid): lane utilisation, leaf fraction and the
share of pair slots a conservative tile-level MAC pre-test could classify, for different ways of
forming tiles.  python scripts/tile_model.py [preset] [bodies] [stride]"""
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200sim  # noqa
from b200sim import presets
from oracle import oracle as orc

so = "/tmp/libtilemodel.so"
subprocess.run(["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "scripts", "tile_model.c"), "-lm"], check=True)
L = C.CDLL(so)
FIELDS = ["tiles", "slots", "lanepairs", "child_evals", "interactions", "visits_open", "leaf_child_evals", "leafpair_slots",
          "leafpair_lanepairs", "sure_slots", "sure_lanepairs", "sure_full_slots", "sureopen_slots"] + [f"h{i}" for i in range(33)] + \
         ["bodies", "slots_both", "slots_one"] + [f"c{i}" for i in range(16)]


class TS(C.Structure):
    _fields_ = [(f, C.c_int64) for f in FIELDS]


def p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def run(pos, tree, order, tstart, theta, eps, stride, halves=1):
    st = TS()
    order = np.ascontiguousarray(order, np.int64)
    tstart = np.ascontiguousarray(tstart, np.int64)
    L.tile_model(p(pos, C.c_double), p(tree.node_half_sizes, C.c_double), p(tree.node_com, C.c_double),
                 p(tree.node_children, C.c_int32), p(tree.node_body_idx, C.c_int32), p(tree.node_is_leaf, C.c_uint8),
                 p(order, C.c_int64), p(tstart, C.c_int64), C.c_int64(len(tstart) - 1), C.c_int64(stride),
                 C.c_double(theta), C.c_double(eps), C.c_int(halves), C.byref(st))
    return {f: getattr(st, f) for f in FIELDS}


def report(name, s):
    slots, lp = s["slots"], s["lanepairs"]
    print(f"--- {name}: tiles {s['tiles']} bodies/tile {s['bodies'] / s['tiles']:.2f}")
    print(f"  interactions/body {s['interactions'] / s['bodies']:.1f}  child evals/body {s['child_evals'] / s['bodies']:.1f} "
          f"(visits/interactions {s['child_evals'] / s['interactions']:.3f})  leaf share of child evals {s['leaf_child_evals'] / s['child_evals']:.3f}")
    print(f"  pair slots/tile {slots / s['tiles']:.1f}  lane utilisation {lp / (32 * slots):.3f}  "
          f"useful interactions per slot-lane {s['interactions'] / (64 * slots):.3f}")
    print(f"  leaf-pair slots {s['leafpair_slots'] / slots:.3f}  sure-accept slots {s['sure_slots'] / slots:.3f} "
          f"(lane util inside {s['sure_lanepairs'] / max(1, 32 * s['sure_slots']):.3f}; full-mask {s['sure_full_slots'] / slots:.3f})  "
          f"sure-open slots {s['sureopen_slots'] / slots:.3f}")
    h = np.array([s[f"h{i}"] for i in range(33)], float)
    h /= h.sum()
    print("  slots by mask population: 1-8 %.3f  9-16 %.3f  17-24 %.3f  25-31 %.3f  32 %.3f" %
          (h[1:9].sum(), h[9:17].sum(), h[17:25].sum(), h[25:32].sum(), h[32]))
    if s["slots_both"] + s["slots_one"]:
        c = np.array([s[f"c{i}"] for i in range(16)], float).reshape(4, 4)
        c /= c.sum()
        print("  64-tiles, pair records by class of (low half, high half); rows/cols = none, sure+full, sure+masked, unsure:")
        for r in c:
            print("     " + "  ".join(f"{v:.3f}" for v in r))
        print(f"  64-tiles: pair records needed by both halves {s['slots_both'] / (s['slots_both'] + s['slots_one']):.3f}")


def aligned_tiles(keys_sorted, cap=32):
    """Greedy cell-aligned tiles: never straddle the boundary of a cell with more than `cap` bodies;
    consecutive sibling cells are merged while they fit."""
    n = len(keys_sorted)
    starts = []

    def rec(lo, hi, level):
        if hi - lo <= cap or level >= 21:
            for s in range(lo, hi, cap):
                starts.append(s)
            return
        shift = 3 * (20 - level)
        d = (keys_sorted[lo:hi] >> np.uint64(shift)) & np.uint64(7)
        bnd = lo + np.searchsorted(d, np.arange(9))
        # group consecutive children while the total fits
        g0 = bnd[0]
        acc = 0
        for c in range(8):
            a, b = bnd[c], bnd[c + 1]
            cnt = b - a
            if cnt == 0:
                continue
            if cnt > cap:
                if acc:
                    starts.append(g0)
                    acc = 0
                rec(a, b, level + 1)
                g0 = b
            elif acc + cnt > cap:
                starts.append(g0)
                g0, acc = a, cnt
            else:
                if acc == 0:
                    g0 = a
                acc += cnt
        if acc:
            starts.append(g0)

    sys.setrecursionlimit(10000)
    rec(0, n, 0)
    starts = np.array(sorted(set(starts)) + [n], np.int64)
    return starts


def chunks_inside_cells(keys_sorted, M, tile=32):
    """Tiles = chunks of `tile` consecutive bodies that never straddle the boundary of a maximal cell with <= M bodies."""
    n = len(keys_sorted)
    starts = []

    def rec(lo, hi, level):
        if hi - lo <= M or level >= 21:
            starts.extend(range(lo, hi, tile))
            return
        shift = 3 * (20 - level)
        d = (keys_sorted[lo:hi] >> np.uint64(shift)) & np.uint64(7)
        bnd = lo + np.searchsorted(d, np.arange(9))
        for c in range(8):
            if bnd[c + 1] > bnd[c]:
                rec(bnd[c], bnd[c + 1], level + 1)

    rec(0, n, 0)
    return np.array(sorted(set(starts)) + [n], np.int64)


def hilbert_order(keys):
    """Hilbert-curve order of the bodies from their 63-bit Morton keys (Skilling's transpose algorithm)."""
    b = 21
    k = keys.astype(np.uint64)

    def compact(v):   # every third bit -> contiguous
        x = v & np.uint64(0x1249249249249249)
        x = (x | (x >> np.uint64(2))) & np.uint64(0x10c30c30c30c30c3)
        x = (x | (x >> np.uint64(4))) & np.uint64(0x100f00f00f00f00f)
        x = (x | (x >> np.uint64(8))) & np.uint64(0x1f0000ff0000ff)
        x = (x | (x >> np.uint64(16))) & np.uint64(0x1f00000000ffff)
        x = (x | (x >> np.uint64(32))) & np.uint64(0x1fffff)
        return x
    X = [compact(k), compact(k >> np.uint64(1)), compact(k >> np.uint64(2))]
    M = np.uint64(1 << (b - 1))
    Q = M
    one = np.uint64(1)
    while Q > one:
        P = Q - one
        for i in range(3):
            m = (X[i] & Q) != 0
            if i == 0:
                X[0] = np.where(m, X[0] ^ P, X[0])
            else:
                t = np.where(m, np.uint64(0), (X[0] ^ X[i]) & P)
                X[0] = np.where(m, X[0] ^ P, X[0] ^ t)
                X[i] = X[i] ^ t
        Q >>= one
    X[1] ^= X[0]
    X[2] ^= X[1]
    t = np.zeros_like(X[0])
    Q = M
    while Q > one:
        t = np.where((X[2] & Q) != 0, t ^ (Q - one), t)
        Q >>= one
    X = [x ^ t for x in X]

    def spread(v):
        x = v & np.uint64(0x1fffff)
        x = (x | x << np.uint64(32)) & np.uint64(0x1f00000000ffff)
        x = (x | x << np.uint64(16)) & np.uint64(0x1f0000ff0000ff)
        x = (x | x << np.uint64(8)) & np.uint64(0x100f00f00f00f00f)
        x = (x | x << np.uint64(4)) & np.uint64(0x10c30c30c30c30c3)
        x = (x | x << np.uint64(2)) & np.uint64(0x1249249249249249)
        return x
    h = (spread(X[0]) << np.uint64(2)) | (spread(X[1]) << np.uint64(1)) | spread(X[2])
    return np.argsort(h, kind="stable").astype(np.int64)


def main():
    key = sys.argv[1] if len(sys.argv) > 1 else "extreme_50m_galaxy_t07"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
    stride = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    t0 = time.time()
    cfg, pos, vel, mass = presets.generate_preset(key, 0, n)
    keys = orc.morton_keys(pos)
    perm = orc.sort_permutation(keys).astype(np.int64)
    ks = keys[perm]
    tree = orc.build_octree(pos, mass)
    print(f"n={n} preset={key} theta={cfg['theta']} eps={cfg['softening']} nodes={tree.num_nodes} setup {time.time() - t0:.1f}s", flush=True)
    th, eps = cfg["theta"], cfg["softening"]
    t32 = np.arange(0, n + 32, 32, dtype=np.int64); t32[-1] = n
    report("morton-consecutive 32", run(pos, tree, perm, t32, th, eps, stride))
    t64 = np.arange(0, n + 64, 64, dtype=np.int64); t64 = t64[t64 <= n + 63]; t64[-1] = n
    if "--t64" in sys.argv:
        report("morton-consecutive 64 (two halves)", run(pos, tree, perm, t64, th, eps, max(1, stride // 2), halves=2))
    if "--aligned" in sys.argv:
        ta = aligned_tiles(ks, 32)
        report("cell-aligned greedy <=32", run(pos, tree, perm, ta, th, eps, stride))
    if "--only64" in sys.argv:
        return
    for M in (128, 512, 2048):
        tc = chunks_inside_cells(ks, M)
        report(f"chunks of 32 inside cells <= {M}", run(pos, tree, perm, tc, th, eps, stride))
    ho = hilbert_order(keys)
    report("hilbert-consecutive 32", run(pos, tree, ho, t32, th, eps, stride))


if __name__ == "__main__":
    main()
