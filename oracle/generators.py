"""CPU twin of the device initial-condition generators (csrc/generate.cu), in numpy.

TEST INFRASTRUCTURE ONLY (like the rest of oracle/): the product generates on the GPU through
libb200sim.so; this module restates the same laws with the same counter-based random streams so the
GPU tests can compare the device output element by element, and the CPU tests can check the laws
statistically against the reference's generators (tools/presets.py:91-1390, unseeded numpy
RandomState, some with per-body Python loops) through the golden quantile tables of
tests/golden/generators_ref.npz (made by tests/golden/make_golden_generators.py).

Every distribution cites the reference lines whose density / velocity law it follows.  What is NOT
reproduced is the reference's random stream: body i's draws come from Philox4x32-10 with
counter (i, k, 0, 0) and key (seed low, seed high), k = the draw index listed in each law, so any body
can be generated independently (on the device: one thread per body, no state).

Draw k gives two uniforms in (0, 1): u = ((x0 >> 5) * 2^26 + (x1 >> 6) + 0.5) / 2^53 (x2, x3 likewise),
or two standard normals by Box-Muller from those two uniforms.
"""
from __future__ import annotations

import numpy as np

DISTRIBUTIONS = [
    "galaxy", "collision", "spiral", "sphere", "ring", "shell", "cluster", "binary", "elliptical", "bar",
    "stream", "filament", "explosion", "disc", "vortex", "cube", "pleiades", "double_helix", "accretion_disk",
    "torus", "hourglass", "fibonacci", "triple", "rosette", "dyson",
]   # order of tools/presets.py:23-49; the index is the C ABI's distribution id

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
_LO = np.uint64(0xFFFFFFFF)
TWO_PI = 2.0 * np.pi


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al., SC'11) on uint64 arrays holding 32-bit words."""
    c0, c1, c2, c3 = (np.asarray(c, np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & _LO, p1 & _LO, ((p0 >> np.uint64(32)) ^ c3 ^ k1) & _LO, p0 & _LO
        k0, k1 = (k0 + _W0) & _LO, (k1 + _W1) & _LO
    return c0, c1, c2, c3


class Draws:
    """Draw k of bodies idx (global indices) -> two uniforms / normals per body."""

    def __init__(self, seed: int, idx, stream: int = 0):
        self.k0, self.k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
        self.idx = np.asarray(idx, np.uint64)
        self.stream = stream

    def u(self, k: int):
        x0, x1, x2, x3 = philox4x32(self.idx, np.full_like(self.idx, k), np.full_like(self.idx, self.stream),
                                    np.zeros_like(self.idx), self.k0, self.k1)
        f = lambda a, b: ((a >> np.uint64(5)).astype(np.float64) * 67108864.0 + (b >> np.uint64(6)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)
        return f(x0, x1), f(x2, x3)

    def n(self, k: int):
        a, b = self.u(k)
        rad = np.sqrt(-2.0 * np.log(a))
        return rad * np.cos(TWO_PI * b), rad * np.sin(TWO_PI * b)


def _ranks(r, groups):
    """rank[i] = 1-based position of body i in the stable ascending sort of r within its group:
    the enclosed body count of compute_rotation_curve (tools/presets.py:61-67, unit masses)."""
    rank = np.empty(len(r), np.float64)
    for b, e in groups:
        order = np.argsort(r[b:e], kind="stable")
        rk = np.empty(e - b, np.float64)
        rk[order] = np.arange(1, e - b + 1, dtype=np.float64)
        rank[b:e] = rk
    return rank


def _rotation_curve(r, rank, G, softening):
    """tools/presets.py:52-88 for unit masses: enclosed mass = rank."""
    eps2 = (2.0 * softening) ** 2
    r2 = r * r
    v = np.sqrt(G * rank * r2 / (r2 + eps2) ** 1.5)
    return v * np.maximum(r2 / (r2 + eps2), 0.3)


def _soft_disk_radius(u, scale, cap, rmin):
    """exponential radius with the soft cap of tools/presets.py:110-116."""
    r = -np.log(u) * scale
    r = r * (1.0 - np.exp(-cap / (r + 0.01)))
    return np.maximum(r, rmin)


def _iso(phi_u, ct_u):
    """isotropic unit vector in the reference's (sin t cos p, cos t, sin t sin p) convention."""
    ct = 2.0 * ct_u - 1.0
    st = np.sqrt(1.0 - ct * ct)
    ph = TWO_PI * phi_u
    return st * np.cos(ph), ct, st * np.sin(ph)


def _sub_mean(a, b, e):
    if e > b:
        a[b:e] -= a[b:e].mean(axis=0)


def filament_nodes(seed: int, R: float):
    """Node table of the cosmic web (tools/presets.py:609-650): the active nodes of the 8^3 grid, their
    cumulative power-law weights and an orthonormal frame per node.  Drawn from stream 1 (counter = node)."""
    gs = 8
    D = Draws(seed, np.arange(gs ** 3), stream=1)
    act_u, w_u = D.u(0)
    g = np.linspace(-1.25 * R, 1.25 * R, gs)
    ix, iy, iz = np.meshgrid(np.arange(gs), np.arange(gs), np.arange(gs), indexing="ij")
    centers = np.stack([g[ix.ravel()], g[iy.ravel()], g[iz.ravel()]], axis=1)
    active = act_u < 0.35
    if not active.any():
        active[0] = True
    w = np.sqrt(w_u)                       # numpy.random.power(2): pdf 2 x on (0, 1) = sqrt(uniform)
    a0, a1 = D.n(1)
    a2, b0 = D.n(2)
    b1, b2 = D.n(3)
    e = np.stack([a0, a1, a2], axis=1)
    e /= (np.linalg.norm(e, axis=1, keepdims=True) + 1e-10)
    p1 = np.stack([b0, b1, b2], axis=1)
    p1 -= (p1 * e).sum(axis=1, keepdims=True) * e
    p1 /= (np.linalg.norm(p1, axis=1, keepdims=True) + 1e-10)
    p2 = np.cross(e, p1)
    p2 /= (np.linalg.norm(p2, axis=1, keepdims=True) + 1e-10)
    sel = np.nonzero(active)[0]
    cw = np.cumsum(w[sel])
    cw /= cw[-1]
    return centers[sel], cw, e[sel], p1[sel], p2[sel]


def generate(distribution: str, n: int, R: float, G: float, seed: int = 0):
    """-> positions (n,3) f64, velocities (n,3) f64, masses (n) f64."""
    i = np.arange(n, dtype=np.int64)
    D = Draws(seed, i)
    pos = np.zeros((n, 3))
    vel = np.zeros((n, 3))
    mass = np.ones(n)
    X, Y, Z = 0, 1, 2

    if distribution in ("galaxy", "collision", "triple"):
        if distribution == "galaxy":            # tools/presets.py:104-146
            groups, scale_f, cap_f, soft_f, hgt, disp = [(0, n)], 0.3, 1.0, 0.03, 0.012, 0.12
        elif distribution == "collision":       # :148-232
            groups, scale_f, cap_f, soft_f, hgt, disp = [(0, n // 2), (n // 2, n)], 0.25, 0.5, 0.025, 0.01, 0.10
        else:                                   # triple :1147-1210
            t = n // 3
            groups, scale_f, cap_f, soft_f, hgt, disp = [(0, t), (t, 2 * t), (2 * t, n)], 0.20, 0.3, 0.02, 0.01, 0.12
        soft = R * soft_f
        ua, ub = D.u(0)
        r = _soft_disk_radius(ua, R * scale_f, R * cap_f, R * 0.001)
        th = TWO_PI * ub
        na, nb = D.n(1)
        ma, mb = D.n(2)
        rank = _ranks(r, groups)
        vc = _rotation_curve(r, rank, G, soft)
        ng = np.empty(n)
        for b, e in groups:
            ng[b:e] = e - b
        sigma = vc * disp * (r / (r + 2.0 * soft)) + np.sqrt(G * ng * 0.00005)
        if distribution == "triple":
            height = np.full(n, R * 0.01)       # z = normal(0, R * 0.01)
        else:
            height = R * hgt * (1.0 + np.sqrt(r / R) * 0.3)
        pos[:, X] = r * np.cos(th)
        pos[:, Y] = na * height
        pos[:, Z] = r * np.sin(th)
        spin = np.ones(n)
        if distribution == "collision":
            spin[n // 2:] = -1.0
        vel[:, X] = -spin * vc * np.sin(th) + ma * sigma
        vel[:, Z] = spin * vc * np.cos(th) + mb * sigma
        vel[:, Y] = nb * sigma * 0.25
        if distribution == "galaxy":
            _sub_mean(vel, 0, n)
        elif distribution == "collision":
            sep = R * 0.5 * 3.5
            speed = np.sqrt(2.0 * G * (n * 0.001) / sep) * 0.6
            h = n // 2
            pos[:h, X] -= sep / 2
            pos[h:, X] += sep / 2
            pos[h:, Y] += R * 0.15
            vel[:h, X] += speed
            vel[h:, X] -= speed
        else:
            sep = R * 0.8
            common = np.sqrt(G * (n * 0.001) / (sep * np.sqrt(3.0)))
            for gi, (b, e) in enumerate(groups):
                cx, cz = sep * np.cos(gi * TWO_PI / 3.0), sep * np.sin(gi * TWO_PI / 3.0)
                pos[b:e, X] += cx
                pos[b:e, Z] += cz
                vel[b:e, X] += -common * cz / sep
                vel[b:e, Z] += common * cx / sep
            _sub_mean(vel, 0, n)

    elif distribution == "spiral":              # tools/presets.py:234-298
        soft = R * 0.03
        ua, ub = D.u(0)
        r = _soft_disk_radius(ua, R * 0.3, R, R * 0.001)
        arm = np.floor(ub * 4.0)
        na, nb = D.n(1)
        ma, mb = D.n(2)
        ca, _ = D.n(3)
        th = -np.log(r / (R * 0.02) + 1.0) / 0.35 + arm * (TWO_PI / 4.0) + ca * (0.12 + 0.15 * np.sqrt(r / R))
        pos[:, X] = r * np.cos(th)
        pos[:, Z] = r * np.sin(th)
        pos[:, Y] = na * (R * 0.012 * (1.0 + np.sqrt(r / R) * 0.3))
        vc = _rotation_curve(r, _ranks(r, [(0, n)]), G, soft)
        vc = np.maximum(vc, np.sqrt(G * (n * 0.001) / (r + soft)) * 0.7)
        pt = np.arctan2(pos[:, Z], pos[:, X])
        sigma = vc * 0.10 * (r / (r + 2.0 * soft)) + np.sqrt(G * n * 0.00005)
        vel[:, X] = -vc * np.sin(pt) + ma * sigma
        vel[:, Z] = vc * np.cos(pt) + mb * sigma
        vel[:, Y] = nb * sigma * 0.25
        _sub_mean(vel, 0, n)

    elif distribution == "sphere":              # tools/presets.py:1379-1390 (the reference's own radius law)
        ua, ub = D.u(0)
        uc, _ = D.u(1)
        dx, dy, dz = _iso(ua, ub)
        r = (uc * R) ** (1.0 / 3.0) * R
        pos[:] = np.stack([r * dx, r * dy, r * dz], axis=1)
        na, nb = D.n(2)
        nc, _ = D.n(3)
        vel[:] = np.stack([na, nb, nc], axis=1) * 0.5

    elif distribution == "ring":                # tools/presets.py:300-327
        cn = n // 10
        ua, ub = D.u(0)
        uc, _ = D.u(1)
        na, _ = D.n(2)
        core = i < cn
        dx, dy, dz = _iso(ua, ub)
        rc = -np.log(uc) * (R * 0.05)
        rr = R * 0.4 + uc * (R * 0.4)
        th = TWO_PI * ua
        sp = np.sqrt(G * cn * 10 * 0.001 / rr)
        pos[:, X] = np.where(core, rc * dx, rr * np.cos(th))
        pos[:, Y] = np.where(core, rc * dy, na * (R * 0.01))
        pos[:, Z] = np.where(core, rc * dz, rr * np.sin(th))
        vel[:, X] = np.where(core, 0.0, -sp * np.sin(th))
        vel[:, Z] = np.where(core, 0.0, sp * np.cos(th))
        mass[:] = np.where(core, 10.0, 1.0)

    elif distribution == "shell":               # tools/presets.py:329-348
        ua, ub = D.u(0)
        uc, _ = D.u(1)
        ri, ro = R * 0.7, R * 0.9
        r = (ri ** 3 + uc * (ro ** 3 - ri ** 3)) ** (1.0 / 3.0)
        dx, dy, dz = _iso(ua, ub)
        pos[:] = np.stack([r * dx, r * dy, r * dz], axis=1)
        vel[:] = pos * 0.01

    elif distribution in ("cluster", "elliptical"):
        ua, ub = D.u(0)
        uc, ud = D.u(1)
        ue, _ = D.u(2)
        na, _ = D.n(3)
        dx, dy, dz = _iso(ua, ub)
        tm = n * 0.001
        if distribution == "cluster":           # tools/presets.py:350-397 (Plummer)
            a = R * 0.3
            r = np.clip(a / np.sqrt(uc ** (-2.0 / 3.0) - 1.0), 0.0, R * 1.5)
            pos[:] = np.stack([r * dx, r * dy, r * dz], axis=1)
            s2 = G * tm / (6.0 * a)
            sigma = np.sqrt(np.maximum(s2 * (1.0 + (r / a) ** 2) ** (-0.5), s2 * 0.01))
        else:                                   # elliptical :475-534
            a, b, c = R * 0.5, R * 0.4, R * 0.3
            r = np.clip(-np.log(uc) * (R * 0.2), 0.0, R * 0.9)
            pos[:] = np.stack([a * r / R * dx, b * r / R * dy, c * r / R * dz], axis=1)
            reff = np.sqrt((pos[:, X] / a) ** 2 + (pos[:, Y] / b) ** 2 + (pos[:, Z] / c) ** 2) * R
            frac = np.clip((reff / (R * 0.9)) ** 1.5, 0.01, 1.0)
            sigma = np.sqrt(np.maximum(G * tm * frac / (reff + R * 0.05), G * tm / (R * 10.0)))
        vm = np.abs(na * sigma * np.sqrt(3.0))
        vx, vy, vz = _iso(ud, ue)
        vel[:] = np.stack([vm * vx, vm * vy, vm * vz], axis=1)
        _sub_mean(vel, 0, n)

    elif distribution == "binary":              # tools/presets.py:399-473
        n1 = n // 2
        n2 = n - n1
        g2 = i >= n1
        tm = n * 0.001
        sep = R * 0.5
        bspeed = np.sqrt(G * tm / sep)
        ua, ub = D.u(0)
        na, nb = D.n(1)
        nc, nd = D.n(2)
        r = np.clip(-np.log(ua) * (R * 0.12), R * 0.01, R * 0.25)
        th = TWO_PI * ub
        tilt = np.pi / 6.0
        sm = np.where(g2, n2, n1) * 0.001
        sp = np.sqrt(G * sm / (r + R * 0.01))
        pos[:, X] = r * np.cos(th) + np.where(g2, sep / 2, -sep / 2)
        pos[:, Y] = np.where(g2, r * np.sin(th) * np.sin(tilt), na * (R * 0.008))
        pos[:, Z] = np.where(g2, r * np.sin(th) * np.cos(tilt), r * np.sin(th))
        sigma = np.sqrt(G * (n1 * 0.001) / (R * 0.1)) * 0.05
        vel[:, X] = -sp * np.sin(th) + nb * sigma
        vel[:, Y] = np.where(g2, sp * np.cos(th) * np.sin(tilt), 0.0) + nc * sigma
        vel[:, Z] = np.where(g2, sp * np.cos(th) * np.cos(tilt) + bspeed * (n1 / n), sp * np.cos(th) - bspeed * (n2 / n)) + nd * sigma
        _sub_mean(vel, 0, n)

    elif distribution == "bar":                 # tools/presets.py:536-592
        bn = n // 3
        isbar = i < bn
        soft = R * 0.025
        ua, ub = D.u(0)
        uc, _ = D.u(1)
        na, nb = D.n(2)
        nc, nd = D.n(3)
        ne, _ = D.n(4)
        blen = R * 0.4
        br = np.clip(-np.log(ua) * (blen * 0.3), R * 0.01, blen)
        bth = (ub * 2.0 - 1.0) * (np.pi / 6.0)
        dr = np.clip(-np.log(ua) * (R * 0.3), R * 0.25, R * 0.85)
        arm = np.floor(uc * 2.0)
        dth = np.log(dr / (R * 0.1) + 1.0) / 0.4 + arm * np.pi + ne * 0.25
        r = np.where(isbar, br, dr)
        th = np.where(isbar, bth, dth)
        rank = _ranks(r, [(0, bn), (bn, n)])
        sp = _rotation_curve(r, rank, G, soft)
        sigma = sp * 0.12 * (r / (r + 2.0 * soft))
        pos[:, X] = r * np.cos(th)
        pos[:, Y] = na * np.where(isbar, R * 0.02, R * 0.01)
        pos[:, Z] = r * np.sin(th) * np.where(isbar, 0.3, 1.0)
        vel[:, X] = -sp * np.sin(th) + nb * sigma
        vel[:, Y] = nc * sigma * np.where(isbar, 0.3, 0.25)
        vel[:, Z] = sp * np.cos(th) + nd * sigma
        _sub_mean(vel, 0, n)

    elif distribution == "stream":              # tools/presets.py:594-607
        t, _ = D.u(0)
        na, nb = D.n(1)
        nc, nd = D.n(2)
        ne, _ = D.n(3)
        pos[:, X] = (t - 0.5) * (R * 3.0)
        pos[:, Y] = np.sin(t * 4.0 * np.pi) * R * 0.3 + na * (R * 0.03)
        pos[:, Z] = np.cos(t * 4.0 * np.pi) * R * 0.3 + nb * (R * 0.03)
        vel[:, X] = 5.0 + nc * 0.5
        vel[:, Y] = nd * 0.3
        vel[:, Z] = ne * 0.3

    elif distribution == "filament":            # tools/presets.py:609-693
        centers, cw, e, p1, p2 = filament_nodes(seed, R)
        spacing = R * 2.5 / 8
        ua, _ = D.u(0)
        na, nb = D.n(1)
        nc, nd = D.n(2)
        ne, nf = D.n(3)
        node = np.minimum(np.searchsorted(cw, ua, side="right"), len(cw) - 1)
        par = na * (spacing * 0.8)
        q1 = nb * (spacing * 0.12)
        q2 = nc * (spacing * 0.12)
        pos[:] = centers[node] + par[:, None] * e[node] + q1[:, None] * p1[node] + q2[:, None] * p2[node]
        vel[:] = pos * 0.05 + np.stack([nd, ne, nf], axis=1) * 0.3
        mass[:] = 0.1

    elif distribution == "explosion":           # tools/presets.py:695-744
        cn = int(n * 0.15)
        core = i < cn
        ua, ub = D.u(0)
        uc, ud = D.u(1)
        na, nb = D.n(2)
        nc, _ = D.n(3)
        dx, dy, dz = _iso(ua, ub)
        r = np.where(core, np.clip(-np.log(uc) * (R * 0.02), 0.0, R * 0.05), R * 0.05 + uc * (R * 0.2))
        pos[:] = np.stack([r * dx, r * dy, r * dz], axis=1)
        dist = np.sqrt((pos ** 2).sum(axis=1)) + 0.01
        speed = 8.0 * (1.0 + (dist / R) * 2.0) + (-np.log(ud)) * 3.0
        asym = 1.0 + np.stack([na, nb, nc], axis=1) * 0.15
        vel[:] = pos / dist[:, None] * speed[:, None] * asym * np.where(core, 0.6, 1.0)[:, None]
        mass[:] = np.where(core, 2.0, 0.5)

    elif distribution == "disc":                # tools/presets.py:746-760
        ua, ub = D.u(0)
        na, _ = D.n(1)
        r = -np.log(ua) * (R * 0.3)
        th = TWO_PI * ub
        z = na * (R * 0.1)
        pos[:] = np.stack([r * np.cos(th), z, r * np.sin(th)], axis=1)
        ts = 8.0 / (r / R + 0.2)
        vel[:] = np.stack([-ts * np.sin(th), 2.0 * np.sign(z), ts * np.cos(th)], axis=1)

    elif distribution == "vortex":              # tools/presets.py:762-825
        ua, ub = D.u(0)
        uc, _ = D.u(1)
        na, nb = D.n(2)
        nc, _ = D.n(3)
        z = (ua * 2.0 - 1.0) * (R * 0.7)
        hf = np.clip(1.0 - 0.5 * (np.abs(z) / (R * 0.7 + 0.01)) ** 1.5, 0.15, 1.0)
        r = -np.log(ub) * (R * 0.25) * hf
        th = TWO_PI * uc + z * 0.5 / R
        pos[:] = np.stack([r * np.cos(th), z, r * np.sin(th)], axis=1)
        soft = R * 0.02
        sp = _rotation_curve(r, _ranks(r, [(0, n)]), G, soft)
        sp = np.maximum(sp, np.sqrt(G * n * 0.0001 / (r + soft)))
        sigma = sp * 0.03
        vel[:, X] = -sp * np.sin(th) + na * sigma
        vel[:, Z] = sp * np.cos(th) + nb * sigma
        vel[:, Y] = 0.05 * (r / R + 0.05) * sp * np.tanh(z / (R * 0.3)) + nc * sigma * 0.15
        _sub_mean(vel, 0, n)

    elif distribution == "cube":                # tools/presets.py:827-835
        side = int(np.ceil(n ** (1.0 / 3.0)))
        while side ** 3 < n:
            side += 1
        g = np.stack([i // (side * side), (i // side) % side, i % side], axis=1).astype(np.float64)
        pos[:] = (g - side / 2) * (R * 2 / side)
        na, nb = D.n(0)
        nc, _ = D.n(1)
        vel[:] = np.stack([na, nb, nc], axis=1) * 0.1

    elif distribution == "pleiades":            # tools/presets.py:837-866
        cn = n // 5
        core = i < cn
        ua, ub = D.u(0)
        uc, _ = D.u(1)
        na, nb = D.n(2)
        nc, _ = D.n(3)
        dx, dy, dz = _iso(ua, ub)
        r = np.where(core, -np.log(uc) * (R * 0.1), -np.log(uc) * (R * 0.5) + R * 0.1)
        pos[:] = np.stack([r * dx, r * dy * np.where(core, 1.0, 0.5), r * dz], axis=1)
        mass[:] = np.where(core, 5.0, 1.0)
        sigma = np.sqrt(G * cn * 5 * 0.001 / (R * 0.2))
        vel[:] = np.stack([na, nb, nc], axis=1) * (sigma * 0.5)

    elif distribution == "double_helix":        # tools/presets.py:868-905
        half = n // 2
        t = i * (6.0 * np.pi / max(n - 1, 1))
        ph = np.where(i < half, 0.0, np.pi)
        na, nb = D.n(0)
        nc, nd = D.n(1)
        radius, pitch = R * 0.25, R * 2.0
        pos[:, X] = radius * np.cos(t + ph) + na * (R * 0.01)
        pos[:, Y] = (t / (6.0 * np.pi)) * pitch - pitch / 2 + nb * (R * 0.01)
        pos[:, Z] = radius * np.sin(t + ph) + nc * (R * 0.01)
        omega = 0.08
        m = np.sqrt(pos[:, X] ** 2 + pos[:, Z] ** 2) > 0.01
        vel[:, X] = np.where(m, -omega * pos[:, Z], 0.0)
        vel[:, Z] = np.where(m, omega * pos[:, X], 0.0)
        vel[:, Y] = nd * (omega * 0.2)

    elif distribution == "accretion_disk":      # tools/presets.py:907-978
        cn = max(1, n // 100)
        dn = int((n - cn) * 0.85)
        jn = n - cn - dn
        jh = jn // 2
        ua, ub = D.u(0)
        uc, _ = D.u(1)
        na, nb = D.n(2)
        nc, nd = D.n(3)
        ne, nf = D.n(4)
        cen = i < cn
        dsk = (i >= cn) & (i < cn + dn)
        up = (i >= cn + dn) & (i < cn + dn + jh)
        r = np.clip(-np.log(ua) * (R * 0.2), R * 0.05, R * 0.8)
        th = TWO_PI * ub
        vk = np.sqrt(G * 1000.0 / (r + R * 0.05))
        zj = R * 0.2 + uc * (R * 1.0)
        rj = -np.log(ua) * (R * 0.05)
        pos[:, X] = np.where(cen, na * (R * 0.02), np.where(dsk, r * np.cos(th), rj * np.cos(th)))
        pos[:, Y] = np.where(cen, nb * (R * 0.02), np.where(dsk, nc * (R * 0.01), np.where(up, zj, -zj)))
        pos[:, Z] = np.where(cen, nc * (R * 0.02), np.where(dsk, r * np.sin(th), rj * np.sin(th)))
        vel[:, X] = np.where(cen, nd * 0.1, np.where(dsk, -vk * np.sin(th), 0.0))
        vel[:, Y] = np.where(cen, ne * 0.1, np.where(dsk, 0.0, np.where(up, 3.0, -3.0)))
        vel[:, Z] = np.where(cen, nf * 0.1, np.where(dsk, vk * np.cos(th), 0.0))
        mass[:] = np.where(cen, 200.0, np.where(dsk, 0.5, 0.1))
        _sub_mean(pos, 0, cn)
        _sub_mean(vel, 0, cn)

    elif distribution == "torus":               # tools/presets.py:980-1017
        ua, ub = D.u(0)
        na, nb = D.n(1)
        nc, nd = D.n(2)
        major, minor = R * 0.6, R * 0.25
        u, v = TWO_PI * ua, TWO_PI * ub
        rn = 1.0 + na * 0.1
        ring = major + minor * np.cos(u) * rn
        pos[:] = np.stack([ring * np.cos(v), minor * np.sin(u) * rn, ring * np.sin(v)], axis=1)
        rxy = np.sqrt(pos[:, X] ** 2 + pos[:, Z] ** 2)
        omega = np.sqrt(G * n * 0.001 / major)
        m = rxy > 0.01
        safe = np.where(m, rxy, 1.0)
        vel[:, X] = np.where(m, -omega * pos[:, Z] / safe, 0.0) + nb * (omega * 0.05)
        vel[:, Y] = nc * (omega * 0.05)
        vel[:, Z] = np.where(m, omega * pos[:, X] / safe, 0.0) + nd * (omega * 0.05)

    elif distribution == "hourglass":           # tools/presets.py:1019-1111
        bn = max(2, n // 200)
        nn = n - bn
        half = nn // 2
        b1 = bn // 2
        ua, ub = D.u(0)
        na, nb = D.n(1)
        nc, nd = D.n(2)
        ne, nf = D.n(3)
        ng, _ = D.n(4)
        star = i < bn
        s1 = i < b1
        upper = (i >= bn) & (i < bn + half)
        bsep = R * 0.05
        vb = np.sqrt(G * 250.0 / bsep)
        zc = np.where(upper, ua * R, -ua * R)
        rc = np.abs(zc) * 0.5 * (1.0 + na * 0.1)
        th = TWO_PI * ub
        pos[:, X] = np.where(star, np.where(s1, -bsep / 2, bsep / 2) + na * (R * 0.01), rc * np.cos(th))
        pos[:, Y] = np.where(star, nb * (R * 0.01), zc)
        pos[:, Z] = np.where(star, nc * (R * 0.01), rc * np.sin(th))
        vel[:, Y] = np.where(star, nd * 0.05, 0.0)
        vel[:, Z] = np.where(star, np.where(s1, vb, -vb) + ne * 0.05, 0.0)
        _sub_mean(pos, 0, bn)
        _sub_mean(vel, 0, bn)
        rxy = np.sqrt(pos[:, X] ** 2 + pos[:, Z] ** 2)
        r3 = np.sqrt((pos ** 2).sum(axis=1))
        vo = np.sqrt(G * 500.0 / (r3 + R * 0.05))
        m = rxy > 0.01
        safe = np.where(m, rxy, 1.0)
        neb = ~star
        vel[:, X] = np.where(neb, np.where(m, -vo * pos[:, Z] / safe, 0.0) + nd * 0.08, vel[:, X])
        vel[:, Y] = np.where(neb, nb * (vo * (r3 / R) * 0.08) + ne * 0.08, vel[:, Y])
        vel[:, Z] = np.where(neb, np.where(m, vo * pos[:, X] / safe, 0.0) + nf * 0.08, vel[:, Z])
        mass[:] = np.where(star, 100.0, 0.1)

    elif distribution == "fibonacci":           # tools/presets.py:1113-1145
        golden = (1.0 + np.sqrt(5.0)) / 2.0
        th = i * (TWO_PI / (golden ** 2))
        r = np.where(i > 0, R * np.sqrt(i / n), R * 0.01)
        na, nb = D.n(0)
        nc, _ = D.n(1)
        pos[:] = np.stack([r * np.cos(th), (i / n - 0.5) * R * 2.0, r * np.sin(th)], axis=1)
        vo = np.where(r > 0.01, np.sqrt(G * (n * 0.001) / (r + R * 0.05)), 0.0)
        vel[:] = np.stack([-vo * np.sin(th) + na * 0.05, nb * 0.05, vo * np.cos(th) + nc * 0.05], axis=1)

    elif distribution == "rosette":             # tools/presets.py:1212-1258
        ps = n // 5
        petal = np.minimum(i // ps, 4) if ps > 0 else np.full(n, 4)
        ang = petal * (TWO_PI / 5.0)
        ua, ub = D.u(0)
        na, nb = D.n(1)
        nc, nd = D.n(2)
        r = -np.log(ua) * (R * 0.25)
        th = TWO_PI * ub
        xl, zl = r * np.cos(th), r * np.sin(th) * 0.3
        pos[:] = np.stack([xl * np.cos(ang) - zl * np.sin(ang), na * (R * 0.02), xl * np.sin(ang) + zl * np.cos(ang)], axis=1)
        rxy = np.sqrt(pos[:, X] ** 2 + pos[:, Z] ** 2)
        r3 = np.sqrt((pos ** 2).sum(axis=1))
        om = 0.5 * np.sqrt(R * 0.3 / (r3 + R * 0.05))
        m = rxy > 0.01
        safe = np.where(m, rxy, 1.0)
        vel[:, X] = np.where(m, -om * pos[:, Z] / safe, 0.0) + nb * 0.05
        vel[:, Y] = nc * 0.05
        vel[:, Z] = np.where(m, om * pos[:, X] / safe, 0.0) + nd * 0.05

    elif distribution == "dyson":               # tools/presets.py:1260-1377
        cn = max(1, n // 200)
        cen = i < cn
        ua, ub = D.u(0)
        na, nb = D.n(1)
        nc, nd = D.n(2)
        ne, nf = D.n(3)
        dx, dy, dz = _iso(ua, ub)
        r = R * 0.7 + na * (R * 0.03)
        pos[:, X] = np.where(cen, na * (R * 0.01), r * dx)
        pos[:, Y] = np.where(cen, nb * (R * 0.01), r * dy)
        pos[:, Z] = np.where(cen, nc * (R * 0.01), r * dz)
        vel[:, X] = np.where(cen, nd * 0.05, 0.0)
        vel[:, Y] = np.where(cen, ne * 0.05, 0.0)
        vel[:, Z] = np.where(cen, nf * 0.05, 0.0)
        mass[:] = np.where(cen, 500.0, 0.1)
        _sub_mean(pos, 0, cn)
        _sub_mean(vel, 0, cn)
        # enclosed mass = the central bodies + the shell bodies at radius <= r (tools/presets.py:1310-1324)
        # -- and the reference then maps the per-body array "back to original order" although it already is in
        # original order (:1326-1328), so body i ends up with the enclosed mass of the shell body whose index is
        # i's rank; the law is reproduced as written (it widens the speed distribution measurably)
        rk = _ranks(r[cn:], [(0, n - cn)])
        rank = np.zeros(n)
        rank[cn:] = rk[rk.astype(np.int64) - 1]
        menc = 500.0 * cn + 0.1 * rank
        vo = np.sqrt(G * menc / (r + R * 0.01))
        rm = np.sqrt((pos ** 2).sum(axis=1))
        valid = rm > 0.01
        ru = pos / np.where(valid, rm, 1.0)[:, None]
        # tangent = radial x Y axis (poles: radial x X axis), tools/presets.py:1336-1353
        tx, ty, tz = -ru[:, Z], np.zeros(n), ru[:, X]
        tmag = np.sqrt(tx * tx + tz * tz)
        pole = tmag < 0.01
        tx, ty, tz = np.where(pole, 0.0, tx), np.where(pole, ru[:, Z], ty), np.where(pole, -ru[:, Y], tz)
        tmag = np.where(pole, np.sqrt(ty * ty + tz * tz), tmag)
        tu = np.stack([tx, ty, tz], axis=1) / (tmag[:, None] + 1e-10)
        sv = vo[:, None] * tu
        # small component along radial x velocity (:1365-1376)
        vert = np.cross(pos, sv)
        vmag = np.sqrt((vert ** 2).sum(axis=1))
        addv = np.where((vmag > 0.01)[:, None], vert / np.where(vmag > 0.01, vmag, 1.0)[:, None] * (nb * vo * 0.01)[:, None], 0.0)
        shell_v = np.where(valid[:, None], sv + addv, np.stack([nc, nd, ne], axis=1) * 0.01)
        vel[:] = np.where(cen[:, None], vel, shell_v)

    else:
        raise ValueError(f"unknown distribution {distribution!r}")
    return np.ascontiguousarray(pos), np.ascontiguousarray(vel), mass
