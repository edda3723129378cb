// generate.cu -- seeded initial-condition generators on the device.
//
// Replaces generate_distribution(distribution, n, R, G) of the reference (tools/presets.py:91-1390; called by
// tools/record.py and nbody_main.py): 25 density / velocity laws, there drawn from numpy's unseeded global
// RandomState, several with per-body Python loops (cluster :380-393, elliptical :520-530, torus :1004-1014,
// fibonacci :1121-1143, rosette :1246-1256) and one O(n^2) (dyson :1314-1318) -- unusable at 50 M bodies.
// Here body i is one thread: its draws come from Philox4x32-10 with counter (i, k, stream, 0) and key = seed
// (k = the draw index), so a body is generated without any state, any slice of the bodies can be generated
// on any GPU, and the result is reproducible.  The laws that need a body's rank in radius (the enclosed mass of
// compute_rotation_curve, tools/presets.py:52-88, and dyson's enclosed shell mass) get it from the library's
// radix sort on the bit pattern of the radius.  The same laws with the same streams are restated in numpy in
// oracle/generators.py (test infrastructure), which the CPU tests check against the reference's generators.
#include "nbody.cuh"
#include <cmath>
#include <cstring>
#include <vector>

namespace b200 {

namespace {

constexpr double TWO_PI = 6.283185307179586476925286766559;
constexpr double PI = 3.141592653589793238462643383279;

enum Dist {
    D_GALAXY = 0, D_COLLISION, D_SPIRAL, D_SPHERE, D_RING, D_SHELL, D_CLUSTER, D_BINARY, D_ELLIPTICAL, D_BAR,
    D_STREAM, D_FILAMENT, D_EXPLOSION, D_DISC, D_VORTEX, D_CUBE, D_PLEIADES, D_DOUBLE_HELIX, D_ACCRETION_DISK,
    D_TORUS, D_HOURGLASS, D_FIBONACCI, D_TRIPLE, D_ROSETTE, D_DYSON, D_COUNT
};
const char* const DIST_NAMES[D_COUNT] = {
    "galaxy", "collision", "spiral", "sphere", "ring", "shell", "cluster", "binary", "elliptical", "bar",
    "stream", "filament", "explosion", "disc", "vortex", "cube", "pleiades", "double_helix", "accretion_disk",
    "torus", "hourglass", "fibonacci", "triple", "rosette", "dyson"};

struct GenParams {
    int dist;
    int64_t n;
    double R, G;
    uint32_t k0, k1;
    const uint32_t* rank;      // 1-based rank in radius inside the body's group (rotation-curve laws), or null
    // filament node table
    int nodes;
    const double* node_tab;    // [nodes] x {cx, cy, cz, cw, e[3], p1[3], p2[3]} = 13 doubles per node
    int64_t cube_side;
};

struct Body { double p[3], v[3], m, rkey; };

__host__ __device__ inline void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Rng {
    uint32_t k0, k1, stream;
    uint64_t i;
    __host__ __device__ void u(uint32_t k, double& a, double& b) const
    {
        uint32_t x[4];
        philox4x32((uint32_t)i, k, stream, 0u, k0, k1, x);
        a = ((double)(x[0] >> 5) * 67108864.0 + (double)(x[1] >> 6) + 0.5) * (1.0 / 9007199254740992.0);
        b = ((double)(x[2] >> 5) * 67108864.0 + (double)(x[3] >> 6) + 0.5) * (1.0 / 9007199254740992.0);
    }
    __host__ __device__ void n(uint32_t k, double& a, double& b) const
    {
        double ua, ub;
        u(k, ua, ub);
        const double rad = sqrt(-2.0 * log(ua));
        a = rad * cos(TWO_PI * ub);
        b = rad * sin(TWO_PI * ub);
    }
};

__host__ __device__ inline double clampd(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

// tools/presets.py:52-88 for unit masses (enclosed mass = rank)
__device__ inline double rotation_curve(double r, double rank, double G, double softening)
{
    const double eps2 = (2.0 * softening) * (2.0 * softening), r2 = r * r;
    const double v = sqrt(G * rank * r2 / pow(r2 + eps2, 1.5));
    return v * fmax(r2 / (r2 + eps2), 0.3);
}
// exponential radius with the soft cap of tools/presets.py:110-116
__device__ inline double soft_disk_radius(double u, double scale, double cap, double rmin)
{
    double r = -log(u) * scale;
    r = r * (1.0 - exp(-cap / (r + 0.01)));
    return fmax(r, rmin);
}
// isotropic unit vector, the reference's (sin t cos p, cos t, sin t sin p)
__device__ inline void iso(double phi_u, double ct_u, double d[3])
{
    const double ct = 2.0 * ct_u - 1.0, st = sqrt(1.0 - ct * ct), ph = TWO_PI * phi_u;
    d[0] = st * cos(ph); d[1] = ct; d[2] = st * sin(ph);
}

// One body of one law.  Mirrors oracle/generators.py::generate branch by branch (same draw indices).
__device__ void gen_body(const GenParams& P, int64_t i, Body& o)
{
    const double R = P.R, G = P.G;
    const int64_t n = P.n;
    const Rng D{P.k0, P.k1, 0u, (uint64_t)i};
    double* p = o.p;
    double* v = o.v;
    p[0] = p[1] = p[2] = v[0] = v[1] = v[2] = 0.0;
    o.m = 1.0;
    o.rkey = 0.0;
    const double rank = P.rank ? (double)P.rank[i] : 1.0;
    double ua, ub, uc, ud, ue, uf, na, nb, nc, nd, ne, nf, ng, nh;
    (void)uf; (void)nh; (void)ng;
    switch (P.dist) {
    case D_GALAXY: case D_COLLISION: case D_TRIPLE: {   // tools/presets.py:104-146, :148-232, :1147-1210
        double scale_f, cap_f, soft_f, hgt, disp, ngrp, spin = 1.0;
        int grp = 0;
        if (P.dist == D_GALAXY) { scale_f = 0.3; cap_f = 1.0; soft_f = 0.03; hgt = 0.012; disp = 0.12; ngrp = (double)n; }
        else if (P.dist == D_COLLISION) {
            scale_f = 0.25; cap_f = 0.5; soft_f = 0.025; hgt = 0.01; disp = 0.10;
            const int64_t h = n / 2;
            grp = i >= h; ngrp = (double)(grp ? n - h : h); spin = grp ? -1.0 : 1.0;
        } else {
            scale_f = 0.20; cap_f = 0.3; soft_f = 0.02; hgt = 0.01; disp = 0.12;
            const int64_t t = n / 3;
            grp = t > 0 ? (int)(i / t < 2 ? i / t : 2) : 2; ngrp = (double)(grp < 2 ? t : n - 2 * t);
        }
        const double soft = R * soft_f;
        D.u(0, ua, ub);
        const double r = soft_disk_radius(ua, R * scale_f, R * cap_f, R * 0.001), th = TWO_PI * ub;
        o.rkey = r;
        D.n(1, na, nb);
        D.n(2, nc, nd);
        const double vc = rotation_curve(r, rank, G, soft);
        const double sigma = vc * disp * (r / (r + 2.0 * soft)) + sqrt(G * ngrp * 0.00005);
        const double height = P.dist == D_TRIPLE ? R * 0.01 : R * hgt * (1.0 + sqrt(r / R) * 0.3);
        p[0] = r * cos(th); p[1] = na * height; p[2] = r * sin(th);
        v[0] = -spin * vc * sin(th) + nc * sigma;
        v[2] = spin * vc * cos(th) + nd * sigma;
        v[1] = nb * sigma * 0.25;
        if (P.dist == D_COLLISION) {
            const double sep = R * 0.5 * 3.5, speed = sqrt(2.0 * G * ((double)n * 0.001) / sep) * 0.6;
            if (grp) { p[0] += sep / 2; p[1] += R * 0.15; v[0] -= speed; } else { p[0] -= sep / 2; v[0] += speed; }
        } else if (P.dist == D_TRIPLE) {
            const double sep = R * 0.8, common = sqrt(G * ((double)n * 0.001) / (sep * sqrt(3.0)));
            const double cx = sep * cos(grp * TWO_PI / 3.0), cz = sep * sin(grp * TWO_PI / 3.0);
            p[0] += cx; p[2] += cz;
            v[0] += -common * cz / sep; v[2] += common * cx / sep;
        }
        break;
    }
    case D_SPIRAL: {                                   // tools/presets.py:234-298
        const double soft = R * 0.03;
        D.u(0, ua, ub);
        const double r = soft_disk_radius(ua, R * 0.3, R, R * 0.001);
        o.rkey = r;
        const double arm = floor(ub * 4.0);
        D.n(1, na, nb);
        D.n(2, nc, nd);
        D.n(3, ne, nf);
        const double th = -log(r / (R * 0.02) + 1.0) / 0.35 + arm * (TWO_PI / 4.0) + ne * (0.12 + 0.15 * sqrt(r / R));
        p[0] = r * cos(th); p[2] = r * sin(th);
        p[1] = na * (R * 0.012 * (1.0 + sqrt(r / R) * 0.3));
        double vc = rotation_curve(r, rank, G, soft);
        vc = fmax(vc, sqrt(G * ((double)n * 0.001) / (r + soft)) * 0.7);
        const double pt = atan2(p[2], p[0]);
        const double sigma = vc * 0.10 * (r / (r + 2.0 * soft)) + sqrt(G * (double)n * 0.00005);
        v[0] = -vc * sin(pt) + nc * sigma;
        v[2] = vc * cos(pt) + nd * sigma;
        v[1] = nb * sigma * 0.25;
        break;
    }
    case D_SPHERE: {                                   // tools/presets.py:1379-1390
        double d[3];
        D.u(0, ua, ub);
        D.u(1, uc, ud);
        iso(ua, ub, d);
        const double r = pow(uc * R, 1.0 / 3.0) * R;
        D.n(2, na, nb);
        D.n(3, nc, nd);
        p[0] = r * d[0]; p[1] = r * d[1]; p[2] = r * d[2];
        v[0] = na * 0.5; v[1] = nb * 0.5; v[2] = nc * 0.5;
        break;
    }
    case D_RING: {                                     // tools/presets.py:300-327
        const int64_t cn = n / 10;
        double d[3];
        D.u(0, ua, ub);
        D.u(1, uc, ud);
        D.n(2, na, nb);
        if (i < cn) {
            iso(ua, ub, d);
            const double rc = -log(uc) * (R * 0.05);
            p[0] = rc * d[0]; p[1] = rc * d[1]; p[2] = rc * d[2];
            o.m = 10.0;
        } else {
            const double rr = R * 0.4 + uc * (R * 0.4), th = TWO_PI * ua;
            const double sp = sqrt(G * (double)cn * 10 * 0.001 / rr);
            p[0] = rr * cos(th); p[1] = na * (R * 0.01); p[2] = rr * sin(th);
            v[0] = -sp * sin(th); v[2] = sp * cos(th);
        }
        break;
    }
    case D_SHELL: {                                    // tools/presets.py:329-348
        double d[3];
        D.u(0, ua, ub);
        D.u(1, uc, ud);
        const double ri = R * 0.7, ro = R * 0.9;
        const double r = pow(ri * ri * ri + uc * (ro * ro * ro - ri * ri * ri), 1.0 / 3.0);
        iso(ua, ub, d);
        for (int k = 0; k < 3; ++k) { p[k] = r * d[k]; v[k] = p[k] * 0.01; }
        break;
    }
    case D_CLUSTER: case D_ELLIPTICAL: {               // tools/presets.py:350-397 (Plummer), :475-534
        double d[3], w[3];
        D.u(0, ua, ub);
        D.u(1, uc, ud);
        D.u(2, ue, uf);
        D.n(3, na, nb);
        iso(ua, ub, d);
        const double tm = (double)n * 0.001;
        double sigma;
        if (P.dist == D_CLUSTER) {
            const double a = R * 0.3;
            const double r = clampd(a / sqrt(pow(uc, -2.0 / 3.0) - 1.0), 0.0, R * 1.5);
            for (int k = 0; k < 3; ++k) p[k] = r * d[k];
            const double s2 = G * tm / (6.0 * a);
            sigma = sqrt(fmax(s2 * pow(1.0 + (r / a) * (r / a), -0.5), s2 * 0.01));
        } else {
            const double a = R * 0.5, b = R * 0.4, c = R * 0.3;
            const double r = clampd(-log(uc) * (R * 0.2), 0.0, R * 0.9);
            p[0] = a * r / R * d[0]; p[1] = b * r / R * d[1]; p[2] = c * r / R * d[2];
            const double reff = sqrt((p[0] / a) * (p[0] / a) + (p[1] / b) * (p[1] / b) + (p[2] / c) * (p[2] / c)) * R;
            const double frac = clampd(pow(reff / (R * 0.9), 1.5), 0.01, 1.0);
            sigma = sqrt(fmax(G * tm * frac / (reff + R * 0.05), G * tm / (R * 10.0)));
        }
        const double vm = fabs(na * sigma * sqrt(3.0));
        iso(ud, ue, w);
        for (int k = 0; k < 3; ++k) v[k] = vm * w[k];
        break;
    }
    case D_BINARY: {                                   // tools/presets.py:399-473
        const int64_t n1 = n / 2, n2 = n - n1;
        const bool g2 = i >= n1;
        const double tm = (double)n * 0.001, sep = R * 0.5, bspeed = sqrt(G * tm / sep);
        D.u(0, ua, ub);
        D.n(1, na, nb);
        D.n(2, nc, nd);
        const double r = clampd(-log(ua) * (R * 0.12), R * 0.01, R * 0.25), th = TWO_PI * ub, tilt = PI / 6.0;
        const double sm = (double)(g2 ? n2 : n1) * 0.001;
        const double sp = sqrt(G * sm / (r + R * 0.01));
        p[0] = r * cos(th) + (g2 ? sep / 2 : -sep / 2);
        p[1] = g2 ? r * sin(th) * sin(tilt) : na * (R * 0.008);
        p[2] = g2 ? r * sin(th) * cos(tilt) : r * sin(th);
        const double sigma = sqrt(G * ((double)n1 * 0.001) / (R * 0.1)) * 0.05;
        v[0] = -sp * sin(th) + nb * sigma;
        v[1] = (g2 ? sp * cos(th) * sin(tilt) : 0.0) + nc * sigma;
        v[2] = (g2 ? sp * cos(th) * cos(tilt) + bspeed * ((double)n1 / (double)n) : sp * cos(th) - bspeed * ((double)n2 / (double)n)) + nd * sigma;
        break;
    }
    case D_BAR: {                                      // tools/presets.py:536-592
        const int64_t bn = n / 3;
        const bool isbar = i < bn;
        const double soft = R * 0.025;
        D.u(0, ua, ub);
        D.u(1, uc, ud);
        D.n(2, na, nb);
        D.n(3, nc, nd);
        D.n(4, ne, nf);
        const double blen = R * 0.4;
        double r, th;
        if (isbar) {
            r = clampd(-log(ua) * (blen * 0.3), R * 0.01, blen);
            th = (ub * 2.0 - 1.0) * (PI / 6.0);
        } else {
            r = clampd(-log(ua) * (R * 0.3), R * 0.25, R * 0.85);
            th = log(r / (R * 0.1) + 1.0) / 0.4 + floor(uc * 2.0) * PI + ne * 0.25;
        }
        o.rkey = r;
        const double sp = rotation_curve(r, rank, G, soft);
        const double sigma = sp * 0.12 * (r / (r + 2.0 * soft));
        p[0] = r * cos(th);
        p[1] = na * (isbar ? R * 0.02 : R * 0.01);
        p[2] = r * sin(th) * (isbar ? 0.3 : 1.0);
        v[0] = -sp * sin(th) + nb * sigma;
        v[1] = nc * sigma * (isbar ? 0.3 : 0.25);
        v[2] = sp * cos(th) + nd * sigma;
        break;
    }
    case D_STREAM: {                                   // tools/presets.py:594-607
        D.u(0, ua, ub);
        D.n(1, na, nb);
        D.n(2, nc, nd);
        D.n(3, ne, nf);
        const double t = ua;
        p[0] = (t - 0.5) * (R * 3.0);
        p[1] = sin(t * 4.0 * PI) * R * 0.3 + na * (R * 0.03);
        p[2] = cos(t * 4.0 * PI) * R * 0.3 + nb * (R * 0.03);
        v[0] = 5.0 + nc * 0.5; v[1] = nd * 0.3; v[2] = ne * 0.3;
        break;
    }
    case D_FILAMENT: {                                 // tools/presets.py:609-693
        const double spacing = R * 2.5 / 8;
        D.u(0, ua, ub);
        D.n(1, na, nb);
        D.n(2, nc, nd);
        D.n(3, ne, nf);
        int lo = 0, hi = P.nodes;                      // first node with cw > ua (searchsorted side = right)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (P.node_tab[13 * mid + 3] > ua) hi = mid; else lo = mid + 1;
        }
        const double* t = P.node_tab + 13 * (lo < P.nodes - 1 ? lo : P.nodes - 1);
        const double par = na * (spacing * 0.8), q1 = nb * (spacing * 0.12), q2 = nc * (spacing * 0.12);
        const double nz[3] = {nd, ne, nf};
        for (int k = 0; k < 3; ++k) {
            p[k] = t[k] + par * t[4 + k] + q1 * t[7 + k] + q2 * t[10 + k];
            v[k] = p[k] * 0.05 + nz[k] * 0.3;
        }
        o.m = 0.1;
        break;
    }
    case D_EXPLOSION: {                                // tools/presets.py:695-744
        const int64_t cn = (int64_t)((double)n * 0.15);
        const bool core = i < cn;
        double d[3];
        D.u(0, ua, ub);
        D.u(1, uc, ud);
        D.n(2, na, nb);
        D.n(3, nc, nd);
        iso(ua, ub, d);
        const double r = core ? clampd(-log(uc) * (R * 0.02), 0.0, R * 0.05) : R * 0.05 + uc * (R * 0.2);
        for (int k = 0; k < 3; ++k) p[k] = r * d[k];
        const double dist = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) + 0.01;
        const double speed = 8.0 * (1.0 + (dist / R) * 2.0) + (-log(ud)) * 3.0;
        const double nz[3] = {na, nb, nc};
        for (int k = 0; k < 3; ++k) v[k] = p[k] / dist * speed * (1.0 + nz[k] * 0.15) * (core ? 0.6 : 1.0);
        o.m = core ? 2.0 : 0.5;
        break;
    }
    case D_DISC: {                                     // tools/presets.py:746-760
        D.u(0, ua, ub);
        D.n(1, na, nb);
        const double r = -log(ua) * (R * 0.3), th = TWO_PI * ub, z = na * (R * 0.1);
        p[0] = r * cos(th); p[1] = z; p[2] = r * sin(th);
        const double ts = 8.0 / (r / R + 0.2);
        v[0] = -ts * sin(th); v[1] = 2.0 * (z > 0.0 ? 1.0 : (z < 0.0 ? -1.0 : 0.0)); v[2] = ts * cos(th);
        break;
    }
    case D_VORTEX: {                                   // tools/presets.py:762-825
        D.u(0, ua, ub);
        D.u(1, uc, ud);
        D.n(2, na, nb);
        D.n(3, nc, nd);
        const double z = (ua * 2.0 - 1.0) * (R * 0.7);
        const double hf = clampd(1.0 - 0.5 * pow(fabs(z) / (R * 0.7 + 0.01), 1.5), 0.15, 1.0);
        const double r = -log(ub) * (R * 0.25) * hf;
        o.rkey = r;
        const double th = TWO_PI * uc + z * 0.5 / R;
        p[0] = r * cos(th); p[1] = z; p[2] = r * sin(th);
        const double soft = R * 0.02;
        double sp = rotation_curve(r, rank, G, soft);
        sp = fmax(sp, sqrt(G * (double)n * 0.0001 / (r + soft)));
        const double sigma = sp * 0.03;
        v[0] = -sp * sin(th) + na * sigma;
        v[2] = sp * cos(th) + nb * sigma;
        v[1] = 0.05 * (r / R + 0.05) * sp * tanh(z / (R * 0.3)) + nc * sigma * 0.15;
        break;
    }
    case D_CUBE: {                                     // tools/presets.py:827-835
        const int64_t side = P.cube_side;
        const double g[3] = {(double)(i / (side * side)), (double)((i / side) % side), (double)(i % side)};
        D.n(0, na, nb);
        D.n(1, nc, nd);
        for (int k = 0; k < 3; ++k) p[k] = (g[k] - (double)side / 2) * (R * 2 / (double)side);
        v[0] = na * 0.1; v[1] = nb * 0.1; v[2] = nc * 0.1;
        break;
    }
    case D_PLEIADES: {                                 // tools/presets.py:837-866
        const int64_t cn = n / 5;
        const bool core = i < cn;
        double d[3];
        D.u(0, ua, ub);
        D.u(1, uc, ud);
        D.n(2, na, nb);
        D.n(3, nc, nd);
        iso(ua, ub, d);
        const double r = core ? -log(uc) * (R * 0.1) : -log(uc) * (R * 0.5) + R * 0.1;
        p[0] = r * d[0]; p[1] = r * d[1] * (core ? 1.0 : 0.5); p[2] = r * d[2];
        o.m = core ? 5.0 : 1.0;
        const double sigma = sqrt(G * (double)cn * 5 * 0.001 / (R * 0.2));
        v[0] = na * (sigma * 0.5); v[1] = nb * (sigma * 0.5); v[2] = nc * (sigma * 0.5);
        break;
    }
    case D_DOUBLE_HELIX: {                             // tools/presets.py:868-905
        const int64_t half = n / 2;
        const double t = (double)i * (6.0 * PI / (double)(n - 1 > 1 ? n - 1 : 1)), ph = i < half ? 0.0 : PI;
        D.n(0, na, nb);
        D.n(1, nc, nd);
        const double radius = R * 0.25, pitch = R * 2.0, omega = 0.08;
        p[0] = radius * cos(t + ph) + na * (R * 0.01);
        p[1] = (t / (6.0 * PI)) * pitch - pitch / 2 + nb * (R * 0.01);
        p[2] = radius * sin(t + ph) + nc * (R * 0.01);
        const bool m = sqrt(p[0] * p[0] + p[2] * p[2]) > 0.01;
        v[0] = m ? -omega * p[2] : 0.0;
        v[2] = m ? omega * p[0] : 0.0;
        v[1] = nd * (omega * 0.2);
        break;
    }
    case D_ACCRETION_DISK: {                           // tools/presets.py:907-978
        const int64_t cn = n / 100 > 1 ? n / 100 : 1;
        const int64_t dn = (int64_t)((double)(n - cn) * 0.85), jn = n - cn - dn, jh = jn / 2;
        D.u(0, ua, ub);
        D.u(1, uc, ud);
        D.n(2, na, nb);
        D.n(3, nc, nd);
        D.n(4, ne, nf);
        const double th = TWO_PI * ub;
        if (i < cn) {
            p[0] = na * (R * 0.02); p[1] = nb * (R * 0.02); p[2] = nc * (R * 0.02);
            v[0] = nd * 0.1; v[1] = ne * 0.1; v[2] = nf * 0.1;
            o.m = 200.0;
        } else if (i < cn + dn) {
            const double r = clampd(-log(ua) * (R * 0.2), R * 0.05, R * 0.8);
            const double vk = sqrt(G * 1000.0 / (r + R * 0.05));
            p[0] = r * cos(th); p[1] = nc * (R * 0.01); p[2] = r * sin(th);
            v[0] = -vk * sin(th); v[2] = vk * cos(th);
            o.m = 0.5;
        } else {
            const bool up = i < cn + dn + jh;
            const double zj = R * 0.2 + uc * (R * 1.0), rj = -log(ua) * (R * 0.05);
            p[0] = rj * cos(th); p[1] = up ? zj : -zj; p[2] = rj * sin(th);
            v[1] = up ? 3.0 : -3.0;
            o.m = 0.1;
        }
        break;
    }
    case D_TORUS: {                                    // tools/presets.py:980-1017
        D.u(0, ua, ub);
        D.n(1, na, nb);
        D.n(2, nc, nd);
        const double major = R * 0.6, minor = R * 0.25, u = TWO_PI * ua, w = TWO_PI * ub, rn = 1.0 + na * 0.1;
        const double ring = major + minor * cos(u) * rn;
        p[0] = ring * cos(w); p[1] = minor * sin(u) * rn; p[2] = ring * sin(w);
        const double rxy = sqrt(p[0] * p[0] + p[2] * p[2]), omega = sqrt(G * (double)n * 0.001 / major);
        const bool m = rxy > 0.01;
        v[0] = (m ? -omega * p[2] / rxy : 0.0) + nb * (omega * 0.05);
        v[1] = nc * (omega * 0.05);
        v[2] = (m ? omega * p[0] / rxy : 0.0) + nd * (omega * 0.05);
        break;
    }
    case D_HOURGLASS: {                                // tools/presets.py:1019-1111 (binary's centring: finish pass)
        const int64_t bn = n / 200 > 2 ? n / 200 : 2, nn = n - bn, half = nn / 2, b1 = bn / 2;
        D.u(0, ua, ub);
        D.n(1, na, nb);
        D.n(2, nc, nd);
        D.n(3, ne, nf);
        const double bsep = R * 0.05, vb = sqrt(G * 250.0 / bsep);
        if (i < bn) {
            const bool s1 = i < b1;
            p[0] = (s1 ? -bsep / 2 : bsep / 2) + na * (R * 0.01); p[1] = nb * (R * 0.01); p[2] = nc * (R * 0.01);
            v[1] = nd * 0.05; v[2] = (s1 ? vb : -vb) + ne * 0.05;
            o.m = 100.0;
        } else {
            const bool upper = i < bn + half;
            const double zc = upper ? ua * R : -ua * R, rc = fabs(zc) * 0.5 * (1.0 + na * 0.1), th = TWO_PI * ub;
            p[0] = rc * cos(th); p[1] = zc; p[2] = rc * sin(th);
            const double rxy = sqrt(p[0] * p[0] + p[2] * p[2]), r3 = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
            const double vo = sqrt(G * 500.0 / (r3 + R * 0.05));
            const bool m = rxy > 0.01;
            v[0] = (m ? -vo * p[2] / rxy : 0.0) + nd * 0.08;
            v[1] = nb * (vo * (r3 / R) * 0.08) + ne * 0.08;
            v[2] = (m ? vo * p[0] / rxy : 0.0) + nf * 0.08;
            o.m = 0.1;
        }
        break;
    }
    case D_FIBONACCI: {                                // tools/presets.py:1113-1145
        const double golden = (1.0 + sqrt(5.0)) / 2.0;
        const double th = (double)i * (TWO_PI / (golden * golden));
        const double r = i > 0 ? R * sqrt((double)i / (double)n) : R * 0.01;
        D.n(0, na, nb);
        D.n(1, nc, nd);
        p[0] = r * cos(th); p[1] = ((double)i / (double)n - 0.5) * R * 2.0; p[2] = r * sin(th);
        const double vo = r > 0.01 ? sqrt(G * ((double)n * 0.001) / (r + R * 0.05)) : 0.0;
        v[0] = -vo * sin(th) + na * 0.05; v[1] = nb * 0.05; v[2] = vo * cos(th) + nc * 0.05;
        break;
    }
    case D_ROSETTE: {                                  // tools/presets.py:1212-1258
        const int64_t ps = n / 5;
        const int64_t petal = ps > 0 ? (i / ps < 4 ? i / ps : 4) : 4;
        const double ang = (double)petal * (TWO_PI / 5.0);
        D.u(0, ua, ub);
        D.n(1, na, nb);
        D.n(2, nc, nd);
        const double r = -log(ua) * (R * 0.25), th = TWO_PI * ub;
        const double xl = r * cos(th), zl = r * sin(th) * 0.3;
        p[0] = xl * cos(ang) - zl * sin(ang); p[1] = na * (R * 0.02); p[2] = xl * sin(ang) + zl * cos(ang);
        const double rxy = sqrt(p[0] * p[0] + p[2] * p[2]), r3 = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
        const double om = 0.5 * sqrt(R * 0.3 / (r3 + R * 0.05));
        const bool m = rxy > 0.01;
        v[0] = (m ? -om * p[2] / rxy : 0.0) + nb * 0.05;
        v[1] = nc * 0.05;
        v[2] = (m ? om * p[0] / rxy : 0.0) + nd * 0.05;
        break;
    }
    case D_DYSON: {                                    // tools/presets.py:1260-1377 (centre's centring: finish pass)
        const int64_t cn = n / 200 > 1 ? n / 200 : 1;
        double d[3];
        D.u(0, ua, ub);
        D.n(1, na, nb);
        D.n(2, nc, nd);
        D.n(3, ne, nf);
        if (i < cn) {
            p[0] = na * (R * 0.01); p[1] = nb * (R * 0.01); p[2] = nc * (R * 0.01);
            v[0] = nd * 0.05; v[1] = ne * 0.05; v[2] = nf * 0.05;
            o.m = 500.0;
            break;
        }
        iso(ua, ub, d);
        const double r = R * 0.7 + na * (R * 0.03);
        o.rkey = r;
        for (int k = 0; k < 3; ++k) p[k] = r * d[k];
        o.m = 0.1;
        // enclosed mass: the reference maps the per-body array through the inverse sort permutation although
        // it already is in body order (:1326-1328): body i gets the enclosed mass of the shell body whose index
        // is i's rank.  Reproduced as written.
        double rk2 = 1.0;
        if (P.rank) rk2 = (double)P.rank[cn + (int64_t)P.rank[i] - 1];
        const double menc = 500.0 * (double)cn + 0.1 * rk2;
        const double vo = sqrt(G * menc / (r + R * 0.01));
        const double rm = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
        if (!(rm > 0.01)) { v[0] = nc * 0.01; v[1] = nd * 0.01; v[2] = ne * 0.01; break; }
        const double ru[3] = {p[0] / rm, p[1] / rm, p[2] / rm};
        double t[3] = {-ru[2], 0.0, ru[0]};              // radial x Y axis (poles: radial x X axis), :1336-1353
        double tmag = sqrt(t[0] * t[0] + t[2] * t[2]);
        if (tmag < 0.01) { t[0] = 0.0; t[1] = ru[2]; t[2] = -ru[1]; tmag = sqrt(t[1] * t[1] + t[2] * t[2]); }
        double sv[3];
        for (int k = 0; k < 3; ++k) sv[k] = vo * (t[k] / (tmag + 1e-10));
        const double vert[3] = {p[1] * sv[2] - p[2] * sv[1], p[2] * sv[0] - p[0] * sv[2], p[0] * sv[1] - p[1] * sv[0]};
        const double vmag = sqrt(vert[0] * vert[0] + vert[1] * vert[1] + vert[2] * vert[2]);
        for (int k = 0; k < 3; ++k) v[k] = sv[k] + (vmag > 0.01 ? vert[k] / vmag * (nb * vo * 0.01) : 0.0);
        break;
    }
    default: break;
    }
}

// phase 0: only the radius the rotation-curve laws rank by -> sort keys (bit pattern of a non-negative double)
__global__ void __launch_bounds__(256) gen_keys_kernel(GenParams P, uint64_t* __restrict__ keys)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    Body b;
    gen_body(P, i, b);
    keys[i] = (uint64_t)__double_as_longlong(fmax(b.rkey, 0.0));
}

__global__ void __launch_bounds__(256) gen_rank_kernel(const uint32_t* __restrict__ sorted_local, int64_t begin, int64_t count,
                                                       uint32_t* __restrict__ rank)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < count) rank[begin + sorted_local[j]] = (uint32_t)(j + 1);
}

__global__ void __launch_bounds__(256) gen_bodies_kernel(GenParams P, double* __restrict__ pos, double* __restrict__ vel,
                                                         double* __restrict__ mass)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    Body b;
    gen_body(P, i, b);
    for (int k = 0; k < 3; ++k) { pos[3 * i + k] = b.p[k]; vel[3 * i + k] = b.v[k]; }
    mass[i] = b.m;
}

// ---- deterministic mean of rows [begin, end) of an (n, 3) array, then subtraction (centre-of-mass shifts)
constexpr int MEAN_CHUNK = 4096;
__global__ void __launch_bounds__(256) mean_partial_kernel(const double* __restrict__ a, int64_t begin, int64_t end, double* __restrict__ partial)
{
    __shared__ double sh[256][3];
    const int64_t lo = begin + (int64_t)blockIdx.x * MEAN_CHUNK, hi = lo + MEAN_CHUNK < end ? lo + MEAN_CHUNK : end;
    double s[3] = {0.0, 0.0, 0.0};
    for (int64_t i = lo + threadIdx.x; i < hi; i += 256)
        for (int k = 0; k < 3; ++k) s[k] += a[3 * i + k];
    for (int k = 0; k < 3; ++k) sh[threadIdx.x][k] = s[k];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w)
            for (int k = 0; k < 3; ++k) sh[threadIdx.x][k] += sh[threadIdx.x + w][k];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int k = 0; k < 3; ++k) partial[3 * (int64_t)blockIdx.x + k] = sh[0][k];
}
__global__ void __launch_bounds__(256) mean_final_kernel(double* __restrict__ partial, int blocks, double count)
{
    __shared__ double sh[256][3];
    double s[3] = {0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < blocks; b += 256)
        for (int k = 0; k < 3; ++k) s[k] += partial[3 * (int64_t)b + k];
    for (int k = 0; k < 3; ++k) sh[threadIdx.x][k] = s[k];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w)
            for (int k = 0; k < 3; ++k) sh[threadIdx.x][k] += sh[threadIdx.x + w][k];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int k = 0; k < 3; ++k) partial[k] = sh[0][k] / count;   // (every block's partial has been read before the barrier)
}
__global__ void __launch_bounds__(256) mean_sub_kernel(double* __restrict__ a, int64_t begin, int64_t end, const double* __restrict__ mean)
{
    const int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < end)
        for (int k = 0; k < 3; ++k) a[3 * i + k] -= mean[k];
}

void subtract_mean(double* a, int64_t begin, int64_t end, double* partial, cudaStream_t st)
{
    if (end <= begin) return;
    const int blocks = div_up(end - begin, MEAN_CHUNK);
    mean_partial_kernel<<<blocks, 256, 0, st>>>(a, begin, end, partial);
    mean_final_kernel<<<1, 256, 0, st>>>(partial, blocks, (double)(end - begin));
    mean_sub_kernel<<<div_up(end - begin, 256), 256, 0, st>>>(a, begin, end, partial);
}

// Node table of the cosmic web (tools/presets.py:609-650), drawn on the host from stream 1 (counter = node).
std::vector<double> filament_nodes(uint32_t k0, uint32_t k1, double R)
{
    const int gs = 8, total = gs * gs * gs;
    std::vector<double> tab;
    std::vector<double> w;
    bool any = false;
    for (int pass = 0; pass < 2 && !any; ++pass) {
        for (int node = 0; node < total; ++node) {
            const Rng D{k0, k1, 1u, (uint64_t)node};
            double act, wu, a0, a1, a2, b0, b1, b2;
            D.u(0, act, wu);
            const bool active = pass == 0 ? act < 0.35 : node == 0;   // (no active node at all: node 0, like the twin)
            if (!active) continue;
            any = true;
            D.n(1, a0, a1);
            D.n(2, a2, b0);
            D.n(3, b1, b2);
            const int ix = node / (gs * gs), iy = (node / gs) % gs, iz = node % gs;
            const double step = 2.5 * R / (gs - 1);
            const double c[3] = {-1.25 * R + step * ix, -1.25 * R + step * iy, -1.25 * R + step * iz};
            double e[3] = {a0, a1, a2}, p1[3] = {b0, b1, b2}, p2[3];
            double nrm = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) + 1e-10;
            for (double& x : e) x /= nrm;
            const double dot = p1[0] * e[0] + p1[1] * e[1] + p1[2] * e[2];
            for (int k = 0; k < 3; ++k) p1[k] -= dot * e[k];
            nrm = std::sqrt(p1[0] * p1[0] + p1[1] * p1[1] + p1[2] * p1[2]) + 1e-10;
            for (double& x : p1) x /= nrm;
            p2[0] = e[1] * p1[2] - e[2] * p1[1]; p2[1] = e[2] * p1[0] - e[0] * p1[2]; p2[2] = e[0] * p1[1] - e[1] * p1[0];
            nrm = std::sqrt(p2[0] * p2[0] + p2[1] * p2[1] + p2[2] * p2[2]) + 1e-10;
            for (double& x : p2) x /= nrm;
            w.push_back(std::sqrt(wu));                  // numpy.random.power(2)
            const double row[13] = {c[0], c[1], c[2], 0.0, e[0], e[1], e[2], p1[0], p1[1], p1[2], p2[0], p2[1], p2[2]};
            tab.insert(tab.end(), row, row + 13);
        }
    }
    double cum = 0.0;
    for (size_t k = 0; k < w.size(); ++k) { cum += w[k]; tab[13 * k + 3] = cum; }
    for (size_t k = 0; k < w.size(); ++k) tab[13 * k + 3] /= cum;
    return tab;
}

}  // namespace

int generator_id(const char* name)
{
    for (int d = 0; d < D_COUNT; ++d)
        if (name && std::strcmp(name, DIST_NAMES[d]) == 0) return d;
    return D_SPHERE;   // tools/presets.py:1379: an unknown name falls through to the sphere
}

// Fills device arrays pos (n,3), vel (n,3), mass (n) [fp64] on `stream`.  Temporary device memory (sort
// buffers for the rank laws: 28 B/body) is allocated and released inside; the call synchronises the stream.
void generate_device(int dist, int64_t n, double R, double G, uint64_t seed, double* pos, double* vel, double* mass,
                     cudaStream_t st, int sm_count)
{
    B200_REQUIRE(n >= 0 && n < (int64_t)1 << 31, "generate: n out of range");
    if (n == 0) return;
    GenParams P{};
    P.dist = dist; P.n = n; P.R = R; P.G = G;
    P.k0 = (uint32_t)seed; P.k1 = (uint32_t)(seed >> 32);
    int64_t side = (int64_t)std::ceil(std::pow((double)n, 1.0 / 3.0));   // int(np.ceil(n ** (1/3))), tools/presets.py:829
    while (side * side * side < n) ++side;
    P.cube_side = side;
    const int blocks = div_up(n, 256);
    double* d_tab = nullptr;
    if (dist == D_FILAMENT) {
        const std::vector<double> tab = filament_nodes(P.k0, P.k1, R);
        d_tab = dev_alloc<double>(tab.size());
        B200_CHECK(cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        B200_CHECK(cudaStreamSynchronize(st));   // (tab is a local)
        P.nodes = (int)(tab.size() / 13);
        P.node_tab = d_tab;
    }
    // rank groups [begin, end): bodies ranked by radius among themselves
    std::vector<std::pair<int64_t, int64_t>> groups;
    switch (dist) {
    case D_GALAXY: case D_SPIRAL: case D_VORTEX: groups = {{0, n}}; break;
    case D_COLLISION: groups = {{0, n / 2}, {n / 2, n}}; break;
    case D_TRIPLE: groups = {{0, n / 3}, {n / 3, 2 * (n / 3)}, {2 * (n / 3), n}}; break;
    case D_BAR: groups = {{0, n / 3}, {n / 3, n}}; break;
    case D_DYSON: { const int64_t cn = n / 200 > 1 ? n / 200 : 1; groups = {{cn, n}}; break; }
    default: break;
    }
    uint64_t* keys[2] = {nullptr, nullptr};
    uint32_t* vals[2] = {nullptr, nullptr};
    uint32_t* rank = nullptr;
    rsort::Sorter<uint64_t> sorter;
    if (!groups.empty()) {
        for (int k = 0; k < 2; ++k) { keys[k] = dev_alloc<uint64_t>(n); vals[k] = dev_alloc<uint32_t>(n); }
        rank = dev_alloc<uint32_t>(n);
        sorter.init((int)n);
        B200_CHECK(cudaMemsetAsync(rank, 0, n * sizeof(uint32_t), st));
        gen_keys_kernel<<<blocks, 256, 0, st>>>(P, keys[0]);
        for (const auto& g : groups) {
            const int64_t cnt = g.second - g.first;
            if (cnt <= 0) continue;
            uint64_t* gk[2] = {keys[0] + g.first, keys[1] + g.first};
            uint32_t* gv[2] = {vals[0] + g.first, vals[1] + g.first};
            const int slot = sorter.sort(gk, gv, 0, (int)cnt, 0, 63, true, st, sm_count);
            gen_rank_kernel<<<div_up(cnt, 256), 256, 0, st>>>(gv[slot], g.first, cnt, rank);
        }
        P.rank = rank;
    }
    gen_bodies_kernel<<<blocks, 256, 0, st>>>(P, pos, vel, mass);
    // centre-of-mass shifts (all of them over bodies of equal mass: plain means)
    double* partial = dev_alloc<double>(3 * (size_t)div_up(n, MEAN_CHUNK) + 3);
    switch (dist) {
    case D_GALAXY: case D_SPIRAL: case D_CLUSTER: case D_BINARY: case D_ELLIPTICAL: case D_BAR: case D_VORTEX: case D_TRIPLE:
        subtract_mean(vel, 0, n, partial, st);
        break;
    case D_ACCRETION_DISK: { const int64_t cn = n / 100 > 1 ? n / 100 : 1; subtract_mean(pos, 0, cn, partial, st); subtract_mean(vel, 0, cn, partial, st); break; }
    case D_HOURGLASS: { const int64_t bn = std::min<int64_t>(n, n / 200 > 2 ? n / 200 : 2); subtract_mean(pos, 0, bn, partial, st); subtract_mean(vel, 0, bn, partial, st); break; }
    case D_DYSON: { const int64_t cn = n / 200 > 1 ? n / 200 : 1; subtract_mean(pos, 0, cn, partial, st); subtract_mean(vel, 0, cn, partial, st); break; }
    default: break;
    }
    B200_CHECK(cudaGetLastError());
    B200_CHECK(cudaStreamSynchronize(st));
    cudaFree(partial);
    if (d_tab) cudaFree(d_tab);
    if (rank) {
        sorter.destroy();
        cudaFree(rank);
        for (int k = 0; k < 2; ++k) { cudaFree(keys[k]); cudaFree(vals[k]); }
    }
}

}  // namespace b200
