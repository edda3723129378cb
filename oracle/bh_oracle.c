/*
 * oracle/bh_oracle.c -- CPU restatement of the reference's Barnes-Hut step (fp64).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package
 * 3d-spatial-sim-for-boid-and-nbody_b200/ or its CUDA library) may import, link or call
 * this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the CPU baseline.
 *
 * Parity pinning: the reference ships no golden vectors, so this
 * restatement is pinned against outputs of the reference itself, imported unmodified
 * in the authoring container by tests/golden/make_golden.py (fixtures committed under
 * tests/golden/, checked by tests/test_oracle_golden.py).
 *
 * Each function cites the reference lines it follows (paths relative to the
 * reference repository root).  All arithmetic is IEEE fp64, compiled without
 * -ffast-math; the reference's numba fastmath=True only permits reassociation, so the
 * two agree to fp64 rounding.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* nbody/simulation.py:38-49  get_octant: bit0 = x>=cx, bit1 = y>=cy, bit2 = z>=cz */
static inline int orc_octant(double px, double py, double pz, double cx, double cy, double cz)
{
    int o = 0;
    if (px >= cx) o |= 1;
    if (py >= cy) o |= 2;
    if (pz >= cz) o |= 4;
    return o;
}

/* nbody/simulation.py:52-60  get_octant_center: child centre = centre +- half/2 */
static inline void orc_octant_center(int o, double cx, double cy, double cz, double hs,
                                     double* ox, double* oy, double* oz)
{
    double q = hs * 0.5;
    *ox = (o & 1) ? cx + q : cx - q;
    *oy = (o & 2) ? cy + q : cy - q;
    *oz = (o & 4) ? cz + q : cz - q;
}

/* nbody/simulation.py:308-317  compute_bounds: 1.1 * max|coord| + 10.
 * The reference compiles this with numba fastmath=True, which contracts the final
 * multiply-add into one FMA on every FMA-capable x86 host (verified against the
 * fixtures: the unfused form is 1 ulp off for cluster_2k); fma() reproduces that. */
ORC_API double orc_compute_bounds(const double* pos, int64_t n)
{
    double m = 0.0;
    for (int64_t i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) {
            double e = fabs(pos[3 * i + d]);
            if (e > m) m = e;
        }
    return fma(m, 1.1, 10.0);
}

/*
 * nbody/simulation.py:63-198  build_octree: sequential top-down insertion into a cube
 * [-bounds,+bounds]^3 centred on the origin; internal nodes keep a running-mean centre
 * of mass.  `max_nodes` plays the role of MAX_TREE_NODES (:35,:141,:176): pass the
 * reference's value to reproduce its truncation, or the allocation size for an
 * uncapped tree.  The caller pre-fills children=-1, body_idx=-1, is_leaf=1 as the
 * reference's callers do (tools/record.py:838-840).
 * Deviation (documented): the reference has no depth limit, so two bodies at exactly
 * the same position subdivide until the node cap; here insertion of that body stops
 * after `max_depth` levels (<=0 means 4096) instead of looping.
 * Returns the node count.
 */
ORC_API int64_t orc_build_octree(const double* pos, const double* mass, int64_t n, double bounds,
                                 int64_t max_nodes, int max_depth,
                                 double* centers, double* half, double* nmass, double* com,
                                 int32_t* children, int32_t* body_idx, uint8_t* is_leaf)
{
    if (max_depth <= 0) max_depth = 4096;
    centers[0] = centers[1] = centers[2] = 0.0;
    half[0] = bounds;
    nmass[0] = 0.0;
    com[0] = com[1] = com[2] = 0.0;
    body_idx[0] = -1;
    is_leaf[0] = 1;
    for (int c = 0; c < 8; ++c) children[c] = -1;
    int64_t num_nodes = 1;

    for (int64_t i = 0; i < n; ++i) {
        const double px = pos[3 * i], py = pos[3 * i + 1], pz = pos[3 * i + 2];
        const double m = mass[i];
        int64_t cur = 0;
        int depth = 0;
        for (;;) {
            const double cx = centers[3 * cur], cy = centers[3 * cur + 1], cz = centers[3 * cur + 2];
            const double hs = half[cur];
            if (is_leaf[cur]) {
                if (body_idx[cur] == -1) {          /* :114-121 empty leaf */
                    body_idx[cur] = (int32_t)i;
                    nmass[cur] = m;
                    com[3 * cur] = px; com[3 * cur + 1] = py; com[3 * cur + 2] = pz;
                    break;
                }
                /* :122-158 occupied leaf: push the resident body one level down */
                const int32_t old = body_idx[cur];
                const double opx = pos[3 * old], opy = pos[3 * old + 1], opz = pos[3 * old + 2];
                const double om = mass[old];
                is_leaf[cur] = 0;
                body_idx[cur] = -1;
                const int o = orc_octant(opx, opy, opz, cx, cy, cz);
                if (children[8 * cur + o] == -1) {
                    const int64_t ch = num_nodes;
                    num_nodes += 1;
                    if (num_nodes >= max_nodes) break;      /* :141-142 */
                    children[8 * cur + o] = (int32_t)ch;
                    orc_octant_center(o, cx, cy, cz, hs, &centers[3 * ch], &centers[3 * ch + 1], &centers[3 * ch + 2]);
                    half[ch] = hs * 0.5;
                    nmass[ch] = om;
                    com[3 * ch] = opx; com[3 * ch + 1] = opy; com[3 * ch + 2] = opz;
                    body_idx[ch] = old;
                    is_leaf[ch] = 1;
                    for (int c = 0; c < 8; ++c) children[8 * ch + c] = -1;
                }
                /* fall through to the internal-node branch on the next iteration (:159) */
            } else {
                /* :160-167 internal: running-mean COM */
                const double tm = nmass[cur] + m;
                if (tm > 0) {
                    com[3 * cur]     = (com[3 * cur]     * nmass[cur] + px * m) / tm;
                    com[3 * cur + 1] = (com[3 * cur + 1] * nmass[cur] + py * m) / tm;
                    com[3 * cur + 2] = (com[3 * cur + 2] * nmass[cur] + pz * m) / tm;
                }
                nmass[cur] = tm;
                const int o = orc_octant(px, py, pz, cx, cy, cz);
                if (children[8 * cur + o] == -1) {          /* :172-193 new leaf child */
                    const int64_t ch = num_nodes;
                    num_nodes += 1;
                    if (num_nodes >= max_nodes) break;      /* :176-177 */
                    children[8 * cur + o] = (int32_t)ch;
                    orc_octant_center(o, cx, cy, cz, hs, &centers[3 * ch], &centers[3 * ch + 1], &centers[3 * ch + 2]);
                    half[ch] = hs * 0.5;
                    nmass[ch] = m;
                    com[3 * ch] = px; com[3 * ch + 1] = py; com[3 * ch + 2] = pz;
                    body_idx[ch] = (int32_t)i;
                    is_leaf[ch] = 1;
                    for (int c = 0; c < 8; ++c) children[8 * ch + c] = -1;
                    break;
                }
                cur = children[8 * cur + o];                /* :195-196 */
                if (++depth > max_depth) break;             /* deviation: see header */
            }
        }
    }
    return num_nodes;
}

/*
 * nbody/simulation.py:201-278  compute_forces_barnes_hut: per-body DFS, softened
 * distance in the MAC (node_size / sqrt(r^2+eps^2) < theta), monopole accumulate,
 * leaf holding the body itself is skipped, r == 0 contributes nothing.
 * `stack_cap`: 64 reproduces the reference's fixed stack (children are silently not
 * pushed when full, :272); 0 = never drop (stack grows as needed).
 * Targets: bodies [t0, t1) if target_idx == NULL, else target_idx[0..nt).
 * Optional outputs (may be NULL): per-call totals of accepted interactions, visited
 * nodes, peak stack occupancy and dropped pushes.
 */
ORC_API void orc_compute_forces(const double* pos, double* acc,
                                const double* half, const double* nmass, const double* com,
                                const int32_t* children, const int32_t* body_idx, const uint8_t* is_leaf,
                                int64_t num_nodes, const int64_t* target_idx, int64_t nt,
                                double theta, double G, double softening, int stack_cap,
                                int64_t* out_interactions, int64_t* out_visits,
                                int32_t* out_peak_stack, int64_t* out_drops)
{
    const double eps2 = softening * softening;
    int64_t tot_inter = 0, tot_visit = 0, tot_drop = 0;
    int32_t peak = 0;
#pragma omp parallel reduction(+ : tot_inter, tot_visit, tot_drop) reduction(max : peak)
    {
        int64_t cap = stack_cap > 0 ? stack_cap : 1024;
        int32_t* stack = (int32_t*)malloc(sizeof(int32_t) * (size_t)cap);
#pragma omp for schedule(dynamic, 256)
        for (int64_t t = 0; t < nt; ++t) {
            const int64_t i = target_idx ? target_idx[t] : t;
            const double px = pos[3 * i], py = pos[3 * i + 1], pz = pos[3 * i + 2];
            double ax = 0.0, ay = 0.0, az = 0.0;
            int64_t sp = 0;
            stack[sp++] = 0;
            while (sp > 0) {
                const int32_t node = stack[--sp];
                if (node < 0 || node >= num_nodes) continue;            /* :241 */
                if (is_leaf[node] && body_idx[node] == i) continue;     /* :245 */
                ++tot_visit;
                const double dx = com[3 * node] - px, dy = com[3 * node + 1] - py, dz = com[3 * node + 2] - pz;
                const double d2 = dx * dx + dy * dy + dz * dz + eps2;
                const double d = sqrt(d2);
                const double size = half[node] * 2.0;
                if (is_leaf[node] || (size / d < theta)) {              /* :258 */
                    if (nmass[node] > 0 && d2 > eps2) {                 /* :260 */
                        const double f = G * nmass[node] * (1.0 / (d * d2));
                        ax += dx * f; ay += dy * f; az += dz * f;
                        ++tot_inter;
                    }
                } else {
                    for (int c = 0; c < 8; ++c) {                       /* :270-274 */
                        const int32_t ch = children[8 * (int64_t)node + c];
                        if (ch < 0) continue;
                        if (stack_cap > 0) {
                            if (sp < stack_cap) stack[sp++] = ch; else ++tot_drop;
                        } else {
                            if (sp == cap) { cap *= 2; stack = (int32_t*)realloc(stack, sizeof(int32_t) * (size_t)cap); }
                            stack[sp++] = ch;
                        }
                    }
                    if (sp > peak) peak = (int32_t)sp;
                }
            }
            acc[3 * t] = ax; acc[3 * t + 1] = ay; acc[3 * t + 2] = az;
        }
        free(stack);
    }
    if (out_interactions) *out_interactions = tot_inter;
    if (out_visits) *out_visits = tot_visit;
    if (out_peak_stack) *out_peak_stack = peak;
    if (out_drops) *out_drops = tot_drop;
}

/* nbody/simulation.py:281-305  update_positions_velocities: v += a dt; v *= damping; x += v dt */
ORC_API void orc_update(double* pos, double* vel, const double* acc, double damping, double dt, int64_t n)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < 3 * n; ++i) {
        double v = vel[i];
        v += acc[i] * dt;
        v *= damping;
        vel[i] = v;
        pos[i] += v * dt;
    }
}

/* nbody/simulation.py:320-400  compute_colors_by_velocity (CUDA twin gpu_backend.py:259-325) */
ORC_API void orc_colors(const double* vel, float* colors, int64_t n, double max_speed)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const double vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];
        const double speed = sqrt(vx * vx + vy * vy + vz * vz);
        const double t = fmin(1.0, speed / max_speed);
        double r, g, b;
        if (t < 0.55) {
            if (t < 0.15) { const double s = t / 0.15; r = 0.4 - 0.2 * s; g = 0.2 + 0.2 * s; b = 0.8 + 0.1 * s; }
            else if (t < 0.30) { const double s = (t - 0.15) / 0.15; r = 0.2 + 0.1 * s; g = 0.4 + 0.1 * s; b = 0.9 + 0.05 * s; }
            else {
                const double s = (t - 0.30) / 0.25;
                if (s < 0.6) { const double s2 = s / 0.6; r = 0.3 - 0.1 * s2; g = 0.5 + 0.3 * s2; b = 0.95 + 0.05 * s2; }
                else { const double s2 = (s - 0.6) / 0.4; r = 0.2 + 0.8 * s2; g = 0.8 + 0.2 * s2; b = 1.0; }
            }
        } else if (t < 0.90) { r = 1.0; g = 1.0; b = 1.0; }
        else if (t < 0.95) { const double s = (t - 0.90) / 0.05; r = 1.0; g = 1.0 - 0.05 * s; b = 1.0 - 1.0 * s; }
        else if (t < 0.99) { const double s = (t - 0.95) / 0.04; r = 1.0; g = 0.95 - 0.45 * s; b = 0.0; }
        else { const double s = (t - 0.99) / 0.01; r = 1.0; g = 0.5 - 0.5 * s; b = 0.0; }
        colors[3 * i] = (float)r; colors[3 * i + 1] = (float)g; colors[3 * i + 2] = (float)b;
    }
}

/*
 * fp64 direct sum with the same softened monopole kernel as nbody/simulation.py:249-267
 * (and the CUDA brute-force twin nbody/gpu_backend.py:145-174): a_i = G sum_j m_j r_ij /
 * (r^2 + eps^2)^{3/2}, j != i.  Targets given by index list (or all when NULL).
 */
ORC_API void orc_direct_sum(const double* pos, const double* mass, int64_t n,
                            const int64_t* target_idx, int64_t nt,
                            double G, double softening, double* acc)
{
    const double eps2 = softening * softening;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t t = 0; t < nt; ++t) {
        const int64_t i = target_idx ? target_idx[t] : t;
        const double px = pos[3 * i], py = pos[3 * i + 1], pz = pos[3 * i + 2];
        double ax = 0, ay = 0, az = 0;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i) continue;
            const double dx = pos[3 * j] - px, dy = pos[3 * j + 1] - py, dz = pos[3 * j + 2] - pz;
            const double d2 = dx * dx + dy * dy + dz * dz + eps2;
            const double f = mass[j] / (d2 * sqrt(d2));
            ax += dx * f; ay += dy * f; az += dz * f;
        }
        acc[3 * t] = G * ax; acc[3 * t + 1] = G * ay; acc[3 * t + 2] = G * az;
    }
}

/*
 * Reference-equivalent 63-bit Morton key (the reference has no keys; SURVEY.md section 7
 * defines them).  The key is the sequence of octants the reference's insertion descent
 * would take for this body through 21 levels of the cube [-bounds,bounds]^3, using the
 * reference's own arithmetic: octant test p >= c (nbody/simulation.py:38-49) and child
 * centre c +- half/2 computed level by level in fp64 (:52-60).  Level-1 octant sits in
 * bits 62..60; inside each 3-bit group z is the MSB and x the LSB.
 */
ORC_API void orc_morton_keys(const double* pos, int64_t n, double bounds, uint64_t* keys)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const double px = pos[3 * i], py = pos[3 * i + 1], pz = pos[3 * i + 2];
        double cx = 0.0, cy = 0.0, cz = 0.0, hs = bounds;
        uint64_t k = 0;
        for (int l = 0; l < 21; ++l) {
            const int o = orc_octant(px, py, pz, cx, cy, cz);
            k = (k << 3) | (uint64_t)o;
            orc_octant_center(o, cx, cy, cz, hs, &cx, &cy, &cz);
            hs *= 0.5;
        }
        keys[i] = k;
    }
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int t)
{
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}
