/*
 * oracle/boids_oracle.c -- CPU restatement of the reference's boids neighbour-rule
 * update (fp64).  TEST INFRASTRUCTURE ONLY (see bh_oracle.c header for the rule).
 *
 * Pinned against outputs of the reference's own Flock.update, imported unmodified by
 * tests/golden/make_golden.py (fixtures under tests/golden/).
 * Citations are boids/flock.py line numbers in the reference repository.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* boids/flock.py:16-27  get_cell_index: int() truncates toward zero, then clamp */
static inline int32_t orc_cell_coord(double p, double cell, int32_t dim, double offset)
{
    int32_t c = (int32_t)((p + offset) / cell);
    if (c > dim - 1) c = dim - 1;
    if (c < 0) c = 0;
    return c;
}

/* boids/flock.py:30-44  assign_cells */
ORC_API void orc_boids_assign_cells(const double* pos, int32_t* cell_idx, double cell, int32_t dim,
                                    double offset, int64_t n)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const int32_t cx = orc_cell_coord(pos[3 * i], cell, dim, offset);
        const int32_t cy = orc_cell_coord(pos[3 * i + 1], cell, dim, offset);
        const int32_t cz = orc_cell_coord(pos[3 * i + 2], cell, dim, offset);
        cell_idx[i] = cx + cy * dim + cz * dim * dim;
    }
}

/*
 * boids/flock.py:618-619 sorts with np.argsort (unstable; intra-cell order unspecified).
 * The restatement uses a stable counting sort: a valid instance of that order.
 * boids/flock.py:47-65  build_cell_lists: first sorted position and count per cell.
 */
ORC_API void orc_boids_sort_and_lists(const int32_t* cell_idx, int32_t* sorted_idx,
                                      int32_t* cell_starts, int32_t* cell_counts,
                                      int64_t n, int64_t num_cells)
{
    for (int64_t c = 0; c < num_cells; ++c) { cell_starts[c] = -1; cell_counts[c] = 0; }
    for (int64_t i = 0; i < n; ++i) cell_counts[cell_idx[i]] += 1;
    int32_t run = 0;
    int32_t* cursor = (int32_t*)malloc(sizeof(int32_t) * (size_t)num_cells);
    for (int64_t c = 0; c < num_cells; ++c) {
        cursor[c] = run;
        if (cell_counts[c] > 0) cell_starts[c] = run;
        run += cell_counts[c];
    }
    for (int64_t i = 0; i < n; ++i) sorted_idx[cursor[cell_idx[i]]++] = (int32_t)i;
    free(cursor);
}

static inline void orc_steer(double* x, double* y, double* z, double vx, double vy, double vz,
                             double max_speed, double max_force, double weight, double* out, int* wrote)
{
    /* boids/flock.py:179-193 (and the identical alignment :200-214 / cohesion :220-234 tails) */
    double mag = sqrt(*x * *x + *y * *y + *z * *z);
    *wrote = 0;
    if (mag > 0) {
        double sx = (*x / mag) * max_speed - vx;
        double sy = (*y / mag) * max_speed - vy;
        double sz = (*z / mag) * max_speed - vz;
        mag = sqrt(sx * sx + sy * sy + sz * sz);
        if (mag > max_force) {
            sx = (sx / mag) * max_force;
            sy = (sy / mag) * max_force;
            sz = (sz / mag) * max_force;
        }
        out[0] = sx * weight; out[1] = sy * weight; out[2] = sz * weight;
        *wrote = 1;
    }
}

/*
 * boids/flock.py:68-238  compute_flocking_spatial.  The force arrays are expected
 * zero-filled and avg_colors pre-filled with colors, as Flock.update does (:633-636).
 * Optional neighbor_counts (may be NULL) receives each boid's accepted-neighbour count.
 */
ORC_API void orc_boids_flocking(const double* pos, const double* vel, const double* col,
                                const int32_t* sorted_idx, const int32_t* cell_starts, const int32_t* cell_counts,
                                double* sepf, double* alif, double* cohf, double* avgc,
                                double cell, int32_t dim, double offset,
                                double perception, double separation,
                                double w_sep, double w_ali, double w_coh,
                                double max_speed, double max_force, int64_t n, int32_t* neighbor_counts)
{
    const double per2 = perception * perception;
    const double sep2 = separation * separation;
    const int32_t range = (int32_t)ceil(perception / cell);
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i) {
        const double px = pos[3 * i], py = pos[3 * i + 1], pz = pos[3 * i + 2];
        const double vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];
        const int32_t cx = orc_cell_coord(px, cell, dim, offset);
        const int32_t cy = orc_cell_coord(py, cell, dim, offset);
        const int32_t cz = orc_cell_coord(pz, cell, dim, offset);
        double sx = 0, sy = 0, sz = 0, ax = 0, ay = 0, az = 0, hx = 0, hy = 0, hz = 0, cr = 0, cg = 0, cb = 0;
        int32_t nsep = 0, nnb = 0;
        for (int32_t dcx = -range; dcx <= range; ++dcx) {
            const int32_t ncx = cx + dcx;
            if (ncx < 0 || ncx >= dim) continue;
            for (int32_t dcy = -range; dcy <= range; ++dcy) {
                const int32_t ncy = cy + dcy;
                if (ncy < 0 || ncy >= dim) continue;
                for (int32_t dcz = -range; dcz <= range; ++dcz) {
                    const int32_t ncz = cz + dcz;
                    if (ncz < 0 || ncz >= dim) continue;
                    const int64_t c = ncx + (int64_t)ncy * dim + (int64_t)ncz * dim * dim;
                    const int32_t start = cell_starts[c];
                    if (start == -1) continue;
                    const int32_t cnt = cell_counts[c];
                    for (int32_t k = 0; k < cnt; ++k) {
                        const int32_t j = sorted_idx[start + k];
                        if (j == i) continue;
                        const double dx = px - pos[3 * j], dy = py - pos[3 * j + 1], dz = pz - pos[3 * j + 2];
                        const double d2 = dx * dx + dy * dy + dz * dz;
                        if (d2 < per2 && d2 > 0.0001) {                 /* :150 */
                            const double d = sqrt(d2);
                            if (d2 < sep2) {                            /* :153-158 */
                                const double inv = 1.0 / d;
                                sx += dx * inv / d; sy += dy * inv / d; sz += dz * inv / d;
                                ++nsep;
                            }
                            ax += vel[3 * j]; ay += vel[3 * j + 1]; az += vel[3 * j + 2];
                            hx += pos[3 * j]; hy += pos[3 * j + 1]; hz += pos[3 * j + 2];
                            cr += col[3 * j]; cg += col[3 * j + 1]; cb += col[3 * j + 2];
                            ++nnb;
                        }
                    }
                }
            }
        }
        int wrote;
        if (nsep > 0) {                                                 /* :174-193 */
            sx /= nsep; sy /= nsep; sz /= nsep;
            orc_steer(&sx, &sy, &sz, vx, vy, vz, max_speed, max_force, w_sep, &sepf[3 * i], &wrote);
        }
        if (nnb > 0) {                                                  /* :195-238 */
            ax /= nnb; ay /= nnb; az /= nnb;
            orc_steer(&ax, &ay, &az, vx, vy, vz, max_speed, max_force, w_ali, &alif[3 * i], &wrote);
            hx = hx / nnb - px; hy = hy / nnb - py; hz = hz / nnb - pz;
            orc_steer(&hx, &hy, &hz, vx, vy, vz, max_speed, max_force, w_coh, &cohf[3 * i], &wrote);
            avgc[3 * i]     = (cr + col[3 * i])     / (nnb + 1);
            avgc[3 * i + 1] = (cg + col[3 * i + 1]) / (nnb + 1);
            avgc[3 * i + 2] = (cb + col[3 * i + 2]) / (nnb + 1);
        }
        if (neighbor_counts) neighbor_counts[i] = nnb;
    }
}

/* boids/flock.py:241-308  update_physics_numba */
ORC_API void orc_boids_physics(double* pos, double* vel, double* col,
                               const double* sepf, const double* alif, const double* cohf, const double* avgc,
                               double bounds, double margin, double wall_force, double max_speed,
                               double blend, double dt, int64_t n)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double a[3];
        for (int d = 0; d < 3; ++d) a[d] = sepf[3 * i + d] + alif[3 * i + d] + cohf[3 * i + d];
        for (int d = 0; d < 3; ++d) {
            const double p = pos[3 * i + d];
            const double dp = p - (bounds - margin);
            if (dp > 0) a[d] -= fmin(dp / margin * 2.0, 1.0) * wall_force;
            const double dn = (-bounds + margin) - p;
            if (dn > 0) a[d] += fmin(dn / margin * 2.0, 1.0) * wall_force;
        }
        for (int d = 0; d < 3; ++d) vel[3 * i + d] += a[d] * dt;
        const double speed = sqrt(vel[3 * i] * vel[3 * i] + vel[3 * i + 1] * vel[3 * i + 1] + vel[3 * i + 2] * vel[3 * i + 2]);
        if (speed > max_speed) {
            const double s = max_speed / speed;
            for (int d = 0; d < 3; ++d) vel[3 * i + d] *= s;
        }
        for (int d = 0; d < 3; ++d) pos[3 * i + d] += vel[3 * i + d] * dt;
        for (int d = 0; d < 3; ++d) col[3 * i + d] += (avgc[3 * i + d] - col[3 * i + d]) * blend;
    }
}
