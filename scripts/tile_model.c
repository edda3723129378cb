/* tile_model.c -- CPU model of the warp-cooperative traversal (development aid, not product code).
 *
 * Walks the oracle's octree for TILES of bodies exactly the way the CUDA traversal does (per-lane
 * MAC, a cell's children evaluated by the lanes that opened it, two children per pair slot) and
 * reports what bounds the kernel: evaluated pair slots, lane utilisation, the histogram of mask
 * populations, leaf fraction, and how many slots a conservative tile-level test (bounding sphere of
 * the tile against the cell's MAC radius) could classify as "accepted by every lane" up front.
 *
 *   gcc -O3 -fopenmp -shared -fPIC -o /tmp/libtilemodel.so scripts/tile_model.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int64_t tiles, slots, lanepairs, child_evals, interactions, visits_open;
    int64_t leaf_child_evals, leafpair_slots, leafpair_lanepairs;
    int64_t sure_slots, sure_lanepairs, sure_full_slots;        /* both children surely accepted by the whole tile */
    int64_t sureopen_slots;                                      /* both children surely opened by every lane of the tile */
    int64_t hist[33];                                            /* slots by mask population */
    int64_t bodies;
    /* 64-body tiles (two 32-halves sharing the walk): slots needed by both halves / one half */
    int64_t slots_both, slots_one;
    int64_t combo[16];   /* 64-tiles: pair records by (class of the low half) * 4 + (class of the high half); 0 none 1 sure+full 2 sure+masked 3 unsure */
} TileStats;

static inline int eff_node(const int32_t* children, const uint8_t* is_leaf, int node)
{   /* collapse single-child chains to their deepest cell (same mass / COM; its size decides the MAC) */
    for (;;) {
        if (is_leaf[node]) return node;
        int cnt = 0, last = -1;
        for (int c = 0; c < 8; ++c) { const int ch = children[8 * (int64_t)node + c]; if (ch >= 0) { ++cnt; last = ch; } }
        if (cnt != 1) return node;
        node = last;
    }
}

/* order: body indices in tile order; tstart[t]..tstart[t+1]: bodies of tile t (<= 32 each, or <= 64 when halves = 2) */
void tile_model(const double* pos, const double* half, const double* com, const int32_t* children,
                const int32_t* body_idx, const uint8_t* is_leaf, const int64_t* order, const int64_t* tstart,
                int64_t ntiles, int64_t stride, double theta, double softening, int halves, TileStats* out)
{
    const double eps2 = softening * softening;
    TileStats tot;
    memset(&tot, 0, sizeof(tot));
#pragma omp parallel
    {
        TileStats st;
        memset(&st, 0, sizeof(st));
        int cap = 4096;
        int32_t* snode = (int32_t*)malloc(sizeof(int32_t) * cap);
        uint64_t* smask = (uint64_t*)malloc(sizeof(uint64_t) * cap);
#pragma omp for schedule(dynamic, 16)
        for (int64_t tt = 0; tt < ntiles; tt += stride) {
            const int64_t b0 = tstart[tt], nb = tstart[tt + 1] - b0;
            if (nb <= 0) continue;
            double px[64], py[64], pz[64];
            int64_t id[64];
            double cx = 0, cy = 0, cz = 0;
            double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
            for (int l = 0; l < nb; ++l) {
                id[l] = order[b0 + l];
                px[l] = pos[3 * id[l]]; py[l] = pos[3 * id[l] + 1]; pz[l] = pos[3 * id[l] + 2];
                if (px[l] < lo[0]) lo[0] = px[l]; if (px[l] > hi[0]) hi[0] = px[l];
                if (py[l] < lo[1]) lo[1] = py[l]; if (py[l] > hi[1]) hi[1] = py[l];
                if (pz[l] < lo[2]) lo[2] = pz[l]; if (pz[l] > hi[2]) hi[2] = pz[l];
            }
            cx = 0.5 * (lo[0] + hi[0]); cy = 0.5 * (lo[1] + hi[1]); cz = 0.5 * (lo[2] + hi[2]);
            /* per-half bounding spheres (64-body tiles) */
            double hc[2][3] = {{0, 0, 0}, {0, 0, 0}}, hR[2] = {0, 0}, hlo[2][3] = {{0}}, hhi[2][3] = {{0}};
            for (int h = 0; h < 2; ++h) {
                double l2[3] = {1e300, 1e300, 1e300}, h2[3] = {-1e300, -1e300, -1e300};
                const int a = 32 * h, b = nb < 32 * (h + 1) ? (int)nb : 32 * (h + 1);
                if (b <= a) continue;
                for (int l = a; l < b; ++l) {
                    if (px[l] < l2[0]) l2[0] = px[l]; if (px[l] > h2[0]) h2[0] = px[l];
                    if (py[l] < l2[1]) l2[1] = py[l]; if (py[l] > h2[1]) h2[1] = py[l];
                    if (pz[l] < l2[2]) l2[2] = pz[l]; if (pz[l] > h2[2]) h2[2] = pz[l];
                }
                for (int d = 0; d < 3; ++d) { hc[h][d] = 0.5 * (l2[d] + h2[d]); hlo[h][d] = l2[d]; hhi[h][d] = h2[d]; }
                for (int l = a; l < b; ++l) {
                    const double d = sqrt((px[l] - hc[h][0]) * (px[l] - hc[h][0]) + (py[l] - hc[h][1]) * (py[l] - hc[h][1]) + (pz[l] - hc[h][2]) * (pz[l] - hc[h][2]));
                    if (d > hR[h]) hR[h] = d;
                }
            }
            double R = 0;
            for (int l = 0; l < nb; ++l) {
                const double d = sqrt((px[l] - cx) * (px[l] - cx) + (py[l] - cy) * (py[l] - cy) + (pz[l] - cz) * (pz[l] - cz));
                if (d > R) R = d;
            }
            ++st.tiles;
            st.bodies += nb;
            const uint64_t full = nb >= 64 ? ~0ull : ((1ull << nb) - 1ull);
            int sp = 0;
            snode[sp] = eff_node(children, is_leaf, 0); smask[sp] = full; ++sp;
            /* the root itself is evaluated as child 0 of pair 0 = {root, dummy}: one slot, all lanes */
            {
                const int root = snode[0];
                st.slots += 1; st.lanepairs += nb; st.hist[nb > 32 ? 32 : nb] += 1;
                if (halves == 2) st.slots_both += 1;
                uint64_t open = 0;
                const double size = half[root] * 2.0;
                for (int l = 0; l < nb; ++l) {
                    const double dx = com[3 * root] - px[l], dy = com[3 * root + 1] - py[l], dz = com[3 * root + 2] - pz[l];
                    const double d2 = dx * dx + dy * dy + dz * dz + eps2;
                    ++st.child_evals;
                    if (is_leaf[root] || size / sqrt(d2) < theta) { if (d2 > eps2) ++st.interactions; }
                    else open |= 1ull << l;
                }
                sp = 0;
                if (open) { snode[sp] = root; smask[sp] = open; ++sp; st.visits_open += __builtin_popcountll(open); }
            }
            while (sp > 0) {
                --sp;
                const int node = snode[sp];
                const uint64_t m = smask[sp];
                const int pm = __builtin_popcountll(m);
                int kids[8], nk = 0;
                if (is_leaf[node]) continue;
                for (int c = 0; c < 8; ++c) {
                    const int ch = children[8 * (int64_t)node + c];
                    if (ch >= 0) kids[nk++] = eff_node(children, is_leaf, ch);
                }
                /* evaluate all children for the lanes in m */
                uint64_t openm[8];
                int sure_acc[8], sure_open[8], leaf[8], sure_h[2][8];
                for (int k = 0; k < nk; ++k) {
                    const int c = kids[k];
                    leaf[k] = is_leaf[c];
                    const double size = half[c] * 2.0;
                    uint64_t om = 0;
                    for (int l = 0; l < nb; ++l) {
                        if (!((m >> l) & 1)) continue;
                        if (leaf[k] && body_idx[c] == id[l]) { ++st.child_evals; ++st.leaf_child_evals; continue; }   /* self: evaluated, contributes 0 */
                        const double dx = com[3 * c] - px[l], dy = com[3 * c + 1] - py[l], dz = com[3 * c + 2] - pz[l];
                        const double d2 = dx * dx + dy * dy + dz * dz + eps2;
                        ++st.child_evals;
                        if (leaf[k]) ++st.leaf_child_evals;
                        if (leaf[k] || size / sqrt(d2) < theta) { if (d2 > eps2) ++st.interactions; }
                        else om |= 1ull << l;
                    }
                    openm[k] = om;
                    st.visits_open += __builtin_popcountll(om);
                    /* conservative tile-level tests against the bounding sphere (centre c*, radius R) */
                    const double dc = sqrt((com[3 * c] - cx) * (com[3 * c] - cx) + (com[3 * c + 1] - cy) * (com[3 * c + 1] - cy) +
                                           (com[3 * c + 2] - cz) * (com[3 * c + 2] - cz));
                    const double T = size * size / (theta * theta);
                    const double S = T > eps2 ? sqrt(T - eps2) : -1.0;   /* accept iff r > S (r = unsoftened distance) */
                    sure_acc[k] = leaf[k] || S < 0 || (dc - R > S * 1.00001);
                    sure_open[k] = !leaf[k] && S >= 0 && (dc + R < S * 0.99999);
                    for (int h = 0; h < 2; ++h) {
                        const double dh = sqrt((com[3 * c] - hc[h][0]) * (com[3 * c] - hc[h][0]) + (com[3 * c + 1] - hc[h][1]) * (com[3 * c + 1] - hc[h][1]) +
                                               (com[3 * c + 2] - hc[h][2]) * (com[3 * c + 2] - hc[h][2]));
                        sure_h[h][k] = leaf[k] || S < 0 || (dh - hR[h] > S * 1.00001);
                        if (getenv("TILE_AABB")) {   /* axis-aligned box of the half instead of its sphere */
                            double m2 = 0;
                            for (int d = 0; d < 3; ++d) {
                                const double cc = com[3 * c + d];
                                const double q = cc < hlo[h][d] ? hlo[h][d] - cc : (cc > hhi[h][d] ? cc - hhi[h][d] : 0.0);
                                m2 += q * q;
                            }
                            sure_h[h][k] = leaf[k] || S < 0 || (m2 > S * S * 1.00002);
                        }
                    }
                    if (om) {
                        if (sp + 1 >= cap) { cap *= 2; snode = (int32_t*)realloc(snode, sizeof(int32_t) * cap); smask = (uint64_t*)realloc(smask, sizeof(uint64_t) * cap); }
                        snode[sp] = c; smask[sp] = om; ++sp;
                    }
                }
                const int npairs = (nk + 1) / 2;
                for (int p = 0; p < npairs; ++p) {
                    const int k0 = 2 * p, k1 = 2 * p + 1;
                    const int has1 = k1 < nk;
                    if (halves == 2) {
                        const int plo = __builtin_popcountll(m & 0xffffffffull), phi = __builtin_popcountll(m >> 32);
                        const uint64_t flo = full & 0xffffffffull, fhi = full >> 32;
                        int cls[2];
                        for (int h = 0; h < 2; ++h) {
                            const uint64_t mh = h ? (m >> 32) : (m & 0xffffffffull), fh = h ? fhi : flo;
                            const int sure = sure_h[h][k0] && (!has1 || sure_h[h][k1]);
                            cls[h] = mh == 0 ? 0 : (!sure ? 3 : (mh == fh ? 1 : 2));
                        }
                        st.combo[cls[0] * 4 + cls[1]] += 1;
                        if (plo && phi) { st.slots += 2; st.slots_both += 1; st.hist[plo] += 1; st.hist[phi] += 1; }
                        else { st.slots += 1; st.slots_one += 1; st.hist[plo + phi] += 1; }
                    } else {
                        st.slots += 1;
                        st.hist[pm] += 1;
                    }
                    st.lanepairs += pm;
                    const int lp = leaf[k0] && (!has1 || leaf[k1]);
                    if (lp) { st.leafpair_slots += 1; st.leafpair_lanepairs += pm; }
                    const int sa = sure_acc[k0] && (!has1 || sure_acc[k1]);
                    if (sa) { st.sure_slots += 1; st.sure_lanepairs += pm; if (m == full) st.sure_full_slots += 1; }
                    if (sure_open[k0] && has1 && sure_open[k1]) st.sureopen_slots += 1;
                }
            }
        }
        free(snode); free(smask);
#pragma omp critical
        {
            int64_t* a = (int64_t*)&tot; const int64_t* b = (const int64_t*)&st;
            for (size_t i = 0; i < sizeof(TileStats) / sizeof(int64_t); ++i) a[i] += b[i];
        }
    }
    *out = tot;
}
