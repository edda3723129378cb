// nbody.cu -- Barnes-Hut step on sm_100a: bounds, Morton keys, radix sort, LBVH + octree
// records, warp-cooperative theta-MAC traversal, fused integrate.  See nbody.cuh for the layout.
#include "nbody.cuh"

namespace b200 {

// ============================================================================ bounds
// max |coord| over all bodies (nbody/simulation.py:308-317).  Non-negative doubles order
// like their bit patterns, so the cross-block reduce is an integer atomicMax.
__global__ void __launch_bounds__(256) absmax_kernel(const double* __restrict__ pos, int64_t count,
                                                     unsigned long long* __restrict__ out)
{
    double m = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) m = fmax(m, fabs(pos[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ double wm[8];
    if (lane_id() == 0) wm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? wm[threadIdx.x] : 0.0;
        for (int o = 4; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
    }
}

// MAC threshold of a cell at an octree level (defined with the records below); keygen tabulates it.
__device__ float cell_threshold(int level, double bounds, double theta, float eps2);

// ============================================================================ Morton keys
// The key is the octant path the reference's insertion descent takes for this body through
// 21 levels of the cube [-bounds, bounds]^3, with the reference's own fp64 arithmetic:
// octant bit = (p >= c) (nbody/simulation.py:38-49), child centre = c +- half/2 level by
// level (:52-60).  bounds = fma(max|coord|, 1.1, 10) (:317; numba fastmath contracts it).
// Fast path: the descent is floor((p + bounds) / (2 bounds) * 2^21) per axis evaluated with level-by-level
// ROUNDED centres.  Rounding moves a cell face by at most 21 ulp(bounds)/2 = 2.4e-9 finest cells and the
// closed form below is off by < 1e-9 finest cells, so whenever the body is farther than 2^-20 finest
// cells from every integer grid coordinate (every face of every level sits on one) both give the same
// cell and the key is the bit interleave of the three integers.  Bodies closer than that (a few per
// million) take the reference's exact descent.
__device__ __forceinline__ uint64_t spread3(uint32_t v)   // bit j -> bit 3 j (21 bits)
{
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// Keys of the bodies at current positions [first, n) are written to keys[0 .. n - first).
__global__ void __launch_bounds__(256) keygen_kernel(const double* __restrict__ pos, int first, int n,
                                                     const unsigned long long* __restrict__ maxabs_bits,
                                                     uint64_t* __restrict__ keys, double* __restrict__ bounds_out,
                                                     float* __restrict__ ttab, double theta, float eps2)
{
    const int i = first + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    keys -= first;
    const double bounds = __fma_rn(__longlong_as_double((long long)*maxabs_bits), 1.1, 10.0);
    if (blockIdx.x == 0 && threadIdx.x == 0) *bounds_out = bounds;
    if (blockIdx.x == 0 && threadIdx.x <= MORTON_LEVELS) ttab[threadIdx.x] = cell_threshold((int)threadIdx.x, bounds, theta, eps2);
    if (i >= n) return;
    const double px = pos[3 * (int64_t)i], py = pos[3 * (int64_t)i + 1], pz = pos[3 * (int64_t)i + 2];
    const double inv = 1048576.0 / bounds;   // 2^21 / (2 bounds)
    const double gx = (px + bounds) * inv, gy = (py + bounds) * inv, gz = (pz + bounds) * inv;
    const double fx = floor(gx), fy = floor(gy), fz = floor(gz);
    constexpr double EDGE = 1.0 / 1048576.0;
    const double rx = gx - fx, ry = gy - fy, rz = gz - fz;
    const bool safe = rx > EDGE && rx < 1.0 - EDGE && ry > EDGE && ry < 1.0 - EDGE && rz > EDGE && rz < 1.0 - EDGE &&
                      fx >= 0.0 && fy >= 0.0 && fz >= 0.0 && fx < 2097152.0 && fy < 2097152.0 && fz < 2097152.0;
    if (safe) {
        keys[i] = spread3((uint32_t)fx) | spread3((uint32_t)fy) << 1 | spread3((uint32_t)fz) << 2;
        return;
    }
    double cx = 0.0, cy = 0.0, cz = 0.0, hs = bounds;
    uint64_t k = 0;
#pragma unroll
    for (int l = 0; l < MORTON_LEVELS; ++l) {
        const double q = hs * 0.5;
        const bool bx = px >= cx, by = py >= cy, bz = pz >= cz;
        k = (k << 3) | (uint64_t)((bx ? 1 : 0) | (by ? 2 : 0) | (bz ? 4 : 0));
        cx = bx ? __dadd_rn(cx, q) : __dsub_rn(cx, q);
        cy = by ? __dadd_rn(cy, q) : __dsub_rn(cy, q);
        cz = bz ? __dadd_rn(cz, q) : __dsub_rn(cz, q);
        hs = q;
    }
    keys[i] = k;
}

// ============================================================================ gather
// Physically reorder positions, masses and ids into the new Morton order (nearly the identity after
// the first step, so the reads stay close to coalesced) and emit the float4 view.  Inside step() the
// velocities are NOT moved here: the traversal's epilogue fetches them through the permutation when it
// integrates (vel_in == nullptr); the other callers (getters, the split multi-GPU step) reorder them too.
struct GatherArgs {
    const uint32_t* __restrict__ perm;
    const double* __restrict__ pos_in; const double* __restrict__ vel_in; const double* __restrict__ mass_in;
    const uint32_t* __restrict__ id_in;
    double* __restrict__ pos_out; double* __restrict__ vel_out; double* __restrict__ mass_out;
    uint32_t* __restrict__ id_out;
    float4* __restrict__ posm;
};

__device__ __forceinline__ void gather_block(int block, const GatherArgs& a, int n)
{
    const uint32_t* __restrict__ perm = a.perm;
    const double* __restrict__ pos_in = a.pos_in; const double* __restrict__ vel_in = a.vel_in;
    const double* __restrict__ mass_in = a.mass_in; const uint32_t* __restrict__ id_in = a.id_in;
    double* __restrict__ pos_out = a.pos_out; double* __restrict__ vel_out = a.vel_out;
    double* __restrict__ mass_out = a.mass_out; uint32_t* __restrict__ id_out = a.id_out;
    float4* __restrict__ posm = a.posm;
    const int k = block * 256 + (int)threadIdx.x;
    if (k >= n) return;
    const int64_t j = perm[k];
    const double x = pos_in[3 * j], y = pos_in[3 * j + 1], z = pos_in[3 * j + 2];
    const double m = mass_in[j];
    const int64_t o = 3 * (int64_t)k;
    pos_out[o] = x; pos_out[o + 1] = y; pos_out[o + 2] = z;
    if (vel_in) {
        const double vx = vel_in[3 * j], vy = vel_in[3 * j + 1], vz = vel_in[3 * j + 2];
        vel_out[o] = vx; vel_out[o + 1] = vy; vel_out[o + 2] = vz;
    }
    mass_out[k] = m;
    id_out[k] = id_in[j];
    posm[k] = make_float4((float)x, (float)y, (float)z, (float)m);
}

// ============================================================================ LBVH build
// Common-prefix metric between sorted keys i and i+1 (larger = more similar); equal keys are
// disambiguated by index (Karras 2012), which makes every metric inside a node's range
// strictly larger than the two at its boundary.
__device__ __forceinline__ int cpl(const uint64_t* __restrict__ keys, int i, int n)
{
    if (i < 0 || i >= n - 1) return -1;
    const uint64_t x = keys[i] ^ keys[i + 1];
    if (x) return __clzll((long long)x);                       // 1..63 (bit 63 of a key is 0)
    return 64 + __clz((int)((unsigned)i ^ (unsigned)(i + 1)));  // 64..95
}
// Octree level of the smallest cell containing a node whose metric is c:
// common prefix = c-1 bits = floor((c-1)/3) whole octant digits; >= 21 => one finest cell.
__device__ __forceinline__ int level_of(int c)
{
    const int l = (c - 1) / 3;
    return l > MORTON_LEVELS ? MORTON_LEVELS : l;
}

// delta(i, j): common-prefix metric between sorted keys i and j (any distance apart), -1 outside
// the array.  Equal keys are told apart by their positions (Karras 2012, section 4).
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, uint64_t ki, int i, int64_t j, int n)
{
    if (j < 0 || j >= n) return -1;
    const uint64_t x = ki ^ keys[j];
    if (x) return __clzll((long long)x);
    return 64 + __clz((int)((unsigned)i ^ (unsigned)j));
}

// Karras (2012) binary radix tree: every internal node i in [0, n-2] finds its own key range and
// split by binary search on delta -- independent threads, no atomics, no fences.  Node i covers
// the sorted bodies [range.x, range.y] (i is one of the two ends); children are internal nodes
// (>= 0) or leaves (~k).  Node 0 is the root.  lvl = octree level of the node's smallest cell.
struct KarrasArgs {
    const uint64_t* __restrict__ keys;
    int* __restrict__ childL; int* __restrict__ childR; int* __restrict__ parent;
    int2* __restrict__ range;
    signed char* __restrict__ lvl;
};

__device__ __forceinline__ void karras_block(int block, const KarrasArgs& a, int n)
{
    const uint64_t* __restrict__ keys = a.keys;
    int* __restrict__ childL = a.childL; int* __restrict__ childR = a.childR; int* __restrict__ parent = a.parent;
    int2* __restrict__ range = a.range;
    signed char* __restrict__ lvl = a.lvl;
    const int i = block * 256 + (int)threadIdx.x;
    if (i >= n - 1) return;
    const uint64_t ki = keys[i];
    const int dr = delta(keys, ki, i, (int64_t)i + 1, n), dl = delta(keys, ki, i, (int64_t)i - 1, n);
    const int d = dr > dl ? 1 : -1;
    const int dmin = dr > dl ? dl : dr;
    int64_t lmax = 2;
    while (delta(keys, ki, i, i + lmax * d, n) > dmin) lmax <<= 1;
    int64_t l = 0;
    for (int64_t t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, ki, i, i + (l + t) * d, n) > dmin) l += t;
    const int j = (int)(i + l * d);
    const int dnode = delta(keys, ki, i, j, n);
    int64_t sft = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, ki, i, i + (sft + t) * d, n) > dnode) sft += t;
    } while (t > 1);
    const int gamma = (int)(i + sft * d + (d < 0 ? -1 : 0));
    const int lo = min(i, j), hi = max(i, j);
    const int left = (lo == gamma) ? ~gamma : gamma;
    const int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    childL[i] = left;
    childR[i] = right;
    if (left >= 0) parent[left] = i;
    if (right >= 0) parent[right] = i;
    if (i == 0) parent[0] = -1;
    range[i] = make_int2(lo, hi);
    lvl[i] = (signed char)level_of(dnode);
}

// The physical reorder (HBM-bound, 5 % issue utilisation) and the radix-tree topology (issue-bound,
// 10 % DRAM utilisation) only depend on the sort, not on each other: one launch interleaves their
// CTAs (even blocks gather, odd blocks build topology) so both run on every SM at the same time.
__global__ void __launch_bounds__(256) gather_karras_kernel(GatherArgs g, KarrasArgs k, int n)
{
    const int block = (int)(blockIdx.x >> 1);
    if (blockIdx.x & 1u) karras_block(block, k, n);
    else gather_block(block, g, n);
}

// Mass sums of every node = differences of a prefix sum over the sorted bodies of
// (m x, m y, m z, m) in fp64 -- the bottom-up equivalent of the reference's running-mean COM
// (nbody/simulation.py:160-167) with no tree walk.  The prefix is BLOCKED (1024 bodies per
// block: ploc = prefix inside the block, bex = sum of all earlier blocks) so that small cells
// difference small numbers: the cancellation error of a cell's sum is <= 1e-16 * (block sum)
// for cells inside a block and 1e-16 * (global sum) / (>= 1024 bodies) otherwise, i.e. <= 1e-8
// absolute on a centre of mass at the 50 M preset's scale, 4 orders below fp32 rounding.
constexpr int PFX_BLOCK = 1024;

__device__ __forceinline__ D4 d4_add(const D4& a, const D4& b) { return D4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
__device__ __forceinline__ D4 d4_sub(const D4& a, const D4& b) { return D4{a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
__device__ __forceinline__ D4 d4_shfl_up(const D4& v, int o)
{
    return D4{__shfl_up_sync(0xffffffffu, v.x, o), __shfl_up_sync(0xffffffffu, v.y, o), __shfl_up_sync(0xffffffffu, v.z, o),
              __shfl_up_sync(0xffffffffu, v.w, o)};
}
__device__ __forceinline__ void d4_store(D4* p, const D4& v)
{
    reinterpret_cast<double2*>(p)[0] = make_double2(v.x, v.y);
    reinterpret_cast<double2*>(p)[1] = make_double2(v.z, v.w);
}
__device__ __forceinline__ D4 d4_load(const D4* p)
{
    const double2 a = reinterpret_cast<const double2*>(p)[0], b = reinterpret_cast<const double2*>(p)[1];
    return D4{a.x, a.y, b.x, b.y};
}

__global__ void __launch_bounds__(256) prefix_kernel(const double* __restrict__ pos, const double* __restrict__ mass, int n,
                                                     D4* __restrict__ ploc, D4* __restrict__ btot)
{
    __shared__ D4 wtot[8];
    const int64_t base = (int64_t)blockIdx.x * PFX_BLOCK + 4 * threadIdx.x;
    D4 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int64_t k = base + q;
        if (k < n) {
            const double m = mass[k];
            v[q] = D4{m * pos[3 * k], m * pos[3 * k + 1], m * pos[3 * k + 2], m};
        } else {
            v[q] = D4{0.0, 0.0, 0.0, 0.0};
        }
        if (q) v[q] = d4_add(v[q - 1], v[q]);
    }
    D4 inc = v[3];
    for (int o = 1; o < 32; o <<= 1) {
        const D4 t = d4_shfl_up(inc, o);
        if (lane_id() >= (unsigned)o) inc = d4_add(t, inc);
    }
    const int warp = threadIdx.x >> 5;
    if (lane_id() == 31) wtot[warp] = inc;
    __syncthreads();
    D4 off{0.0, 0.0, 0.0, 0.0};
    for (int w = 0; w < warp; ++w) off = d4_add(off, wtot[w]);
    const D4 excl = d4_add(off, d4_sub(inc, v[3]));
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (base + q < n) d4_store(&ploc[base + q], d4_add(excl, v[q]));
    if (threadIdx.x == 255) d4_store(&btot[blockIdx.x], d4_add(excl, v[3]));
}

// exclusive scan of the block totals, in place (one CTA; <= 50 k blocks at 50 M bodies)
__global__ void __launch_bounds__(1024) prefix_blocks_kernel(D4* __restrict__ btot, int nb)
{
    __shared__ D4 wtot[32];
    __shared__ D4 carry_s;
    if (threadIdx.x == 0) carry_s = D4{0.0, 0.0, 0.0, 0.0};
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int b = base + threadIdx.x;
        const D4 v = b < nb ? d4_load(&btot[b]) : D4{0.0, 0.0, 0.0, 0.0};
        D4 inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            const D4 t = d4_shfl_up(inc, o);
            if (lane_id() >= (unsigned)o) inc = d4_add(t, inc);
        }
        const int warp = threadIdx.x >> 5;
        if (lane_id() == 31) wtot[warp] = inc;
        __syncthreads();
        D4 off = carry_s;
        for (int w = 0; w < warp; ++w) off = d4_add(off, wtot[w]);
        if (b < nb) d4_store(&btot[b], d4_add(off, d4_sub(inc, v)));
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = d4_add(off, inc);
        __syncthreads();
    }
}

// (sum m x, sum m y, sum m z, sum m) over the sorted bodies [l, r]
__device__ __forceinline__ D4 segment_sum(const D4* __restrict__ ploc, const D4* __restrict__ bex, int l, int r)
{
    const int bl = l / PFX_BLOCK, br = r / PFX_BLOCK;
    D4 S = d4_load(&ploc[r]);
    if (l % PFX_BLOCK) S = d4_sub(S, d4_load(&ploc[l - 1]));
    if (bl != br) S = d4_add(d4_sub(d4_load(&bex[br]), d4_load(&bex[bl])), S);
    return S;
}

// ---------------------------------------------------------------------------- octree records
// A binary node is an octree cell ("head") iff its level differs from its parent's; binary
// nodes with the parent's level are sub-octant groupings and are flattened away.  This is the
// reference's octree with single-child chains collapsed to their deepest cell, whose size is
// the one that decides the reference's MAC (all chain cells share mass and COM).
// Pass 1 walks the <= 7 same-level binary descendants of every head once, stores the <= 8
// octree children it finds (kids[8 i .. 8 i + 7], in key = octant order), allocates the cell's
// pair block and packs what a parent needs to know about the cell into ONE 16-byte word
// meta[i] = {range lo, range hi, first pair, children << 5 | level}; pass 2 reads the list back
// and writes whole 64-byte pair records (one meta sector + two prefix-sum sectors per child cell).
// Cells at the finest level (bodies sharing all 63 key bits) are buckets: their children are the
// leaves of their range, any number of them, and are written by a plain loop.
constexpr int KIDS = 8;

struct ChildrenArgs {
    const int* __restrict__ childL; const int* __restrict__ childR; const int* __restrict__ parent;
    const int2* __restrict__ range;
    const signed char* __restrict__ lvl;
    int4* __restrict__ meta;
    unsigned char* __restrict__ ishead;
    int4* __restrict__ kids;
    unsigned* alloc; unsigned capacity; unsigned* error; unsigned* children_total;
    // locally essential tree: null boxes = every cell is materialised
    const int* __restrict__ boxes; const uint64_t* __restrict__ keys; const double* __restrict__ bounds; const float* __restrict__ ttab;
    float eps2;
    int shard_begin, shard_end;
};

// order-preserving float <-> int map (signed integer compare == float compare)
__device__ __forceinline__ int let_f2ord(float f) { const int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }
__device__ __forceinline__ float let_ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

__global__ void __launch_bounds__(256) let_boxes_init_kernel(int* __restrict__ boxes)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < LET_BOXES * 6) boxes[t] = (t % 6) < 3 ? 0x7f7fffff : -0x7f7fffff;   // lo = +FLT_MAX, hi = -FLT_MAX (empty)
}

// Level-3 boxes: bounding box of the shard's bodies (the traversal's targets, sorted positions [begin, end)) per
// 9-bit key prefix.  Neighbours in the sorted order share the prefix: a thread accumulates a running box while the
// prefix stays the same, a CTA (4096 consecutive bodies) combines its threads in shared memory, and the global
// atomics (six per CTA and prefix) stay far from the serialisation limit of the hot boxes of the dense core.
constexpr int LET_L3 = 1 + 8 + 64;   // offset of the level-3 boxes
constexpr int LET_CHUNK = 4096;
__device__ __forceinline__ void let_flush(int* __restrict__ boxes, unsigned pre, const int mn[3], const int mx[3])
{
    for (int d = 0; d < 3; ++d) { atomicMin(&boxes[6 * (LET_L3 + pre) + d], mn[d]); atomicMax(&boxes[6 * (LET_L3 + pre) + 3 + d], mx[d]); }
}
__global__ void __launch_bounds__(256) let_boxes_kernel(const float4* __restrict__ posm, const uint64_t* __restrict__ keys, int begin, int end,
                                                        int* __restrict__ boxes)
{
    __shared__ unsigned s_pre[256];
    __shared__ int s_box[256][6];
    const int base = begin + (int)blockIdx.x * LET_CHUNK;
    unsigned cur = 0xffffffffu;
    int mn[3] = {0x7f7fffff, 0x7f7fffff, 0x7f7fffff}, mx[3] = {-0x7f7fffff, -0x7f7fffff, -0x7f7fffff};
    for (int j = 0; j < LET_CHUNK / 256; ++j) {
        const int k = base + j * 256 + (int)threadIdx.x;
        if (k >= end) break;
        const float4 p = posm[k];
        const unsigned pre = (unsigned)(keys[k] >> (63 - 3 * LET_LEVELS));
        if (pre != cur) {
            if (cur != 0xffffffffu) let_flush(boxes, cur, mn, mx);   // rare: the chunk crosses a prefix boundary
            cur = pre;
            for (int d = 0; d < 3; ++d) { mn[d] = 0x7f7fffff; mx[d] = -0x7f7fffff; }
        }
        const int o[3] = {let_f2ord(p.x), let_f2ord(p.y), let_f2ord(p.z)};
        for (int d = 0; d < 3; ++d) { mn[d] = min(mn[d], o[d]); mx[d] = max(mx[d], o[d]); }
    }
    s_pre[threadIdx.x] = cur;
    for (int d = 0; d < 3; ++d) { s_box[threadIdx.x][d] = mn[d]; s_box[threadIdx.x][3 + d] = mx[d]; }
    __syncthreads();
    // threads holding the same prefix as thread 0 are combined by the first six threads; the others flush themselves
    const unsigned p0 = s_pre[0];
    if (cur != 0xffffffffu && cur != p0) let_flush(boxes, cur, mn, mx);
    if (threadIdx.x < 6 && p0 != 0xffffffffu) {
        const int d = threadIdx.x;
        int v = d < 3 ? 0x7f7fffff : -0x7f7fffff;
        for (int t = 0; t < 256; ++t)
            if (s_pre[t] == p0) v = d < 3 ? min(v, s_box[t][d]) : max(v, s_box[t][d]);
        if (d < 3) atomicMin(&boxes[6 * (LET_L3 + p0) + d], v); else atomicMax(&boxes[6 * (LET_L3 + p0) + d], v);
    }
}

// levels 2, 1, 0 = unions of the eight children (one CTA)
__global__ void __launch_bounds__(512) let_boxes_up_kernel(int* __restrict__ boxes)
{
    const int off[4] = {0, 1, 9, 73};
    for (int l = LET_LEVELS - 1; l >= 0; --l) {
        const int cells = 1 << (3 * l);
        for (int t = threadIdx.x; t < cells * 6; t += blockDim.x) {
            const int c = t / 6, d = t % 6;
            int v = d < 3 ? 0x7f7fffff : -0x7f7fffff;
            for (int q = 0; q < 8; ++q) {
                const int w = boxes[6 * (off[l + 1] + 8 * c + q) + d];
                v = d < 3 ? min(v, w) : max(v, w);
            }
            boxes[6 * (off[l] + c) + d] = v;
        }
        __syncthreads();
    }
}

// every third bit of a 63-bit key, compacted into 21 bits (the inverse of spread3)
__device__ __forceinline__ unsigned let_compact3(uint64_t x)
{
    x &= 0x1249249249249249ull;
    x = (x ^ (x >> 2)) & 0x10c30c30c30c30c3ull;
    x = (x ^ (x >> 4)) & 0x100f00f00f00f00full;
    x = (x ^ (x >> 8)) & 0x001f0000ff0000ffull;
    x = (x ^ (x >> 16)) & 0x001f00000000ffffull;
    x = (x ^ (x >> 32)) & 0x00000000001fffffull;
    return (unsigned)x;
}

// May a body of the shard open the cell (level L, first sorted body's key)?  A body opens a cell iff
// d^2(com, p) + eps^2 <= T.  The centre of mass lies in the cell's cube (the key prefix of its bodies) and p in one of
// the level-3 boxes, so dist^2(cube, box) + eps^2 > T for every box means no body of the shard ever asks for the
// cell's children.  The boxes form an octree (a parent box contains its children): the cell walks it from the top and
// stops at the first level-3 box in range.  Conservative margins: the cube is widened by 1e-6 * bounds (float rounding
// of positions and centres of mass is 6e-8 relative) and T by 1e-5.
__device__ __forceinline__ bool let_needed(const float* __restrict__ sbox, int L, uint64_t key, double bounds, float T, float eps2)
{
    // cell index per axis at level L = the top L of the 21 bits of every third key bit
    const unsigned ix = let_compact3(key) >> (MORTON_LEVELS - L), iy = let_compact3(key >> 1) >> (MORTON_LEVELS - L),
                   iz = let_compact3(key >> 2) >> (MORTON_LEVELS - L);
    const double size = ldexp(2.0 * bounds, -L), pad = 1e-6 * bounds;
    const float clo[3] = {(float)(-bounds + ix * size - pad), (float)(-bounds + iy * size - pad), (float)(-bounds + iz * size - pad)};
    const float chi[3] = {(float)(-bounds + (ix + 1) * size + pad), (float)(-bounds + (iy + 1) * size + pad), (float)(-bounds + (iz + 1) * size + pad)};
    const float Tm = T * 1.00001f;
    auto near = [&](int box) {   // (an empty box has lo = +FLT_MAX, hi = -FLT_MAX: never near)
        const float* bx = sbox + 6 * box;
        float d2 = eps2;
        for (int d = 0; d < 3; ++d) {
            const float q = fmaxf(fmaxf(bx[d] - chi[d], clo[d] - bx[3 + d]), 0.f);
            d2 = fmaf(q, q, d2);
        }
        return d2 <= Tm;
    };
    if (!near(0)) return false;
    for (int a = 0; a < 8; ++a) {
        if (!near(1 + a)) continue;
        for (int b = 0; b < 8; ++b) {
            if (!near(9 + 8 * a + b)) continue;
            for (int c = 0; c < 8; ++c)
                if (near(73 + 64 * a + 8 * b + c)) return true;
        }
    }
    return false;
}

__device__ __forceinline__ void count_children_block(int block, int n, const ChildrenArgs& a, const float* __restrict__ sbox)
{
    const int* __restrict__ childL = a.childL; const int* __restrict__ childR = a.childR; const int* __restrict__ parent = a.parent;
    const int2* __restrict__ range = a.range;
    const signed char* __restrict__ lvl = a.lvl;
    int4* __restrict__ meta = a.meta;
    unsigned char* __restrict__ ishead = a.ishead;
    int4* __restrict__ kids = a.kids;
    unsigned* alloc = a.alloc; const unsigned capacity = a.capacity; unsigned* error = a.error;
    unsigned* children_total = a.children_total;
    const int i = block * 256 + (int)threadIdx.x;
    int cnt = 0;
    int Li = 0;
    int2 rg = make_int2(0, 0);
    bool head = false;
    bool pruned = false;
    if (i < n - 1) {
        Li = lvl[i];
        const int par = parent[i];
        head = par < 0 || lvl[par] != Li;
        if (head) {
            rg = range[i];
            // (a cell holding bodies of the shard is kept without a test: keeping is always safe)
            if (sbox && par >= 0 && (rg.y < a.shard_begin || rg.x >= a.shard_end))
                pruned = !let_needed(sbox, Li, a.keys[rg.x], *a.bounds, a.ttab[Li], a.eps2);
            if (pruned) {
                cnt = 0;
            } else if (Li >= MORTON_LEVELS) {
                cnt = rg.y - rg.x + 1;
            } else {
                int kid[KIDS];
#pragma unroll
                for (int q = 0; q < KIDS; ++q) kid[q] = 0;
                int stack[4];
                int sp = 0;
                int c = childL[i];
                int pending = childR[i];
                // in-order walk without recursion: same-level subtrees are at most 3 deep
                for (;;) {
                    if (c >= 0 && lvl[c] == Li) {
                        stack[sp++] = childR[c];
                        c = childL[c];
                        continue;
                    }
#pragma unroll
                    for (int q = 0; q < KIDS; ++q)
                        if (q == cnt) kid[q] = c;
                    ++cnt;
                    if (sp > 0) c = stack[--sp];
                    else if (pending != 0x7fffffff) { c = pending; pending = 0x7fffffff; }
                    else break;
                }
                kids[2 * (int64_t)i] = make_int4(kid[0], kid[1], kid[2], kid[3]);
                kids[2 * (int64_t)i + 1] = make_int4(kid[4], kid[5], kid[6], kid[7]);
            }
        }
    }
    // block-aggregated allocation of the pair blocks: ONE atomic per CTA on the shared counter
    // (a per-thread atomicAdd on one address serialises in L2: 21 ms at 50 M bodies).
    __shared__ unsigned wsum[8];
    __shared__ unsigned wkids[8];
    __shared__ unsigned s_base;
    unsigned nk = (unsigned)cnt;
    for (int o = 16; o > 0; o >>= 1) nk += __shfl_xor_sync(0xffffffffu, nk, o);
    if (lane_id() == 0) wkids[threadIdx.x >> 5] = nk;
    const unsigned npair = (unsigned)(cnt + 1) >> 1;   // children are stored two per 64-byte pair record
    unsigned inc = npair;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane_id() >= (unsigned)o) inc += t;
    }
    const int warp = threadIdx.x >> 5;
    if (lane_id() == 31) wsum[warp] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        unsigned ktot = 0;
        for (int w = 0; w < 8; ++w) { const unsigned t = wsum[w]; wsum[w] = tot; tot += t; ktot += wkids[w]; }
        if (ktot) atomicAdd(children_total, ktot);
        unsigned base = tot ? atomicAdd(alloc, tot) : 0u;
        if (base + tot > capacity) { atomicOr(error, (unsigned)ERR_RECORD_OVERFLOW); base = 0xffffffffu; }
        s_base = base;
    }
    __syncthreads();
    if (i < n - 1) ishead[i] = head ? 1 : 0;
    // one 16-byte record per octree cell: everything write_records needs about a child cell
    if (!head) return;
    if (s_base == 0xffffffffu) cnt = 0;
    // a pruned cell keeps its own record in the parent (mass, centre, threshold) and says so in the child count
    meta[i] = make_int4(rg.x, rg.y, cnt ? (int)(s_base + wsum[warp] + inc - npair) : -1,
                        (int)(((pruned ? PRUNED_NCHILD : (unsigned)cnt) << 5) | (unsigned)Li));
}

__global__ void __launch_bounds__(256) count_children_kernel(int n, ChildrenArgs ca)
{
    count_children_block((int)blockIdx.x, n, ca, nullptr);
}

// sharded step with the locally essential tree test: persistent CTAs load the shard's boxes into shared memory once
__global__ void __launch_bounds__(256) count_children_let_kernel(int n, int nblocks, ChildrenArgs ca)
{
    __shared__ float sbox[6 * LET_BOXES];
    for (int t = threadIdx.x; t < 6 * LET_BOXES; t += 256) sbox[t] = let_ord2f(ca.boxes[t]);
    __syncthreads();
    for (int block = (int)blockIdx.x; block < nblocks; block += (int)gridDim.x) {
        count_children_block(block, n, ca, sbox);
        __syncthreads();   // the block's shared scratch is reused by the next iteration
    }
}

// Pair records: children 2j and 2j+1 of a cell share one 64-byte record laid out for packed
// fp32x2 math in the traversal (component c = child & 1):
//   q0 = {x0, x1, y0, y1}   q1 = {z0, z1, m0, m1}   q2 = {T0, T1, -, -}
//   q3 = {first0, first1, nchild0, nchild1}
// (q2.z is where the traversal's staging copy carries the pair's lane mask)
// T = max(size^2/theta^2, eps^2) for a cell (clamped to REC_T_MAX; theta = 0 => REC_T_MAX, "always
// open"), eps^2 for a leaf.  An odd child count is padded with a massless dummy far away.
constexpr float REC_T_MAX = 1e30f;        // above any real squared distance, below the sentinels'
constexpr float REC_LANE_SENTINEL = 1e18f;   // x of a lane that is not in an entry's mask
constexpr float REC_DUMMY_X = 3e18f;      // x of a padding child

struct ChildRec { float x, y, z, m, T; int first, nchild, body; };

__device__ __forceinline__ ChildRec dummy_child()
{
    return ChildRec{REC_DUMMY_X, 0.f, 0.f, 0.f, 0.f, 0, 0, -1};
}

__device__ float cell_threshold(int level, double bounds, double theta, float eps2)
{
    // cell size = 2*bounds / 2^level; MAC  size/d < theta  <=>  d^2 > size^2/theta^2
    const double size = ldexp(2.0 * bounds, -level);
    float T = REC_T_MAX;
    if (theta > 0.0) T = fminf(fmaxf((float)fmin((size * size) / (theta * theta), 1e31), eps2), REC_T_MAX);
    return T;
}

__device__ __forceinline__ ChildRec cell_child(const D4& S, float T, int first, int nchild)
{
    const double inv = S.w > 0.0 ? 1.0 / S.w : 0.0;
    return ChildRec{(float)(S.x * inv), (float)(S.y * inv), (float)(S.z * inv), (float)S.w,
                    T, first, nchild, -1};
}

struct TreeView {
    const int4* __restrict__ meta;
    const D4* __restrict__ ploc;
    const D4* __restrict__ bex;
    const float4* __restrict__ posm;
    const float* __restrict__ ttab;   // MAC threshold per octree level (written by keygen)
};

__device__ __forceinline__ ChildRec load_child(int c, const TreeView& tv, float eps2)
{
    if (c < 0) {
        const int k = ~c;
        const float4 b = tv.posm[k];
        return ChildRec{b.x, b.y, b.z, b.w, eps2, 0, 0, k};
    }
    const int4 m = __ldg(&tv.meta[c]);
    return cell_child(segment_sum(tv.ploc, tv.bex, m.x, m.y), __ldg(&tv.ttab[m.w & 31]), m.z, m.w >> 5);
}

__device__ __forceinline__ void store_pair(float4* __restrict__ recs, int64_t pair, const ChildRec& a, const ChildRec& b)
{
    float4* q = recs + 4 * pair;
    q[0] = make_float4(a.x, b.x, a.y, b.y);
    q[1] = make_float4(a.z, b.z, a.m, b.m);
    q[2] = make_float4(a.T, b.T, 0.f, 0.f);
    q[3] = make_float4(__int_as_float(a.first), __int_as_float(b.first), __int_as_float(a.nchild), __int_as_float(b.nchild));
}

// Eight lanes per octree cell, one per child: all of a cell's children are fetched at once (meta ->
// two prefix-sum entries, or the leaf's float4), neighbouring lanes swap their results and the even
// lane writes the 64-byte pair record.  A warp covers 32 consecutive binary nodes and serves its
// cells four at a time.
__device__ __forceinline__ ChildRec shfl_down_child(unsigned mask, const ChildRec& r)
{
    ChildRec o;
    o.x = __shfl_down_sync(mask, r.x, 1); o.y = __shfl_down_sync(mask, r.y, 1); o.z = __shfl_down_sync(mask, r.z, 1);
    o.m = __shfl_down_sync(mask, r.m, 1); o.T = __shfl_down_sync(mask, r.T, 1);
    o.first = __shfl_down_sync(mask, r.first, 1); o.nchild = __shfl_down_sync(mask, r.nchild, 1);
    o.body = -1;
    return o;
}

__global__ void __launch_bounds__(256, 6) write_records_kernel(int n, TreeView tv, const unsigned char* __restrict__ ishead,
                                                            const int* __restrict__ kids, float eps2,
                                                            float4* __restrict__ recs)
{
    const unsigned lane = lane_id();
    const int w0 = (int)((blockIdx.x * blockDim.x + threadIdx.x) & ~31u);
    const int i = w0 + (int)lane;
    int4 mi = make_int4(0, 0, 0, 0);
    bool h = false;
    if (i < n - 1 && ishead[i]) {
        mi = __ldg(&tv.meta[i]);
        const unsigned nc = (unsigned)mi.w >> 5;
        h = nc != 0u && nc != PRUNED_NCHILD;
    }
    unsigned hm = __ballot_sync(0xffffffffu, h);
    if (!hm) return;
    const int sub = (int)(lane >> 3), c = (int)(lane & 7u);
    const unsigned gmask = 0xffu << (8 * sub);
    while (hm) {
        const unsigned src = __fns(hm, 0, sub + 1);   // lane of the sub-th remaining cell, or 0xffffffff
        const bool have = src != 0xffffffffu;
        const int sl = have ? (int)src : 0;
        const int4 m = make_int4(__shfl_sync(0xffffffffu, mi.x, sl), __shfl_sync(0xffffffffu, mi.y, sl),
                                 __shfl_sync(0xffffffffu, mi.z, sl), __shfl_sync(0xffffffffu, mi.w, sl));
        if (have) {
            const int node = w0 + sl;
            const int nc = m.w >> 5, Li = m.w & 31;
            const int64_t base = m.z;
            const bool bucket = Li >= MORTON_LEVELS;   // cell at the finest level: its children are the leaves of its range
            if (node == 0 && c == 0)   // node 0 is the root; pair 0 = {root cell, dummy}
                store_pair(recs, 0, cell_child(segment_sum(tv.ploc, tv.bex, 0, n - 1), __ldg(&tv.ttab[Li]), (int)base, nc),
                           dummy_child());
            for (int c0 = 0; c0 < nc; c0 += 8) {
                const int cc = c0 + c;
                ChildRec rec = dummy_child();
                if (cc < nc) {
                    const int kid = bucket ? ~(m.x + cc) : __ldg(&kids[8 * (int64_t)node + cc]);
                    rec = load_child(kid, tv, eps2);
                }
                const ChildRec nxt = shfl_down_child(gmask, rec);
                if ((cc & 1) == 0 && cc < nc) store_pair(recs, base + (cc >> 1), rec, nxt);
            }
        }
        hm &= hm - 1; hm &= hm - 1; hm &= hm - 1; hm &= hm - 1;   // (x & (x-1) of 0 stays 0)
    }
}

// single body: the root record is that leaf
__global__ void single_body_record_kernel(const float4* __restrict__ posm, float eps2, float4* __restrict__ recs)
{
    const float4 b = posm[0];
    store_pair(recs, 0, ChildRec{b.x, b.y, b.z, b.w, eps2, 0, 0, 0}, dummy_child());
}

#include "traverse.cuh"   // the traversal kernels (theta-MAC walk)

// ============================================================================ integrate
// v = (v + a dt) * damping; x += v dt (nbody/simulation.py:281-305; CUDA twin
// gpu_backend.py:242-257), fp64 state, with the next step's max|coord| reduce fused in.
__global__ void __launch_bounds__(256) integrate_kernel(double* __restrict__ pos, double* __restrict__ vel,
                                                        const float4* __restrict__ acc, int n, double dt, double damping,
                                                        unsigned long long* __restrict__ maxabs_out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    double m = 0.0;
    if (k < n) {
        const float4 a = acc[k];
        const int64_t o = 3 * (int64_t)k;
        const double a3[3] = {(double)a.x, (double)a.y, (double)a.z};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double v = vel[o + d];
            v += a3[d] * dt;
            v *= damping;
            vel[o + d] = v;
            const double x = pos[o + d] + v * dt;
            pos[o + d] = x;
            m = fmax(m, fabs(x));
        }
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ double wm[8];
    if (lane_id() == 0) wm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < 8 ? wm[threadIdx.x] : 0.0;
        for (int o = 4; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) atomicMax(maxabs_out, (unsigned long long)__double_as_longlong(m));
    }
}

// ============================================================================ colours / egress
// speed -> RGB heat map (nbody/simulation.py:320-400; gpu_backend.py:259-325), written in
// creation order through id[].
__device__ __forceinline__ void speed_color(double t, float& r, float& g, float& b)
{
    double R, Gc, B;
    if (t < 0.55) {
        if (t < 0.15) { const double s = t / 0.15; R = 0.4 - 0.2 * s; Gc = 0.2 + 0.2 * s; B = 0.8 + 0.1 * s; }
        else if (t < 0.30) { const double s = (t - 0.15) / 0.15; R = 0.2 + 0.1 * s; Gc = 0.4 + 0.1 * s; B = 0.9 + 0.05 * s; }
        else {
            const double s = (t - 0.30) / 0.25;
            if (s < 0.6) { const double s2 = s / 0.6; R = 0.3 - 0.1 * s2; Gc = 0.5 + 0.3 * s2; B = 0.95 + 0.05 * s2; }
            else { const double s2 = (s - 0.6) / 0.4; R = 0.2 + 0.8 * s2; Gc = 0.8 + 0.2 * s2; B = 1.0; }
        }
    } else if (t < 0.90) { R = 1.0; Gc = 1.0; B = 1.0; }
    else if (t < 0.95) { const double s = (t - 0.90) / 0.05; R = 1.0; Gc = 1.0 - 0.05 * s; B = 1.0 - 1.0 * s; }
    else if (t < 0.99) { const double s = (t - 0.95) / 0.04; R = 1.0; Gc = 0.95 - 0.45 * s; B = 0.0; }
    else { const double s = (t - 0.99) / 0.01; R = 1.0; Gc = 0.5 - 0.5 * s; B = 0.0; }
    r = (float)R; g = (float)Gc; b = (float)B;
}

__global__ void __launch_bounds__(256) colors_kernel(const double* __restrict__ vel, const uint32_t* __restrict__ id,
                                                     float* __restrict__ colors, int n, double max_speed)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t o = 3 * (int64_t)k;
    const double vx = vel[o], vy = vel[o + 1], vz = vel[o + 2];
    const double speed = sqrt(vx * vx + vy * vy + vz * vz);
    const double t = fmin(1.0, speed / max_speed);
    float r, g, b;
    speed_color(t, r, g, b);
    const int64_t w = 3 * (int64_t)id[k];
    colors[w] = r; colors[w + 1] = g; colors[w + 2] = b;
}

template <typename T>
__global__ void __launch_bounds__(256) unpermute3_kernel(const double* __restrict__ src, const uint32_t* __restrict__ id,
                                                         T* __restrict__ dst, int n)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t o = 3 * (int64_t)k, w = 3 * (int64_t)id[k];
    dst[w] = (T)src[o]; dst[w + 1] = (T)src[o + 1]; dst[w + 2] = (T)src[o + 2];
}

__global__ void __launch_bounds__(256) unpermute_acc_kernel(const float4* __restrict__ acc, const uint32_t* __restrict__ id,
                                                            float* __restrict__ dst, int n)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 a = acc[k];
    const int64_t w = 3 * (int64_t)id[k];
    dst[w] = a.x; dst[w + 1] = a.y; dst[w + 2] = a.z;
}

__global__ void __launch_bounds__(256) iota_kernel(uint32_t* __restrict__ p, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}


// ============================================================================ host side
template <typename T>
static T* alloc_counted(NBodySim& s, size_t count)
{
    s.bytes_allocated += (count ? count : 1) * sizeof(T);
    return dev_alloc<T>(count);
}

void nbody_alloc(NBodySim& s, int n)
{
    s.n = n;
    B200_CHECK(cudaSetDevice(s.device));
    cudaDeviceProp prop;
    B200_CHECK(cudaGetDeviceProperties(&prop, s.device));
    s.sm_count = prop.multiProcessorCount;
    B200_CHECK(cudaStreamCreateWithFlags(&s.own_stream, cudaStreamNonBlocking));
    s.stream = s.own_stream;
    const size_t N = (size_t)n;
    for (int b = 0; b < 3; ++b) s.pos[b] = alloc_counted<double>(s, 3 * N);
    for (int b = 0; b < 2; ++b) {
        s.vel[b] = alloc_counted<double>(s, 3 * N);
        s.mass[b] = alloc_counted<double>(s, N);
        s.id[b] = alloc_counted<uint32_t>(s, N);
        s.keys[b] = alloc_counted<uint64_t>(s, N);
        s.vals[b] = alloc_counted<uint32_t>(s, N);
    }
    s.mass0 = alloc_counted<double>(s, N);
    s.sorter.init(n);
    s.bytes_allocated += s.sorter.bytes();
    s.posm = alloc_counted<float4>(s, N);
    s.acc_capacity = (int64_t)N + 64 * 64;   // room for padded equal slices (whole 64-body tiles) up to 64 ranks
    s.acc = alloc_counted<float4>(s, (size_t)s.acc_capacity);
    s.childL = alloc_counted<int>(s, N);
    s.childR = alloc_counted<int>(s, N);
    s.parent = alloc_counted<int>(s, N);
    s.range = alloc_counted<int2>(s, N);
    s.ploc = alloc_counted<D4>(s, N);
    s.bex = alloc_counted<D4>(s, N / PFX_BLOCK + 2);
    s.lvl = alloc_counted<signed char>(s, N);
    s.kids = alloc_counted<int4>(s, 2 * N);
    s.meta = alloc_counted<int4>(s, N);
    s.ishead = alloc_counted<unsigned char>(s, N);
    B200_REQUIRE(N < ((size_t)1 << 26), "too many bodies for the packed child count of a cell");
    // pair records: <= (children + cells) / 2 <= 1.5 N pairs, + the root pair
    s.rec_capacity = (3 * (int64_t)N) / 2 + 16;
    B200_REQUIRE(s.rec_capacity < (int64_t)(1u << 29), "too many bodies for the 29-bit pair index of a stack entry");
    s.recs = alloc_counted<float4>(s, 4 * (size_t)s.rec_capacity);
    s.colors = alloc_counted<float>(s, 3 * N);
    s.stage = alloc_counted<double>(s, 3 * N);
    s.d_maxabs = alloc_counted<unsigned long long>(s, 2);
    s.d_bounds = alloc_counted<double>(s, 1);
    s.d_ttab = alloc_counted<float>(s, 32);
    s.d_boxes = alloc_counted<int>(s, 6 * LET_BOXES);
    s.d_root = alloc_counted<int>(s, 1);
    s.d_alloc = alloc_counted<unsigned>(s, 1);
    s.d_tile_counter = alloc_counted<unsigned>(s, 1);
    s.d_children = alloc_counted<unsigned>(s, 1);
    B200_CHECK(cudaMemset(s.d_children, 0, sizeof(unsigned)));
    s.d_interactions = alloc_counted<unsigned long long>(s, TRAV_COUNTERS);
    s.d_error = alloc_counted<unsigned>(s, 1);
    B200_CHECK(cudaMemset(s.d_error, 0, sizeof(unsigned)));
    B200_CHECK(cudaHostAlloc(&s.h_error, sizeof(unsigned), cudaHostAllocDefault));
    *s.h_error = 0;
    B200_CHECK(cudaMemset(s.d_interactions, 0, TRAV_COUNTERS * sizeof(unsigned long long)));
    B200_CHECK(cudaFuncSetAttribute(traverse_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRAV_SMEM_BYTES));
    B200_CHECK(cudaFuncSetAttribute(traverse_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRAV_SMEM_BYTES));
    B200_CHECK(cudaFuncSetAttribute(traverse_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRAV_SMEM_BYTES));
    B200_CHECK(cudaFuncSetAttribute(traverse_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRAV_SMEM_BYTES));
    {
        // B200_TRAV = 32 | 64 forces a walk; default: chosen per launch (see nbody_traverse)
        // (tests shrink these to exercise the bucketed un-permute at small n)
        if (const char* v = getenv("B200_UNPERM_MIN_N")) s.unperm_min_n = atoi(v);
        if (const char* v = getenv("B200_UNPERM_SHIFT")) s.unperm_shift = max(1, min(30, atoi(v)));
        if (const char* v = getenv("B200_LET")) s.let_enabled = atoi(v) != 0;
        if (const char* v = getenv("B200_REC_CAPACITY")) s.rec_capacity_override = atoll(v);   // tests: force a record pool overflow
        const char* ng = getenv("B200_NO_GRAPH");   // plain launches instead of the captured step (debugging, A/B timing)
        s.use_graph = !(ng && ng[0] == '1');
        const char* mode = getenv("B200_TRAV");
        s.trav_mode = !mode ? 0 : (mode[0] == '6' ? 64 : 32);
    }
    B200_CHECK(cudaFuncSetAttribute(traverse64c_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRAV64C_SMEM_BYTES));
    B200_CHECK(cudaFuncSetAttribute(traverse64c_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRAV64C_SMEM_BYTES));
    B200_CHECK(cudaFuncSetAttribute(traverse64c_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRAV64C_SMEM_BYTES));
    B200_CHECK(cudaFuncSetAttribute(traverse64c_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRAV64C_SMEM_BYTES));
    B200_CHECK(cudaMemset(s.colors, 0, 3 * N * sizeof(float)));
    s.shard_begin = 0;
    s.shard_end = n;
    s.timer.init();
}

static void async_free(NBodySim& s);

void nbody_free(NBodySim& s)
{
    cudaSetDevice(s.device);
    if (s.stream) cudaStreamSynchronize(s.stream);
    nbody_graphs_reset(s);
    for (int b = 0; b < 2; ++b) {
        cudaFree(s.vel[b]); cudaFree(s.mass[b]); cudaFree(s.id[b]);
        cudaFree(s.keys[b]); cudaFree(s.vals[b]);
    }
    for (int b = 0; b < 3; ++b) cudaFree(s.pos[b]);
    cudaFree(s.mass0);
    s.sorter.destroy();
    if (s.ms_keys) { cudaFree(s.ms_keys); cudaFree(s.ms_vals); s.ms_keys = nullptr; s.ms_vals = nullptr; }
    cudaFree(s.posm); cudaFree(s.acc); cudaFree(s.childL); cudaFree(s.childR); cudaFree(s.parent);
    cudaFree(s.range); cudaFree(s.ploc); cudaFree(s.bex); cudaFree(s.meta); cudaFree(s.ishead);
    cudaFree(s.lvl); cudaFree(s.kids);
    cudaFree(s.recs); cudaFree(s.colors); cudaFree(s.stage); cudaFree(s.d_maxabs); cudaFree(s.d_bounds); cudaFree(s.d_ttab); cudaFree(s.d_boxes);
    cudaFree(s.d_children);
    cudaFree(s.d_root); cudaFree(s.d_alloc); cudaFree(s.d_tile_counter); cudaFree(s.d_interactions);
    cudaFree(s.d_error);
    if (s.h_error) { cudaFreeHost(s.h_error); s.h_error = nullptr; }
    s.timer.destroy();
    if (s.vis_counts) { cudaFree(s.vis_counts); cudaFreeHost(s.vis_total_host); s.vis_counts = nullptr; s.vis_total_host = nullptr; }
    if (s.vis_pos) { cudaFree(s.vis_pos); cudaFree(s.vis_col); s.vis_pos = s.vis_col = nullptr; }
    async_free(s);
    if (s.own_stream) cudaStreamDestroy(s.own_stream);
    s.own_stream = nullptr;
    s.stream = nullptr;
}

static void recompute_maxabs(NBodySim& s)
{
    s.maxabs_slot = 0;
    B200_CHECK(cudaMemsetAsync(s.d_maxabs, 0, 2 * sizeof(unsigned long long), s.stream));
    if (s.n > 0) {
        const int64_t count = 3 * (int64_t)s.n;
        const int64_t want = (count + 255) / 256, cap = (int64_t)s.sm_count * 16;
        const int blocks = (int)(want < cap ? want : cap);
        absmax_kernel<<<blocks, 256, 0, s.stream>>>(s.pos[s.pcur], count, s.d_maxabs);
        ++s.launches;
        B200_CHECK(cudaGetLastError());
    }
}

void nbody_upload(NBodySim& s, const double* pos, const double* vel, const double* mass)
{
    B200_CHECK(cudaSetDevice(s.device));
    const size_t N = (size_t)s.n;
    s.pcur = 0;
    s.vcur = 0;
    B200_CHECK(cudaMemcpyAsync(s.pos[0], pos, 3 * N * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    B200_CHECK(cudaMemcpyAsync(s.vel[0], vel, 3 * N * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    B200_CHECK(cudaMemcpyAsync(s.mass[0], mass, N * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    B200_CHECK(cudaMemcpyAsync(s.mass0, s.mass[0], N * sizeof(double), cudaMemcpyDeviceToDevice, s.stream));   // creation order, kept
    if (s.n > 0) iota_kernel<<<div_up(s.n, 256), 256, 0, s.stream>>>(s.id[0], s.n);
    recompute_maxabs(s);
    s.tree_valid = false;
    B200_CHECK(cudaStreamSynchronize(s.stream));
}

// Initial state drawn on the device, straight into the state buffers (generate.cu).
void nbody_generate(NBodySim& s, int dist, double R, double G_dist, uint64_t seed)
{
    B200_CHECK(cudaSetDevice(s.device));
    s.pcur = 0;
    s.vcur = 0;
    generate_device(dist, s.n, R, G_dist, seed, s.pos[0], s.vel[0], s.mass[0], s.stream, s.sm_count);
    B200_CHECK(cudaMemcpyAsync(s.mass0, s.mass[0], (size_t)s.n * sizeof(double), cudaMemcpyDeviceToDevice, s.stream));
    if (s.n > 0) iota_kernel<<<div_up(s.n, 256), 256, 0, s.stream>>>(s.id[0], s.n);
    recompute_maxabs(s);
    s.tree_valid = false;
    B200_CHECK(cudaStreamSynchronize(s.stream));
}

// New positions / velocities for the same bodies, given in creation order.
void nbody_upload_state(NBodySim& s, const double* pos, const double* vel)
{
    B200_CHECK(cudaSetDevice(s.device));
    const size_t N = (size_t)s.n;
    // bring masses back to creation order alongside: simplest is to scatter through id[]
    // into the other buffer, then continue from there with id = iota.
    const int po = (s.pcur + 1) % 3, o = s.vcur ^ 1;
    B200_CHECK(cudaMemcpyAsync(s.pos[po], pos, 3 * N * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    B200_CHECK(cudaMemcpyAsync(s.vel[o], vel, 3 * N * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    if (s.n > 0) {
        // masses back in creation order: a copy of the kept array (a scatter through id[] costs 2 ms at 50 M)
        B200_CHECK(cudaMemcpyAsync(s.mass[o], s.mass0, (size_t)s.n * sizeof(double), cudaMemcpyDeviceToDevice, s.stream));
        iota_kernel<<<div_up(s.n, 256), 256, 0, s.stream>>>(s.id[o], s.n);
    }
    s.pcur = po;
    s.vcur = o;
    recompute_maxabs(s);
    s.tree_valid = false;
    B200_CHECK(cudaStreamSynchronize(s.stream));
}

static void build_after_sort(NBodySim& s, bool for_step);
static void build_tree_impl(NBodySim& s, bool for_step);

void nbody_build_tree(NBodySim& s) { build_tree_impl(s, false); }

static void build_tree_impl(NBodySim& s, bool for_step)
{
    B200_CHECK(cudaSetDevice(s.device));
    const int n = s.n;
    if (n == 0) { s.tree_valid = true; return; }
    const int grid = div_up(n, 256);
    cudaStream_t st = s.stream;
    s.timer.begin(st);
    // ---- keys
    keygen_kernel<<<grid, 256, 0, st>>>(s.pos[s.pcur], 0, n, s.d_maxabs + s.maxabs_slot, s.keys[0], s.d_bounds, s.d_ttab, s.theta,
                                        (float)(s.softening * s.softening));
    ++s.launches;
    B200_CHECK(cudaGetLastError());
    s.timer.mark(st);
    // ---- sort (key, position)
    s.sorted_slot = s.sorter.sort(s.keys, s.vals, 0, n, 0, 64, /*iota=*/true, st, s.sm_count);
    s.launches += s.sorter.last_launches;
    s.timer.mark(st);
    build_after_sort(s, for_step);
}

// ---------------------------------------------------------------------------- sharded sort (multi-GPU)
// Every rank holds the full state in last step's Morton order.  Rank r generates keys for and sorts only
// the bodies at current positions [r S, (r+1) S) (S = slice); the sorted slices are all-gathered by the
// host plumbing (12 B/body) and merged here by counting: the global rank of an element of run r is its
// local rank + the number of elements of every earlier run that are <= it + of every later run that
// are < it -- the result of a stable sort of the whole array, bit-identical on every rank.
__device__ __forceinline__ int count_le(const uint64_t* __restrict__ run, int lo, int hi, uint64_t k)
{   // number of elements <= k in the sorted run, known to lie in [lo, hi]
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (run[mid] <= k) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int count_lt(const uint64_t* __restrict__ run, int lo, int hi, uint64_t k)
{   // number of elements < k in the sorted run, known to lie in [lo, hi]
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (run[mid] < k) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// number of elements <= k (le) or < k (!le) of the sorted run, known to lie in [lo, hi]: the whole warp probes 32
// positions per round (all lanes pass the same arguments and get the same result)
__device__ __forceinline__ int warp_count(const uint64_t* __restrict__ run, int lo, int hi, uint64_t k, bool le)
{
    const int lane = (int)lane_id();
    while (hi - lo > 32) {
        const int step = (hi - lo + 32) / 33;
        const int idx = lo + (lane + 1) * step - 1;
        bool pred = false;
        if (idx < hi) { const uint64_t v = run[idx]; pred = le ? v <= k : v < k; }
        const int c = __popc(__ballot_sync(0xffffffffu, pred));   // pred is monotone along the lanes
        const int nlo = lo + c * step;
        if (c < 32) hi = min(hi, lo + (c + 1) * step - 1);
        lo = nlo;
    }
    bool pred = false;
    if (lo + lane < hi) { const uint64_t v = run[lo + lane]; pred = le ? v <= k : v < k; }
    return lo + __popc(__ballot_sync(0xffffffffu, pred));
}

// A CTA merges MERGE_TILE consecutive elements of one run.  The counts in every other run are monotone
// along the tile, so one thread per other run brackets them with two full binary searches at the tile's
// first and last key; runs that were Morton ranges one step ago barely interleave, the bracket is
// usually empty (count constant over the tile) and otherwise a few elements wide.
constexpr int MERGE_TILE = 4096;
constexpr int MERGE_MAX_WORLD = 64;

__global__ void __launch_bounds__(256) merge_runs_kernel(const uint64_t* __restrict__ rkeys, const uint32_t* __restrict__ rvals,
                                                         int slice, int world, int n, int tiles_per_run,
                                                         uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out)
{
    __shared__ int s_lo[MERGE_MAX_WORLD], s_hi[MERGE_MAX_WORLD];
    const int r = (int)(blockIdx.x / tiles_per_run), tile = (int)(blockIdx.x % tiles_per_run);
    const int begin = r * slice;
    const int len = max(0, min(slice, n - begin));
    const int j0 = tile * MERGE_TILE;
    if (j0 >= len) return;
    const int j1 = min(j0 + MERGE_TILE, len) - 1;     // last element of the tile
    const uint64_t* __restrict__ mine = rkeys + begin;
    // one warp per other run: 32-ary searches (5 dependent rounds of loads instead of 23) for the counts at the tile's
    // first and last key; runs entirely below or above the tile are decided by two loads
    for (int q = (int)(threadIdx.x >> 5); q < world; q += 8) {
        const int qb = q * slice;
        const int ql = max(0, min(slice, n - qb));
        int lo = 0, hi = 0;
        if (q != r && ql > 0) {
            const uint64_t kf = mine[j0], kl = mine[j1];
            const uint64_t* __restrict__ run = rkeys + qb;
            const uint64_t qf = run[0], qlast = run[ql - 1];
            const bool le = q < r;
            if (le ? qlast <= kf : qlast < kf) lo = hi = ql;
            else if (le ? qf > kl : qf >= kl) lo = hi = 0;
            else { lo = warp_count(run, 0, ql, kf, le); hi = warp_count(run, lo, ql, kl, le); }
        }
        if (lane_id() == 0) { s_lo[q] = lo; s_hi[q] = hi; }
    }
    __syncthreads();
    // the constant part of every element's rank (runs whose count does not change over the tile) is summed once
    int base_rank = 0;
    unsigned long long open_runs = 0ull;   // runs with a non-empty bracket: per-element searches
    for (int q = 0; q < world; ++q) {
        if (s_lo[q] == s_hi[q]) base_rank += s_lo[q];
        else open_runs |= 1ull << q;
    }
#pragma unroll 4
    for (int j = j0 + (int)threadIdx.x; j <= j1; j += 256) {
        const uint64_t k = mine[j];
        const uint32_t v = rvals[begin + j];
        int rank = j + base_rank;
        for (unsigned long long m = open_runs; m; m &= m - 1) {
            const int q = __ffsll((long long)m) - 1;
            rank += q < r ? count_le(rkeys + q * slice, s_lo[q], s_hi[q], k) : count_lt(rkeys + q * slice, s_lo[q], s_hi[q], k);
        }
        keys_out[rank] = k;
        vals_out[rank] = (uint32_t)begin + v;   // local sort position -> position in the current arrays
    }
}

void nbody_ms_setup(NBodySim& s, int slice, int world)
{
    B200_CHECK(cudaSetDevice(s.device));
    B200_REQUIRE(slice > 0 && world > 0 && (int64_t)slice * world >= s.n && slice % 32 == 0, "bad slice / world for the sharded sort");
    if (s.ms_keys && s.ms_slice == slice && s.ms_world == world) return;
    B200_CHECK(cudaStreamSynchronize(s.stream));
    if (s.ms_keys) { cudaFree(s.ms_keys); cudaFree(s.ms_vals); }
    s.ms_keys = alloc_counted<uint64_t>(s, (size_t)slice * world);
    s.ms_vals = alloc_counted<uint32_t>(s, (size_t)slice * world);
    s.ms_slice = slice;
    s.ms_world = world;
}

// keygen + local radix sort of the bodies at current positions [rank*slice, (rank+1)*slice) into the
// rank's part of the exchange buffers
void nbody_ms_sort_local(NBodySim& s, int rank)
{
    B200_CHECK(cudaSetDevice(s.device));
    B200_REQUIRE(s.ms_keys && rank >= 0 && rank < s.ms_world, "sharded sort is not set up");
    cudaStream_t st = s.stream;
    const int begin = min(rank * s.ms_slice, s.n);
    const int len = max(0, min(s.ms_slice, s.n - begin));
    s.timer.begin(st);
    uint64_t* k2[2] = {s.ms_keys + (size_t)rank * s.ms_slice, s.keys[1] + begin};
    uint32_t* v2[2] = {s.ms_vals + (size_t)rank * s.ms_slice, s.vals[1] + begin};
    // (rank 0's launch also writes bounds and the threshold table; every rank runs with len >= 0, and
    // a rank with an empty slice still needs them)
    keygen_kernel<<<max(1, div_up(len, 256)), 256, 0, st>>>(s.pos[s.pcur], begin, begin + len, s.d_maxabs + s.maxabs_slot, k2[0],
                                                             s.d_bounds, s.d_ttab, s.theta, (float)(s.softening * s.softening));
    ++s.launches;
    B200_CHECK(cudaGetLastError());
    s.timer.mark(st);
    if (len > 0) {
        const int slot = s.sorter.sort(k2, v2, 0, len, 0, 64, /*iota=*/true, st, s.sm_count);
        B200_REQUIRE(slot == 0, "the sharded sort expects an even number of passes");
        s.launches += s.sorter.last_launches;
    }
    s.timer.mark(st);
}

// after the host plumbing has all-gathered both exchange buffers: merge + the rest of the tree build
void nbody_build_tree_presorted(NBodySim& s)
{
    B200_CHECK(cudaSetDevice(s.device));
    B200_REQUIRE(s.ms_keys, "sharded sort is not set up");
    if (s.n == 0) { s.tree_valid = true; return; }
    cudaStream_t st = s.stream;
    const int tiles_per_run = div_up(s.ms_slice, MERGE_TILE);
    merge_runs_kernel<<<tiles_per_run * s.ms_world, 256, 0, st>>>(s.ms_keys, s.ms_vals, s.ms_slice, s.ms_world, s.n, tiles_per_run,
                                                                  s.keys[0], s.vals[0]);
    ++s.launches;
    B200_CHECK(cudaGetLastError());
    s.sorted_slot = 0;
    build_after_sort(s, false);
}

__global__ void init_build_counters_kernel(unsigned* alloc, unsigned* children)
{
    *alloc = 1u;
    *children = 0u;
}

// for_step: called by step(): the velocities stay where they are (the fused traversal + integration reads them
// through the permutation) and the state buffers are only advanced once that kernel has been launched.
static void build_after_sort(NBodySim& s, bool for_step)
{
    const int n = s.n;
    const int grid = div_up(n, 256);
    cudaStream_t st = s.stream;
    // ---- physical reorder
    const int gp = (s.pcur + 1) % 3, o = s.vcur ^ 1;
    {
        const GatherArgs ga{s.vals[s.sorted_slot], s.pos[s.pcur], for_step ? nullptr : s.vel[s.vcur], s.mass[s.vcur], s.id[s.vcur],
                            s.pos[gp], s.vel[o], s.mass[o], s.id[o], s.posm};
        const KarrasArgs ka{s.keys[s.sorted_slot], s.childL, s.childR, s.parent, s.range, s.lvl};
        gather_karras_kernel<<<2 * grid, 256, 0, st>>>(ga, ka, n);   // phase "gather" = reorder + tree topology
    }
    ++s.launches;
    B200_CHECK(cudaGetLastError());
    const double* pos_sorted = s.pos[gp];
    const double* mass_sorted = s.mass[o];
    if (for_step) {
        s.step_pending = true;     // pos[gp], mass[o], id[o] hold the new order; vel[vcur] still the previous one
    } else {
        s.pcur = gp;               // the whole state is in the new order
        s.vcur = o;
    }
    s.timer.mark(st);
    // ---- blocked prefix sums of (m x, m y, m z, m) = node mass / centre of mass; pass 1 of the cell
    // extraction (children lists, pair-block allocation)
    if (n > 1) {
        init_build_counters_kernel<<<1, 1, 0, st>>>(s.d_alloc, s.d_children);   // pair 0 is the root's
        const int nb = div_up(n, PFX_BLOCK);
        // sharded step: only the cells a body of this rank's shard may open get their children written
        const bool let = for_step && s.let_enabled && s.world > 1 && s.shard_end - s.shard_begin < n && s.shard_end > s.shard_begin;
        if (let) {
            let_boxes_init_kernel<<<div_up(6 * LET_BOXES, 256), 256, 0, st>>>(s.d_boxes);
            let_boxes_kernel<<<div_up(s.shard_end - s.shard_begin, LET_CHUNK), 256, 0, st>>>(s.posm, s.keys[s.sorted_slot], s.shard_begin, s.shard_end, s.d_boxes);
            let_boxes_up_kernel<<<1, 512, 0, st>>>(s.d_boxes);
            s.launches += 3;
        }
        const ChildrenArgs ca{s.childL, s.childR, s.parent, s.range, s.lvl, s.meta, s.ishead, s.kids,
                              s.d_alloc, (unsigned)(s.rec_capacity_override > 0 ? min(s.rec_capacity, s.rec_capacity_override) : s.rec_capacity),
                              s.d_error, s.d_children,
                              let ? s.d_boxes : nullptr, s.keys[s.sorted_slot], s.d_bounds, s.d_ttab, (float)(s.softening * s.softening),
                              s.shard_begin, s.shard_end};
        prefix_kernel<<<nb, 256, 0, st>>>(pos_sorted, mass_sorted, n, s.ploc, s.bex);
        prefix_blocks_kernel<<<1, 1024, 0, st>>>(s.bex, nb);
        if (let) count_children_let_kernel<<<min(div_up(n - 1, 256), 8 * s.sm_count), 256, 0, st>>>(n, div_up(n - 1, 256), ca);
        else count_children_kernel<<<div_up(n - 1, 256), 256, 0, st>>>(n, ca);
        s.launches += 3;
        B200_CHECK(cudaGetLastError());
    }
    s.timer.mark(st);
    // ---- octree records
    const float eps2f = (float)(s.softening * s.softening);
    if (n > 1) {
        const int g1 = div_up(n - 1, 256);
        const TreeView tv{s.meta, s.ploc, s.bex, s.posm, s.d_ttab};
        write_records_kernel<<<g1, 256, 0, st>>>(n, tv, s.ishead, reinterpret_cast<const int*>(s.kids), eps2f, s.recs);
        ++s.launches;
    } else {
        single_body_record_kernel<<<1, 1, 0, st>>>(s.posm, eps2f, s.recs);
        ++s.launches;
    }
    B200_CHECK(cudaGetLastError());
    s.timer.mark(st);
    s.tree_valid = true;
}

// Launches the walk over sorted bodies [begin, end).  so == nullptr: forces only (s.acc); otherwise the kernel's
// epilogue also integrates the bodies it traversed and writes the next state (see StepOut in traverse.cuh).
template <bool COUNT, bool INTEG>
static void traverse_launch_t(NBodySim& s, int mode, int begin, int end, const StepOut& so)
{
    cudaStream_t st = s.stream;
    const float eps2 = (float)(s.softening * s.softening);
    const int stop = min(end, s.n);
    if (mode == 64) {
        const int tiles64 = div_up(stop - begin, 64);
        const int blocks = min(div_up(tiles64, TRAV_WARPS), s.sm_count * B200_TRAV64_CTAS);
        traverse64c_kernel<COUNT, INTEG><<<blocks, TRAV_BLOCK, TRAV64C_SMEM_BYTES, st>>>(s.recs, s.posm, s.acc, begin, stop, eps2, (float)s.G,
                                                                                        s.d_tile_counter, s.d_interactions, s.d_error, so);
    } else {
        const int tile_begin = begin / 32, tile_end = div_up(end, 32);
        const int blocks = min(div_up(tile_end - tile_begin, TRAV_WARPS), s.sm_count * 4);
        traverse_kernel<COUNT, INTEG><<<blocks, TRAV_BLOCK, TRAV_SMEM_BYTES, st>>>(s.recs, s.posm, s.acc, tile_begin, tile_end, stop, eps2, (float)s.G,
                                                                                  s.d_tile_counter, s.d_interactions, s.d_error, so);
    }
}

static void traverse_launch(NBodySim& s, int begin, int end, const StepOut* so)
{
    B200_CHECK(cudaSetDevice(s.device));
    cudaStream_t st = s.stream;
    if (end > begin) {
        B200_REQUIRE(begin % 32 == 0, "traversal range must start on a 32-body tile");
        B200_CHECK(cudaMemsetAsync(s.d_tile_counter, 0, sizeof(unsigned), st));
        // Two bodies per lane (64-body tiles, the classed walk) win from about a million bodies up: measured on
        // B200 4.79 vs 5.23 ms at 8 M / theta 0.7, 26.0 vs 30.6 at 50 M, a tie at 1 M / theta 0.5; below that the
        // 32-body tiles give the machine twice as many warps to schedule.
        const int mode = s.trav_mode ? s.trav_mode : (s.n >= 1000000 ? 64 : 32);
        B200_REQUIRE(mode == 32 || begin % 64 == 0, "the 64-body walk needs a range that starts on a 64-body tile");
        s.last_trav_kernel = mode;
        const StepOut none{};
        if (so) {
            if (s.count_interactions) traverse_launch_t<true, true>(s, mode, begin, end, *so);
            else traverse_launch_t<false, true>(s, mode, begin, end, *so);
        } else {
            if (s.count_interactions) traverse_launch_t<true, false>(s, mode, begin, end, none);
            else traverse_launch_t<false, false>(s, mode, begin, end, none);
        }
        ++s.launches;
        B200_CHECK(cudaGetLastError());
    }
    s.timer.mark(st);
}

void nbody_traverse(NBodySim& s, int begin, int end) { traverse_launch(s, begin, end, nullptr); }

void nbody_integrate(NBodySim& s, double dt)
{
    B200_CHECK(cudaSetDevice(s.device));
    cudaStream_t st = s.stream;
    if (s.n > 0) {
        const int next = s.maxabs_slot ^ 1;
        B200_CHECK(cudaMemsetAsync(s.d_maxabs + next, 0, sizeof(unsigned long long), st));
        integrate_kernel<<<div_up(s.n, 256), 256, 0, st>>>(s.pos[s.pcur], s.vel[s.vcur], s.acc, s.n, dt, s.damping,
                                                           s.d_maxabs + next);
        ++s.launches;
        B200_CHECK(cudaGetLastError());
        s.maxabs_slot = next;
    }
    s.timer.mark(st);
    s.tree_valid = false;
    ++s.steps;
}

void nbody_step_begin(NBodySim& s)
{
    nbody_build_tree(s);
    nbody_traverse(s, s.shard_begin, s.shard_end);
}

void nbody_step_begin_sorted(NBodySim& s)
{
    nbody_build_tree_presorted(s);
    nbody_traverse(s, s.shard_begin, s.shard_end);
}

void nbody_step_end(NBodySim& s, double dt)
{
    s.timer.mark(s.stream);   // exchange phase (the caller's collective); empty on one GPU
    nbody_integrate(s, dt);
    if (s.timer.enabled) {
        B200_CHECK(cudaStreamSynchronize(s.stream));
        s.timer.collect();
    }
}

// One whole step with the integration fused into the traversal (single GPU): keys -> sort -> gather of
// positions / masses / ids -> tree -> records -> traversal whose epilogue integrates each tile's bodies and
// writes the next state.  Phases "exchange" and "integrate" are empty (kept so the phase table keeps its shape).
static void step_fused(NBodySim& s, double dt)
{
    if (s.n == 0) { ++s.steps; return; }
    build_tree_impl(s, true);
    cudaStream_t st = s.stream;
    const int gp = (s.pcur + 1) % 3, np = (s.pcur + 2) % 3, o = s.vcur ^ 1;
    const int next = s.maxabs_slot ^ 1;
    B200_CHECK(cudaMemsetAsync(s.d_maxabs + next, 0, sizeof(unsigned long long), st));
    StepOut so{};
    so.perm = s.vals[s.sorted_slot];
    so.vel_prev = s.vel[s.vcur];
    so.pos_new = s.pos[gp];
    so.pos_out[0] = s.pos[np];
    so.vel_out[0] = s.vel[o];
    so.world = 1;
    so.dt = dt;
    so.damping = s.damping;
    so.maxabs = s.d_maxabs + next;
    traverse_launch(s, 0, s.n, &so);
    s.timer.mark(st);   // exchange: nothing
    s.timer.mark(st);   // integrate: fused into the traversal
    s.pcur = np;
    s.vcur = o;
    s.maxabs_slot = next;
    s.step_pending = false;
    s.tree_valid = false;
    ++s.steps;
    if (s.timer.enabled) {
        B200_CHECK(cudaStreamSynchronize(st));
        s.timer.collect();
    }
}

// ---------------------------------------------------------------------------- sharded fused step (pieces)
void nbody_shard_build(NBodySim& s, bool presorted)
{
    B200_CHECK(cudaSetDevice(s.device));
    if (s.n == 0) return;
    if (presorted) {
        B200_REQUIRE(s.ms_keys, "sharded sort is not set up");
        cudaStream_t st = s.stream;
        const int tiles_per_run = div_up(s.ms_slice, MERGE_TILE);
        merge_runs_kernel<<<tiles_per_run * s.ms_world, 256, 0, st>>>(s.ms_keys, s.ms_vals, s.ms_slice, s.ms_world, s.n, tiles_per_run,
                                                                      s.keys[0], s.vals[0]);
        ++s.launches;
        B200_CHECK(cudaGetLastError());
        s.sorted_slot = 0;
        build_after_sort(s, true);
    } else {
        build_tree_impl(s, true);
    }
}

unsigned long long* nbody_shard_maxabs_next(NBodySim& s) { return s.d_maxabs + (s.maxabs_slot ^ 1); }

void nbody_shard_traverse(NBodySim& s, double dt, const ShardPeers& peers)
{
    B200_CHECK(cudaSetDevice(s.device));
    if (s.n == 0) return;
    cudaStream_t st = s.stream;
    const int gp = (s.pcur + 1) % 3, np = (s.pcur + 2) % 3, o = s.vcur ^ 1;
    B200_CHECK(cudaMemsetAsync(nbody_shard_maxabs_next(s), 0, sizeof(unsigned long long), st));
    StepOut so{};
    so.perm = s.vals[s.sorted_slot];
    so.vel_prev = s.vel[s.vcur];
    so.pos_new = s.pos[gp];
    B200_REQUIRE(peers.world >= 1 && peers.world <= TRAV_MAX_PEERS, "too many ranks for the fused broadcast");
    for (int r = 0; r < peers.world; ++r) { so.pos_out[r] = peers.pos[r][np]; so.vel_out[r] = peers.vel[r][o]; }
    so.world = peers.world;
    so.dt = dt;
    so.damping = s.damping;
    so.maxabs = nbody_shard_maxabs_next(s);
    traverse_launch(s, s.shard_begin, s.shard_end, &so);
}

void nbody_shard_finish(NBodySim& s)
{
    if (s.n > 0) {
        s.timer.mark(s.stream);   // exchange: the caller's all-reduce (the barrier that ends the step)
        s.timer.mark(s.stream);   // integrate: fused into the traversal
        s.pcur = (s.pcur + 2) % 3;
        s.vcur ^= 1;
        s.maxabs_slot ^= 1;
    }
    s.step_pending = false;
    s.tree_valid = false;
    ++s.steps;
    if (s.timer.enabled) {
        B200_CHECK(cudaStreamSynchronize(s.stream));
        s.timer.collect();
    }
}

static void step_plain(NBodySim& s, double dt)
{
    if (s.shard_begin == 0 && s.shard_end == s.n) {
        step_fused(s, dt);
    } else {   // a shard is set (split multi-GPU step driven from outside): forces of the shard, then integrate all
        nbody_step_begin(s);
        nbody_step_end(s, dt);
    }
}

// ---------------------------------------------------------------------------- captured step
// A step is ~25 launches and memsets; below a few hundred thousand bodies it is launch-bound (the
// reference's own CPU-runnable presets: 10 K - 100 K bodies).  On the handle's own stream the whole step
// is captured into a CUDA graph and replayed.  The kernel arguments depend on which of the two state
// buffers is current and on the parameters, so graphs are cached by that key (the buffers alternate:
// two graphs in steady state); any change (dt, theta, a new state, a shard) captures a new one.
struct StepGraphKey {
    int pcur, vcur, maxabs_slot, shard_begin, shard_end, trav_mode;
    double dt, G, softening, damping, theta;
    cudaStream_t stream;
};

static StepGraphKey step_graph_key(const NBodySim& s, double dt)
{
    StepGraphKey k;
    memset(&k, 0, sizeof(k));
    k.pcur = s.pcur; k.vcur = s.vcur; k.maxabs_slot = s.maxabs_slot; k.shard_begin = s.shard_begin; k.shard_end = s.shard_end;
    k.trav_mode = s.trav_mode;
    k.dt = dt; k.G = s.G; k.softening = s.softening; k.damping = s.damping; k.theta = s.theta;
    k.stream = s.stream;
    return k;
}

void nbody_graphs_reset(NBodySim& s)
{
    for (int i = 0; i < NBodySim::MAX_STEP_GRAPHS; ++i) {
        if (s.step_graph[i]) cudaGraphExecDestroy(s.step_graph[i]);
        s.step_graph[i] = nullptr;
    }
}

void nbody_step(NBodySim& s, double dt)
{
    // (the legacy default stream cannot be captured: a caller that hands in torch's default stream gets plain launches)
    const bool capturable = s.stream != nullptr && s.stream != cudaStreamLegacy && s.stream != cudaStreamPerThread;
    const bool plain = s.timer.enabled || s.count_interactions || !capturable || !s.use_graph || s.n < 2;
    if (plain) {
        step_plain(s, dt);
        return;
    }
    B200_CHECK(cudaSetDevice(s.device));
    static_assert(sizeof(StepGraphKey) <= sizeof(s.step_graph_key[0]), "graph key storage too small");
    const StepGraphKey key = step_graph_key(s, dt);
    int slot = -1;
    for (int i = 0; i < NBodySim::MAX_STEP_GRAPHS; ++i)
        if (s.step_graph[i] && memcmp(s.step_graph_key[i], &key, sizeof(key)) == 0) slot = i;
    const int pcur0 = s.pcur, vcur0 = s.vcur, mslot0 = s.maxabs_slot;
    if (slot < 0) {
        slot = (int)(s.step_graph_next++ % NBodySim::MAX_STEP_GRAPHS);
        if (s.step_graph[slot]) { cudaGraphExecDestroy(s.step_graph[slot]); s.step_graph[slot] = nullptr; }
        const int64_t launches0 = s.launches, steps0 = s.steps;
        cudaGraph_t g = nullptr;
        B200_CHECK(cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
        try {
            step_plain(s, dt);
        } catch (...) {
            cudaStreamEndCapture(s.stream, &g);
            if (g) cudaGraphDestroy(g);
            s.pcur = pcur0; s.vcur = vcur0; s.maxabs_slot = mslot0; s.launches = launches0; s.steps = steps0;
            s.step_pending = false;
            throw;
        }
        B200_CHECK(cudaStreamEndCapture(s.stream, &g));
        const cudaError_t e = cudaGraphInstantiate(&s.step_graph[slot], g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { s.step_graph[slot] = nullptr; s.pcur = pcur0; s.vcur = vcur0; s.maxabs_slot = mslot0; }
        B200_CHECK(e);
        memcpy(s.step_graph_key[slot], &key, sizeof(key));
        s.step_graph_launches[slot] = s.launches - launches0;
        s.step_graph_sorted_slot[slot] = s.sorted_slot;
        s.step_graph_pcur[slot] = s.pcur;
        s.step_graph_vcur[slot] = s.vcur;
        // the capture ran the host side of the step (buffer flips, counters); undo the counters, keep the flips
        s.launches = launches0;
        s.steps = steps0;
    } else {
        // replay: the host-side effects of the step
        s.pcur = s.step_graph_pcur[slot];
        s.vcur = s.step_graph_vcur[slot];
        s.maxabs_slot = mslot0 ^ 1;
        s.sorted_slot = s.step_graph_sorted_slot[slot];
    }
    B200_CHECK(cudaGraphLaunch(s.step_graph[slot], s.stream));
    s.launches += s.step_graph_launches[slot];
    ++s.steps;
    s.tree_valid = false;
}

// ---------------------------------------------------------------------------- FP32 peak probe
// Dependent-free FFMA chains: 8 accumulators x 4096 iterations per thread, enough CTAs to
// fill the machine.  Gives the measured denominator of the traversal's roofline.
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, float a, float b, int iters)
{
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    const float r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (r == 12345.678f) out[0] = r;
}

double fp32_peak_tflops(int device)
{
    B200_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    B200_CHECK(cudaGetDeviceProperties(&prop, device));
    float* d = dev_alloc<float>(1);
    cudaEvent_t e0, e1;
    B200_CHECK(cudaEventCreate(&e0));
    B200_CHECK(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        B200_CHECK(cudaEventRecord(e0, 0));
        ffma_peak_kernel<<<blocks, 256>>>(d, 1.0000001f, 1e-7f, iters);
        B200_CHECK(cudaEventRecord(e1, 0));
        B200_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        B200_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8.0 * iters * 256.0 * blocks;
        if (rep > 0 && ms > 0) best = fmax(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return best;
}

// ---------------------------------------------------------------------------- device error flags
// The traversal (stack overflow) and the record allocator (pool overflow) raise sticky flags in
// d_error instead of dropping work silently (the reference silently drops, nbody/simulation.py:176,272).
// Every reference-facing synchronisation point -- sync(), the getters, frame_wait() -- reads them
// (one 4-byte copy on a stream it already waits for) and fails with B200_ERR_STATE.
static void throw_if_flagged(unsigned flags)
{
    if (!flags) return;
    char b[256];
    snprintf(b, sizeof(b), "device error flags %#x (1 = traversal stack overflow, 2 = octree record pool overflow, 4 = a body "
                           "asked for a cell the locally essential tree had pruned): forces of the affected step(s) are incomplete", flags);
    throw StateError{std::string(b)};
}

void nbody_check_errors(NBodySim& s)
{
    unsigned flags = 0;
    B200_CHECK(cudaMemcpy(&flags, s.d_error, sizeof(flags), cudaMemcpyDeviceToHost));
    throw_if_flagged(flags);
}

// end of a getter: the flags travel with the data on the same stream
static void sync_and_check(NBodySim& s)
{
    B200_CHECK(cudaMemcpyAsync(s.h_error, s.d_error, sizeof(unsigned), cudaMemcpyDeviceToHost, s.stream));
    B200_CHECK(cudaStreamSynchronize(s.stream));
    throw_if_flagged(*s.h_error);
}

void nbody_compute_colors(NBodySim& s, double max_speed)
{
    B200_CHECK(cudaSetDevice(s.device));
    if (s.n > 0) {
        colors_kernel<<<div_up(s.n, 256), 256, 0, s.stream>>>(s.vel[s.vcur], s.id[s.vcur], s.colors, s.n, max_speed);
        B200_CHECK(cudaGetLastError());
    }
}

void nbody_get_positions(NBodySim& s, float* out)
{
    B200_CHECK(cudaSetDevice(s.device));
    if (s.n == 0) return;
    unpermute3_kernel<float><<<div_up(s.n, 256), 256, 0, s.stream>>>(s.pos[s.pcur], s.id[s.vcur], (float*)s.stage, s.n);
    B200_CHECK(cudaGetLastError());
    B200_CHECK(cudaMemcpyAsync(out, s.stage, 3 * (size_t)s.n * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    sync_and_check(s);
}

void nbody_get_positions_f64(NBodySim& s, double* out)
{
    B200_CHECK(cudaSetDevice(s.device));
    if (s.n == 0) return;
    unpermute3_kernel<double><<<div_up(s.n, 256), 256, 0, s.stream>>>(s.pos[s.pcur], s.id[s.vcur], (double*)s.stage, s.n);
    B200_CHECK(cudaGetLastError());
    B200_CHECK(cudaMemcpyAsync(out, s.stage, 3 * (size_t)s.n * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    sync_and_check(s);
}

void nbody_get_velocities(NBodySim& s, double* out)
{
    B200_CHECK(cudaSetDevice(s.device));
    if (s.n == 0) return;
    unpermute3_kernel<double><<<div_up(s.n, 256), 256, 0, s.stream>>>(s.vel[s.vcur], s.id[s.vcur], (double*)s.stage, s.n);
    B200_CHECK(cudaGetLastError());
    B200_CHECK(cudaMemcpyAsync(out, s.stage, 3 * (size_t)s.n * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    sync_and_check(s);
}

void nbody_get_colors(NBodySim& s, float* out)
{
    B200_CHECK(cudaSetDevice(s.device));
    if (s.n == 0) return;
    B200_CHECK(cudaMemcpyAsync(out, s.colors, 3 * (size_t)s.n * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    sync_and_check(s);
}

void nbody_get_accelerations(NBodySim& s, float* out)
{
    B200_CHECK(cudaSetDevice(s.device));
    if (s.n == 0) return;
    const bool timing = s.timer.enabled, counting = s.count_interactions;
    s.timer.enabled = false;
    s.count_interactions = true;
    if (!s.tree_valid) nbody_build_tree(s);
    nbody_traverse(s, 0, s.n);
    s.timer.enabled = timing;
    s.count_interactions = counting;
    unpermute_acc_kernel<<<div_up(s.n, 256), 256, 0, s.stream>>>(s.acc, s.id[s.vcur], (float*)s.stage, s.n);
    B200_CHECK(cudaGetLastError());
    B200_CHECK(cudaMemcpyAsync(out, s.stage, 3 * (size_t)s.n * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    sync_and_check(s);
}

int64_t nbody_count_interactions(NBodySim& s)
{
    B200_CHECK(cudaSetDevice(s.device));
    if (s.n == 0) return 0;
    const bool timing = s.timer.enabled, counting = s.count_interactions;
    s.timer.enabled = false;
    s.count_interactions = true;
    if (!s.tree_valid) nbody_build_tree(s);
    unsigned long long before = 0, after = 0;
    B200_CHECK(cudaMemcpyAsync(&before, s.d_interactions, sizeof(before), cudaMemcpyDeviceToHost, s.stream));
    nbody_traverse(s, s.shard_begin, s.shard_end);
    s.timer.enabled = timing;
    s.count_interactions = counting;
    B200_CHECK(cudaMemcpyAsync(&after, s.d_interactions, sizeof(after), cudaMemcpyDeviceToHost, s.stream));
    sync_and_check(s);
    return (int64_t)(after - before);
}

// sum over bodies of mix(creation index) * (bit patterns of the three components): commutative, so equal for
// any order of the bodies, and sensitive to a single flipped bit
__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
    return x;
}

__global__ void __launch_bounds__(256) checksum_kernel(const double* __restrict__ pos, const double* __restrict__ vel,
                                                       const uint32_t* __restrict__ id, int n, unsigned long long* __restrict__ out)
{
    unsigned long long a = 0, b = 0;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long h = mix64(0x9e3779b97f4a7c15ull + id[k]) | 1ull;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            a += mix64(h + d) * (unsigned long long)__double_as_longlong(pos[3 * k + d]);
            b += mix64(h + 7 + d) * (unsigned long long)__double_as_longlong(vel[3 * k + d]);
        }
    }
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (lane_id() == 0) { atomicAdd(&out[0], a); atomicAdd(&out[1], b); }
}

void nbody_state_checksum(NBodySim& s, uint64_t out[2])
{
    B200_CHECK(cudaSetDevice(s.device));
    out[0] = out[1] = 0;
    if (s.n == 0) return;
    unsigned long long* d = reinterpret_cast<unsigned long long*>(s.stage);   // getter staging: free between calls
    B200_CHECK(cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), s.stream));
    checksum_kernel<<<min(div_up(s.n, 256), s.sm_count * 8), 256, 0, s.stream>>>(s.pos[s.pcur], s.vel[s.vcur], s.id[s.vcur], s.n, d);
    B200_CHECK(cudaGetLastError());
    unsigned long long h[2] = {0, 0};
    B200_CHECK(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, s.stream));
    sync_and_check(s);
    out[0] = h[0]; out[1] = h[1];
}

void nbody_get_keys(NBodySim& s, uint64_t* out)
{
    B200_CHECK(cudaSetDevice(s.device));
    if (s.n == 0) return;
    if (!s.tree_valid) nbody_build_tree(s);
    B200_CHECK(cudaMemcpyAsync(out, s.keys[s.sorted_slot], (size_t)s.n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s.stream));
    sync_and_check(s);
}

void nbody_get_perm(NBodySim& s, uint32_t* out)
{
    B200_CHECK(cudaSetDevice(s.device));
    if (s.n == 0) return;
    if (!s.tree_valid) nbody_build_tree(s);
    B200_CHECK(cudaMemcpyAsync(out, s.id[s.vcur], (size_t)s.n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
    sync_and_check(s);
}

// ---------------------------------------------------------------------------- asynchronous host traffic
// The reference's frame loop reads positions and colours back every frame (tools/record.py:826-832)
// and its getters block (gpu_backend.py:394-404).  These entry points keep PCIe busy in both
// directions while the next step computes: the device-side staging is produced on the compute
// stream, the copies run on their own streams, events order the hand-offs.
static void async_init(NBodySim& s)
{
    if (s.up_stream) return;
    const size_t N = (size_t)s.n;
    B200_CHECK(cudaStreamCreateWithFlags(&s.up_stream, cudaStreamNonBlocking));
    B200_CHECK(cudaStreamCreateWithFlags(&s.down_stream, cudaStreamNonBlocking));
    cudaEvent_t* evs[4] = {&s.ev_frame_ready, &s.ev_frame_done, &s.ev_upload_done, &s.ev_upload_consumed};
    for (cudaEvent_t* e : evs) B200_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    s.frame_pos = alloc_counted<float>(s, 3 * N);
    s.frame_col = alloc_counted<float>(s, 3 * N);
    s.unperm_counts = alloc_counted<unsigned>(s, 2 * 64);
    s.up_pos = alloc_counted<double>(s, 3 * (N + 64));   // + padding rows: equal all-gather slices up to 64 ranks
    s.up_vel = alloc_counted<double>(s, 3 * (N + 64));
}

static void async_free(NBodySim& s)
{
    if (!s.up_stream) return;
    cudaStreamSynchronize(s.up_stream);
    cudaStreamSynchronize(s.down_stream);
    cudaFree(s.frame_pos); cudaFree(s.frame_col); cudaFree(s.up_pos); cudaFree(s.up_vel); cudaFree(s.unperm_counts);
    if (s.frame_dpos) { cudaFree(s.frame_dpos); cudaFree(s.frame_dcol); cudaFree(s.frame_pos2); cudaFree(s.frame_col2); s.frame_dpos = s.frame_dcol = nullptr; }
    cudaEventDestroy(s.ev_frame_ready); cudaEventDestroy(s.ev_frame_done);
    cudaEventDestroy(s.ev_upload_done); cudaEventDestroy(s.ev_upload_consumed);
    cudaStreamDestroy(s.up_stream); cudaStreamDestroy(s.down_stream);
    s.up_stream = s.down_stream = nullptr;
}

__global__ void __launch_bounds__(256) frame_kernel(const double* __restrict__ pos, const double* __restrict__ vel,
                                                    const uint32_t* __restrict__ id, float* __restrict__ fpos,
                                                    float* __restrict__ fcol, int n, double max_speed)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t o = 3 * (int64_t)k, w = 3 * (int64_t)id[k];
    const double vx = vel[o], vy = vel[o + 1], vz = vel[o + 2];
    float r, g, b;
    speed_color(fmin(1.0, sqrt(vx * vx + vy * vy + vz * vz) / max_speed), r, g, b);
    fpos[w] = (float)pos[o]; fpos[w + 1] = (float)pos[o + 1]; fpos[w + 2] = (float)pos[o + 2];
    fcol[w] = r; fcol[w + 1] = g; fcol[w + 2] = b;
}

// rows [row_begin, row_end) of the frame only (sharded egress: a rank copies out 1 / world of the rows, so it
// un-permutes only those: every thread reads the 4-byte creation index, one in `world` does the rest)
__global__ void __launch_bounds__(256) frame_rows_kernel(const double* __restrict__ pos, const double* __restrict__ vel,
                                                         const uint32_t* __restrict__ id, float* __restrict__ fpos,
                                                         float* __restrict__ fcol, int n, double max_speed, uint32_t row_begin,
                                                         uint32_t row_end)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t c = id[k];
    if (c < row_begin || c >= row_end) return;
    const int64_t o = 3 * (int64_t)k, w = 3 * (int64_t)c;
    const double vx = vel[o], vy = vel[o + 1], vz = vel[o + 2];
    float r, g, b;
    speed_color(fmin(1.0, sqrt(vx * vx + vy * vy + vz * vz) / max_speed), r, g, b);
    fpos[w] = (float)pos[o]; fpos[w + 1] = (float)pos[o + 1]; fpos[w + 2] = (float)pos[o + 2];
    fcol[w] = r; fcol[w + 1] = g; fcol[w + 2] = b;
}

// ---------------------------------------------------------------------------- bucketed un-permute (large n)
// frame_kernel's un-permute to creation order is a random 12-byte scatter: every store is a partial
// sector, read-modify-written in DRAM (6.5 ms at 50 M).  For large n the frame is produced in three
// coalesced passes instead: the bodies are partitioned by the top bits of their creation index into
// <= 64 buckets of float4 records {x, y, z, id} {r, g, b, -} (scratch: the pair-record pool, free
// between steps), then scattered bucket by bucket -- a bucket's output window (<= 24 MB) stays in the
// 126 MB L2, so the partial stores merge into whole sectors before they reach DRAM.
constexpr int UNP_ITEMS = 16;                 // bodies per thread (4096 per CTA)
constexpr int UNP_BUCKETS = 64;

__global__ void __launch_bounds__(256) unperm_hist_kernel(const uint32_t* __restrict__ id, int n, int shift, unsigned* __restrict__ counts)
{
    __shared__ unsigned h[UNP_BUCKETS];
    if (threadIdx.x < UNP_BUCKETS) h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * (256 * UNP_ITEMS) + threadIdx.x;
#pragma unroll
    for (int i = 0; i < UNP_ITEMS; ++i) {
        const int64_t k = base + (int64_t)i * 256;
        if (k < n) atomicAdd(&h[id[k] >> shift], 1u);
    }
    __syncthreads();
    if (threadIdx.x < UNP_BUCKETS && h[threadIdx.x]) atomicAdd(&counts[threadIdx.x], h[threadIdx.x]);
}

__global__ void __launch_bounds__(256) unperm_partition_kernel(const double* __restrict__ pos, const double* __restrict__ vel,
                                                               const uint32_t* __restrict__ id, int n, int shift, double max_speed,
                                                               const unsigned* __restrict__ counts, unsigned* __restrict__ cursors,
                                                               float4* __restrict__ rec_a, float4* __restrict__ rec_b)
{
    __shared__ unsigned h[UNP_BUCKETS];
    __shared__ unsigned gbase[UNP_BUCKETS];
    if (threadIdx.x < UNP_BUCKETS) h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * (256 * UNP_ITEMS) + threadIdx.x;
    uint32_t ids[UNP_ITEMS];
#pragma unroll
    for (int i = 0; i < UNP_ITEMS; ++i) {
        const int64_t k = base + (int64_t)i * 256;
        ids[i] = k < n ? id[k] : 0xffffffffu;
        if (k < n) atomicAdd(&h[ids[i] >> shift], 1u);
    }
    __syncthreads();
    if (threadIdx.x < UNP_BUCKETS) {
        unsigned start = 0;
        for (int b = 0; b < (int)threadIdx.x; ++b) start += counts[b];   // bucket starts: exclusive scan of the global counts
        const unsigned mine = h[threadIdx.x];
        gbase[threadIdx.x] = start + (mine ? atomicAdd(&cursors[threadIdx.x], mine) : 0u);
        h[threadIdx.x] = 0;   // now the CTA-local cursor
    }
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < UNP_ITEMS; ++i) {
        const int64_t k = base + (int64_t)i * 256;
        if (k >= n) continue;
        const unsigned b = ids[i] >> shift;
        const unsigned dst = gbase[b] + atomicAdd(&h[b], 1u);
        const int64_t o = 3 * k;
        const double vx = vel[o], vy = vel[o + 1], vz = vel[o + 2];
        float r, g, bl;
        speed_color(fmin(1.0, sqrt(vx * vx + vy * vy + vz * vz) / max_speed), r, g, bl);
        rec_a[dst] = make_float4((float)pos[o], (float)pos[o + 1], (float)pos[o + 2], __uint_as_float(ids[i]));
        rec_b[dst] = make_float4(r, g, bl, 0.f);
    }
}

__global__ void __launch_bounds__(256) unperm_scatter_kernel(const float4* __restrict__ rec_a, const float4* __restrict__ rec_b, int n,
                                                             float* __restrict__ fpos, float* __restrict__ fcol)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = rec_a[i], b = rec_b[i];
    const int64_t w = 3 * (int64_t)__float_as_uint(a.w);
    fpos[w] = a.x; fpos[w + 1] = a.y; fpos[w + 2] = a.z;
    fcol[w] = b.x; fcol[w + 1] = b.y; fcol[w + 2] = b.z;
}

// colours + creation-order float32 positions of the current state into (fpos, fcol), on the handle's stream
static void launch_frame(NBodySim& s, float* fpos, float* fcol, double max_speed, int row_begin = 0, int row_end = -1)
{
    const int n = s.n;
    cudaStream_t st = s.stream;
    if (row_end >= 0 && (row_begin > 0 || row_end < n) && 2 * (int64_t)(row_end - row_begin) <= n) {
        frame_rows_kernel<<<div_up(n, 256), 256, 0, st>>>(s.pos[s.pcur], s.vel[s.vcur], s.id[s.vcur], fpos, fcol, n, max_speed,
                                                          (uint32_t)row_begin, (uint32_t)row_end);
        ++s.launches;
    } else if (n < s.unperm_min_n || !s.unperm_counts) {
        frame_kernel<<<div_up(n, 256), 256, 0, st>>>(s.pos[s.pcur], s.vel[s.vcur], s.id[s.vcur], fpos, fcol, n, max_speed);
        ++s.launches;
    } else {
        int shift = s.unperm_shift;   // 2^20 creation indices per bucket (24 MB of output), more when n > 64 M
        while (((int64_t)(n - 1) >> shift) >= UNP_BUCKETS) ++shift;
        float4* rec_a = s.recs;                    // the pair-record pool (6 n float4) is free between steps
        float4* rec_b = s.recs + (size_t)n;
        s.tree_valid = false;
        B200_CHECK(cudaMemsetAsync(s.unperm_counts, 0, 2 * UNP_BUCKETS * sizeof(unsigned), st));
        const int blocks = div_up(n, 256 * UNP_ITEMS);
        unperm_hist_kernel<<<blocks, 256, 0, st>>>(s.id[s.vcur], n, shift, s.unperm_counts);
        unperm_partition_kernel<<<blocks, 256, 0, st>>>(s.pos[s.pcur], s.vel[s.vcur], s.id[s.vcur], n, shift, max_speed, s.unperm_counts,
                                                        s.unperm_counts + UNP_BUCKETS, rec_a, rec_b);
        unperm_scatter_kernel<<<div_up(n, 256), 256, 0, st>>>(rec_a, rec_b, n, fpos, fcol);
        s.launches += 3;
    }
    B200_CHECK(cudaGetLastError());
}

void nbody_frame_begin(NBodySim& s, double max_speed, float* host_pos, float* host_col)
{
    nbody_frame_begin_rows(s, max_speed, host_pos, host_col, 0, s.n);
}

// rows [row_begin, row_end) of the frame only (creation order): sharded frame egress, one slice per rank
void nbody_frame_begin_rows(NBodySim& s, double max_speed, float* host_pos, float* host_col, int row_begin, int row_end)
{
    B200_CHECK(cudaSetDevice(s.device));
    B200_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= s.n, "frame rows out of range");
    if (s.n == 0) return;
    async_init(s);
    if (s.frame_pending) B200_CHECK(cudaStreamWaitEvent(s.stream, s.ev_frame_done, 0));   // staging still being read
    launch_frame(s, s.frame_pos, s.frame_col, max_speed, row_begin, row_end);
    const bool partial_frame = row_begin > 0 || row_end < s.n;   // (a partial frame cannot seed a delta frame)
    B200_CHECK(cudaEventRecord(s.ev_frame_ready, s.stream));
    B200_CHECK(cudaStreamWaitEvent(s.down_stream, s.ev_frame_ready, 0));
    const size_t off = 3 * (size_t)row_begin, bytes = 3 * (size_t)(row_end - row_begin) * sizeof(float);
    if (bytes) {
        B200_CHECK(cudaMemcpyAsync(host_pos + off, s.frame_pos + off, bytes, cudaMemcpyDeviceToHost, s.down_stream));
        B200_CHECK(cudaMemcpyAsync(host_col + off, s.frame_col + off, bytes, cudaMemcpyDeviceToHost, s.down_stream));
    }
    B200_CHECK(cudaMemcpyAsync(s.h_error, s.d_error, sizeof(unsigned), cudaMemcpyDeviceToHost, s.down_stream));
    B200_CHECK(cudaEventRecord(s.ev_frame_done, s.down_stream));
    s.frame_pending = true;
    s.frame_has_prev = !partial_frame;
}

// ---------------------------------------------------------------------------- delta frames (frame codec)
// The recorder's on-disk format 2 stores a frame as int16((frame - previous frame) * 1000) per component
// (tools/record.py:254-262), computed on float32 frames.  Producing the deltas on the device halves the
// device-to-host bytes of a frame (12 instead of 24 B/body); the float32 staging of the last frame is
// the "previous frame" and is updated in place.
__device__ __forceinline__ short delta_i16(float cur, float prev)
{
    // numpy: ((cur - prev) * 1000).astype(int16) on float32 arrays = truncation through a 32-bit integer
    // (cvttss2si: "integer indefinite" 0x80000000 outside the int32 range), then the low 16 bits
    const float x = __fmul_rn(__fsub_rn(cur, prev), 1000.0f);
    const int v = fabsf(x) < 2147483648.0f ? __float2int_rz(x) : (int)0x80000000;
    return (short)(unsigned short)((unsigned)v & 0xffffu);
}

// Elementwise over the two creation-order float32 frames (fully coalesced; the random un-permute is done
// once, by frame_kernel, into the other staging buffer).  count = 3 n components, two per thread.
__global__ void __launch_bounds__(256) frame_delta_kernel(const float2* __restrict__ cur_pos, const float2* __restrict__ prev_pos,
                                                          const float2* __restrict__ cur_col, const float2* __restrict__ prev_col,
                                                          short2* __restrict__ dpos, short2* __restrict__ dcol, int64_t pairs,
                                                          int64_t count)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < pairs) {
        const float2 a = cur_pos[i], b = prev_pos[i], c = cur_col[i], d = prev_col[i];
        dpos[i] = make_short2(delta_i16(a.x, b.x), delta_i16(a.y, b.y));
        dcol[i] = make_short2(delta_i16(c.x, d.x), delta_i16(c.y, d.y));
    } else if (i == pairs && (count & 1)) {   // odd tail component
        const float* cp = reinterpret_cast<const float*>(cur_pos); const float* pp = reinterpret_cast<const float*>(prev_pos);
        const float* cc = reinterpret_cast<const float*>(cur_col); const float* pc = reinterpret_cast<const float*>(prev_col);
        reinterpret_cast<short*>(dpos)[count - 1] = delta_i16(cp[count - 1], pp[count - 1]);
        reinterpret_cast<short*>(dcol)[count - 1] = delta_i16(cc[count - 1], pc[count - 1]);
    }
}

void nbody_frame_delta_begin(NBodySim& s, double max_speed, short* host_dpos, short* host_dcol)
{
    B200_CHECK(cudaSetDevice(s.device));
    if (s.n == 0) return;
    B200_REQUIRE(s.frame_has_prev, "frame_delta_begin needs a previous frame (call frame_begin first)");
    const size_t N = (size_t)s.n;
    if (!s.frame_dpos) {   // second float32 staging (the frames alternate between the two) + the int16 payloads
        s.frame_pos2 = alloc_counted<float>(s, 3 * N);
        s.frame_col2 = alloc_counted<float>(s, 3 * N);
        s.frame_dpos = alloc_counted<short>(s, 3 * N);
        s.frame_dcol = alloc_counted<short>(s, 3 * N);
    }
    if (s.frame_pending) B200_CHECK(cudaStreamWaitEvent(s.stream, s.ev_frame_done, 0));   // staging still being read
    float* prev_pos = s.frame_pos; float* prev_col = s.frame_col;
    float* cur_pos = s.frame_pos2; float* cur_col = s.frame_col2;
    launch_frame(s, cur_pos, cur_col, max_speed);
    const int64_t count = 3 * (int64_t)N, pairs = count / 2;
    frame_delta_kernel<<<(unsigned)((pairs + 1 + 255) / 256), 256, 0, s.stream>>>(
        reinterpret_cast<const float2*>(cur_pos), reinterpret_cast<const float2*>(prev_pos),
        reinterpret_cast<const float2*>(cur_col), reinterpret_cast<const float2*>(prev_col),
        reinterpret_cast<short2*>(s.frame_dpos), reinterpret_cast<short2*>(s.frame_dcol), pairs, count);
    s.launches += 1;
    B200_CHECK(cudaGetLastError());
    // the new frame becomes the "previous frame" (and the staging frame_begin writes to)
    s.frame_pos = cur_pos; s.frame_col = cur_col;
    s.frame_pos2 = prev_pos; s.frame_col2 = prev_col;
    B200_CHECK(cudaEventRecord(s.ev_frame_ready, s.stream));
    B200_CHECK(cudaStreamWaitEvent(s.down_stream, s.ev_frame_ready, 0));
    B200_CHECK(cudaMemcpyAsync(host_dpos, s.frame_dpos, 3 * N * sizeof(short), cudaMemcpyDeviceToHost, s.down_stream));
    B200_CHECK(cudaMemcpyAsync(host_dcol, s.frame_dcol, 3 * N * sizeof(short), cudaMemcpyDeviceToHost, s.down_stream));
    B200_CHECK(cudaMemcpyAsync(s.h_error, s.d_error, sizeof(unsigned), cudaMemcpyDeviceToHost, s.down_stream));
    B200_CHECK(cudaEventRecord(s.ev_frame_done, s.down_stream));
    s.frame_pending = true;
}

// ---------------------------------------------------------------------------- live-viewer path
// The reference's viewer copies every position and colour to the host each frame (nbody/simulation.py:809-817),
// tests all of them against the view frustum on the CPU (compute_visibility_points, :403-434), gathers the visible
// ones with a boolean mask (:926-927) and uploads those to two VBOs.  Here the test and the gather run on the
// device over the creation-order fp32 frame (the reference tests float32 positions widened to float64: :816), and
// only the visible bodies leave the GPU -- or none do, when the caller passes a mapped VBO.
constexpr int VIS_ITEMS = 4;
constexpr int VIS_TILE = 256 * VIS_ITEMS;

__device__ __forceinline__ bool frustum_visible(const float* __restrict__ fpos, int64_t i, const Camera& c)
{
    // the arithmetic of compute_visibility_points, operation by operation, without contraction
    const double dx = __dsub_rn((double)fpos[3 * i], c.pos[0]), dy = __dsub_rn((double)fpos[3 * i + 1], c.pos[1]),
                 dz = __dsub_rn((double)fpos[3 * i + 2], c.pos[2]);
    const double z = __dadd_rn(__dadd_rn(__dmul_rn(dx, c.forward[0]), __dmul_rn(dy, c.forward[1])), __dmul_rn(dz, c.forward[2]));
    if (z < 0.1 || z > c.far_dist) return false;
    const double x = __dadd_rn(__dadd_rn(__dmul_rn(dx, c.right[0]), __dmul_rn(dy, c.right[1])), __dmul_rn(dz, c.right[2]));
    const double y = __dadd_rn(__dadd_rn(__dmul_rn(dx, c.up[0]), __dmul_rn(dy, c.up[1])), __dmul_rn(dz, c.up[2]));
    const double hw = __dmul_rn(__dmul_rn(z, c.tan_h), 1.2), hh = __dmul_rn(__dmul_rn(z, c.tan_v), 1.2);
    return fabs(x) < hw && fabs(y) < hh;
}

__global__ void __launch_bounds__(256) vis_count_kernel(const float* __restrict__ fpos, int n, Camera cam, unsigned* __restrict__ counts)
{
    const int64_t base = (int64_t)blockIdx.x * VIS_TILE + (int64_t)threadIdx.x * VIS_ITEMS;
    int c = 0;
#pragma unroll
    for (int q = 0; q < VIS_ITEMS; ++q)
        if (base + q < n && frustum_visible(fpos, base + q, cam)) ++c;
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ int sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        counts[blockIdx.x] = (unsigned)t;
    }
}

// exclusive scan of the block counts in place (one CTA; the total lands in counts[blocks])
__global__ void __launch_bounds__(1024) vis_scan_kernel(unsigned* __restrict__ counts, int blocks)
{
    __shared__ unsigned sh[32];
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    for (int b0 = 0; b0 < blocks; b0 += 1024) {
        const int i = b0 + (int)threadIdx.x;
        const unsigned v = i < blocks ? counts[i] : 0u;
        unsigned inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
            if ((int)(threadIdx.x & 31) >= o) inc += u;
        }
        if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned w = sh[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned u = __shfl_up_sync(0xffffffffu, w, o);
                if ((int)threadIdx.x >= o) w += u;
            }
            sh[threadIdx.x] = w;
        }
        __syncthreads();
        const unsigned before = carry + (threadIdx.x >= 32 ? sh[(threadIdx.x >> 5) - 1] : 0u) + inc - v;
        if (i < blocks) counts[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[blocks] = carry;
}

__global__ void __launch_bounds__(256) vis_compact_kernel(const float* __restrict__ fpos, const float* __restrict__ fcol, int n, Camera cam,
                                                          const unsigned* __restrict__ offsets, float* __restrict__ out_pos,
                                                          float* __restrict__ out_col)
{
    const int64_t base = (int64_t)blockIdx.x * VIS_TILE + (int64_t)threadIdx.x * VIS_ITEMS;
    bool vis[VIS_ITEMS];
    int c = 0;
#pragma unroll
    for (int q = 0; q < VIS_ITEMS; ++q) {
        vis[q] = base + q < n && frustum_visible(fpos, base + q, cam);
        c += vis[q];
    }
    int inc = c;
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)(threadIdx.x & 31) >= o) inc += u;
    }
    __shared__ int sh[8];
    if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = inc;
    __syncthreads();
    int before = inc - c;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += sh[w];
    int64_t o = (int64_t)offsets[blockIdx.x] + before;
#pragma unroll
    for (int q = 0; q < VIS_ITEMS; ++q)
        if (vis[q]) {
            const int64_t i = base + q;
            out_pos[3 * o] = fpos[3 * i]; out_pos[3 * o + 1] = fpos[3 * i + 1]; out_pos[3 * o + 2] = fpos[3 * i + 2];
            out_col[3 * o] = fcol[3 * i]; out_col[3 * o + 1] = fcol[3 * i + 1]; out_col[3 * o + 2] = fcol[3 * i + 2];
            ++o;
        }
}

int64_t nbody_visible_frame(NBodySim& s, double max_speed, const Camera& cam, float* out_pos, float* out_col, bool to_device)
{
    B200_CHECK(cudaSetDevice(s.device));
    if (s.n == 0) return 0;
    async_init(s);
    const int blocks = div_up(s.n, VIS_TILE);
    if (!s.vis_counts) {
        s.vis_counts = alloc_counted<unsigned>(s, (size_t)blocks + 1);
        B200_CHECK(cudaMallocHost(&s.vis_total_host, sizeof(unsigned)));
    }
    if (!to_device && !s.vis_pos) {
        s.vis_pos = alloc_counted<float>(s, 3 * (size_t)s.n);
        s.vis_col = alloc_counted<float>(s, 3 * (size_t)s.n);
    }
    if (s.frame_pending) B200_CHECK(cudaStreamWaitEvent(s.stream, s.ev_frame_done, 0));   // staging still being read
    launch_frame(s, s.frame_pos, s.frame_col, max_speed);
    s.frame_has_prev = true;
    float* dpos = to_device ? out_pos : s.vis_pos;
    float* dcol = to_device ? out_col : s.vis_col;
    vis_count_kernel<<<blocks, 256, 0, s.stream>>>(s.frame_pos, s.n, cam, s.vis_counts);
    vis_scan_kernel<<<1, 1024, 0, s.stream>>>(s.vis_counts, blocks);
    vis_compact_kernel<<<blocks, 256, 0, s.stream>>>(s.frame_pos, s.frame_col, s.n, cam, s.vis_counts, dpos, dcol);
    s.launches += 3;
    B200_CHECK(cudaGetLastError());
    B200_CHECK(cudaMemcpyAsync(s.vis_total_host, s.vis_counts + blocks, sizeof(unsigned), cudaMemcpyDeviceToHost, s.stream));
    sync_and_check(s);
    const int64_t count = (int64_t)*s.vis_total_host;
    if (!to_device && count > 0) {
        B200_CHECK(cudaMemcpyAsync(out_pos, s.vis_pos, 3 * (size_t)count * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
        B200_CHECK(cudaMemcpyAsync(out_col, s.vis_col, 3 * (size_t)count * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
        B200_CHECK(cudaStreamSynchronize(s.stream));
    }
    return count;
}

void nbody_frame_wait(NBodySim& s)
{
    B200_CHECK(cudaSetDevice(s.device));
    if (!s.frame_pending) return;
    B200_CHECK(cudaEventSynchronize(s.ev_frame_done));
    s.frame_pending = false;
    throw_if_flagged(*s.h_error);   // the frame was produced by steps that overflowed: do not hand it out silently
}

void nbody_set_state_begin(NBodySim& s, const double* pos, const double* vel)
{
    nbody_set_state_begin_rows(s, pos, vel, 0, s.n);
}

void nbody_upload_staging(NBodySim& s, double** pos, double** vel)
{
    B200_CHECK(cudaSetDevice(s.device));
    async_init(s);
    *pos = s.up_pos;
    *vel = s.up_vel;
}

// the compute stream waits (on the device) for the pending upload: whatever the caller enqueues next on
// that stream -- e.g. an all-gather of the staging slices -- sees the uploaded rows
void nbody_upload_wait(NBodySim& s)
{
    B200_CHECK(cudaSetDevice(s.device));
    B200_REQUIRE(s.upload_pending || s.n == 0, "upload_wait without set_state_begin");
    if (s.n == 0) return;
    B200_CHECK(cudaStreamWaitEvent(s.stream, s.ev_upload_done, 0));
}

// rows [row_begin, row_end) only: sharded upload, the caller completes the staging buffers on the
// compute stream (upload_wait + all-gather) before set_state_commit
void nbody_set_state_begin_rows(NBodySim& s, const double* pos, const double* vel, int row_begin, int row_end)
{
    B200_CHECK(cudaSetDevice(s.device));
    B200_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= s.n, "upload rows out of range");
    B200_REQUIRE(!s.upload_pending, "set_state_begin: the previous upload was not committed");
    if (s.n == 0) return;
    async_init(s);
    // the staging may still be read by the previous commit's copies on the compute stream
    if (s.ev_upload_consumed) B200_CHECK(cudaStreamWaitEvent(s.up_stream, s.ev_upload_consumed, 0));
    const size_t off = 3 * (size_t)row_begin, bytes = 3 * (size_t)(row_end - row_begin) * sizeof(double);
    if (bytes) {
        B200_CHECK(cudaMemcpyAsync(s.up_pos + off, pos + off, bytes, cudaMemcpyHostToDevice, s.up_stream));
        B200_CHECK(cudaMemcpyAsync(s.up_vel + off, vel + off, bytes, cudaMemcpyHostToDevice, s.up_stream));
    }
    B200_CHECK(cudaEventRecord(s.ev_upload_done, s.up_stream));
    s.upload_pending = true;
}

void nbody_set_state_commit(NBodySim& s)
{
    B200_CHECK(cudaSetDevice(s.device));
    B200_REQUIRE(s.upload_pending || s.n == 0, "set_state_commit without set_state_begin");
    if (s.n == 0) return;
    const size_t bytes = 3 * (size_t)s.n * sizeof(double);
    const int po = (s.pcur + 1) % 3, o = s.vcur ^ 1;
    // host-side wait: once commit returns the caller may reuse its host arrays (the copy was started a
    // step earlier, so this normally returns at once)
    B200_CHECK(cudaEventSynchronize(s.ev_upload_done));
    B200_CHECK(cudaMemcpyAsync(s.pos[po], s.up_pos, bytes, cudaMemcpyDeviceToDevice, s.stream));
    B200_CHECK(cudaMemcpyAsync(s.vel[o], s.up_vel, bytes, cudaMemcpyDeviceToDevice, s.stream));
    B200_CHECK(cudaEventRecord(s.ev_upload_consumed, s.stream));
    B200_CHECK(cudaMemcpyAsync(s.mass[o], s.mass0, (size_t)s.n * sizeof(double), cudaMemcpyDeviceToDevice, s.stream));
    iota_kernel<<<div_up(s.n, 256), 256, 0, s.stream>>>(s.id[o], s.n);
    s.launches += 1;
    B200_CHECK(cudaGetLastError());
    s.pcur = po;
    s.vcur = o;
    recompute_maxabs(s);
    s.tree_valid = false;
    s.upload_pending = false;
}

}  // namespace b200
