// traverse.cuh -- theta-MAC force traversal kernels of the Barnes-Hut step (included by nbody.cu, inside
// namespace b200, after the pair-record definitions it reads).  See DESIGN.md section 4.
#pragma once

// ============================================================================ traversal
// One warp owns 32 consecutive sorted bodies (one per lane) and walks the octree for all of them
// at once.  Work is organised in BATCHES so that the tree-walk bookkeeping is done by 32 lanes in
// parallel and the inner loop is nothing but arithmetic:
//   select  pop as many entries (cell's child block, pair count, lane mask) from the warp's stack
//           as fit 32 pair slots (one lane per entry + a warp prefix sum)
//   load    the pair records of all selected cells: coalesced 16 B/lane loads, staged in shared
//           memory as SoA {x0,x1,y0,y1} {z0,z1,m0,m1} {T0,T1,mask} {first0,first1,n0,n1}
//   eval    every lane evaluates every staged pair, TWO children per iteration with Blackwell's
//           packed fp32x2 instructions (FADD2/FMUL2/FFMA2, sm_100+) from shared-memory broadcasts:
//             d2 = |com - p|^2 + eps^2;   open iff d2 <= T;   otherwise a += m (com - p) d2^-3/2
//           lane 0 records the two ballots of "open" per pair
//   expand  lane j turns pair j's ballots into new stack entries (cells with children that some
//           lane must open, with the mask of exactly those lanes) -- a ballot/popc prefix sum
// T = max(size^2/theta^2, eps^2) (leaf: eps^2).  This is the reference's "accept iff
// size/d < theta, add iff d^2 > eps^2" (nbody/simulation.py:252-267): for size^2/theta^2 >= eps^2
// the tests coincide; otherwise the cell is always accepted and only d2 == eps^2 (zero distance:
// the body itself) is excluded -- it "opens" nothing because leaves have no children.
// Every lane makes the reference's own per-body MAC decision; there is no group MAC.  Lanes
// outside a pair's mask take x = 1e18: d2 ~ 1e36 > T, so they never open, and their
// "contribution" m * d2^-3/2 underflows to exactly 0; no per-child mask logic is needed.
// The stack is depth-first in batches: while it holds more than TRAV_DFS_MARK entries only the
// top entry is popped per batch, which bounds it by MARK + 64 + 7 * 21 < TRAV_CAP.
struct __align__(16) WarpShared {
    unsigned stk_first[TRAV_CAP];        // first pair of the child block | (pairs - 1) << 29
    unsigned stk_mask[TRAV_CAP];         // lanes that must open the cell
    // XY[32] ZM[32] TM[32] FN[32] OP[32]; areas are TRAV_AREA = 34 entries apart so that the four
    // 16-byte quarters of a pair record, stored by four neighbouring lanes, fall in different banks
    float4 stage[5 * TRAV_AREA];
    unsigned d_first[TRAV_BATCH];        // pair slot -> global pair index
    unsigned d_mask[TRAV_BATCH];         //           -> lane mask
};
static_assert(sizeof(WarpShared) % 16 == 0, "WarpShared must keep float4 alignment");
constexpr size_t TRAV_SMEM_BYTES = sizeof(WarpShared) * TRAV_WARPS;
constexpr unsigned TRAV_FIRST_MASK = (1u << 29) - 1u;
constexpr int TRAV_CHUNK_PAIRS = 8;      // an entry holds <= 8 pairs (3 bits); larger buckets are split

__device__ __forceinline__ float rsqrt_approx(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// The MAC of both children of a pair for one lane: a lane inside the pair's mask accepts a child iff
// d2 > T (one FSETP with the mask bit as its predicate operand); r = accept ? d2^-1/2 : 0.  The
// ballots are of "not accepted" and still contain the lanes outside the mask: expand() removes them.
__device__ __forceinline__ void mac_pair(float2 d2, float T0, float T1, bool in, unsigned& om0, unsigned& om1, float2& r)
{
    const bool a0 = in && d2.x > T0, a1 = in && d2.y > T1;
    om0 = __ballot_sync(0xffffffffu, !a0);
    om1 = __ballot_sync(0xffffffffu, !a1);
    r.x = a0 ? rsqrt_approx(d2.x) : 0.f;
    r.y = a1 ? rsqrt_approx(d2.y) : 0.f;
}

// ---------------------------------------------------------------------------- fused integrate + broadcast
// The warp that has just finished the forces of a tile also finishes the step for its bodies: it fetches the
// body's velocity through the sort permutation (the velocities are never reordered by a pass of their own),
// applies v = (v + a dt) damping, x += v dt (nbody/simulation.py:281-305) in fp64 and writes the new state at the
// body's NEW sorted position into the next-state buffers of EVERY rank -- its own and, over NVLink peer
// mappings, the other GPUs' -- so with several GPUs the position exchange rides inside the compute-bound
// traversal instead of following it as a collective.  The next step's max |coord| is reduced on the way.
constexpr int TRAV_MAX_PEERS = 8;
struct StepOut {
    const uint32_t* perm;          // sorted position -> position in the previous order
    const double* vel_prev;        // previous order
    const double* pos_new;         // sorted order (gathered by the tree build)
    double* pos_out[TRAV_MAX_PEERS];   // next-state buffers, sorted order; [0 .. world)
    double* vel_out[TRAV_MAX_PEERS];
    int world;
    double dt, damping;
    unsigned long long* maxabs;    // bit pattern of max |coord| over the bodies this launch integrated
};

// one body: v = (v + a dt) damping, x += v dt in fp64; m accumulates max |coord| of the new position
__device__ __forceinline__ void integrate_body(const StepOut& o, int k, float ax, float ay, float az, double x[3], double v[3], double& m)
{
    const int64_t j = 3 * (int64_t)o.perm[k], w = 3 * (int64_t)k;
    const double a3[3] = {(double)ax, (double)ay, (double)az};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        v[d] = o.vel_prev[j + d];
        v[d] += a3[d] * o.dt;
        v[d] *= o.damping;
        x[d] = o.pos_new[w + d] + v[d] * o.dt;
        m = fmax(m, fabs(x[d]));
    }
}

__device__ __forceinline__ void finish_body(const StepOut& o, int k, float ax, float ay, float az, double& m)
{
    double v[3], x[3];
    integrate_body(o, k, ax, ay, az, x, v, m);
    const int64_t w = 3 * (int64_t)k;
    for (int r = 0; r < o.world; ++r) {
        double* __restrict__ po = o.pos_out[r];
        double* __restrict__ vo = o.vel_out[r];
#pragma unroll
        for (int d = 0; d < 3; ++d) { po[w + d] = x[d]; vo[w + d] = v[d]; }
    }
}

// A full 64-body tile: the new state of the tile (2 x 1536 contiguous bytes) is staged in the warp's shared memory
// (the walk's stack, idle by now) and written with 16-byte stores, 512 contiguous bytes per warp instruction, to every
// rank.  Over NVLink this matters: per-body 8-byte stores at a 24-byte stride reach a fraction of the link rate (at 8
// GPUs a rank sends 7 x 48 B per body; the scattered form held the 8-GPU traversal at 5.2 ms against 3.3 of compute).
__device__ __forceinline__ void finish_tile64(const StepOut& o, int64_t base, unsigned lane, double* stage,
                                              const double xa[3], const double va[3], const double xb[3], const double vb[3])
{
    double* sp = stage;            // [192] positions of the tile
    double* sv = stage + 192;      // [192] velocities
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        sp[3 * lane + d] = xa[d]; sp[3 * (lane + 32) + d] = xb[d];
        sv[3 * lane + d] = va[d]; sv[3 * (lane + 32) + d] = vb[d];
    }
    __syncwarp();
    const double2* sp2 = reinterpret_cast<const double2*>(sp);
    const double2* sv2 = reinterpret_cast<const double2*>(sv);
    double2 p2[3], v2[3];
#pragma unroll
    for (int it = 0; it < 3; ++it) { p2[it] = sp2[lane + 32 * it]; v2[it] = sv2[lane + 32 * it]; }
    for (int r = 0; r < o.world; ++r) {
        double2* __restrict__ po = reinterpret_cast<double2*>(o.pos_out[r] + 3 * base);
        double2* __restrict__ vo = reinterpret_cast<double2*>(o.vel_out[r] + 3 * base);
#pragma unroll
        for (int it = 0; it < 3; ++it) { po[lane + 32 * it] = p2[it]; vo[lane + 32 * it] = v2[it]; }
    }
    __syncwarp();                  // the next tile reuses the stack
}

__device__ __forceinline__ void finish_warp(const StepOut& o, double m)
{
    if (o.world > 1) __threadfence_system();   // the stores into the peers' buffers are performed before the kernel ends
    for (int s = 16; s > 0; s >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, s));
    if (lane_id() == 0) atomicMax(o.maxabs, (unsigned long long)__double_as_longlong(m));
}

template <bool COUNT, bool INTEG>
__global__ void __launch_bounds__(TRAV_BLOCK, 4) traverse_kernel(const float4* __restrict__ recs, const float4* __restrict__ posm,
                                                                 float4* __restrict__ acc, int tile_begin, int tile_end, int n,
                                                                 float eps2, float G, unsigned* tile_counter,
                                                                 unsigned long long* counters, unsigned* error, const StepOut so)
{
    extern __shared__ __align__(16) unsigned char trav_smem[];
    const unsigned lane = lane_id();
    const unsigned lanebit = 1u << lane;
    const unsigned lt = lanemask_lt();
    double w_maxabs = 0.0;
    WarpShared& ws = reinterpret_cast<WarpShared*>(trav_smem)[threadIdx.x >> 5];
    const float4* sXY = ws.stage;
    const float4* sZM = ws.stage + TRAV_AREA;
    const float4* sTM = ws.stage + 2 * TRAV_AREA;
    const float4* sFN = ws.stage + 3 * TRAV_AREA;
    uint4* sOP = reinterpret_cast<uint4*>(ws.stage + 4 * TRAV_AREA);   // .x/.y = ballots of "open", child 0 / 1
    const float2 eps22 = make_float2(eps2, eps2);
    unsigned long long w_inter = 0, w_slots = 0, w_lanepairs = 0, w_batches = 0;
    int w_spmax = 0;

    for (;;) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(tile_counter, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        const int tile = tile_begin + (int)t;
        if (tile >= tile_end) break;
        const int k = tile * 32 + (int)lane;
        const bool valid = k < n;
        const float4 p = valid ? posm[k] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float2 npx = make_float2(-p.x, -p.x), npy = make_float2(-p.y, -p.y), npz = make_float2(-p.z, -p.z);
        float2 ax = make_float2(0.f, 0.f), ay = ax, az = ax;   // (even, odd) children accumulate separately
        int cnt = 0, lanepairs = 0, slots = 0;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) { ws.stk_first[0] = 0u; ws.stk_mask[0] = vmask; }   // pair 0 = {root, dummy}
        int sp = 1;
        __syncwarp();
        while (sp > 0) {
            // ---- select: lane l looks at the l-th entry from the top
            const int idx = sp - 1 - (int)lane;
            unsigned ef = 0, em = 0;
            int np = 0;
            if (idx >= 0) { ef = ws.stk_first[idx]; em = ws.stk_mask[idx]; np = (int)(ef >> 29) + 1; }
            int incl = np;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += u;
            }
            const int limit = sp > TRAV_DFS_MARK ? 1 : 32;
            const bool take = idx >= 0 && incl <= TRAV_BATCH && (int)lane < limit;
            const int E = __popc(__ballot_sync(0xffffffffu, take));   // take is a prefix of the lanes; E >= 1
            // (the max-reduce leaves P in a uniform register: the loops below are provably convergent)
            const int P = __reduce_max_sync(0xffffffffu, (int)lane == E - 1 ? incl : 0);
            sp -= E;
            if ((int)lane < E) {
                const int base = incl - np;
                const unsigned f = ef & TRAV_FIRST_MASK;
                for (int q = 0; q < np; ++q) { ws.d_first[base + q] = f + q; ws.d_mask[base + q] = em; }
            }
            __syncwarp();
            // ---- load: 8 pair records (4 x 16 B each) per warp-wide load
#pragma unroll
            for (int it = 0; it < TRAV_BATCH / 8; ++it) {
                const int slot = it * 8 + (int)(lane >> 2);
                if (slot < P) {
                    float4 v = __ldg(&recs[4 * (int64_t)ws.d_first[slot] + (lane & 3u)]);
                    if ((lane & 3u) == 2u) v.z = __uint_as_float(ws.d_mask[slot]);
                    ws.stage[(lane & 3u) * TRAV_AREA + slot] = v;
                }
            }
            __syncwarp();
            // ---- eval
#pragma unroll 4
            for (int j = 0; j < P; ++j) {
                const float4 XY = sXY[j];
                const float4 ZM = sZM[j];
                const float4 TM = sTM[j];
                const bool inbit = (__float_as_uint(TM.z) & lanebit) != 0u;
                const float2 dx = __fadd2_rn(make_float2(XY.x, XY.y), npx);
                const float2 dy = __fadd2_rn(make_float2(XY.z, XY.w), npy);
                const float2 dz = __fadd2_rn(make_float2(ZM.x, ZM.y), npz);
                const float2 d2 = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, __ffma2_rn(dz, dz, eps22)));
                unsigned om0, om1;
                float2 r;
                mac_pair(d2, TM.x, TM.y, inbit, om0, om1, r);
                if (lane == 0) *reinterpret_cast<uint2*>(&sOP[j]) = make_uint2(om0, om1);
                if (COUNT) {
                    if (inbit) {
                        ++lanepairs;
                        if (r.x != 0.f && XY.x < 2e18f) ++cnt;
                        if (r.y != 0.f && XY.y < 2e18f) ++cnt;
                    }
                }
                const float2 f = __fmul2_rn(make_float2(ZM.z, ZM.w), __fmul2_rn(__fmul2_rn(r, r), r));
                ax = __ffma2_rn(dx, f, ax);
                ay = __ffma2_rn(dy, f, ay);
                az = __ffma2_rn(dz, f, az);
            }
            slots += P;
            __syncwarp();
            // ---- expand: lane j owns pair slot j
            bool c0 = false, c1 = false;
            unsigned f0 = 0, f1 = 0, n0 = 0, n1 = 0;
            uint2 om = make_uint2(0u, 0u);
            if ((int)lane < P) {
                om = *reinterpret_cast<const uint2*>(&sOP[lane]);
                const unsigned mask = __float_as_uint(sTM[lane].z);   // the ballots include the lanes outside the mask
                om.x &= mask; om.y &= mask;
                const float4 fn = sFN[lane];
                f0 = __float_as_uint(fn.x); f1 = __float_as_uint(fn.y);
                n0 = __float_as_uint(fn.z); n1 = __float_as_uint(fn.w);
                c0 = om.x != 0u && n0 != 0u;
                c1 = om.y != 0u && n1 != 0u;
                if ((c0 && n0 == PRUNED_NCHILD) || (c1 && n1 == PRUNED_NCHILD)) {   // see traverse64c_kernel
                    atomicOr(error, (unsigned)ERR_PRUNED_CELL_OPENED);
                    c0 = c0 && n0 != PRUNED_NCHILD;
                    c1 = c1 && n1 != PRUNED_NCHILD;
                }
            }
            const int np0 = (int)((n0 + 1u) >> 1), np1 = (int)((n1 + 1u) >> 1);
            const bool big = (c0 && np0 > TRAV_CHUNK_PAIRS) || (c1 && np1 > TRAV_CHUNK_PAIRS);
            if (!__any_sync(0xffffffffu, big)) {
                const unsigned b0 = __ballot_sync(0xffffffffu, c0), b1 = __ballot_sync(0xffffffffu, c1);
                int pos = sp + __popc(b0 & lt) + __popc(b1 & lt);
                if (c0) { ws.stk_first[pos] = f0 | ((unsigned)(np0 - 1) << 29); ws.stk_mask[pos] = om.x; ++pos; }
                if (c1) { ws.stk_first[pos] = f1 | ((unsigned)(np1 - 1) << 29); ws.stk_mask[pos] = om.y; }
                sp += __popc(b0) + __popc(b1);
            } else {
                // rare: a bucket of > 16 bodies sharing one finest-level cell is pushed in chunks
                const int e0 = c0 ? (np0 + TRAV_CHUNK_PAIRS - 1) / TRAV_CHUNK_PAIRS : 0;
                const int e1 = c1 ? (np1 + TRAV_CHUNK_PAIRS - 1) / TRAV_CHUNK_PAIRS : 0;
                int inc2 = e0 + e1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, inc2, o);
                    if ((int)lane >= o) inc2 += u;
                }
                const int total = __shfl_sync(0xffffffffu, inc2, 31);
                if (sp + total > TRAV_CAP) {   // never drop silently
                    if (lane == 0) atomicOr(error, (unsigned)ERR_STACK_OVERFLOW);
                    sp = 0;
                } else {
                    int pos = sp + inc2 - (e0 + e1);
                    for (int q = 0; q < e0; ++q, ++pos) {
                        const int r = min(np0 - TRAV_CHUNK_PAIRS * q, TRAV_CHUNK_PAIRS);
                        ws.stk_first[pos] = (f0 + (unsigned)(TRAV_CHUNK_PAIRS * q)) | ((unsigned)(r - 1) << 29);
                        ws.stk_mask[pos] = om.x;
                    }
                    for (int q = 0; q < e1; ++q, ++pos) {
                        const int r = min(np1 - TRAV_CHUNK_PAIRS * q, TRAV_CHUNK_PAIRS);
                        ws.stk_first[pos] = (f1 + (unsigned)(TRAV_CHUNK_PAIRS * q)) | ((unsigned)(r - 1) << 29);
                        ws.stk_mask[pos] = om.y;
                    }
                    sp += total;
                }
            }
            if (COUNT) { ++w_batches; w_spmax = max(w_spmax, sp); }
            __syncwarp();
        }
        // acc.w: exact interaction count (COUNT) or the tile's evaluated pair slots (a cost proxy)
        if (valid) {
            const float fx = G * (ax.x + ax.y), fy = G * (ay.x + ay.y), fz = G * (az.x + az.y);
            acc[k] = make_float4(fx, fy, fz, __int_as_float(COUNT ? cnt : slots));
            if (INTEG) finish_body(so, k, fx, fy, fz, w_maxabs);
        }
        if (COUNT) {
            unsigned c32 = valid ? (unsigned)cnt : 0u, l32 = (unsigned)lanepairs;
            for (int o = 16; o > 0; o >>= 1) {
                c32 += __shfl_xor_sync(0xffffffffu, c32, o);
                l32 += __shfl_xor_sync(0xffffffffu, l32, o);
            }
            w_inter += c32;
            w_lanepairs += l32;
            w_slots += (unsigned)slots;
        }
    }
    if (INTEG) finish_warp(so, w_maxabs);
    if (COUNT && lane == 0) {
        if (w_inter) atomicAdd(&counters[0], w_inter);
        atomicAdd(&counters[1], w_slots);
        atomicAdd(&counters[2], w_lanepairs);
        atomicAdd(&counters[3], w_batches);
        atomicMax(&counters[4], (unsigned long long)w_spmax);
    }
}

// ---------------------------------------------------------------------------- two bodies per lane
// A warp owns 64 consecutive sorted bodies (lane l: bodies l and l + 32 of the tile), a stack entry carries
// one lane mask per 32-body half.  A staged pair record is read from shared memory once for both halves and
// the batch bookkeeping is shared.  The evaluated (pair, half) set is exactly that of two independent 32-body
// walks; every lane still makes the reference's per-body MAC decision.
struct EvalBody {
    float npx, npy, npz;                 // -position
    float2 ax, ay, az;                   // (even, odd) children accumulate separately
    int cnt, lanepairs;
};

template <bool COUNT>
__device__ __forceinline__ void eval_pair(const float4& XY, const float4& ZM, float T0, float T1, unsigned mask, unsigned lanebit,
                                          float2 eps22, EvalBody& b, unsigned& om0, unsigned& om1)
{
    const bool inbit = (mask & lanebit) != 0u;
    const float2 dx = __fadd2_rn(make_float2(XY.x, XY.y), make_float2(b.npx, b.npx));
    const float2 dy = __fadd2_rn(make_float2(XY.z, XY.w), make_float2(b.npy, b.npy));
    const float2 dz = __fadd2_rn(make_float2(ZM.x, ZM.y), make_float2(b.npz, b.npz));
    const float2 d2 = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, __ffma2_rn(dz, dz, eps22)));
    float2 r;
    mac_pair(d2, T0, T1, inbit, om0, om1, r);
    if (COUNT) {
        if (inbit) {
            ++b.lanepairs;
            if (r.x != 0.f && XY.x < 2e18f) ++b.cnt;
            if (r.y != 0.f && XY.y < 2e18f) ++b.cnt;
        }
    }
    const float2 f = __fmul2_rn(make_float2(ZM.z, ZM.w), __fmul2_rn(__fmul2_rn(r, r), r));
    b.ax = __ffma2_rn(dx, f, b.ax);
    b.ay = __ffma2_rn(dy, f, b.ay);
    b.az = __ffma2_rn(dz, f, b.az);
}

// ---------------------------------------------------------------------------- 64-body walk, classed
// The per-lane MAC costs more than the arithmetic it guards: on B200 every ALU-pipe instruction (FSETP,
// VOTE, LOP3) takes about as long as a packed FFMA2 -- the both-halves loop above runs at 36 clk per (pair,
// half) evaluation, the same loop without the MAC at 27 (scripts/evalbench.cu).  But for most pairs the
// outcome of the MAC is known for the whole tile before any lane looks: if the axis-aligned box of a
// 32-body half is farther from a child's centre of mass than that child's MAC radius, EVERY body of the half
// accepts the child -- each lane would have decided exactly that (the test is conservative by a relative
// 1e-5, four orders above the fp32 rounding of the lanes' own d^2).  So each lane loads ONE pair record of
// the batch, classifies it against the two boxes, and the batch is sorted into five classes with a loop each:
//   FF  both halves need the pair, both surely accept, both masks full  -> arithmetic only
//   SS  both halves surely accept, partial masks                       -> accumulation predicated by the mask bit
//   SA / SB  only one half needs the pair, surely accepts
//   UU  anything else                                                   -> the per-lane MAC loop (as above)
// Only UU pairs produce "open" ballots and are looked at by the expand step.  The evaluated (pair, half) set
// and every lane's accept / open decisions are exactly those of the walk above.
struct __align__(16) WarpShared64C {
    unsigned stk_first[TRAV_CAP];        // (stk_first + stk_lo double as the 3 KB staging of finish_tile64)
    unsigned stk_lo[TRAV_CAP];
    unsigned stk_hi[TRAV_CAP];
    float4 stage[5 * TRAV_AREA];         // class-sorted: XY ZM {T0,T1,mask lo,mask hi} {open lo 0, lo 1, hi 0, hi 1}; selection order: FN
    float4 box[4];                       // {Alo.xyz, -} {Ahi.xyz, -} {Blo.xyz, -} {Bhi.xyz, -}
};
constexpr size_t TRAV64C_SMEM_BYTES = sizeof(WarpShared64C) * TRAV_WARPS + B200_TRAV64_PAD;
static_assert(2 * TRAV_CAP * sizeof(unsigned) >= 2 * 192 * sizeof(double), "the stack must hold a tile's staged state");
constexpr float TRAV_SURE_MARGIN = 1.00001f;

// order-preserving float <-> int map (signed integer compare == float compare), for redux.sync.min/max
__device__ __forceinline__ int f2ord(float f) { const int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// squared distance from the box [lo, hi] to the two children's coordinates along one axis, added to acc
__device__ __forceinline__ float2 box_axis(float2 c, float lo, float hi, float2 acc)
{
    const float q0 = fmaxf(fmaxf(lo - c.x, c.x - hi), 0.f), q1 = fmaxf(fmaxf(lo - c.y, c.y - hi), 0.f);
    return __ffma2_rn(make_float2(q0, q1), make_float2(q0, q1), acc);
}

template <bool COUNT, int MODE>   // MODE 1: mask-predicated, no MAC   2: no mask, no MAC
__device__ __forceinline__ void eval_sure(const float4& XY, const float4& ZM, bool in, float2 eps22, EvalBody& b)
{
    const float2 dx = __fadd2_rn(make_float2(XY.x, XY.y), make_float2(b.npx, b.npx));
    const float2 dy = __fadd2_rn(make_float2(XY.z, XY.w), make_float2(b.npy, b.npy));
    const float2 dz = __fadd2_rn(make_float2(ZM.x, ZM.y), make_float2(b.npz, b.npz));
    const float2 d2 = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, __ffma2_rn(dz, dz, eps22)));
    float2 r;
    r.x = rsqrt_approx(d2.x);
    r.y = rsqrt_approx(d2.y);
    const float2 f = __fmul2_rn(make_float2(ZM.z, ZM.w), __fmul2_rn(__fmul2_rn(r, r), r));
    if (MODE == 2 || in) {
        b.ax = __ffma2_rn(dx, f, b.ax);
        b.ay = __ffma2_rn(dy, f, b.ay);
        b.az = __ffma2_rn(dz, f, b.az);
    }
    if (COUNT) {
        if (in) {
            ++b.lanepairs;
            if (XY.x < 2e18f) ++b.cnt;
            if (XY.y < 2e18f) ++b.cnt;
        }
    }
}

template <bool COUNT, bool INTEG>
__global__ void __launch_bounds__(TRAV_BLOCK, B200_TRAV64_CTAS) traverse64c_kernel(const float4* __restrict__ recs, const float4* __restrict__ posm,
                                                                    float4* __restrict__ acc, int begin, int end,
                                                                    float eps2, float G, unsigned* tile_counter,
                                                                    unsigned long long* counters, unsigned* error, const StepOut so)
{
    double w_maxabs = 0.0;
    extern __shared__ __align__(16) unsigned char trav_smem[];
    const unsigned lane = lane_id();
    const unsigned lanebit = 1u << lane;
    const unsigned lt = lanemask_lt();
    WarpShared64C& ws = reinterpret_cast<WarpShared64C*>(trav_smem)[threadIdx.x >> 5];
    float4* sXY = ws.stage;
    float4* sZM = ws.stage + TRAV_AREA;
    float4* sTM = ws.stage + 2 * TRAV_AREA;
    uint4* sOP = reinterpret_cast<uint4*>(ws.stage + 3 * TRAV_AREA);
    float4* sFN = ws.stage + 4 * TRAV_AREA;
    const float2 eps22 = make_float2(eps2, eps2);
    unsigned long long w_inter = 0, w_slots = 0, w_lanepairs = 0, w_batches = 0, w_both = 0, w_sure = 0;
    int w_spmax = 0;

    for (;;) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(tile_counter, 1u);
        t = __reduce_max_sync(0xffffffffu, t);
        const int64_t base = (int64_t)begin + 64 * (int64_t)t;
        if (base >= end) break;
        const int ka = (int)base + (int)lane, kb = ka + 32;
        const bool va = ka < end, vb = kb < end;
        EvalBody A, B;
        unsigned vma, vmb;
        {
            const float4 pa = va ? posm[ka] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 pb = vb ? posm[kb] : make_float4(0.f, 0.f, 0.f, 0.f);
            A.npx = -pa.x; A.npy = -pa.y; A.npz = -pa.z;
            B.npx = -pb.x; B.npy = -pb.y; B.npz = -pb.z;
            vma = __ballot_sync(0xffffffffu, va);
            vmb = __ballot_sync(0xffffffffu, vb);
            // boxes of the two halves (over their valid bodies)
            const int big = 0x7f7fffff;   // f2ord(FLT_MAX)
            int lo[6], hi[6];
            const float pv[6] = {pa.x, pa.y, pa.z, pb.x, pb.y, pb.z};
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const bool v = q < 3 ? va : vb;
                const int o = f2ord(pv[q]);
                lo[q] = __reduce_min_sync(0xffffffffu, v ? o : big);
                hi[q] = __reduce_max_sync(0xffffffffu, v ? o : -big);
            }
            if (lane == 0) {
                ws.box[0] = make_float4(ord2f(lo[0]), ord2f(lo[1]), ord2f(lo[2]), 0.f);
                ws.box[1] = make_float4(ord2f(hi[0]), ord2f(hi[1]), ord2f(hi[2]), 0.f);
                ws.box[2] = make_float4(ord2f(lo[3]), ord2f(lo[4]), ord2f(lo[5]), 0.f);
                ws.box[3] = make_float4(ord2f(hi[3]), ord2f(hi[4]), ord2f(hi[5]), 0.f);
                ws.stk_first[0] = 0u; ws.stk_lo[0] = vma; ws.stk_hi[0] = vmb;   // pair 0 = {root, dummy}
            }
        }
        A.ax = A.ay = A.az = B.ax = B.ay = B.az = make_float2(0.f, 0.f);
        A.cnt = A.lanepairs = B.cnt = B.lanepairs = 0;
        int slots_a = 0, slots_b = 0;
        int sp = 1;
        __syncwarp();
        while (sp > 0) {
            // ---- select: the top entries, one per lane
            const int idx = sp - 1 - (int)lane;
            unsigned ef = 0, ml = 0, mh = 0;
            if (idx >= 0) { ef = ws.stk_first[idx]; ml = ws.stk_lo[idx]; mh = ws.stk_hi[idx]; }
            const unsigned multi = __ballot_sync(0xffffffffu, (ef >> 29) != 0u);
            const bool chunk = (multi & 1u) != 0u;
            const int room = (TRAV_CAP - TRAV_RESERVE - sp) / 7;
            int E = min(min(sp, TRAV_BATCH), max(room, 1));
            if (multi) E = min(E, __ffs(multi) - 1);
            int P = E;
            if (chunk) {   // a chunk of a bucket is a batch of its own: P consecutive pairs with the entry's masks
                const unsigned ef0 = __shfl_sync(0xffffffffu, ef, 0);
                ml = __shfl_sync(0xffffffffu, ml, 0);
                mh = __shfl_sync(0xffffffffu, mh, 0);
                E = 1;
                P = (int)(ef0 >> 29) + 1;
                ef = (ef0 & TRAV_FIRST_MASK) + lane;
            }
            P = __reduce_max_sync(0xffffffffu, P);
            sp -= E;
            // ---- load + classify: lane j owns pair slot j
            const bool mine = (int)lane < P;
            int cls = 5;   // 0 FF, 1 SS, 2 SA, 3 SB, 4 UU, 5 none
            float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0, q2 = q0, q3 = q0;
            if (mine) {
                const float4* r = recs + 4 * (int64_t)(ef & TRAV_FIRST_MASK);
                q0 = __ldg(r); q1 = __ldg(r + 1); q2 = __ldg(r + 2); q3 = __ldg(r + 3);
                const float T0 = q2.x * TRAV_SURE_MARGIN, T1 = q2.y * TRAV_SURE_MARGIN;
                bool sureA = true, sureB = true;
                if (ml) {
                    const float4 blo = ws.box[0], bhi = ws.box[1];
                    float2 m2 = box_axis(make_float2(q0.x, q0.y), blo.x, bhi.x, eps22);
                    m2 = box_axis(make_float2(q0.z, q0.w), blo.y, bhi.y, m2);
                    m2 = box_axis(make_float2(q1.x, q1.y), blo.z, bhi.z, m2);
                    sureA = m2.x > T0 && m2.y > T1;
                }
                if (mh) {
                    const float4 blo = ws.box[2], bhi = ws.box[3];
                    float2 m2 = box_axis(make_float2(q0.x, q0.y), blo.x, bhi.x, eps22);
                    m2 = box_axis(make_float2(q0.z, q0.w), blo.y, bhi.y, m2);
                    m2 = box_axis(make_float2(q1.x, q1.y), blo.z, bhi.z, m2);
                    sureB = m2.x > T0 && m2.y > T1;
                }
                if (eps2 <= 0.f) sureA = sureB = false;   // a body's own leaf must go through the MAC's d2 > eps2 test
                if (!(sureA && sureB)) cls = 4;
                else if (ml && mh) cls = (ml == vma && mh == vmb) ? 0 : 1;
                else cls = ml ? 2 : 3;
            }
            const unsigned b0 = __ballot_sync(0xffffffffu, cls == 0), b1 = __ballot_sync(0xffffffffu, cls == 1);
            const unsigned b2 = __ballot_sync(0xffffffffu, cls == 2), b3 = __ballot_sync(0xffffffffu, cls == 3);
            const unsigned b4 = __ballot_sync(0xffffffffu, cls == 4);
            const int n0 = __popc(b0), n1 = n0 + __popc(b1), n2 = n1 + __popc(b2), n3 = n2 + __popc(b3);
            const unsigned mycls = cls == 0 ? b0 : cls == 1 ? b1 : cls == 2 ? b2 : cls == 3 ? b3 : b4;
            const int cbase = cls == 0 ? 0 : cls == 1 ? n0 : cls == 2 ? n1 : cls == 3 ? n2 : n3;
            const int pos = cbase + __popc(mycls & lt);
            if (mine) {
                sXY[pos] = q0;
                sZM[pos] = q1;
                sTM[pos] = make_float4(q2.x, q2.y, __uint_as_float(ml), __uint_as_float(mh));
                sFN[lane] = q3;
            }
            __syncwarp();
            // ---- eval, one loop per class (warp-uniform trip counts)
            const int e0 = n0, e1 = n1, e2 = n2, e3 = n3;
#pragma unroll 2
            for (int j = 0; j < e0; ++j) {
                const float4 XY = sXY[j];
                const float4 ZM = sZM[j];
                eval_sure<COUNT, 2>(XY, ZM, va, eps22, A);   // (va / vb only matter to the counters of a ragged last tile)
                eval_sure<COUNT, 2>(XY, ZM, vb, eps22, B);
            }
#pragma unroll 2
            for (int j = e0; j < e1; ++j) {
                const float4 XY = sXY[j];
                const float4 ZM = sZM[j];
                const float4 TM = sTM[j];
                eval_sure<COUNT, 1>(XY, ZM, (__float_as_uint(TM.z) & lanebit) != 0u, eps22, A);
                eval_sure<COUNT, 1>(XY, ZM, (__float_as_uint(TM.w) & lanebit) != 0u, eps22, B);
            }
#pragma unroll 2
            for (int j = e1; j < e2; ++j) {
                const float4 XY = sXY[j];
                const float4 ZM = sZM[j];
                const float4 TM = sTM[j];
                eval_sure<COUNT, 1>(XY, ZM, (__float_as_uint(TM.z) & lanebit) != 0u, eps22, A);
            }
#pragma unroll 2
            for (int j = e2; j < e3; ++j) {
                const float4 XY = sXY[j];
                const float4 ZM = sZM[j];
                const float4 TM = sTM[j];
                eval_sure<COUNT, 1>(XY, ZM, (__float_as_uint(TM.w) & lanebit) != 0u, eps22, B);
            }
#pragma unroll 2
            for (int j = e3; j < P; ++j) {
                const float4 XY = sXY[j];
                const float4 ZM = sZM[j];
                const float4 TM = sTM[j];
                uint4 om;
                eval_pair<COUNT>(XY, ZM, TM.x, TM.y, __float_as_uint(TM.z), lanebit, eps22, A, om.x, om.y);
                eval_pair<COUNT>(XY, ZM, TM.x, TM.y, __float_as_uint(TM.w), lanebit, eps22, B, om.z, om.w);
                if (lane == 0) sOP[j] = om;
            }
            if (COUNT) {
                // UU pairs: count the halves that are really needed (masks non-empty), like the class-sorted walk
                const unsigned ua = __ballot_sync(0xffffffffu, cls == 4 && ml != 0u), ub = __ballot_sync(0xffffffffu, cls == 4 && mh != 0u);
                slots_a += e2 + __popc(ua);
                slots_b += e1 + (e3 - e2) + __popc(ub);
                w_both += (unsigned)(e1 + __popc(ua & ub));
                w_sure += (unsigned)(e1 + e3);   // (pair, half) evaluations that skipped the MAC: FF and SS count twice
            } else {   // cost proxy for the sharding: evaluated (pair, half) slots
                slots_a += e2 + (P - e3);
                slots_b += e1 + (e3 - e2) + (P - e3);
            }
            __syncwarp();
            // ---- expand: only UU pairs can have opened children
            unsigned f0 = 0, f1 = 0, c0n = 0, c1n = 0;
            uint4 om = make_uint4(0u, 0u, 0u, 0u);
            if (cls == 4) {
                om = sOP[pos];
                om.x &= ml; om.y &= ml; om.z &= mh; om.w &= mh;   // the ballots include the lanes outside the masks
                const float4 fn = sFN[lane];
                f0 = __float_as_uint(fn.x); f1 = __float_as_uint(fn.y);
                c0n = __float_as_uint(fn.z); c1n = __float_as_uint(fn.w);
            }
            bool c0 = (om.x | om.z) != 0u && c0n != 0u, c1 = (om.y | om.w) != 0u && c1n != 0u;
            if ((c0 && c0n == PRUNED_NCHILD) || (c1 && c1n == PRUNED_NCHILD)) {   // cannot happen (locally essential tree test is conservative); never drop silently
                atomicOr(error, (unsigned)ERR_PRUNED_CELL_OPENED);
                c0 = c0 && c0n != PRUNED_NCHILD;
                c1 = c1 && c1n != PRUNED_NCHILD;
            }
            const int np0 = c0 ? (int)((c0n + 1u) >> 1) : 0, np1 = c1 ? (int)((c1n + 1u) >> 1) : 0;
            const bool bigc = np0 > TRAV_CELL_PAIRS || np1 > TRAV_CELL_PAIRS;
            if (!__any_sync(0xffffffffu, bigc)) {
                const unsigned k = (unsigned)(np0 + np1);
                const unsigned k0 = __ballot_sync(0xffffffffu, (k & 1u) != 0u), k1 = __ballot_sync(0xffffffffu, (k & 2u) != 0u);
                const unsigned k2 = __ballot_sync(0xffffffffu, (k & 4u) != 0u), k3 = __ballot_sync(0xffffffffu, (k & 8u) != 0u);
                const int total = __popc(k0) + 2 * __popc(k1) + 4 * __popc(k2) + 8 * __popc(k3);
                if (sp + total > TRAV_CAP) {   // cannot happen (see the stack bound above); never drop silently
                    if (lane == 0) atomicOr(error, (unsigned)ERR_STACK_OVERFLOW);
                    sp = 0;
                } else {
                    int p = sp + __popc(k0 & lt) + 2 * __popc(k1 & lt) + 4 * __popc(k2 & lt) + 8 * __popc(k3 & lt);
                    for (int q = 0; q < np0; ++q, ++p) { ws.stk_first[p] = f0 + (unsigned)q; ws.stk_lo[p] = om.x; ws.stk_hi[p] = om.z; }
                    for (int q = 0; q < np1; ++q, ++p) { ws.stk_first[p] = f1 + (unsigned)q; ws.stk_lo[p] = om.y; ws.stk_hi[p] = om.w; }
                    sp += total;
                }
            } else {
                // rare: a bucket of > 8 bodies sharing one finest-level cell is pushed in chunks of <= 8 pairs
                const int x0 = (np0 + TRAV_CHUNK_PAIRS - 1) / TRAV_CHUNK_PAIRS;
                const int x1 = (np1 + TRAV_CHUNK_PAIRS - 1) / TRAV_CHUNK_PAIRS;
                int inc2 = x0 + x1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, inc2, o);
                    if ((int)lane >= o) inc2 += u;
                }
                const int total = __reduce_add_sync(0xffffffffu, x0 + x1);
                if (sp + total > TRAV_CAP) {   // never drop silently
                    if (lane == 0) atomicOr(error, (unsigned)ERR_STACK_OVERFLOW);
                    sp = 0;
                } else {
                    int p = sp + inc2 - (x0 + x1);
                    for (int q = 0; q < x0; ++q, ++p) {
                        const int r = min(np0 - TRAV_CHUNK_PAIRS * q, TRAV_CHUNK_PAIRS);
                        ws.stk_first[p] = (f0 + (unsigned)(TRAV_CHUNK_PAIRS * q)) | ((unsigned)(r - 1) << 29);
                        ws.stk_lo[p] = om.x; ws.stk_hi[p] = om.z;
                    }
                    for (int q = 0; q < x1; ++q, ++p) {
                        const int r = min(np1 - TRAV_CHUNK_PAIRS * q, TRAV_CHUNK_PAIRS);
                        ws.stk_first[p] = (f1 + (unsigned)(TRAV_CHUNK_PAIRS * q)) | ((unsigned)(r - 1) << 29);
                        ws.stk_lo[p] = om.y; ws.stk_hi[p] = om.w;
                    }
                    sp += total;
                }
            }
            sp = __reduce_max_sync(0xffffffffu, sp);   // uniform register: the walk loop is provably convergent
            if (COUNT) { ++w_batches; w_spmax = max(w_spmax, sp); }
            __syncwarp();
        }
        // acc.w: exact interaction count (COUNT) or the half-tile's evaluated pair slots (a cost proxy)
        {
            const float fax = G * (A.ax.x + A.ax.y), fay = G * (A.ay.x + A.ay.y), faz = G * (A.az.x + A.az.y);
            const float fbx = G * (B.ax.x + B.ax.y), fby = G * (B.ay.x + B.ay.y), fbz = G * (B.az.x + B.az.y);
            if (va) acc[ka] = make_float4(fax, fay, faz, __int_as_float(COUNT ? A.cnt : slots_a));
            if (vb) acc[kb] = make_float4(fbx, fby, fbz, __int_as_float(COUNT ? B.cnt : slots_b));
            if (INTEG) {
                if (vmb == 0xffffffffu) {   // full tile (vb for every lane implies va)
                    double xa[3], wa[3], xb[3], wb[3];
                    integrate_body(so, ka, fax, fay, faz, xa, wa, w_maxabs);
                    integrate_body(so, kb, fbx, fby, fbz, xb, wb, w_maxabs);
                    finish_tile64(so, base, lane, reinterpret_cast<double*>(ws.stk_first), xa, wa, xb, wb);
                } else {
                    if (va) finish_body(so, ka, fax, fay, faz, w_maxabs);
                    if (vb) finish_body(so, kb, fbx, fby, fbz, w_maxabs);
                }
            }
        }
        if (COUNT) {
            unsigned c32 = (va ? (unsigned)A.cnt : 0u) + (vb ? (unsigned)B.cnt : 0u), l32 = (unsigned)(A.lanepairs + B.lanepairs);
            for (int o = 16; o > 0; o >>= 1) {
                c32 += __shfl_xor_sync(0xffffffffu, c32, o);
                l32 += __shfl_xor_sync(0xffffffffu, l32, o);
            }
            w_inter += c32;
            w_lanepairs += l32;
            w_slots += (unsigned)(slots_a + slots_b);
        }
    }
    if (INTEG) finish_warp(so, w_maxabs);
    if (COUNT && lane == 0) {
        if (w_inter) atomicAdd(&counters[0], w_inter);
        atomicAdd(&counters[1], w_slots);
        atomicAdd(&counters[2], w_lanepairs);
        atomicAdd(&counters[3], w_batches);
        atomicMax(&counters[4], (unsigned long long)w_spmax);
        atomicAdd(&counters[5], w_both);
        atomicAdd(&counters[6], w_sure);
    }
}
