// radix_sort.cuh -- hand-written stable LSD radix sort (8-bit digits) for sm_100a.
//
// One up-front kernel histograms every digit of every pass in a single read of the keys;
// each pass is then ONE kernel ("onesweep"): a CTA ranks a tile of keys with warp
// match-any ballots, obtains its global digit offsets from its predecessors by decoupled
// look-back over a per-tile status table, stages the tile in shared memory in sorted
// order and writes coalesced runs.  Per pass the keys and values are read once and
// written once: 2*(sizeof(K)+4) bytes per element, plus sizeof(K) for the histogram read.
//
// Stability: tile t covers input [t*TILE, (t+1)*TILE); inside a tile elements are ranked in
// input order, so equal digits keep input order (ties in the full key are broken by
// input position -- the property the bit-exact permutation check relies on).
#pragma once
#include "common.cuh"

namespace b200 {
namespace rsort {

constexpr int RADIX = 256;
constexpr int BLOCK = 256;
constexpr int WARPS = BLOCK / 32;
constexpr unsigned FLAG_AGG = 1u << 30;
constexpr unsigned FLAG_INC = 2u << 30;
constexpr unsigned FLAG_MASK = 3u << 30;
constexpr unsigned VAL_MASK = ~FLAG_MASK;
constexpr int MAX_PASSES = 8;
constexpr int LOOKBACK = 8;          // status words fetched per look-back round

#ifndef RSORT_ITEMS64
#define RSORT_ITEMS64 12
#endif
#ifndef RSORT_MINB
#define RSORT_MINB 4
#endif
template <typename K> struct Tile { static constexpr int ITEMS = RSORT_ITEMS64; };
template <> struct Tile<uint32_t> { static constexpr int ITEMS = 16; };

// ---------------------------------------------------------------- histogram of all passes
template <typename K>
__global__ void __launch_bounds__(BLOCK) hist_kernel(const K* __restrict__ keys, int n, int begin_bit, int npass,
                                                     unsigned* __restrict__ ghist)
{
    __shared__ unsigned sh[MAX_PASSES * RADIX];
    for (int i = threadIdx.x; i < npass * RADIX; i += BLOCK) sh[i] = 0;
    __syncthreads();
    // whole warps walk the keys 32 at a time.  The input is nearly sorted (last step's order), so
    // in the high digits all 32 lanes usually share one bin: one +32 instead of 32 conflicting
    // atomics; low digits are close to random and conflict-free.
    const int64_t nwarps = (int64_t)gridDim.x * (BLOCK / 32);
    for (int64_t w = (int64_t)blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5); w * 32 < n; w += nwarps) {
        const int64_t i = w * 32 + lane_id();
        if (w * 32 + 32 <= n) {
            const K k = keys[i];
            for (int p = 0; p < npass; ++p) {
                const unsigned d = (unsigned)(k >> (begin_bit + 8 * p)) & 255u;
                const unsigned d0 = __shfl_sync(0xffffffffu, d, 0);
                if (__all_sync(0xffffffffu, d == d0)) {
                    if (lane_id() == 0) atomicAdd(&sh[p * RADIX + d0], 32u);
                } else {
                    atomicAdd(&sh[p * RADIX + d], 1u);
                }
            }
        } else if (i < n) {
            const K k = keys[i];
            for (int p = 0; p < npass; ++p) atomicAdd(&sh[p * RADIX + ((unsigned)(k >> (begin_bit + 8 * p)) & 255u)], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * RADIX; i += BLOCK)
        if (sh[i]) atomicAdd(&ghist[i], sh[i]);
}

// exclusive scan of each pass's 256 bins, in place: ghist[p][d] -> first output slot of digit d
static __global__ void __launch_bounds__(RADIX) scan_hist_kernel(unsigned* ghist)
{
    __shared__ unsigned wsum[RADIX / 32];
    unsigned* h = ghist + blockIdx.x * RADIX;
    const unsigned v = h[threadIdx.x];
    unsigned inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane_id() >= (unsigned)o) inc += t;
    }
    if (lane_id() == 31) wsum[threadIdx.x >> 5] = inc;
    __syncthreads();
    unsigned base = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) base += wsum[w];
    h[threadIdx.x] = base + inc - v;
}

// lanes of the warp holding the same 8-bit digit: eight ballots (one per bit) instead of MATCH.ANY,
// whose latency grows with the number of distinct values in the warp
__device__ __forceinline__ unsigned match_digit(unsigned d)
{
    unsigned peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const unsigned bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
    }
    return peers;
}

// ---------------------------------------------------------------- one pass
template <typename K, int ITEMS, bool IOTA, bool BALLOTS>
__global__ void __launch_bounds__(BLOCK, RSORT_MINB) onesweep_kernel(const K* __restrict__ keys_in, K* __restrict__ keys_out,
                                                         const uint32_t* __restrict__ vals_in,
                                                         uint32_t* __restrict__ vals_out, int n, int shift,
                                                         const unsigned* __restrict__ digit_start,
                                                         unsigned* status, unsigned* ticket)
{
    constexpr int TILE = BLOCK * ITEMS;
    __shared__ unsigned warp_hist[WARPS][RADIX];
    __shared__ K skeys[TILE];
    __shared__ uint32_t svals[TILE];
    __shared__ unsigned tile_off[RADIX];
    __shared__ unsigned gbase[RADIX];
    __shared__ unsigned wsum[WARPS];
    __shared__ unsigned s_tile;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const unsigned lane = lane_id();
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);   // dynamic tile order => predecessors are already running
    for (int i = tid; i < WARPS * RADIX; i += BLOCK) (&warp_hist[0][0])[i] = 0;
    tile_off[tid] = 0;                              // doubles as the tile's early digit histogram
    __syncthreads();
    const unsigned tile = s_tile;
    const int64_t base = (int64_t)tile * TILE;
    const int valid = (int)min((int64_t)TILE, (int64_t)n - base);

    // ---- load (warp-striped: item j of lane l is element warp*ITEMS*32 + j*32 + l of the tile)
    K key[ITEMS];
    unsigned rank[ITEMS];
    const int woff = warp * ITEMS * 32 + (int)lane;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const int t = woff + j * 32;
        key[j] = (t < valid) ? keys_in[base + t] : (K)~(K)0;   // padding sorts to the tile's end
    }
    // ---- early digit counts of the tile, published at once so successors never wait on our ranking
    unsigned same = 0;   // bit j: all 32 lanes hold the same digit in item j (nearly sorted input, high digits)
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const unsigned d = (unsigned)(key[j] >> shift) & 255u;
        const unsigned d0 = __shfl_sync(0xffffffffu, d, 0);
        if (__all_sync(0xffffffffu, d == d0)) {
            same |= 1u << j;
            if (lane == 0) atomicAdd(&tile_off[d0], 32u);
        } else {
            atomicAdd(&tile_off[d], 1u);
        }
    }
    __syncthreads();
    const int d = tid;
    const unsigned count = tile_off[d];
    unsigned* st = status + (size_t)tile * RADIX + d;
    st_relaxed_u32(st, (tile > 0 ? FLAG_AGG : FLAG_INC) | count);

    // ---- rank inside (warp, digit): all match-any first (independent), then the running counters
    const unsigned lt = lanemask_lt();
    unsigned peers[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const unsigned dj = (unsigned)(key[j] >> shift) & 255u;
        // measured on B200: MATCH.ANY wins when many tiles keep the SMs busy (50 M keys: 4.2 vs 4.6 ms per
        // sort), the ballots win when the sort is latency-bound (1 M 32-bit keys: 0.072 vs 0.092 ms)
        if (BALLOTS) peers[j] = (same >> j) & 1u ? 0xffffffffu : match_digit(dj);
        else peers[j] = __match_any_sync(0xffffffffu, dj);
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const unsigned dj = (unsigned)(key[j] >> shift) & 255u;
        const int leader = __ffs(peers[j]) - 1;
        unsigned prev = 0;
        if ((int)lane == leader) prev = atomicAdd(&warp_hist[warp][dj], (unsigned)__popc(peers[j]));
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[j] = prev + __popc(peers[j] & lt);
    }
    __syncthreads();

    // ---- per digit: exclusive offsets of each warp; exclusive scan of the tile's digit counts
    {
        unsigned run = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const unsigned t = warp_hist[w][d];
            warp_hist[w][d] = run;
            run += t;
        }
    }
    unsigned inc = count;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    unsigned wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += wsum[w];
    const unsigned toff = wbase + inc - count;
    tile_off[d] = toff;
    __syncthreads();

    // ---- stage keys and values in sorted order
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const unsigned dd = (unsigned)(key[j] >> shift) & 255u;
        const unsigned p = tile_off[dd] + warp_hist[warp][dd] + rank[j];
        rank[j] = p;
        skeys[p] = key[j];
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const int t = woff + j * 32;
        uint32_t v;
        if (IOTA) v = (uint32_t)(base + t);
        else v = (t < valid) ? vals_in[base + t] : 0u;
        svals[rank[j]] = v;
    }

    // ---- decoupled look-back for digit d (after the staging, so the wait overlaps useful work)
    // LOOKBACK status words are fetched per round (independent loads in flight) and consumed in order up to
    // the first inclusive prefix or the first word that is not published yet: the walk over the tiles
    // that run concurrently is latency-bound, not the wait for any single predecessor.
    unsigned excl = 0;
    if (tile > 0) {
        int t = (int)tile - 1;
        bool done = false;
        while (!done) {
            unsigned sw[LOOKBACK];
#pragma unroll
            for (int k = 0; k < LOOKBACK; ++k)
                sw[k] = t - k >= 0 ? ld_relaxed_u32(status + (size_t)(t - k) * RADIX + d) : FLAG_INC;
            bool go = true;   // false after the inclusive prefix or the first word that is not published yet
#pragma unroll
            for (int k = 0; k < LOOKBACK; ++k) {
                go = go && (sw[k] & FLAG_MASK) != 0u;
                if (go) {
                    excl += sw[k] & VAL_MASK;
                    --t;
                    done = (sw[k] & FLAG_INC) != 0u;
                    go = !done;
                }
            }
        }
        st_relaxed_u32(st, FLAG_INC | (excl + count));
    }
    gbase[d] = digit_start[d] + excl - toff;
    __syncthreads();

    // ---- coalesced runs out
    for (int q = tid; q < valid; q += BLOCK) {
        const K k = skeys[q];
        const unsigned dd = (unsigned)(k >> shift) & 255u;
        const unsigned g = gbase[dd] + q;
        keys_out[g] = k;
        vals_out[g] = svals[q];
    }
}

// ---------------------------------------------------------------- host driver
template <typename K>
struct Sorter {
    int n_max = 0;
    int tiles_max = 0;
    unsigned* ghist = nullptr;     // [MAX_PASSES][256]
    unsigned* status = nullptr;    // [MAX_PASSES][tiles_max][256]
    unsigned* tickets = nullptr;   // [MAX_PASSES]
    size_t scratch_bytes = 0;
    unsigned char* scratch = nullptr;
    int last_launches = 0;         // kernels launched by the last sort()

    static constexpr int ITEMS = Tile<K>::ITEMS;
    static constexpr int TILE = BLOCK * ITEMS;

    void init(int n_max_)
    {
        n_max = n_max_;
        tiles_max = div_up(n_max > 0 ? n_max : 1, TILE);
        const size_t words = (size_t)MAX_PASSES * RADIX + (size_t)MAX_PASSES * tiles_max * RADIX + MAX_PASSES;
        scratch_bytes = words * sizeof(unsigned);
        scratch = dev_alloc<unsigned char>(scratch_bytes);
        ghist = reinterpret_cast<unsigned*>(scratch);
        status = ghist + (size_t)MAX_PASSES * RADIX;
        tickets = status + (size_t)MAX_PASSES * tiles_max * RADIX;
    }
    void destroy()
    {
        if (scratch) cudaFree(scratch);
        scratch = nullptr;
    }
    size_t bytes() const { return scratch_bytes; }

    // Sorts n (key, value) pairs on bits [begin_bit, end_bit).  Buffers ping-pong between
    // (keys[0], vals[0]) and (keys[1], vals[1]); the input is in slot `cur`, the returned int
    // is the slot holding the result.  iota: values are the input positions 0..n-1 and
    // vals[cur] is not read.
    int sort(K* keys[2], uint32_t* vals[2], int cur, int n, int begin_bit, int end_bit, bool iota,
             cudaStream_t stream, int sm_count)
    {
        B200_REQUIRE(n <= n_max, "radix sort: n exceeds workspace");
        last_launches = 0;
        if (n <= 1 && !iota) return cur;
        const int npass = (end_bit - begin_bit + 7) / 8;
        B200_REQUIRE(npass >= 1 && npass <= MAX_PASSES, "radix sort: bad bit range");
        const int tiles = div_up(n > 0 ? n : 1, TILE);
        // one memset covers histograms, the used part of the status table is cleared per pass region
        B200_CHECK(cudaMemsetAsync(ghist, 0, (size_t)MAX_PASSES * RADIX * sizeof(unsigned), stream));
        B200_CHECK(cudaMemsetAsync(tickets, 0, MAX_PASSES * sizeof(unsigned), stream));
        for (int p = 0; p < npass; ++p)
            B200_CHECK(cudaMemsetAsync(status + (size_t)p * tiles_max * RADIX, 0,
                                       (size_t)tiles * RADIX * sizeof(unsigned), stream));
        if (n > 0) {
            const int hblocks = min(div_up(n, BLOCK * 8), sm_count * 8);
            hist_kernel<K><<<hblocks, BLOCK, 0, stream>>>(keys[cur], n, begin_bit, npass, ghist);
            scan_hist_kernel<<<npass, RADIX, 0, stream>>>(ghist);
            const bool ballots = tiles < 8 * sm_count;   // fewer than two waves of resident CTAs: latency-bound
            for (int p = 0; p < npass; ++p) {
                const int shift = begin_bit + 8 * p;
                unsigned* st = status + (size_t)p * tiles_max * RADIX;
                auto launch = [&](auto kernel) {
                    kernel<<<tiles, BLOCK, 0, stream>>>(keys[cur], keys[cur ^ 1], vals[cur], vals[cur ^ 1], n, shift,
                                                         ghist + p * RADIX, st, tickets + p);
                };
                if (iota && p == 0) {
                    if (ballots) launch(onesweep_kernel<K, ITEMS, true, true>);
                    else launch(onesweep_kernel<K, ITEMS, true, false>);
                } else {
                    if (ballots) launch(onesweep_kernel<K, ITEMS, false, true>);
                    else launch(onesweep_kernel<K, ITEMS, false, false>);
                }
                cur ^= 1;
            }
            B200_CHECK(cudaGetLastError());
            last_launches = npass + 2;
        }
        return cur;
    }
};

}  // namespace rsort
}  // namespace b200
