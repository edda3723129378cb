// multi.cu -- several GPUs behind one C ABI: NCCL (loaded at run time) for the two small collectives of the
// sharded sort and for the barrier that ends a step, peer-mapped state buffers for everything else.
//
// Two ways to get there, one step function:
//   * one process per GPU (torchrun): b200_nccl_unique_id on rank 0, the 128 bytes travel by any means (the
//     Python wrapper broadcasts them with torch.distributed), b200_nbody_comm_init on every rank.  The state
//     buffers of the other ranks are mapped with CUDA IPC.
//   * one process, several GPUs (b200_nbody_create_multi with a device mask): the replicas live in one handle,
//     ncclCommInitAll, direct peer access.  This is how a caller without torch -- the reference's recorder
//     through create_gpu_simulation -- uses more than one GPU.
// Step (every rank r of N; S = slice = whole 64-body tiles):
//   keygen + radix sort of the bodies at current positions [r S, (r+1) S)            (1/N of the sort)
//   all-gather of the sorted runs: 8 B key + 4 B position per body                   (NCCL over NVLink)
//   merge by counting, gather of positions / masses / ids, tree, records            (replicated, deterministic)
//   traversal of the rank's shard; its epilogue integrates those bodies and stores the new positions and
//     velocities into the next-state buffers of ALL ranks                           (P2P stores, inside the kernel)
//   all-reduce (max) of the next bounds = the barrier after which every replica holds the complete next state.
// The reference has no multi-GPU code: this is new design.
#include "nbody.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace b200 {

// ---------------------------------------------------------------------------- NCCL, loaded lazily
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi& nccl()
{
    static NcclApi api;
    if (api.lib) return api;
    // (a process that already loaded a libnccl.so.2 -- torch bundles one -- gets that one back: same soname)
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) throw StateError{std::string("NCCL is not available (dlopen libnccl.so.2 failed): ") + dlerror()};
    auto sym = [&](const char* name) {
        void* p = dlsym(api.lib, name);
        if (!p) throw StateError{std::string("libnccl lacks ") + name};
        return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    return api;
}

#define B200_NCCL(expr)                                                                           \
    do {                                                                                          \
        ncclResult_t _r = (expr);                                                                 \
        if (_r != ncclSuccess) {                                                                  \
            char _b[512];                                                                         \
            snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, nccl().GetErrorString(_r), __FILE__, __LINE__); \
            throw ::b200::CudaError{std::string(_b)};                                             \
        }                                                                                         \
    } while (0)

void nccl_unique_id(void* out128)
{
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    B200_NCCL(nccl().GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
}

// ---------------------------------------------------------------------------- group of replicas
// sims[i] is rank ranks[i] of `world`; one entry per process with torchrun, all of them in one process with a
// device mask.
struct NBodyGroup {
    int world = 1;
    std::vector<NBodySim*> sims;
    std::vector<int> ranks;
    std::vector<ncclComm_t> comms;
    std::vector<ShardPeers> peers;                 // per local replica: every rank's buffers as mapped on its device
    std::vector<void*> ipc_opened;                 // cudaIpcOpenMemHandle results to close
    int slice = 0;
    bool in_morton_order = false;                  // the state is in last step's Morton order: slices are key ranges
    // cost-weighted Morton split: the traversal leaves a per-body cost (evaluated pair slots of its
    // half tile) in acc.w; every `rebalance_every` steps the costs are summed per chunk of COST_CHUNK sorted
    // positions, all-reduced (exact integer sums: every rank computes the same split) and the shard boundaries
    // move to equal-cost chunk boundaries.  The sort slices stay equal-count.
    std::vector<unsigned long long*> d_cost;       // per local replica: [nchunks]
    unsigned long long* h_cost = nullptr;          // pinned
    int nchunks = 0;
    int rebalance_every = 8;
    int64_t steps_since_reset = 0;
    std::vector<int> split;                        // world + 1 boundaries (sorted positions) of the current shards
};
constexpr int COST_CHUNK = 4096;                   // a multiple of the 64-body tiles: shards are whole tiles

static int slice_size(int n, int world)
{
    const int64_t tiles = (n + 63) / 64;
    return (int)(((tiles + world - 1) / world) * 64);
}

static void group_set_shards(NBodyGroup& g)
{
    for (size_t i = 0; i < g.sims.size(); ++i) {
        NBodySim& s = *g.sims[i];
        g.slice = slice_size(s.n, g.world);
        const int r = g.ranks[i];
        s.rank = r;
        s.world = g.world;
        s.shard_begin = (int)min((int64_t)r * g.slice, (int64_t)s.n);
        s.shard_end = (int)min((int64_t)(r + 1) * g.slice, (int64_t)s.n);
        nbody_ms_setup(s, g.slice, g.world);
    }
    g.steps_since_reset = 0;
    if (!g.sims.empty()) {
        const int n = g.sims[0]->n;
        g.split.assign(g.world + 1, 0);
        for (int r = 0; r <= g.world; ++r) g.split[r] = (int)min((int64_t)r * g.slice, (int64_t)n);
        if (const char* e = getenv("B200_REBALANCE")) g.rebalance_every = atoi(e);
        const int nchunks = div_up(n > 0 ? n : 1, COST_CHUNK);
        if (g.world > 1 && g.rebalance_every > 0 && g.d_cost.empty()) {
            g.nchunks = nchunks;
            for (NBodySim* s : g.sims) {
                B200_CHECK(cudaSetDevice(s->device));
                g.d_cost.push_back(dev_alloc<unsigned long long>((size_t)nchunks));
            }
            B200_CHECK(cudaMallocHost(&g.h_cost, (size_t)nchunks * sizeof(unsigned long long)));
        }
    }
}

// cost of the rank's shard per chunk of sorted positions (zero outside the shard): one CTA per chunk
__global__ void __launch_bounds__(256) shard_cost_kernel(const float4* __restrict__ acc, int begin, int end, unsigned long long* __restrict__ cost)
{
    const int lo = max(begin, (int)blockIdx.x * COST_CHUNK), hi = min(end, ((int)blockIdx.x + 1) * COST_CHUNK);
    unsigned long long c = 0;
    for (int k = lo + (int)threadIdx.x; k < hi; k += 256) c += (unsigned long long)(unsigned)__float_as_int(acc[k].w);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    __shared__ unsigned long long sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        cost[blockIdx.x] = t;
    }
}

// equal-cost shard boundaries at chunk granularity from the all-reduced chunk costs (identical on every rank);
// pure host arithmetic, exported as b200_cost_weighted_split for the CPU tests
std::vector<int64_t> cost_weighted_split(const unsigned long long* cost, int C, int64_t chunk, int64_t n, int W)
{
    std::vector<int64_t> split(W + 1, 0);
    split[W] = n;
    unsigned long long total = 0;
    for (int c = 0; c < C; ++c) total += cost[c];
    if (total == 0 || C < W) {   // nothing to weigh: equal chunk counts
        for (int r = 1; r < W; ++r) split[r] = std::min<int64_t>((int64_t)((int64_t)C * r / W) * chunk, n);
        return split;
    }
    unsigned long long run = 0;
    int c = 0;
    for (int r = 1; r < W; ++r) {
        const unsigned long long target = (unsigned long long)((long double)total * r / W);
        while (c < C && run + cost[c] / 2 < target) run += cost[c++];       // the chunk boundary nearest to the target
        const int lo_chunk = (int)(split[r - 1] / chunk) + 1;                // every rank keeps at least one chunk
        const int hi_chunk = C - (W - r);
        const int b = c < lo_chunk ? lo_chunk : (c > hi_chunk ? hi_chunk : c);
        while (c < b) run += cost[c++];
        split[r] = std::min<int64_t>((int64_t)b * chunk, n);
    }
    return split;
}

static void apply_cost_split(NBodyGroup& g, const unsigned long long* cost)
{
    const std::vector<int64_t> split = cost_weighted_split(cost, g.nchunks, COST_CHUNK, g.sims[0]->n, g.world);
    g.split.assign(split.begin(), split.end());
    for (size_t i = 0; i < g.sims.size(); ++i) {
        NBodySim& s = *g.sims[i];
        s.shard_begin = (int)split[g.ranks[i]];
        s.shard_end = (int)split[g.ranks[i] + 1];
    }
}

struct IpcBlock {                                  // what a rank tells the others about its state buffers
    cudaIpcMemHandle_t pos[3], vel[2];
    int n, device;
};

// one process per GPU: create the communicator, then exchange CUDA IPC handles of the five state buffers
NBodyGroup* group_create_rank(NBodySim& s, const void* id128, int rank, int world)
{
    B200_REQUIRE(world >= 1 && world <= SHARD_MAX_WORLD && rank >= 0 && rank < world, "bad rank / world (at most 8 ranks)");
    B200_CHECK(cudaSetDevice(s.device));
    auto* g = new NBodyGroup();
    try {
        g->world = world;
        g->sims.push_back(&s);
        g->ranks.push_back(rank);
        g->comms.resize(1);
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        B200_NCCL(nccl().CommInitRank(&g->comms[0], world, id, rank));
        // all-gather of the IPC handles through device memory
        IpcBlock mine;
        memset(&mine, 0, sizeof(mine));
        mine.n = s.n;
        mine.device = s.device;
        if (s.n > 0) {
            for (int b = 0; b < 3; ++b) B200_CHECK(cudaIpcGetMemHandle(&mine.pos[b], s.pos[b]));
            for (int b = 0; b < 2; ++b) B200_CHECK(cudaIpcGetMemHandle(&mine.vel[b], s.vel[b]));
        }
        IpcBlock* d_all = nullptr;
        B200_CHECK(cudaMalloc(&d_all, sizeof(IpcBlock) * world));
        B200_CHECK(cudaMemcpyAsync(d_all + rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, s.stream));
        B200_NCCL(nccl().AllGather(d_all + rank, d_all, sizeof(IpcBlock), ncclChar, g->comms[0], s.stream));
        std::vector<IpcBlock> all(world);
        B200_CHECK(cudaMemcpyAsync(all.data(), d_all, sizeof(IpcBlock) * world, cudaMemcpyDeviceToHost, s.stream));
        B200_CHECK(cudaStreamSynchronize(s.stream));
        cudaFree(d_all);
        ShardPeers p;
        memset(&p, 0, sizeof(p));
        p.world = world;
        for (int r = 0; r < world; ++r) {
            B200_REQUIRE(all[r].n == s.n, "the ranks disagree on the number of bodies");
            for (int b = 0; b < 3; ++b) {
                if (r == rank) p.pos[r][b] = s.pos[b];
                else if (s.n > 0) {
                    void* ptr = nullptr;
                    B200_CHECK(cudaIpcOpenMemHandle(&ptr, all[r].pos[b], cudaIpcMemLazyEnablePeerAccess));
                    g->ipc_opened.push_back(ptr);
                    p.pos[r][b] = (double*)ptr;
                }
            }
            for (int b = 0; b < 2; ++b) {
                if (r == rank) p.vel[r][b] = s.vel[b];
                else if (s.n > 0) {
                    void* ptr = nullptr;
                    B200_CHECK(cudaIpcOpenMemHandle(&ptr, all[r].vel[b], cudaIpcMemLazyEnablePeerAccess));
                    g->ipc_opened.push_back(ptr);
                    p.vel[r][b] = (double*)ptr;
                }
            }
        }
        g->peers.push_back(p);
        group_set_shards(*g);
    } catch (...) {
        delete g;
        throw;
    }
    return g;
}

// one process, several GPUs: replicas already created on their devices; direct peer access
NBodyGroup* group_create_local(const std::vector<NBodySim*>& sims)
{
    const int world = (int)sims.size();
    B200_REQUIRE(world >= 1 && world <= SHARD_MAX_WORLD, "between 1 and 8 devices");
    auto* g = new NBodyGroup();
    try {
        g->world = world;
        g->sims = sims;
        std::vector<int> devs(world);
        for (int r = 0; r < world; ++r) { g->ranks.push_back(r); devs[r] = sims[r]->device; }
        g->comms.resize(world);
        if (world > 1) B200_NCCL(nccl().CommInitAll(g->comms.data(), world, devs.data()));
        for (int r = 0; r < world; ++r) {
            B200_CHECK(cudaSetDevice(devs[r]));
            for (int q = 0; q < world; ++q) {
                if (q == r) continue;
                int can = 0;
                B200_CHECK(cudaDeviceCanAccessPeer(&can, devs[r], devs[q]));
                B200_REQUIRE(can, "the devices of the mask cannot access each other's memory (no NVLink / P2P)");
                const cudaError_t e = cudaDeviceEnablePeerAccess(devs[q], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) B200_CHECK(e);
                cudaGetLastError();
            }
            ShardPeers p;
            memset(&p, 0, sizeof(p));
            p.world = world;
            for (int q = 0; q < world; ++q) {
                for (int b = 0; b < 3; ++b) p.pos[q][b] = sims[q]->pos[b];
                for (int b = 0; b < 2; ++b) p.vel[q][b] = sims[q]->vel[b];
            }
            g->peers.push_back(p);
        }
        group_set_shards(*g);
    } catch (...) {
        delete g;
        throw;
    }
    return g;
}

void group_destroy(NBodyGroup* g)
{
    if (!g) return;
    for (NBodySim* s : g->sims) {
        cudaSetDevice(s->device);
        cudaStreamSynchronize(s->stream);
    }
    for (size_t i = 0; i < g->d_cost.size(); ++i) { cudaSetDevice(g->sims[i]->device); cudaFree(g->d_cost[i]); }
    if (g->h_cost) cudaFreeHost(g->h_cost);
    for (void* p : g->ipc_opened) cudaIpcCloseMemHandle(p);
    for (ncclComm_t c : g->comms)
        if (c) nccl().CommDestroy(c);
    delete g;
}

int group_world(const NBodyGroup* g) { return g ? g->world : 1; }

// a new state was uploaded (creation order): the next step sorts everything on every rank, and the shards go back
// to equal counts until the new order has been costed
void group_state_replaced(NBodyGroup* g)
{
    if (!g) return;
    g->in_morton_order = false;
    group_set_shards(*g);
}

void group_step(NBodyGroup& g, double dt)
{
    const size_t L = g.sims.size();
    const bool sharded_sort = g.world > 1 && g.in_morton_order;
    // 1. sort: the rank's slice (state in Morton order) or everything (first step after an upload: inside the build)
    if (sharded_sort) {
        for (size_t i = 0; i < L; ++i) nbody_ms_sort_local(*g.sims[i], g.ranks[i]);
        B200_NCCL(nccl().GroupStart());
        for (size_t i = 0; i < L; ++i) {
            NBodySim& s = *g.sims[i];
            const size_t off = (size_t)g.ranks[i] * g.slice;
            B200_NCCL(nccl().AllGather(s.ms_keys + off, s.ms_keys, (size_t)g.slice, ncclUint64, g.comms[i], s.stream));
            B200_NCCL(nccl().AllGather(s.ms_vals + off, s.ms_vals, (size_t)g.slice, ncclUint32, g.comms[i], s.stream));
        }
        B200_NCCL(nccl().GroupEnd());
    }
    // 2. tree (replicated), 3. forces + integration + broadcast of the shard
    for (size_t i = 0; i < L; ++i) nbody_shard_build(*g.sims[i], sharded_sort);
    for (size_t i = 0; i < L; ++i) nbody_shard_traverse(*g.sims[i], dt, g.peers[i]);
    // 4. the barrier that ends the step doubles as the reduction of the next bounds (and, on rebalancing steps, of
    //    the shards' chunk costs)
    ++g.steps_since_reset;
    const bool rebalance = g.world > 1 && !g.d_cost.empty() && g.sims[0]->n > 0 &&
                           (g.steps_since_reset == 1 || g.steps_since_reset % g.rebalance_every == 0);
    if (rebalance)
        for (size_t i = 0; i < L; ++i) {
            NBodySim& s = *g.sims[i];
            B200_CHECK(cudaSetDevice(s.device));
            shard_cost_kernel<<<g.nchunks, 256, 0, s.stream>>>(s.acc, s.shard_begin, s.shard_end, g.d_cost[i]);
            ++s.launches;
            B200_CHECK(cudaGetLastError());
        }
    if (g.world > 1) {
        B200_NCCL(nccl().GroupStart());
        for (size_t i = 0; i < L; ++i) {
            NBodySim& s = *g.sims[i];
            unsigned long long* m = nbody_shard_maxabs_next(s);
            B200_NCCL(nccl().AllReduce(m, m, 1, ncclUint64, ncclMax, g.comms[i], s.stream));
            if (rebalance) B200_NCCL(nccl().AllReduce(g.d_cost[i], g.d_cost[i], (size_t)g.nchunks, ncclUint64, ncclSum, g.comms[i], s.stream));
        }
        B200_NCCL(nccl().GroupEnd());
    }
    for (size_t i = 0; i < L; ++i) nbody_shard_finish(*g.sims[i]);
    if (rebalance) {   // one small read-back every `rebalance_every` steps: the shard bounds are launch arguments
        NBodySim& s = *g.sims[0];
        B200_CHECK(cudaSetDevice(s.device));
        B200_CHECK(cudaMemcpyAsync(g.h_cost, g.d_cost[0], (size_t)g.nchunks * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
        B200_CHECK(cudaStreamSynchronize(s.stream));
        apply_cost_split(g, g.h_cost);
    }
    g.in_morton_order = true;
}

}  // namespace b200
