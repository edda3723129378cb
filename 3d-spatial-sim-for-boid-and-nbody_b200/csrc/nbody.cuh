// nbody.cuh -- device state and kernels of the Barnes-Hut step (sm_100a).
//
// Replaces, on device, the reference's per-substep sequence (tools/record.py:835-858):
//   compute_bounds (nbody/simulation.py:308-317) -> build_octree (:63-198) ->
//   compute_forces_barnes_hut (:201-278) -> update_positions_velocities (:281-305)
// and compute_colors_by_velocity (:320-400).
//
// Data layout in HBM (N bodies):
//   state[2]   double-buffered master state in the CURRENT Morton order:
//                pos (N,3) f64, vel (N,3) f64, mass (N) f64, id (N) u32 = creation index
//   keys[2]/vals[2]  63-bit Morton keys + permutation, radix-sort ping-pong
//   posm       (N) float4 {x,y,z,m} of the sorted bodies (traversal targets / leaf records)
//   binary radix tree over the sorted keys (N-1 internal nodes): childL/childR/parent,
//                range, lvl; ploc/bex = blocked fp64 prefix sums of (m x, m y, m z, m) over the
//                sorted bodies, so any node's mass / centre of mass is a difference of two entries
//   recs       octree "pair records", 64 B each = 4 x float4, holding children 2j and 2j+1 of
//              a cell side by side for packed fp32x2 math:
//                {x0,x1,y0,y1} {z0,z1,m0,m1} {T0,T1,-,-} {first0,first1,nchild0,nchild1}
//              T = max(size^2/theta^2, eps^2) (leaf: eps^2); the children of a cell are
//              CONTIGUOUS pairs, so opening a cell is one coalesced 16-byte-per-lane load.
//   acc        (N) float4 {ax, ay, az, interaction count} in sorted order
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"
#include <vector>

namespace b200 {

struct __align__(16) D4 { double x, y, z, w; };

constexpr int MORTON_LEVELS = 21;
// (build-time knobs of the traversal, for A/B builds: scripts/build_variants.sh; the defaults are the shipped kernel)
#ifndef B200_TRAV_BLOCK
#define B200_TRAV_BLOCK 256
#endif
#ifndef B200_TRAV_CAP
#define B200_TRAV_CAP 512
#endif
#ifndef B200_TRAV64_CTAS
#define B200_TRAV64_CTAS 3           // resident CTAs per SM of the 64-body walk (launch bounds and grid size)
#endif
#ifndef B200_TRAV64_PAD
#define B200_TRAV64_PAD 0            // extra dynamic shared memory per CTA (occupancy experiments)
#endif
constexpr int TRAV_BLOCK = B200_TRAV_BLOCK;
constexpr int TRAV_WARPS = TRAV_BLOCK / 32;
constexpr int TRAV_BATCH = 32;       // pair slots evaluated per batch
constexpr int TRAV_AREA = TRAV_BATCH + 2;   // entries between staging areas (bank offset of 32 B)
constexpr int TRAV_CAP = B200_TRAV_CAP;        // stack entries per warp (only cells that must be opened are pushed)
constexpr int TRAV_DFS_MARK = 256;   // 32-body walk: above this many entries a batch pops only the top one
constexpr int TRAV_RESERVE = 160;    // stack room kept free by the batch size limit (>= the depth-first bound 7 * 21)
constexpr int TRAV_CELL_PAIRS = 4;    // pairs of an ordinary cell (<= 8 children); more = a finest-level bucket
constexpr int TRAV_COUNTERS = 8;     // interactions, pair slots, lane-pairs, batches, stack high-water

enum NBodyPhase { PH_KEYGEN = 0, PH_SORT, PH_GATHER, PH_BUILD, PH_EXTRACT, PH_TRAVERSE, PH_EXCHANGE, PH_INTEGRATE, PH_COUNT };

enum NBodyError { ERR_STACK_OVERFLOW = 1, ERR_RECORD_OVERFLOW = 2, ERR_PRUNED_CELL_OPENED = 4 };
// locally essential tree test: bounding boxes of the shard's bodies per key prefix of 0, 1, 2 and 3 octree levels
// (1 + 8 + 64 + 512 boxes, a small octree of boxes walked per cell)
constexpr int LET_LEVELS = 3;
constexpr int LET_BOXES = 1 + 8 + 64 + 512;
constexpr unsigned PRUNED_NCHILD = 0x3ffffffu; // "children not materialised on this rank" in a record's child count

struct NBodySim {
    int n = 0;
    int device = 0;
    int sm_count = 148;
    double G = 0, softening = 0, damping = 1, theta = 0.5;
    cudaStream_t stream = nullptr;

    // Master state.  Positions rotate through THREE buffers -- previous state (keygen / gather input) ->
    // the same bodies in the new Morton order (written by the gather, read by the prefix sums and by the fused
    // integration) -> next state (written by the traversal's epilogue, on every rank) -- so a peer GPU may
    // already write step t's result while this GPU still reads step t's inputs.  Velocities, masses and ids
    // alternate between two buffers; velocities are never reordered by a pass of their own.
    double* pos[3] = {nullptr, nullptr, nullptr};
    double* vel[2] = {nullptr, nullptr};
    double* mass[2] = {nullptr, nullptr};
    double* mass0 = nullptr;                  // masses in creation order (what set_state restores without a scatter)
    uint32_t* id[2] = {nullptr, nullptr};
    int pcur = 0, vcur = 0;                   // buffers holding the current state (pos / vel, mass, id)
    bool step_pending = false;                // between a step's tree build and its fused traversal + integration

    uint64_t* keys[2] = {nullptr, nullptr};
    uint32_t* vals[2] = {nullptr, nullptr};
    int sorted_slot = 0;             // which keys/vals slot holds the last sort's output
    rsort::Sorter<uint64_t> sorter;

    float4* posm = nullptr;
    float4* acc = nullptr;
    int *childL = nullptr, *childR = nullptr, *parent = nullptr;
    int2* range = nullptr;
    D4 *ploc = nullptr, *bex = nullptr;       // blocked fp64 prefix sums of (m x, m y, m z, m)
    unsigned char* ishead = nullptr;          // binary node is an octree cell
    int4* meta = nullptr;                     // per octree cell: {range lo, range hi, first pair, children << 5 | level}
    signed char* lvl = nullptr;               // octree level of every binary node
    int4* kids = nullptr;                     // per head: its <= 8 octree children (2 x int4)
    float4* recs = nullptr;
    int64_t rec_capacity = 0;

    // small device scalars
    unsigned long long* d_maxabs = nullptr;   // [2] bit patterns of max |coord|
    int maxabs_slot = 0;
    double* d_bounds = nullptr;
    float* d_ttab = nullptr;                  // MAC threshold per octree level of the current tree
    int* d_root = nullptr;
    unsigned* d_alloc = nullptr;              // pair-record allocator
    unsigned* d_children = nullptr;           // octree children (cells + leaves) of the last tree
    unsigned* d_tile_counter = nullptr;
    unsigned long long* d_interactions = nullptr;
    unsigned* d_error = nullptr;
    unsigned* h_error = nullptr;              // pinned copy of d_error taken with every frame (checked by frame_wait)
    int64_t rec_capacity_override = 0;        // B200_REC_CAPACITY: shrinks the record pool (tests force an overflow)

    float* colors = nullptr;                  // (N,3) f32, creation order
    void* stage = nullptr;                    // (N,3) f64-sized staging for getters
    bool tree_valid = false;                  // keys/perm/tree describe the current positions
    // captured steps (nbody_step): CUDA graphs cached by (current buffer, parameters)
    static constexpr int MAX_STEP_GRAPHS = 8;   // (3 position buffers x 2 velocity buffers: six graphs in steady state)
    cudaGraphExec_t step_graph[MAX_STEP_GRAPHS] = {};
    alignas(8) unsigned char step_graph_key[MAX_STEP_GRAPHS][96] = {};
    int64_t step_graph_launches[MAX_STEP_GRAPHS] = {};
    int step_graph_sorted_slot[MAX_STEP_GRAPHS] = {};
    int step_graph_pcur[MAX_STEP_GRAPHS] = {}, step_graph_vcur[MAX_STEP_GRAPHS] = {};   // state buffers after the step
    unsigned step_graph_next = 0;
    bool use_graph = true;
    int trav_mode = 0;                        // 0: per launch (64 for large N and theta, else 32); 32 / 64: forced
    int last_trav_kernel = 0;                 // 32 / 64: walk of the last traversal launch
    bool count_interactions = false;          // exact per-body interaction counts in the traversal (slower)

    // multi-GPU: this rank traverses sorted bodies [shard_begin, shard_end) (multiples of 32)
    int rank = 0, world = 1;
    int shard_begin = 0, shard_end = 0;
    // locally essential tree (sharded step only): children records are written only for cells some body of the
    // rank's shard may open (conservative box test); B200_LET=0 writes the whole tree on every rank
    bool let_enabled = true;
    int* d_boxes = nullptr;                   // [LET_BOXES][6] order-preserving ints: lo x,y,z, hi x,y,z of the shard's bodies per key prefix

    // asynchronous host traffic (frame egress / state prefetch), allocated on first use
    cudaStream_t up_stream = nullptr, down_stream = nullptr;
    cudaEvent_t ev_frame_ready = nullptr, ev_frame_done = nullptr;     // device staging filled / D2H finished
    cudaEvent_t ev_upload_done = nullptr, ev_upload_consumed = nullptr; // H2D finished / staging read by the commit
    float *frame_pos = nullptr, *frame_col = nullptr;                  // (N,3) f32 staging, creation order
    short *frame_dpos = nullptr, *frame_dcol = nullptr;                // (N,3) int16 delta staging (frame codec), lazily allocated
    float *frame_pos2 = nullptr, *frame_col2 = nullptr;                // second f32 staging: delta frames alternate between the two
    int unperm_min_n = 4 * 1024 * 1024;                                // below this n the single scatter kernel is used
    int unperm_shift = 20;
    unsigned* unperm_counts = nullptr;                                 // bucket counts + cursors of the bucketed un-permute
    bool frame_has_prev = false;                                       // frame_pos / frame_col hold the previous frame
    double *up_pos = nullptr, *up_vel = nullptr;                       // (N,3) f64 staging of a prefetched state
    bool frame_pending = false, upload_pending = false;
    // live-viewer path: frustum cull + ordered compaction of the frame (lazily allocated)
    float *vis_pos = nullptr, *vis_col = nullptr;                      // (N,3) f32 compacted visible bodies
    unsigned* vis_counts = nullptr;                                    // per-block visible counts -> exclusive offsets; [blocks] = total
    unsigned* vis_total_host = nullptr;                                // pinned

    // sharded sort (multi-GPU): padded exchange buffers of world * slice sorted (key, local position) pairs
    uint64_t* ms_keys = nullptr;
    uint32_t* ms_vals = nullptr;
    int ms_slice = 0, ms_world = 0;

    PhaseTimer timer;
    int64_t steps = 0;
    int64_t launches = 0;                     // kernels launched by this handle (bench: gpu_launches)
    cudaStream_t own_stream = nullptr;
    int64_t acc_capacity = 0;                 // entries in acc (padded so equal shard slices fit)
    size_t bytes_allocated = 0;
};

void nbody_alloc(NBodySim& s, int n);
void nbody_free(NBodySim& s);
void nbody_upload(NBodySim& s, const double* pos, const double* vel, const double* mass);
void nbody_upload_state(NBodySim& s, const double* pos, const double* vel);
// generate.cu: seeded initial conditions on the device (tools/presets.py:91-1390, generate_distribution)
int generator_id(const char* name);
void generate_device(int dist, int64_t n, double R, double G, uint64_t seed, double* pos, double* vel, double* mass,
                     cudaStream_t stream, int sm_count);
// initial state of the handle generated in place (no host traffic)
void nbody_generate(NBodySim& s, int dist, double R, double G_dist, uint64_t seed);
// keygen .. extract: leaves keys/perm/tree valid for the current positions; the state is physically in the new
// Morton order afterwards (positions, velocities, masses, ids)
void nbody_build_tree(NBodySim& s);
// traversal of sorted bodies [begin, end) into s.acc (forces only)
void nbody_traverse(NBodySim& s, int begin, int end);
// sharded sort: setup the exchange buffers; keygen + local sort of one slice; merge + rest of the tree
void nbody_ms_setup(NBodySim& s, int slice, int world);
void nbody_ms_sort_local(NBodySim& s, int rank);
void nbody_build_tree_presorted(NBodySim& s);
void nbody_integrate(NBodySim& s, double dt);
void nbody_step(NBodySim& s, double dt);
void nbody_graphs_reset(NBodySim& s);
// split step for sharded runs: begin = tree + traversal of [shard_begin, shard_end); the caller
// then exchanges acc slices between ranks on the same stream; end = integrate of all bodies.
void nbody_step_begin(NBodySim& s);
void nbody_step_begin_sorted(NBodySim& s);   // after nbody_ms_sort_local + the all-gather of the exchange buffers
void nbody_step_end(NBodySim& s, double dt);
// ---- sharded fused step (multi-GPU inside the library, see multi.cu): every rank holds the full replicated
// state, sorts one slice, builds the whole tree, traverses and integrates only its shard, and its traversal
// kernel writes the new positions / velocities of those bodies into the next-state buffers of EVERY rank
// (peer-mapped memory over NVLink).  The caller supplies the collectives between the pieces.
constexpr int SHARD_MAX_WORLD = 8;
struct ShardPeers {                            // next-state buffers of every rank, as seen from this device
    double* pos[SHARD_MAX_WORLD][3];
    double* vel[SHARD_MAX_WORLD][2];
    int world = 1;
};
void nbody_shard_build(NBodySim& s, bool presorted);   // (merge of the all-gathered runs +) gather, tree, records
void nbody_shard_traverse(NBodySim& s, double dt, const ShardPeers& peers);   // forces + integration + broadcast of the shard
void nbody_shard_finish(NBodySim& s);          // host side: advance the state buffers (after the caller's all-reduce of d_maxabs)
unsigned long long* nbody_shard_maxabs_next(NBodySim& s);
// multi.cu
struct NBodyGroup;
void nccl_unique_id(void* out128);
NBodyGroup* group_create_rank(NBodySim& s, const void* id128, int rank, int world);
NBodyGroup* group_create_local(const std::vector<NBodySim*>& sims);
void group_destroy(NBodyGroup* g);
int group_world(const NBodyGroup* g);
void group_state_replaced(NBodyGroup* g);
void group_step(NBodyGroup& g, double dt);
std::vector<int64_t> cost_weighted_split(const unsigned long long* cost, int nchunks, int64_t chunk, int64_t n, int world);
double fp32_peak_tflops(int device);
void nbody_compute_colors(NBodySim& s, double max_speed);
// after a synchronisation: throws StateError if a kernel raised a device error flag (sticky)
void nbody_check_errors(NBodySim& s);
void nbody_get_positions(NBodySim& s, float* out);
void nbody_get_positions_f64(NBodySim& s, double* out);
void nbody_get_velocities(NBodySim& s, double* out);
void nbody_get_colors(NBodySim& s, float* out);
void nbody_get_accelerations(NBodySim& s, float* out);   // creation order, runs build+traverse if needed
int64_t nbody_count_interactions(NBodySim& s);           // one counting force pass over the shard, current state
void nbody_state_checksum(NBodySim& s, uint64_t out[2]);
void nbody_get_keys(NBodySim& s, uint64_t* out);
void nbody_get_perm(NBodySim& s, uint32_t* out);
// frame egress: colours + creation-order fp32 positions on the compute stream, D2H on a second stream
void nbody_frame_begin(NBodySim& s, double max_speed, float* host_pos, float* host_col);
void nbody_frame_begin_rows(NBodySim& s, double max_speed, float* host_pos, float* host_col, int row_begin, int row_end);
void nbody_frame_wait(NBodySim& s);
// live-viewer path (nbody/simulation.py:403-434 compute_visibility_points + the mask gathers of draw() :926-927):
// colours, creation-order fp32 frame, frustum test, stable compaction; the visible bodies go to host buffers or
// (to_device) straight into device memory such as a mapped OpenGL VBO.  Returns the visible count.  Blocking.
struct Camera { double pos[3], forward[3], right[3], up[3], tan_h, tan_v, far_dist; };
int64_t nbody_visible_frame(NBodySim& s, double max_speed, const Camera& cam, float* out_pos, float* out_col, bool to_device);
void nbody_frame_delta_begin(NBodySim& s, double max_speed, short* host_dpos, short* host_dcol);
// state prefetch: H2D on a third stream into staging; commit swaps it in on the compute stream
void nbody_set_state_begin(NBodySim& s, const double* pos, const double* vel);
void nbody_set_state_begin_rows(NBodySim& s, const double* pos, const double* vel, int row_begin, int row_end);
void nbody_upload_staging(NBodySim& s, double** pos, double** vel);
void nbody_upload_wait(NBodySim& s);
void nbody_set_state_commit(NBodySim& s);

}  // namespace b200
