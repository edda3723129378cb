/*
 * b200sim.h -- C ABI of libb200sim.so: the B200-native (sm_100a) Barnes-Hut step and boids
 * neighbour-rule update behind the reference's backend interface.
 *
 * Plain pointers and sizes only.  Every function returns 0 on success and a non-zero status
 * on failure; b200_last_error() then returns a message (thread-local).  Host pointers are
 * ordinary (pageable or pinned) host memory; arrays are C-contiguous.
 *
 * Reference interfaces replaced (paths relative to the reference repository):
 *   nbody/gpu_backend.py:336-409   class CUDASimulation (step / compute_colors / get_* / sync)
 *   nbody/gpu_backend.py:623-679   create_gpu_simulation(...)
 *   nbody/simulation.py:201-218    compute_forces_barnes_hut(...)  (the CPU "force entry point")
 *   tools/record.py:835-858        the CPU substep sequence the device step replaces
 *   boids/flock.py:627-678         Flock.update(dt)
 */
#ifndef B200SIM_H
#define B200SIM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_CUDA 1
#define B200_ERR_ARG 2
#define B200_ERR_STATE 3

typedef struct b200_nbody b200_nbody;   /* opaque handle */
typedef struct b200_boids b200_boids;   /* opaque handle */

/* Per-phase device time of step(), accumulated while profiling is on (CUDA events). */
#define B200_NBODY_PHASES 8
/* order: keygen, sort, gather, build, extract, traverse, exchange, integrate */

typedef struct b200_nbody_stats {
    int64_t n;                 /* bodies */
    int64_t steps;             /* step() calls so far */
    int64_t records;           /* octree records of the last tree: root + cells + leaves */
    int64_t interactions;      /* accepted body-node interactions since the last reset (device-counted) */
    double  bounds;            /* root half-size of the last tree: fma(max|coord|, 1.1, 10) */
    uint32_t error_flags;      /* 0 = clean; 1 = traversal stack overflow; 2 = record pool overflow; 4 = pruned cell opened */
    int32_t  sm_count;
    int64_t bytes_allocated;   /* device bytes owned by the handle */
    int64_t timed_steps;       /* steps accumulated in phase_ms */
    double  phase_ms[B200_NBODY_PHASES];
    int64_t pair_records;      /* 64-byte pair records allocated for the last tree */
    /* traversal work counters since the last reset (only while counting is on) */
    int64_t trav_pair_slots;   /* (pair record, 32-body half tile) evaluations (each = 2 children x 32 lanes) */
    int64_t trav_lane_pairs;   /* (lane, pair) evaluations where the lane was in the pair's mask */
    int64_t trav_batches;      /* select/load/eval/expand rounds */
    int64_t trav_stack_max;    /* high-water mark of a warp's stack (capacity 512) */
    int64_t trav_shared_pairs; /* pair records evaluated for both 32-body halves of a 64-body tile from one staging */
    int32_t trav_kernel;       /* walk of the last traversal launch: 32 = traverse_kernel (one body per lane), 64 = traverse64_kernel */
    int32_t reserved0;
    int64_t trav_sure_pairs;   /* (pair, half) evaluations the tile-level box test let skip the per-lane MAC (classed walk) */
} b200_nbody_stats;

const char* b200_last_error(void);
int b200_device_count(int* count);
/* name buffer receives "<name> (CC x.y, N GB)" like nbody/gpu_backend.py:58-70 prints */
int b200_device_info(int device, char* name, int name_len);

/* ---- n-body ------------------------------------------------------------------------------
 * create: replaces CUDASimulation.__init__ (nbody/gpu_backend.py:339-366) + theta of the
 * Metal Barnes-Hut twin.  Copies pos (n,3), vel (n,3), mass (n) [fp64] to the device.
 * 0 <= n < 2^26 bodies per handle (B200_ERR_ARG otherwise: the record format packs a cell's child count
 * into 26 bits); the reference itself stops at an 8 M-node pool (nbody/simulation.py:35). */
int b200_nbody_create(int64_t n, const double* pos, const double* vel, const double* mass,
                      double G, double softening, double damping, double theta,
                      int device, b200_nbody** out);
int b200_nbody_destroy(b200_nbody* h);
/* Seeded initial conditions generated ON THE DEVICE: replaces generate_distribution(distribution,
 * n, R, G) of tools/presets.py:91-1390 (the 25 laws of tools/presets.py:23-49; an unknown name gives the sphere,
 * like the reference's final else).  Body i is drawn from Philox4x32-10 with counter (i, draw, stream, 0) and
 * key = seed: reproducible, independent of n, no per-body host loops.  Host outputs: pos (n,3), vel (n,3),
 * mass (n), fp64. */
int b200_generate_distribution(const char* distribution, int64_t n, double R, double G, uint64_t seed, int device,
                               double* pos, double* vel, double* mass);
/* create() with the initial state drawn straight into the handle's device buffers (no host arrays, no upload):
 * generate_distribution(distribution, n, R, G_dist) followed by create(..., G, softening, damping, theta). */
int b200_nbody_create_generated(const char* distribution, int64_t n, double R, double G_dist, uint64_t seed,
                                double G, double softening, double damping, double theta, int device, b200_nbody** out);
/* step: replaces CUDASimulation.step (nbody/gpu_backend.py:368-386): one force evaluation +
 * kick-drift.  Asynchronous with respect to the host. */
int b200_nbody_step(b200_nbody* h, double dt);
int b200_nbody_step_n(b200_nbody* h, double dt, int nsteps);
/* Barnes-Hut accelerations of the current state, creation order, (n,3) fp32, no integrate:
 * the device twin of compute_forces_barnes_hut (nbody/simulation.py:201-278). */
int b200_nbody_compute_accelerations(b200_nbody* h, float* out);
/* compute_colors / get_* / sync: nbody/gpu_backend.py:388-409.  get_* return creation order. */
int b200_nbody_compute_colors(b200_nbody* h, double max_speed);
int b200_nbody_get_positions(b200_nbody* h, float* out);     /* (n,3) fp32 */
int b200_nbody_get_positions_f64(b200_nbody* h, double* out);/* (n,3) fp64 master state */
int b200_nbody_get_velocities(b200_nbody* h, double* out);   /* (n,3) fp64 */
int b200_nbody_get_colors(b200_nbody* h, float* out);        /* (n,3) fp32 */
int b200_nbody_sync(b200_nbody* h);
/* Replace positions and velocities (creation order, fp64) keeping masses: the restore path
 * of tools/record.py:718-735 without rebuilding the object. */
int b200_nbody_set_state(b200_nbody* h, const double* pos, const double* vel);
int b200_nbody_set_params(b200_nbody* h, double G, double softening, double damping, double theta);
/* Asynchronous frame egress (replaces the per-frame compute_colors + get_positions +
 * get_colors of tools/record.py:826-832): colours and creation-order fp32 positions are produced on the
 * handle's stream, the two device-to-host copies run on a second stream while later steps compute.
 * pos_out / col_out: (n,3) fp32 host buffers (pinned for a truly asynchronous copy) that must stay
 * untouched until frame_wait returns.  At most one frame is in flight. */
int b200_nbody_frame_begin(b200_nbody* h, double max_speed, float* pos_out, float* col_out);
int b200_nbody_frame_wait(b200_nbody* h);
/* Delta frame for the recorder's on-disk format 2 (the arithmetic of compress_frame,
 * tools/record.py:254-262): int16((frame - previous frame) * 1000) per component, float32 arithmetic, for
 * positions and colours in creation order; "previous frame" = the frame of the last frame_begin /
 * frame_delta_begin.  Half the device-to-host bytes of frame_begin.  Completed by frame_wait. */
int b200_nbody_frame_delta_begin(b200_nbody* h, double max_speed, int16_t* pos_delta_out, int16_t* col_delta_out);
/* Live-viewer frame (replaces, per displayed frame, the compute_colors + get_positions + get_colors of
 * NBodySimulation._update_gpu (nbody/simulation.py:809-817), the CPU frustum test compute_visibility_points (:403-434,
 * called from _compute_visibility :880-904) and the boolean-mask gathers of draw() (:926-927)): colours and the
 * creation-order fp32 frame are produced on the device, every body is tested against the view frustum there (the
 * reference's fp64 arithmetic on float32 positions, margin 1.2, near 0.1), and the visible bodies are compacted IN
 * CREATION ORDER (what positions[mask] gives).  camera = 15 doubles: cam_pos[3], cam_forward[3], cam_right[3],
 * cam_up[3], tan(half_fov_h), tan(half_fov_v), far distance (fog_end).  pos_out / col_out: host buffers with room
 * for (n,3) fp32; *count = number of visible bodies written.  Blocking. */
int b200_nbody_visible_frame(b200_nbody* h, double max_speed, const double* camera, float* pos_out, float* col_out, int64_t* count);
/* Same, but the visible bodies are written to DEVICE memory ((n,3) fp32 each) -- e.g. the two vertex buffer objects
 * of nbody/simulation.py:936-937 mapped with cudaGraphicsResourceGetMappedPointer -- so a displayed frame moves only
 * the 8-byte count over PCIe. */
int b200_nbody_visible_frame_device(b200_nbody* h, double max_speed, const double* camera, void* pos_device, void* col_device,
                                    int64_t* count);
/* Asynchronous set_state: begin starts the host-to-device copies on a third stream into staging;
 * commit waits for them, then makes them the current state on the handle's stream; after commit the
 * host arrays may be reused.  Started one step ahead, the copy overlaps the previous step's kernels. */
int b200_nbody_set_state_begin(b200_nbody* h, const double* pos, const double* vel);
int b200_nbody_set_state_commit(b200_nbody* h);
/* Sharded host traffic (one process per GPU, every rank a replica): a rank uploads / downloads only rows
 * [row_begin, row_end) of the creation-order arrays (the pointers are the FULL arrays' bases).  After
 * set_state_begin_rows the caller makes the handle's stream wait for the copy (upload_wait), completes
 * the device staging buffers (upload_staging: (n + 64, 3) fp64 each, padded for equal slices) with an
 * all-gather over NVLink on that stream, then calls set_state_commit. */
int b200_nbody_set_state_begin_rows(b200_nbody* h, const double* pos, const double* vel, int64_t row_begin, int64_t row_end);
int b200_nbody_upload_staging(b200_nbody* h, void** pos_device_ptr, void** vel_device_ptr);
int b200_nbody_upload_wait(b200_nbody* h);
int b200_nbody_frame_begin_rows(b200_nbody* h, double max_speed, float* pos_out, float* col_out, int64_t row_begin, int64_t row_end);
/* Sorted 63-bit Morton keys of the current state and the sort permutation
 * (perm[k] = creation index of the body at sorted position k). */
int b200_nbody_get_keys(b200_nbody* h, uint64_t* out);
int b200_nbody_get_perm(b200_nbody* h, uint32_t* out);
int b200_nbody_get_stats(b200_nbody* h, b200_nbody_stats* out);
int b200_nbody_reset_stats(b200_nbody* h);
int b200_nbody_set_profiling(b200_nbody* h, int enabled);
/* Exact device-side interaction counting in step() (stats.interactions); off by default because
 * it costs a few instructions in the traversal's inner loop.  compute_accelerations always counts. */
int b200_nbody_set_counting(b200_nbody* h, int enabled);
/* Exact accepted-interaction count of one force pass over the handle's shard on the CURRENT state (tree
 * built if needed, nothing integrated, nothing copied out): what bench.py divides the timed traversal by. */
int b200_nbody_count_interactions(b200_nbody* h, int64_t* interactions);
/* Order-independent 64-bit checksums of the fp64 master state (positions, velocities), keyed by creation
 * index: equal on two handles iff they hold bit-identical states.  bench.py compares replicas with it. */
int b200_nbody_state_checksum(b200_nbody* h, uint64_t out[2]);
/* Runs nsteps steps and returns their device time (CUDA events on the handle's stream). */
int b200_nbody_timed_steps(b200_nbody* h, double dt, int nsteps, float* elapsed_ms);
/* Kernels launched by this handle so far (bench.py's gpu_launches). */
int b200_nbody_launch_count(b200_nbody* h, int64_t* out);

/* ---- several GPUs behind the C ABI -------------------------------------------------------------
 * The sharded step lives inside the library (csrc/multi.cu): every rank holds the full replicated state,
 * radix-sorts one slice of it (the sorted runs are all-gathered with NCCL, loaded at run time, and merged by
 * counting), builds the whole tree, and traverses + integrates only its own Morton range; the traversal
 * kernel stores the new positions and velocities of those bodies straight into the next-state buffers of
 * every rank over NVLink peer mappings, and one 8-byte all-reduce (the next bounds) is the barrier that ends
 * the step.  Replicas stay bit-identical to each other and to a single-GPU run.  At most 8 ranks.
 *
 * (a) one process, several GPUs -- SURVEY.md 8b's `device_mask`: bit d of the mask selects CUDA device d.  The
 *     handle behaves like a single-GPU one (getters and frames are served by the lowest device); this is how
 *     a caller without torch, e.g. the reference's recorder through create_gpu_simulation, uses N GPUs. */
int b200_nbody_create_multi(int64_t n, const double* pos, const double* vel, const double* mass,
                            double G, double softening, double damping, double theta,
                            uint32_t device_mask, b200_nbody** out);
/* (b) one process per GPU (torchrun): rank 0 draws 128 bytes (an ncclUniqueId), the caller carries them to
 *     every rank by any means, and every rank calls comm_init on its own handle (collective).  The other
 *     ranks' state buffers are mapped with CUDA IPC.  From then on step() is the sharded step, and every call
 *     that changes the state (step, set_state, compute_accelerations, ...) must be made by all ranks alike. */
int b200_nccl_unique_id(void* out128);
int b200_nbody_comm_init(b200_nbody* h, const void* id128, int rank, int world);
int b200_nbody_world(b200_nbody* h, int* world);
/* Sorted-position range [begin, end) this rank currently traverses and integrates.  Inside a group the ranges are
 * COST-WEIGHTED: the traversal leaves a per-body cost in the accelerations buffer, every 8 steps
 * (B200_REBALANCE=k; 0 = keep equal counts) the costs are summed per 4096-body chunk, all-reduced, and the
 * boundaries move to equal-cost chunk boundaries -- identical on every rank, results unchanged bit for bit. */
int b200_nbody_get_shard(b200_nbody* h, int64_t* begin, int64_t* end);
/* The split rule itself (pure host arithmetic, no device needed): world + 1 boundaries in sorted positions, multiples
 * of `chunk`, from the per-chunk costs; every rank keeps at least one chunk; equal chunk counts when all costs are 0. */
int b200_cost_weighted_split(const uint64_t* chunk_cost, int nchunks, int64_t chunk, int64_t n, int world, int64_t* split_out);

/* ---- split sharded step (building blocks; the collectives are the caller's) -------------------
 * One process per GPU, every rank holds the full replicated state and builds the full tree;
 * a rank traverses only sorted bodies [begin, end) (begin a multiple of 32).  Between
 * step_begin and step_end the host-side plumbing (torch.distributed / NCCL) all-gathers the
 * accelerations buffer slices on the SAME stream (set_stream), then every rank integrates all
 * bodies, which keeps the replicas bit-identical.  No reference counterpart. */
/* external != 0: run all work of the handle on cuda_stream (a cudaStream_t; 0 is the legacy
 * default stream); external == 0: back to the handle's own non-blocking stream. */
int b200_nbody_set_stream(b200_nbody* h, void* cuda_stream, int external);
int b200_nbody_set_shard(b200_nbody* h, int64_t begin, int64_t end);
int b200_nbody_step_begin(b200_nbody* h);
int b200_nbody_step_end(b200_nbody* h, double dt);
/* Sharded sort (optional, valid from the second step after an upload: the state is then physically in
 * last step's Morton order, so equal slices of it are nearly disjoint key ranges).  Rank r generates
 * keys for and sorts only the bodies at current positions [r*slice, (r+1)*slice) into its part of two
 * padded exchange buffers (world*slice uint64 keys, world*slice uint32 local positions); the host
 * plumbing all-gathers both on the same stream; step_begin_sorted merges the world sorted runs by
 * counting (bit-identical to the single-GPU stable sort) and continues like step_begin. */
int b200_nbody_sharded_sort_setup(b200_nbody* h, int64_t slice, int world, void** keys_device_ptr, void** vals_device_ptr);
int b200_nbody_sort_local(b200_nbody* h, int rank);
int b200_nbody_step_begin_sorted(b200_nbody* h);
/* Device address and capacity (in float4 entries of 16 bytes) of the sorted-order
 * accelerations buffer {ax, ay, az, interaction count}. */
int b200_nbody_acc_buffer(b200_nbody* h, void** device_ptr, int64_t* capacity_entries);

/* Measured FP32 FFMA throughput of the device (TFLOP/s): the traversal's roofline denominator. */
int b200_fp32_peak_tflops(int device, double* tflops);

/* Page-locked host memory for the asynchronous entry points (frame_begin, frame_delta_begin, set_state_begin):
 * with pageable buffers -- what numpy.empty gives the reference's recorder (tools/record.py:828-829) -- the
 * copies are staged by the driver and do not overlap the next step.  A caller without a CUDA binding of its own
 * gets pinned buffers here; free with b200_host_free. */
int b200_host_alloc(int64_t bytes, void** out);
int b200_host_free(void* ptr);

/* ---- boids ---------------------------------------------------------------------------------
 * The reference has no backend layer for boids; the seam is Flock.update(dt)
 * (boids/flock.py:627-678) mutating positions / velocities / colors (n,3) fp64.
 * Parameters: config/boids.py:30-46 (same names). Grid: boids/flock.py:478-481. */
typedef struct b200_boids_params {
    double bounds, max_speed, max_force, wall_margin, wall_weight;
    double perception_radius, separation_radius;
    double separation_weight, alignment_weight, cohesion_weight, color_blend_rate;
} b200_boids_params;

#define B200_BOIDS_PHASES 5
/* order: cells, sort, gather, table, rules(+physics) */
typedef struct b200_boids_stats {
    int64_t n, steps, num_cells;
    int32_t grid_dim, key_bits;
    double  cell_size, grid_offset;
    int64_t neighbor_pairs;      /* accepted (i, j) neighbour pairs since the last reset */
    int64_t bytes_allocated;
    int64_t launches;
    int64_t timed_steps;
    double  phase_ms[B200_BOIDS_PHASES];
} b200_boids_stats;

int b200_boids_create(int64_t n, const double* pos, const double* vel, const double* col,
                      const b200_boids_params* params, int device, b200_boids** out);
int b200_boids_destroy(b200_boids* h);
/* one Flock.update(dt): grid build + rules + physics, state stays on the device */
int b200_boids_step(b200_boids* h, double dt);
/* creation order; any of the three pointers may be NULL to skip that array */
int b200_boids_get_state(b200_boids* h, double* pos, double* vel, double* col);
int b200_boids_set_state(b200_boids* h, const double* pos, const double* vel, const double* col);
/* cell index of every boid of the current state (assign_cells, boids/flock.py:30-44) */
int b200_boids_get_cell_indices(b200_boids* h, int32_t* out);
int b200_boids_get_stats(b200_boids* h, b200_boids_stats* out);
int b200_boids_reset_stats(b200_boids* h);
int b200_boids_set_profiling(b200_boids* h, int enabled);
int b200_boids_timed_steps(b200_boids* h, double dt, int nsteps, float* elapsed_ms);
int b200_boids_sync(b200_boids* h);

#ifdef __cplusplus
}
#endif
#endif /* B200SIM_H */
