// boids.cuh -- device state of the boids neighbour-rule update (sm_100a).
//
// Replaces Flock.update (boids/flock.py:627-678): assign_cells (:30-44), argsort (:618),
// build_cell_lists (:47-65), compute_flocking_spatial (:68-238), update_physics_numba (:241-308).
//
// Data layout in HBM (N boids, C = grid_dim^3 cells):
//   state[2]  pos, vel, col (N,3) f64 + id (N) u32, double-buffered.  Each step the state is
//             physically reordered into cell order (gather A->B) and the fused rules+physics
//             kernel reads B (neighbours are contiguous runs) and writes the updated state to A.
//   keys[2]/vals[2]  cell index per boid (u32) + permutation, radix-sort ping-pong
//   cell_first (C + 1) i32: sorted position of the first boid whose cell index is >= c (lower-bound table)
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"

namespace b200 {

struct BoidsParams {   // config/boids.py:30-46
    double bounds, max_speed, max_force, wall_margin, wall_weight;
    double perception_radius, separation_radius;
    double separation_weight, alignment_weight, cohesion_weight, color_blend_rate;
};

enum BoidsPhase { BP_CELLS = 0, BP_SORT, BP_GATHER, BP_TABLE, BP_RULES, BP_COUNT };

struct BoidsSim {
    int n = 0;
    int device = 0;
    int sm_count = 148;
    BoidsParams p{};
    double cell_size = 0, grid_offset = 0;   // boids/flock.py:478-481
    int grid_dim = 0;
    int64_t num_cells = 0;
    int key_bits = 0;
    int cell_range = 1;
    cudaStream_t stream = nullptr;

    double* pos[2] = {nullptr, nullptr};
    double* vel[2] = {nullptr, nullptr};
    double* col[2] = {nullptr, nullptr};
    uint32_t* id[2] = {nullptr, nullptr};
    int cur = 0;
    uint32_t* keys[2] = {nullptr, nullptr};
    uint32_t* vals[2] = {nullptr, nullptr};
    rsort::Sorter<uint32_t> sorter;
    int* cell_first = nullptr;
    int* chunk_lb = nullptr;     // (table_chunks + 1) lower bound of every 4096-cell chunk of the table
    int table_chunks = 0;
    unsigned long long* d_pairs = nullptr;   // accepted neighbour pairs (device-counted)
    double* stage = nullptr;

    cudaGraphExec_t graph_exec = nullptr;   // the captured step (boids_step)
    double graph_dt = 0.0;
    int64_t graph_launches = 0;
    bool use_graph = true;

    PhaseTimer timer;
    int64_t steps = 0;
    int64_t launches = 0;
    size_t bytes_allocated = 0;
};

void boids_alloc(BoidsSim& s, int n);
void boids_free(BoidsSim& s);
void boids_upload(BoidsSim& s, const double* pos, const double* vel, const double* col);
void boids_step(BoidsSim& s, double dt);
void boids_get_state(BoidsSim& s, double* pos, double* vel, double* col);
void boids_get_cells(BoidsSim& s, int32_t* out);   // creation order, current state

}  // namespace b200
