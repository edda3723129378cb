// boids.cu -- uniform-grid cell sort + fused separation/alignment/cohesion + physics (sm_100a).
// fp64 throughout, the reference's arithmetic (boids/flock.py); see boids.cuh for the layout.
#include "boids.cuh"

namespace b200 {

// boids/flock.py:16-27: int() truncates toward zero, then clamp to [0, dim-1]
__device__ __forceinline__ int cell_coord(double p, double offset, double cell, int dim)
{
    const double q = (p + offset) / cell;
    int c = (q >= 2147483647.0) ? 2147483647 : (q <= -2147483648.0 ? (-2147483647 - 1) : (int)q);
    c = min(c, dim - 1);
    return max(c, 0);
}

__global__ void __launch_bounds__(256) boids_cells_kernel(const double* __restrict__ pos, int n, double offset, double cell,
                                                          int dim, uint32_t* __restrict__ keys)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t o = 3 * (int64_t)i;
    const int cx = cell_coord(pos[o], offset, cell, dim);
    const int cy = cell_coord(pos[o + 1], offset, cell, dim);
    const int cz = cell_coord(pos[o + 2], offset, cell, dim);
    keys[i] = (uint32_t)(cx + cy * dim + cz * dim * dim);   // x fastest (boids/flock.py:27)
}

constexpr int TABLE_CHUNK_LOG2 = 12;
constexpr int TABLE_CHUNK = 1 << TABLE_CHUNK_LOG2;   // cells of the table built by one CTA
constexpr int TABLE_PER_THREAD = TABLE_CHUNK / 256;

__global__ void __launch_bounds__(256) boids_gather_kernel(const uint32_t* __restrict__ perm,
                                                           const double* __restrict__ pos_in, const double* __restrict__ vel_in,
                                                           const double* __restrict__ col_in, const uint32_t* __restrict__ id_in,
                                                           double* __restrict__ pos_out, double* __restrict__ vel_out,
                                                           double* __restrict__ col_out, uint32_t* __restrict__ id_out, int n,
                                                           const uint32_t* __restrict__ sorted_keys, int nchunks,
                                                           int* __restrict__ chunk_lb)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    // lower bound of every TABLE_CHUNK-cell chunk of the cell table (boids_table_kernel), from the
    // boundaries of the sorted keys: chunk_lb[b] = first sorted position whose cell is in chunk >= b
    {
        const int cb = (int)(sorted_keys[k] >> TABLE_CHUNK_LOG2);
        const int pb = k > 0 ? (int)(sorted_keys[k - 1] >> TABLE_CHUNK_LOG2) : -1;
        for (int b = pb + 1; b <= cb; ++b) chunk_lb[b] = k;
        if (k == n - 1)
            for (int b = cb + 1; b <= nchunks; ++b) chunk_lb[b] = n;
    }
    const int64_t j = 3 * (int64_t)perm[k], o = 3 * (int64_t)k;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        pos_out[o + d] = pos_in[j + d];
        vel_out[o + d] = vel_in[j + d];
        col_out[o + d] = col_in[j + d];
    }
    id_out[k] = id_in[perm[k]];
}

// build_cell_lists (boids/flock.py:47-65) as ONE lower-bound table: first[c] = sorted position of the first
// boid whose cell index is >= c, for every cell c in [0, C] (first[C] = n).  The boids of the cells
// [a, b] (consecutive indices = a grid row segment) are then the run [first[a], first[b + 1]): two loads
// per row instead of a start and an end per cell, and no memset.  A CTA owns TABLE_CHUNK consecutive
// cells: its boids are the run between two chunk lower bounds (written by the gather kernel at the
// chunk boundaries of the sorted keys); it marks the cell starts in shared memory and finishes with a
// suffix-min scan seeded by the next chunk's lower bound.

__global__ void __launch_bounds__(256) boids_table_kernel(const uint32_t* __restrict__ sorted_keys, int n, int64_t num_cells,
                                                          const int* __restrict__ chunk_lb, int* __restrict__ first)
{
    __shared__ int sf[TABLE_CHUNK];
    __shared__ int s_warp[8];
    const int64_t c0 = (int64_t)blockIdx.x * TABLE_CHUNK;
    const int tid = threadIdx.x;
    const int kb = chunk_lb[blockIdx.x], ke = chunk_lb[blockIdx.x + 1];   // this chunk's boids (boids_gather_kernel)
    for (int i = tid; i < TABLE_CHUNK; i += 256) sf[i] = 0x7fffffff;
    __syncthreads();
    for (int k = kb + tid; k < ke; k += 256) {
        const uint32_t c = sorted_keys[k];
        if (k == 0 || sorted_keys[k - 1] != c) sf[(int)((int64_t)c - c0)] = k;
    }
    __syncthreads();
    // suffix-min: thread t owns cells [t*16, t*16+16)
    int loc[TABLE_PER_THREAD];
    int run = 0x7fffffff;
#pragma unroll
    for (int i = TABLE_PER_THREAD - 1; i >= 0; --i) {
        run = min(run, sf[tid * TABLE_PER_THREAD + i]);
        loc[i] = run;
    }
    // exclusive suffix-min of the threads' totals: later lanes, later warps, then the next chunk's bound
    int inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_down_sync(0xffffffffu, inc, o);
        if ((int)lane_id() + o < 32) inc = min(inc, v);
    }
    if (lane_id() == 0) s_warp[tid >> 5] = inc;
    int after = __shfl_down_sync(0xffffffffu, inc, 1);
    if (lane_id() == 31) after = 0x7fffffff;
    __syncthreads();
    for (int w = (tid >> 5) + 1; w < 8; ++w) after = min(after, s_warp[w]);
    after = min(after, ke);
#pragma unroll
    for (int i = 0; i < TABLE_PER_THREAD; ++i) {
        const int64_t c = c0 + tid * TABLE_PER_THREAD + i;
        if (c <= num_cells) first[c] = min(loc[i], after);
    }
}

// boids/flock.py:179-193 / :200-214 / :220-234: normalise to max_speed, subtract own velocity,
// clamp to max_force, weight.  Returns zero force when the mean vector is exactly zero.
__device__ __forceinline__ void steer(double x, double y, double z, double vx, double vy, double vz, double max_speed,
                                      double max_force, double weight, double& fx, double& fy, double& fz)
{
    double mag = sqrt(x * x + y * y + z * z);
    fx = fy = fz = 0.0;
    if (mag > 0.0) {
        double sx = (x / mag) * max_speed - vx;
        double sy = (y / mag) * max_speed - vy;
        double sz = (z / mag) * max_speed - vz;
        mag = sqrt(sx * sx + sy * sy + sz * sz);
        if (mag > max_force) {
            sx = (sx / mag) * max_force;
            sy = (sy / mag) * max_force;
            sz = (sz / mag) * max_force;
        }
        fx = sx * weight; fy = sy * weight; fz = sz * weight;
    }
}

// compute_flocking_spatial (boids/flock.py:68-238) fused with update_physics_numba (:241-308).
// One thread per boid in cell order; the (2R+1) cells of a grid row are consecutive cell indices
// (x fastest), so their boids are ONE contiguous run of the sorted state.
template <int RT>   // RT = 1: perception radius <= cell size (the reference's default grid), 3 x 3 rows unrolled with all
                    // table loads issued up front; RT = 0: any cell range R
__global__ void __launch_bounds__(128, 6) boids_rules_kernel(
    const double* __restrict__ pos_in, const double* __restrict__ vel_in, const double* __restrict__ col_in,
    const uint32_t* __restrict__ id_in, const int* __restrict__ cell_first,
    double* __restrict__ pos_out, double* __restrict__ vel_out, double* __restrict__ col_out, uint32_t* __restrict__ id_out,
    int n, BoidsParams P, double offset, double cell, int dim, int R, double dt, double blend,
    unsigned long long* __restrict__ pairs)
{
    __shared__ int2 runs[RT == 1 ? 9 : 1][128];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    int nnb = 0;
    if (k < n) {
        const int64_t o = 3 * (int64_t)k;
        const double px = pos_in[o], py = pos_in[o + 1], pz = pos_in[o + 2];
        double vx = vel_in[o], vy = vel_in[o + 1], vz = vel_in[o + 2];
        const double c0 = col_in[o], c1 = col_in[o + 1], c2 = col_in[o + 2];
        const int cx = cell_coord(px, offset, cell, dim);
        const int cy = cell_coord(py, offset, cell, dim);
        const int cz = cell_coord(pz, offset, cell, dim);
        const double per2 = P.perception_radius * P.perception_radius;
        const double sep2 = P.separation_radius * P.separation_radius;
        double sx = 0, sy = 0, sz = 0, ax = 0, ay = 0, az = 0, hx = 0, hy = 0, hz = 0, cr = 0, cg = 0, cb = 0;
        int nsep = 0;
        const int x0 = max(cx - R, 0), x1 = min(cx + R, dim - 1);
        auto scan_run = [&](int s, int e) {
            for (int j = s; j < e; ++j) {
                if (j == k) continue;
                const int64_t q = 3 * (int64_t)j;
                const double dx = px - pos_in[q], dy = py - pos_in[q + 1], dz = pz - pos_in[q + 2];
                const double d2 = dx * dx + dy * dy + dz * dz;
                if (d2 < per2 && d2 > 0.0001) {                 // :150
                    if (d2 < sep2) {                            // :153-158
                        const double d = sqrt(d2);
                        const double inv = 1.0 / d;
                        sx += dx * inv / d; sy += dy * inv / d; sz += dz * inv / d;
                        ++nsep;
                    }
                    ax += vel_in[q]; ay += vel_in[q + 1]; az += vel_in[q + 2];
                    hx += pos_in[q]; hy += pos_in[q + 1]; hz += pos_in[q + 2];
                    cr += col_in[q]; cg += col_in[q + 1]; cb += col_in[q + 2];
                    ++nnb;
                }
            }
        };
        if (RT == 1) {
            // same visiting order as the loops below (z outer, y inner); rows outside the grid are empty runs
            // (the 18 independent table loads are issued back to back; the runs are parked in shared memory so
            // the scan stays one compact loop)
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const int ncz = cz - 1 + i / 3, ncy = cy - 1 + i % 3;
                const bool ok = ncz >= 0 && ncz < dim && ncy >= 0 && ncy < dim;
                const int64_t row = (int64_t)ncy * dim + (int64_t)ncz * dim * dim;
                runs[i][threadIdx.x] = ok ? make_int2(cell_first[row + x0], cell_first[row + x1 + 1]) : make_int2(0, 0);
            }
#pragma unroll 1
            for (int i = 0; i < 9; ++i) {
                const int2 r = runs[i][threadIdx.x];
                scan_run(r.x, r.y);
            }
        } else {
            for (int ncz = max(cz - R, 0); ncz <= min(cz + R, dim - 1); ++ncz) {
                for (int ncy = max(cy - R, 0); ncy <= min(cy + R, dim - 1); ++ncy) {
                    const int64_t row = (int64_t)ncy * dim + (int64_t)ncz * dim * dim;
                    scan_run(cell_first[row + x0], cell_first[row + x1 + 1]);   // the row segment is one run
                }
            }
        }
        double f[3] = {0.0, 0.0, 0.0};
        double avg0 = c0, avg1 = c1, avg2 = c2;                     // avg_colors <- colors (:636)
        double fx, fy, fz;
        if (nsep > 0) {                                             // :174-193
            steer(sx / nsep, sy / nsep, sz / nsep, vx, vy, vz, P.max_speed, P.max_force, P.separation_weight, fx, fy, fz);
            f[0] = fx; f[1] = fy; f[2] = fz;
        }
        if (nnb > 0) {                                              // :195-238
            double gx, gy, gz;
            steer(ax / nnb, ay / nnb, az / nnb, vx, vy, vz, P.max_speed, P.max_force, P.alignment_weight, gx, gy, gz);
            // update_physics sums sep + align + coh in this order (:260-262)
            f[0] += gx; f[1] += gy; f[2] += gz;
            steer(hx / nnb - px, hy / nnb - py, hz / nnb - pz, vx, vy, vz, P.max_speed, P.max_force, P.cohesion_weight, gx, gy, gz);
            f[0] += gx; f[1] += gy; f[2] += gz;
            avg0 = (cr + c0) / (nnb + 1); avg1 = (cg + c1) / (nnb + 1); avg2 = (cb + c2) / (nnb + 1);
        }
        // ---- physics (:259-308): soft walls, integrate, speed clamp, colour blend
        const double wall_force = P.max_force * P.wall_weight;      // :673
        const double p3[3] = {px, py, pz};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double dp = p3[d] - (P.bounds - P.wall_margin);
            if (dp > 0) f[d] -= fmin(dp / P.wall_margin * 2.0, 1.0) * wall_force;
            const double dn = (-P.bounds + P.wall_margin) - p3[d];
            if (dn > 0) f[d] += fmin(dn / P.wall_margin * 2.0, 1.0) * wall_force;
        }
        vx += f[0] * dt; vy += f[1] * dt; vz += f[2] * dt;
        const double speed = sqrt(vx * vx + vy * vy + vz * vz);
        if (speed > P.max_speed) {
            const double sc = P.max_speed / speed;
            vx *= sc; vy *= sc; vz *= sc;
        }
        vel_out[o] = vx; vel_out[o + 1] = vy; vel_out[o + 2] = vz;
        pos_out[o] = px + vx * dt; pos_out[o + 1] = py + vy * dt; pos_out[o + 2] = pz + vz * dt;
        col_out[o] = c0 + (avg0 - c0) * blend;
        col_out[o + 1] = c1 + (avg1 - c1) * blend;
        col_out[o + 2] = c2 + (avg2 - c2) * blend;
        id_out[k] = id_in[k];
    }
    unsigned c32 = (unsigned)nnb;
    for (int o = 16; o > 0; o >>= 1) c32 += __shfl_xor_sync(0xffffffffu, c32, o);
    if (lane_id() == 0 && c32) atomicAdd(pairs, (unsigned long long)c32);
}

template <typename T>
__global__ void __launch_bounds__(256) boids_unpermute_kernel(const T* __restrict__ src, const uint32_t* __restrict__ id,
                                                              T* __restrict__ dst, int n, int width)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t o = (int64_t)width * k, w = (int64_t)width * id[k];
    for (int d = 0; d < width; ++d) dst[w + d] = src[o + d];
}

__global__ void __launch_bounds__(256) boids_iota_kernel(uint32_t* __restrict__ p, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}

// ============================================================================ host side
template <typename T>
static T* balloc(BoidsSim& s, size_t count)
{
    s.bytes_allocated += (count ? count : 1) * sizeof(T);
    return dev_alloc<T>(count);
}

void boids_alloc(BoidsSim& s, int n)
{
    s.n = n;
    B200_CHECK(cudaSetDevice(s.device));
    cudaDeviceProp prop;
    B200_CHECK(cudaGetDeviceProperties(&prop, s.device));
    s.sm_count = prop.multiProcessorCount;
    B200_CHECK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    // boids/flock.py:478-481
    s.cell_size = s.p.perception_radius;
    s.grid_dim = (int)ceil(s.p.bounds * 2.0 / s.cell_size) + 2;
    s.num_cells = (int64_t)s.grid_dim * s.grid_dim * s.grid_dim;
    s.grid_offset = s.p.bounds + s.cell_size;
    s.cell_range = (int)ceil(s.p.perception_radius / s.cell_size);   // :95
    B200_REQUIRE(s.grid_dim >= 1 && s.num_cells < ((int64_t)1 << 31), "boids grid too large (grid_dim^3 must fit int32)");
    s.key_bits = 1;
    while (((int64_t)1 << s.key_bits) < s.num_cells) ++s.key_bits;
    const size_t N = (size_t)n;
    for (int b = 0; b < 2; ++b) {
        s.pos[b] = balloc<double>(s, 3 * N);
        s.vel[b] = balloc<double>(s, 3 * N);
        s.col[b] = balloc<double>(s, 3 * N);
        s.id[b] = balloc<uint32_t>(s, N);
        s.keys[b] = balloc<uint32_t>(s, N);
        s.vals[b] = balloc<uint32_t>(s, N);
    }
    s.sorter.init(n);
    s.bytes_allocated += s.sorter.bytes();
    s.cell_first = balloc<int>(s, (size_t)s.num_cells + 1);
    s.table_chunks = (int)((s.num_cells + 1 + TABLE_CHUNK - 1) / TABLE_CHUNK);
    s.chunk_lb = balloc<int>(s, (size_t)s.table_chunks + 1);
    s.d_pairs = balloc<unsigned long long>(s, 1);
    s.stage = balloc<double>(s, 3 * N);
    B200_CHECK(cudaMemset(s.d_pairs, 0, sizeof(unsigned long long)));
    s.timer.init();
}

static void boids_graph_reset(BoidsSim& s);

void boids_free(BoidsSim& s)
{
    cudaSetDevice(s.device);
    if (s.stream) cudaStreamSynchronize(s.stream);
    for (int b = 0; b < 2; ++b) {
        cudaFree(s.pos[b]); cudaFree(s.vel[b]); cudaFree(s.col[b]); cudaFree(s.id[b]);
        cudaFree(s.keys[b]); cudaFree(s.vals[b]);
    }
    s.sorter.destroy();
    boids_graph_reset(s);
    cudaFree(s.cell_first); cudaFree(s.chunk_lb); cudaFree(s.d_pairs); cudaFree(s.stage);
    s.timer.destroy();
    if (s.stream) cudaStreamDestroy(s.stream);
    s.stream = nullptr;
}

void boids_upload(BoidsSim& s, const double* pos, const double* vel, const double* col)
{
    B200_CHECK(cudaSetDevice(s.device));
    const size_t B = 3 * (size_t)s.n * sizeof(double);
    s.cur = 0;
    B200_CHECK(cudaMemcpyAsync(s.pos[0], pos, B, cudaMemcpyHostToDevice, s.stream));
    B200_CHECK(cudaMemcpyAsync(s.vel[0], vel, B, cudaMemcpyHostToDevice, s.stream));
    B200_CHECK(cudaMemcpyAsync(s.col[0], col, B, cudaMemcpyHostToDevice, s.stream));
    if (s.n > 0) {
        boids_iota_kernel<<<div_up(s.n, 256), 256, 0, s.stream>>>(s.id[0], s.n);
        ++s.launches;
    }
    B200_CHECK(cudaStreamSynchronize(s.stream));
}

// one step's kernels on the stream (captured into a CUDA graph by boids_step)
static void boids_enqueue(BoidsSim& s, double dt)
{
    const int n = s.n;
    cudaStream_t st = s.stream;
    const int grid = div_up(n, 256);
    s.timer.begin(st);
    boids_cells_kernel<<<grid, 256, 0, st>>>(s.pos[s.cur], n, s.grid_offset, s.cell_size, s.grid_dim, s.keys[0]);
    ++s.launches;
    s.timer.mark(st);
    const int slot = s.sorter.sort(s.keys, s.vals, 0, n, 0, s.key_bits, /*iota=*/true, st, s.sm_count);
    s.launches += s.sorter.last_launches;
    s.timer.mark(st);
    const int o = s.cur ^ 1;
    boids_gather_kernel<<<grid, 256, 0, st>>>(s.vals[slot], s.pos[s.cur], s.vel[s.cur], s.col[s.cur], s.id[s.cur],
                                              s.pos[o], s.vel[o], s.col[o], s.id[o], n, s.keys[slot], s.table_chunks, s.chunk_lb);
    ++s.launches;
    s.timer.mark(st);
    boids_table_kernel<<<s.table_chunks, 256, 0, st>>>(s.keys[slot], n, s.num_cells, s.chunk_lb, s.cell_first);
    ++s.launches;
    s.timer.mark(st);
    const double blend = fmin(1.0, s.p.color_blend_rate * dt);   // boids/flock.py:662
    if (s.cell_range == 1)
        boids_rules_kernel<1><<<div_up(n, 128), 128, 0, st>>>(s.pos[o], s.vel[o], s.col[o], s.id[o], s.cell_first,
                                                              s.pos[s.cur], s.vel[s.cur], s.col[s.cur], s.id[s.cur], n, s.p,
                                                              s.grid_offset, s.cell_size, s.grid_dim, s.cell_range, dt, blend,
                                                              s.d_pairs);
    else
        boids_rules_kernel<0><<<div_up(n, 128), 128, 0, st>>>(s.pos[o], s.vel[o], s.col[o], s.id[o], s.cell_first,
                                                              s.pos[s.cur], s.vel[s.cur], s.col[s.cur], s.id[s.cur], n, s.p,
                                                              s.grid_offset, s.cell_size, s.grid_dim, s.cell_range, dt, blend,
                                                              s.d_pairs);
    ++s.launches;
    B200_CHECK(cudaGetLastError());
    s.timer.mark(st);
}

static void boids_graph_reset(BoidsSim& s)
{
    if (s.graph_exec) { cudaGraphExecDestroy(s.graph_exec); s.graph_exec = nullptr; }
}

// The step is ~14 launches and memsets of a few tens of microseconds each: launch-bound.  The same
// buffers are read and written every step (the rules kernel writes back into the current buffer), so the
// whole step is captured ONCE into a CUDA graph and replayed; a new dt re-captures (dt and the colour
// blend are kernel arguments).  Per-phase profiling runs the plain launches.
void boids_step(BoidsSim& s, double dt)
{
    B200_CHECK(cudaSetDevice(s.device));
    if (s.n == 0) { ++s.steps; return; }
    cudaStream_t st = s.stream;
    if (s.timer.enabled || !s.use_graph) {
        boids_enqueue(s, dt);
    } else {
        if (!s.graph_exec || s.graph_dt != dt) {
            boids_graph_reset(s);
            const int64_t before = s.launches;
            cudaGraph_t g = nullptr;
            B200_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            try {
                boids_enqueue(s, dt);
            } catch (...) {
                cudaStreamEndCapture(st, &g);
                if (g) cudaGraphDestroy(g);
                throw;
            }
            B200_CHECK(cudaStreamEndCapture(st, &g));
            const cudaError_t e = cudaGraphInstantiate(&s.graph_exec, g, 0);
            cudaGraphDestroy(g);
            B200_CHECK(e);
            s.graph_dt = dt;
            s.graph_launches = s.launches - before;
            s.launches = before;
        }
        B200_CHECK(cudaGraphLaunch(s.graph_exec, st));
        s.launches += s.graph_launches;
    }
    ++s.steps;
    if (s.timer.enabled) {
        B200_CHECK(cudaStreamSynchronize(st));
        s.timer.collect();
    }
}

void boids_get_state(BoidsSim& s, double* pos, double* vel, double* col)
{
    B200_CHECK(cudaSetDevice(s.device));
    if (s.n == 0) return;
    const size_t B = 3 * (size_t)s.n * sizeof(double);
    const double* src[3] = {s.pos[s.cur], s.vel[s.cur], s.col[s.cur]};
    double* dst[3] = {pos, vel, col};
    for (int a = 0; a < 3; ++a) {
        if (!dst[a]) continue;
        boids_unpermute_kernel<double><<<div_up(s.n, 256), 256, 0, s.stream>>>(src[a], s.id[s.cur], s.stage, s.n, 3);
        ++s.launches;
        B200_CHECK(cudaMemcpyAsync(dst[a], s.stage, B, cudaMemcpyDeviceToHost, s.stream));
        B200_CHECK(cudaStreamSynchronize(s.stream));
    }
}

void boids_get_cells(BoidsSim& s, int32_t* out)
{
    B200_CHECK(cudaSetDevice(s.device));
    if (s.n == 0) return;
    // cell index of every boid of the CURRENT state, creation order (keys[1]/vals[1] are free between steps)
    boids_cells_kernel<<<div_up(s.n, 256), 256, 0, s.stream>>>(s.pos[s.cur], s.n, s.grid_offset, s.cell_size, s.grid_dim, s.keys[1]);
    boids_unpermute_kernel<uint32_t><<<div_up(s.n, 256), 256, 0, s.stream>>>(s.keys[1], s.id[s.cur], s.vals[1], s.n, 1);
    s.launches += 2;
    B200_CHECK(cudaMemcpyAsync(out, s.vals[1], (size_t)s.n * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
    B200_CHECK(cudaStreamSynchronize(s.stream));
}

}  // namespace b200
