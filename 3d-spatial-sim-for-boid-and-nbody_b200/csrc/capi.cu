// capi.cu -- extern "C" boundary of libb200sim.so (see include/b200sim.h).
#include "../../include/b200sim.h"
#include "nbody.cuh"
#include "boids.cuh"

#include <string.h>

namespace b200 {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
}  // namespace b200

struct b200_nbody {
    b200::NBodySim sim;                        // the (primary) replica: getters, frames, stats
    std::vector<b200::NBodySim*> others;       // further replicas of a device-mask handle (one process, several GPUs)
    b200::NBodyGroup* group = nullptr;         // set by comm_init / create_multi: step() is the sharded fused step
};

// device-mask handles: a getter that rebuilds the tree reorders the state of the replica it runs on; the other
// replicas must follow, or the ranks would disagree on which state buffer is current
static void rebuild_other_replicas(b200_nbody* h)
{
    for (b200::NBodySim* o : h->others)
        if (!o->tree_valid) b200::nbody_build_tree(*o);
}

static void step_any(b200_nbody* h, double dt)
{
    if (h->group) b200::group_step(*h->group, dt);
    else b200::nbody_step(h->sim, dt);
}
struct b200_boids {
    b200::BoidsSim sim;
};
static_assert(sizeof(b200_boids_params) == sizeof(b200::BoidsParams), "boids params layout");

#define B200_API extern "C" __attribute__((visibility("default")))

#define B200_TRY(...)                                    \
    try {                                                \
        __VA_ARGS__;                                     \
        return B200_OK;                                  \
    } catch (const b200::CudaError& e) {                 \
        b200::set_error(e.msg);                          \
        return B200_ERR_CUDA;                            \
    } catch (const b200::StateError& e) {                \
        b200::set_error(e.msg);                          \
        return B200_ERR_STATE;                           \
    } catch (const std::exception& e) {                  \
        b200::set_error(e.what());                       \
        return B200_ERR_STATE;                           \
    }

// constructors: free what was built, then the same status mapping as B200_TRY (nothing may unwind through extern "C")
#define B200_CATCH_CREATE(cleanup)                       \
    catch (const b200::CudaError& e) {                   \
        b200::set_error(e.msg);                          \
        cleanup;                                         \
        return B200_ERR_CUDA;                            \
    } catch (const b200::StateError& e) {                \
        b200::set_error(e.msg);                          \
        cleanup;                                         \
        return B200_ERR_STATE;                           \
    } catch (const std::exception& e) {                  \
        b200::set_error(e.what());                       \
        cleanup;                                         \
        return B200_ERR_STATE;                           \
    }

#define B200_ARG(cond, text)                             \
    if (!(cond)) {                                       \
        b200::set_error(text);                           \
        return B200_ERR_ARG;                             \
    }

B200_API const char* b200_last_error(void) { return b200::g_last_error.c_str(); }

B200_API int b200_device_count(int* count)
{
    B200_ARG(count, "count is null");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        b200::set_error(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
        cudaGetLastError();
        return B200_ERR_CUDA;
    }
    return B200_OK;
}

B200_API int b200_device_info(int device, char* name, int name_len)
{
    B200_ARG(name && name_len > 0, "name buffer is null");
    B200_TRY({
        cudaDeviceProp p;
        B200_CHECK(cudaGetDeviceProperties(&p, device));
        snprintf(name, (size_t)name_len, "%s (CC %d.%d, %dGB)", p.name, p.major, p.minor,
                 (int)(p.totalGlobalMem >> 30));
    })
}

B200_API int b200_nbody_create(int64_t n, const double* pos, const double* vel, const double* mass, double G,
                               double softening, double damping, double theta, int device, b200_nbody** out)
{
    B200_ARG(out, "out handle is null");
    *out = nullptr;
    B200_ARG(n >= 0 && n < (int64_t)1 << 26, "n out of range [0, 2^26) (format limit of an n-body handle)");
    B200_ARG(n == 0 || (pos && vel && mass), "pos/vel/mass is null");
    B200_ARG(theta >= 0.0, "theta must be >= 0");
    b200_nbody* h = new b200_nbody();
    try {
        h->sim.device = device;
        h->sim.G = G;
        h->sim.softening = softening;
        h->sim.damping = damping;
        h->sim.theta = theta;
        b200::nbody_alloc(h->sim, (int)n);
        b200::nbody_upload(h->sim, pos, vel, mass);
    }
    B200_CATCH_CREATE({ b200::nbody_free(h->sim); delete h; })
    *out = h;
    return B200_OK;
}

B200_API int b200_generate_distribution(const char* distribution, int64_t n, double R, double G, uint64_t seed, int device,
                                        double* pos, double* vel, double* mass)
{
    B200_ARG(distribution, "distribution name is null");
    B200_ARG(n >= 0 && n < (int64_t)1 << 30, "n out of range [0, 2^30)");
    B200_ARG(n == 0 || (pos && vel && mass), "pos/vel/mass is null");
    B200_TRY({
        B200_CHECK(cudaSetDevice(device));
        cudaDeviceProp prop;
        B200_CHECK(cudaGetDeviceProperties(&prop, device));
        const size_t N = (size_t)n;
        double* d = b200::dev_alloc<double>(7 * N);
        try {
            b200::generate_device(b200::generator_id(distribution), n, R, G, seed, d, d + 3 * N, d + 6 * N, nullptr,
                                  prop.multiProcessorCount);
            B200_CHECK(cudaMemcpy(pos, d, 3 * N * sizeof(double), cudaMemcpyDeviceToHost));
            B200_CHECK(cudaMemcpy(vel, d + 3 * N, 3 * N * sizeof(double), cudaMemcpyDeviceToHost));
            B200_CHECK(cudaMemcpy(mass, d + 6 * N, N * sizeof(double), cudaMemcpyDeviceToHost));
        } catch (...) {
            cudaFree(d);
            throw;
        }
        cudaFree(d);
    })
}

B200_API int b200_nbody_create_generated(const char* distribution, int64_t n, double R, double G_dist, uint64_t seed,
                                         double G, double softening, double damping, double theta, int device, b200_nbody** out)
{
    B200_ARG(out, "out handle is null");
    *out = nullptr;
    B200_ARG(distribution, "distribution name is null");
    B200_ARG(n >= 0 && n < (int64_t)1 << 26, "n out of range [0, 2^26) (format limit of an n-body handle)");
    B200_ARG(theta >= 0.0, "theta must be >= 0");
    b200_nbody* h = new b200_nbody();
    try {
        h->sim.device = device;
        h->sim.G = G;
        h->sim.softening = softening;
        h->sim.damping = damping;
        h->sim.theta = theta;
        b200::nbody_alloc(h->sim, (int)n);
        b200::nbody_generate(h->sim, b200::generator_id(distribution), R, G_dist, seed);
    }
    B200_CATCH_CREATE({ b200::nbody_free(h->sim); delete h; })
    *out = h;
    return B200_OK;
}

B200_API int b200_nbody_destroy(b200_nbody* h)
{
    if (!h) return B200_OK;
    if (h->group) b200::group_destroy(h->group);
    h->group = nullptr;
    for (b200::NBodySim* o : h->others) { b200::nbody_free(*o); delete o; }
    b200::nbody_free(h->sim);
    delete h;
    return B200_OK;
}

B200_API int b200_nccl_unique_id(void* out128)
{
    B200_ARG(out128, "null argument");
    B200_TRY(b200::nccl_unique_id(out128))
}

B200_API int b200_nbody_comm_init(b200_nbody* h, const void* id128, int rank, int world)
{
    B200_ARG(h && id128, "null argument");
    B200_ARG(!h->group, "the handle already belongs to a group");
    B200_ARG(world >= 1 && world <= b200::SHARD_MAX_WORLD && rank >= 0 && rank < world, "bad rank / world (at most 8 ranks)");
    B200_TRY(h->group = b200::group_create_rank(h->sim, id128, rank, world))
}

B200_API int b200_nbody_world(b200_nbody* h, int* world)
{
    B200_ARG(h && world, "null argument");
    *world = b200::group_world(h->group);
    return B200_OK;
}

B200_API int b200_nbody_create_multi(int64_t n, const double* pos, const double* vel, const double* mass, double G,
                                     double softening, double damping, double theta, uint32_t device_mask, b200_nbody** out)
{
    B200_ARG(out, "out handle is null");
    *out = nullptr;
    B200_ARG(n >= 0 && n < (int64_t)1 << 26, "n out of range [0, 2^26) (format limit of an n-body handle)");
    B200_ARG(n == 0 || (pos && vel && mass), "pos/vel/mass is null");
    B200_ARG(theta >= 0.0, "theta must be >= 0");
    B200_ARG(device_mask != 0u, "empty device mask");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) ndev = 0;
    std::vector<int> devs;
    for (int d = 0; d < 32; ++d)
        if (device_mask & (1u << d)) {
            B200_ARG(d < ndev, "device mask names a device that does not exist");
            devs.push_back(d);
        }
    B200_ARG((int)devs.size() <= b200::SHARD_MAX_WORLD, "at most 8 devices");
    b200_nbody* h = new b200_nbody();
    try {
        std::vector<b200::NBodySim*> sims;
        for (size_t i = 0; i < devs.size(); ++i) {
            b200::NBodySim* s = &h->sim;
            if (i > 0) { s = new b200::NBodySim(); h->others.push_back(s); }
            s->device = devs[i];
            s->G = G; s->softening = softening; s->damping = damping; s->theta = theta;
            b200::nbody_alloc(*s, (int)n);
            b200::nbody_upload(*s, pos, vel, mass);
            sims.push_back(s);
        }
        if (sims.size() > 1) h->group = b200::group_create_local(sims);
    }
    B200_CATCH_CREATE(b200_nbody_destroy(h))
    *out = h;
    return B200_OK;
}

static int visible_frame(b200_nbody* h, double max_speed, const double* camera, float* pos, float* col, int64_t* count, bool to_device)
{
    B200_ARG(h && camera && count, "null argument");
    B200_ARG(h->sim.n == 0 || (pos && col), "output buffer is null");
    B200_TRY({
        b200::Camera c;
        for (int k = 0; k < 3; ++k) { c.pos[k] = camera[k]; c.forward[k] = camera[3 + k]; c.right[k] = camera[6 + k]; c.up[k] = camera[9 + k]; }
        c.tan_h = camera[12]; c.tan_v = camera[13]; c.far_dist = camera[14];
        *count = b200::nbody_visible_frame(h->sim, max_speed, c, pos, col, to_device);
    })
}

B200_API int b200_nbody_visible_frame(b200_nbody* h, double max_speed, const double* camera, float* pos_out, float* col_out, int64_t* count)
{
    return visible_frame(h, max_speed, camera, pos_out, col_out, count, false);
}

B200_API int b200_nbody_visible_frame_device(b200_nbody* h, double max_speed, const double* camera, void* pos_device, void* col_device,
                                            int64_t* count)
{
    return visible_frame(h, max_speed, camera, (float*)pos_device, (float*)col_device, count, true);
}

B200_API int b200_nbody_step(b200_nbody* h, double dt)
{
    B200_ARG(h, "handle is null");
    B200_TRY(step_any(h, dt))
}

B200_API int b200_nbody_step_n(b200_nbody* h, double dt, int nsteps)
{
    B200_ARG(h, "handle is null");
    B200_TRY({
        for (int i = 0; i < nsteps; ++i) step_any(h, dt);
    })
}

B200_API int b200_nbody_compute_accelerations(b200_nbody* h, float* out)
{
    B200_ARG(h && (out || h->sim.n == 0), "null argument");
    B200_TRY({
        rebuild_other_replicas(h);
        b200::nbody_get_accelerations(h->sim, out);
    })
}

B200_API int b200_nbody_compute_colors(b200_nbody* h, double max_speed)
{
    B200_ARG(h, "handle is null");
    B200_TRY(b200::nbody_compute_colors(h->sim, max_speed))
}

B200_API int b200_nbody_get_positions(b200_nbody* h, float* out)
{
    B200_ARG(h && (out || h->sim.n == 0), "null argument");
    B200_TRY(b200::nbody_get_positions(h->sim, out))
}

B200_API int b200_nbody_get_positions_f64(b200_nbody* h, double* out)
{
    B200_ARG(h && (out || h->sim.n == 0), "null argument");
    B200_TRY(b200::nbody_get_positions_f64(h->sim, out))
}

B200_API int b200_nbody_get_velocities(b200_nbody* h, double* out)
{
    B200_ARG(h && (out || h->sim.n == 0), "null argument");
    B200_TRY(b200::nbody_get_velocities(h->sim, out))
}

B200_API int b200_nbody_get_colors(b200_nbody* h, float* out)
{
    B200_ARG(h && (out || h->sim.n == 0), "null argument");
    B200_TRY(b200::nbody_get_colors(h->sim, out))
}

B200_API int b200_nbody_sync(b200_nbody* h)
{
    B200_ARG(h, "handle is null");
    B200_TRY({
        for (b200::NBodySim* o : h->others) {
            B200_CHECK(cudaSetDevice(o->device));
            B200_CHECK(cudaStreamSynchronize(o->stream));
            b200::nbody_check_errors(*o);
        }
        B200_CHECK(cudaSetDevice(h->sim.device));
        B200_CHECK(cudaStreamSynchronize(h->sim.stream));
        b200::nbody_check_errors(h->sim);   // never return truncated forces silently
    })
}

B200_API int b200_nbody_set_state(b200_nbody* h, const double* pos, const double* vel)
{
    B200_ARG(h && ((pos && vel) || h->sim.n == 0), "null argument");
    B200_TRY({
        for (b200::NBodySim* o : h->others) b200::nbody_upload_state(*o, pos, vel);
        b200::nbody_upload_state(h->sim, pos, vel);
        b200::group_state_replaced(h->group);
    })
}

B200_API int b200_nbody_set_params(b200_nbody* h, double G, double softening, double damping, double theta)
{
    B200_ARG(h, "handle is null");
    B200_ARG(theta >= 0.0, "theta must be >= 0");
    std::vector<b200::NBodySim*> all(h->others);
    all.push_back(&h->sim);
    for (b200::NBodySim* s : all) {
        s->G = G;
        s->softening = softening;
        s->damping = damping;
        s->theta = theta;
        s->tree_valid = false;
    }
    return B200_OK;
}

B200_API int b200_nbody_get_keys(b200_nbody* h, uint64_t* out)
{
    B200_ARG(h && (out || h->sim.n == 0), "null argument");
    B200_TRY({
        rebuild_other_replicas(h);
        b200::nbody_get_keys(h->sim, out);
    })
}

B200_API int b200_nbody_get_perm(b200_nbody* h, uint32_t* out)
{
    B200_ARG(h && (out || h->sim.n == 0), "null argument");
    B200_TRY({
        rebuild_other_replicas(h);
        b200::nbody_get_perm(h->sim, out);
    })
}

B200_API int b200_nbody_get_stats(b200_nbody* h, b200_nbody_stats* out)
{
    B200_ARG(h && out, "null argument");
    B200_TRY({
        b200::NBodySim& s = h->sim;
        B200_CHECK(cudaSetDevice(s.device));
        B200_CHECK(cudaStreamSynchronize(s.stream));
        memset(out, 0, sizeof(*out));
        out->n = s.n;
        out->steps = s.steps;
        unsigned alloc = 0, err = 0;
        unsigned long long ctr[b200::TRAV_COUNTERS] = {0};
        B200_CHECK(cudaMemcpy(&alloc, s.d_alloc, sizeof(alloc), cudaMemcpyDeviceToHost));
        B200_CHECK(cudaMemcpy(&err, s.d_error, sizeof(err), cudaMemcpyDeviceToHost));
        B200_CHECK(cudaMemcpy(ctr, s.d_interactions, sizeof(ctr), cudaMemcpyDeviceToHost));
        B200_CHECK(cudaMemcpy(&out->bounds, s.d_bounds, sizeof(double), cudaMemcpyDeviceToHost));
        unsigned kids = 0;
        B200_CHECK(cudaMemcpy(&kids, s.d_children, sizeof(kids), cudaMemcpyDeviceToHost));
        out->records = s.n > 1 ? (int64_t)kids + 1 : s.n;   // root + every cell's children
        out->pair_records = s.n > 1 ? (int64_t)alloc : 1;
        out->interactions = (int64_t)ctr[0];
        out->trav_pair_slots = (int64_t)ctr[1];
        out->trav_lane_pairs = (int64_t)ctr[2];
        out->trav_batches = (int64_t)ctr[3];
        out->trav_stack_max = (int64_t)ctr[4];
        out->trav_shared_pairs = (int64_t)ctr[5];
        out->trav_kernel = s.last_trav_kernel;
        out->trav_sure_pairs = (int64_t)ctr[6];
        out->error_flags = err;
        out->sm_count = s.sm_count;
        out->bytes_allocated = (int64_t)s.bytes_allocated;
        out->timed_steps = s.timer.count;
        for (int i = 0; i < B200_NBODY_PHASES; ++i) out->phase_ms[i] = s.timer.ms[i];
    })
}

B200_API int b200_nbody_reset_stats(b200_nbody* h)
{
    B200_ARG(h, "handle is null");
    B200_TRY({
        b200::NBodySim& s = h->sim;
        B200_CHECK(cudaSetDevice(s.device));
        B200_CHECK(cudaStreamSynchronize(s.stream));
        B200_CHECK(cudaMemset(s.d_interactions, 0, b200::TRAV_COUNTERS * sizeof(unsigned long long)));
        s.timer.reset();
    })
}

B200_API int b200_nbody_set_profiling(b200_nbody* h, int enabled)
{
    B200_ARG(h, "handle is null");
    h->sim.timer.enabled = enabled != 0;
    return B200_OK;
}

B200_API int b200_nbody_set_counting(b200_nbody* h, int enabled)
{
    B200_ARG(h, "handle is null");
    h->sim.count_interactions = enabled != 0;
    return B200_OK;
}

B200_API int b200_nbody_count_interactions(b200_nbody* h, int64_t* interactions)
{
    B200_ARG(h && interactions, "null argument");
    B200_TRY({
        rebuild_other_replicas(h);
        *interactions = b200::nbody_count_interactions(h->sim);
    })
}

B200_API int b200_nbody_state_checksum(b200_nbody* h, uint64_t out[2])
{
    B200_ARG(h && out, "null argument");
    B200_TRY(b200::nbody_state_checksum(h->sim, out))
}

B200_API int b200_nbody_timed_steps(b200_nbody* h, double dt, int nsteps, float* elapsed_ms)
{
    B200_ARG(h && elapsed_ms, "null argument");
    B200_TRY({
        b200::NBodySim& s = h->sim;
        B200_CHECK(cudaSetDevice(s.device));
        cudaEvent_t e0, e1;
        B200_CHECK(cudaEventCreate(&e0));
        B200_CHECK(cudaEventCreate(&e1));
        B200_CHECK(cudaEventRecord(e0, s.stream));
        for (int i = 0; i < nsteps; ++i) step_any(h, dt);
        B200_CHECK(cudaSetDevice(s.device));
        B200_CHECK(cudaEventRecord(e1, s.stream));
        B200_CHECK(cudaEventSynchronize(e1));
        B200_CHECK(cudaEventElapsedTime(elapsed_ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    })
}

B200_API int b200_nbody_launch_count(b200_nbody* h, int64_t* out)
{
    B200_ARG(h && out, "null argument");
    *out = h->sim.launches;
    return B200_OK;
}

B200_API int b200_nbody_frame_begin(b200_nbody* h, double max_speed, float* pos_out, float* col_out)
{
    B200_ARG(h && ((pos_out && col_out) || h->sim.n == 0), "null argument");
    B200_ARG(max_speed > 0.0, "max_speed must be > 0");
    B200_TRY(b200::nbody_frame_begin(h->sim, max_speed, pos_out, col_out))
}

B200_API int b200_nbody_frame_delta_begin(b200_nbody* h, double max_speed, int16_t* pos_delta_out, int16_t* col_delta_out)
{
    B200_ARG(h && ((pos_delta_out && col_delta_out) || h->sim.n == 0), "null argument");
    B200_ARG(max_speed > 0.0, "max_speed must be > 0");
    B200_TRY(b200::nbody_frame_delta_begin(h->sim, max_speed, pos_delta_out, col_delta_out))
}

B200_API int b200_nbody_frame_wait(b200_nbody* h)
{
    B200_ARG(h, "handle is null");
    B200_TRY(b200::nbody_frame_wait(h->sim))
}

B200_API int b200_nbody_set_state_begin(b200_nbody* h, const double* pos, const double* vel)
{
    B200_ARG(h && ((pos && vel) || h->sim.n == 0), "null argument");
    B200_ARG(h->others.empty(), "the asynchronous state prefetch is not available on a device-mask handle (use set_state)");
    B200_TRY(b200::nbody_set_state_begin(h->sim, pos, vel))
}

B200_API int b200_nbody_set_state_begin_rows(b200_nbody* h, const double* pos, const double* vel, int64_t row_begin, int64_t row_end)
{
    B200_ARG(h && ((pos && vel) || h->sim.n == 0), "null argument");
    B200_ARG(row_begin >= 0 && row_begin <= row_end && row_end <= h->sim.n, "rows out of range");
    B200_ARG(h->others.empty(), "the asynchronous state prefetch is not available on a device-mask handle (use set_state)");
    B200_TRY(b200::nbody_set_state_begin_rows(h->sim, pos, vel, (int)row_begin, (int)row_end))
}

B200_API int b200_nbody_upload_staging(b200_nbody* h, void** pos_device_ptr, void** vel_device_ptr)
{
    B200_ARG(h && pos_device_ptr && vel_device_ptr, "null argument");
    B200_TRY({
        double *p = nullptr, *v = nullptr;
        b200::nbody_upload_staging(h->sim, &p, &v);
        *pos_device_ptr = p;
        *vel_device_ptr = v;
    })
}

B200_API int b200_nbody_upload_wait(b200_nbody* h)
{
    B200_ARG(h, "handle is null");
    B200_TRY(b200::nbody_upload_wait(h->sim))
}

B200_API int b200_nbody_frame_begin_rows(b200_nbody* h, double max_speed, float* pos_out, float* col_out, int64_t row_begin,
                                         int64_t row_end)
{
    B200_ARG(h && ((pos_out && col_out) || h->sim.n == 0), "null argument");
    B200_ARG(max_speed > 0.0, "max_speed must be > 0");
    B200_ARG(row_begin >= 0 && row_begin <= row_end && row_end <= h->sim.n, "rows out of range");
    B200_TRY(b200::nbody_frame_begin_rows(h->sim, max_speed, pos_out, col_out, (int)row_begin, (int)row_end))
}

B200_API int b200_nbody_set_state_commit(b200_nbody* h)
{
    B200_ARG(h, "handle is null");
    B200_ARG(h->others.empty(), "the asynchronous state prefetch is not available on a device-mask handle (use set_state)");
    B200_TRY({
        b200::nbody_set_state_commit(h->sim);
        b200::group_state_replaced(h->group);
    })
}

B200_API int b200_nbody_set_stream(b200_nbody* h, void* cuda_stream, int external)
{
    B200_ARG(h, "handle is null");
    B200_TRY({
        b200::NBodySim& s = h->sim;
        B200_CHECK(cudaSetDevice(s.device));
        B200_CHECK(cudaStreamSynchronize(s.stream));
        s.stream = external ? (cudaStream_t)cuda_stream : s.own_stream;
    })
}

B200_API int b200_nbody_set_shard(b200_nbody* h, int64_t begin, int64_t end)
{
    B200_ARG(h, "handle is null");
    B200_ARG(begin >= 0 && begin <= end && end <= h->sim.n && (begin % 32 == 0 || begin == end), "bad shard range");
    h->sim.shard_begin = (int)begin;
    h->sim.shard_end = (int)end;
    return B200_OK;
}

B200_API int b200_cost_weighted_split(const uint64_t* chunk_cost, int nchunks, int64_t chunk, int64_t n, int world, int64_t* split_out)
{
    B200_ARG(chunk_cost && split_out && nchunks > 0 && chunk > 0 && n >= 0 && world >= 1, "bad argument");
    B200_ARG((int64_t)nchunks * chunk >= n, "the chunks do not cover n");
    B200_TRY({
        const std::vector<int64_t> sp = b200::cost_weighted_split((const unsigned long long*)chunk_cost, nchunks, chunk, n, world);
        for (int r = 0; r <= world; ++r) split_out[r] = sp[r];
    })
}

B200_API int b200_nbody_get_shard(b200_nbody* h, int64_t* begin, int64_t* end)
{
    B200_ARG(h && begin && end, "null argument");
    *begin = h->sim.shard_begin;
    *end = h->sim.shard_end;
    return B200_OK;
}

B200_API int b200_nbody_sharded_sort_setup(b200_nbody* h, int64_t slice, int world, void** keys_device_ptr, void** vals_device_ptr)
{
    B200_ARG(h && keys_device_ptr && vals_device_ptr, "null argument");
    B200_ARG(slice > 0 && slice < ((int64_t)1 << 31) && world > 0 && world <= 64, "bad slice / world");
    B200_TRY({
        b200::nbody_ms_setup(h->sim, (int)slice, world);
        *keys_device_ptr = h->sim.ms_keys;
        *vals_device_ptr = h->sim.ms_vals;
    })
}

B200_API int b200_nbody_sort_local(b200_nbody* h, int rank)
{
    B200_ARG(h, "handle is null");
    B200_TRY(b200::nbody_ms_sort_local(h->sim, rank))
}

B200_API int b200_nbody_step_begin_sorted(b200_nbody* h)
{
    B200_ARG(h, "handle is null");
    B200_TRY(b200::nbody_step_begin_sorted(h->sim))
}

B200_API int b200_nbody_step_begin(b200_nbody* h)
{
    B200_ARG(h, "handle is null");
    B200_TRY(b200::nbody_step_begin(h->sim))
}

B200_API int b200_nbody_step_end(b200_nbody* h, double dt)
{
    B200_ARG(h, "handle is null");
    B200_TRY(b200::nbody_step_end(h->sim, dt))
}

B200_API int b200_nbody_acc_buffer(b200_nbody* h, void** device_ptr, int64_t* capacity_entries)
{
    B200_ARG(h && device_ptr && capacity_entries, "null argument");
    *device_ptr = h->sim.acc;
    *capacity_entries = h->sim.acc_capacity;
    return B200_OK;
}

B200_API int b200_fp32_peak_tflops(int device, double* tflops)
{
    B200_ARG(tflops, "null argument");
    B200_TRY(*tflops = b200::fp32_peak_tflops(device))
}

B200_API int b200_host_alloc(int64_t bytes, void** out)
{
    B200_ARG(out, "out pointer is null");
    *out = nullptr;
    B200_ARG(bytes >= 0, "negative size");
    B200_TRY(B200_CHECK(cudaHostAlloc(out, bytes > 0 ? (size_t)bytes : 1, cudaHostAllocPortable)))
}

B200_API int b200_host_free(void* ptr)
{
    if (!ptr) return B200_OK;
    B200_TRY(B200_CHECK(cudaFreeHost(ptr)))
}

// ============================================================================ boids
B200_API int b200_boids_create(int64_t n, const double* pos, const double* vel, const double* col,
                               const b200_boids_params* params, int device, b200_boids** out)
{
    B200_ARG(out, "out handle is null");
    *out = nullptr;
    B200_ARG(params, "params is null");
    B200_ARG(n >= 0 && n < (int64_t)1 << 30, "n out of range [0, 2^30)");
    B200_ARG(n == 0 || (pos && vel && col), "pos/vel/col is null");
    B200_ARG(params->perception_radius > 0 && params->bounds > 0 && params->wall_margin > 0, "bad boids params");
    b200_boids* h = new b200_boids();
    try {
        h->sim.device = device;
        memcpy(&h->sim.p, params, sizeof(b200::BoidsParams));
        b200::boids_alloc(h->sim, (int)n);
        b200::boids_upload(h->sim, pos, vel, col);
    }
    B200_CATCH_CREATE({ b200::boids_free(h->sim); delete h; })
    *out = h;
    return B200_OK;
}

B200_API int b200_boids_destroy(b200_boids* h)
{
    if (!h) return B200_OK;
    b200::boids_free(h->sim);
    delete h;
    return B200_OK;
}

B200_API int b200_boids_step(b200_boids* h, double dt)
{
    B200_ARG(h, "handle is null");
    B200_TRY(b200::boids_step(h->sim, dt))
}

B200_API int b200_boids_get_state(b200_boids* h, double* pos, double* vel, double* col)
{
    B200_ARG(h, "handle is null");
    B200_TRY(b200::boids_get_state(h->sim, pos, vel, col))
}

B200_API int b200_boids_set_state(b200_boids* h, const double* pos, const double* vel, const double* col)
{
    B200_ARG(h && ((pos && vel && col) || h->sim.n == 0), "null argument");
    B200_TRY(b200::boids_upload(h->sim, pos, vel, col))
}

B200_API int b200_boids_get_cell_indices(b200_boids* h, int32_t* out)
{
    B200_ARG(h && (out || h->sim.n == 0), "null argument");
    B200_TRY(b200::boids_get_cells(h->sim, out))
}

B200_API int b200_boids_get_stats(b200_boids* h, b200_boids_stats* out)
{
    B200_ARG(h && out, "null argument");
    B200_TRY({
        b200::BoidsSim& s = h->sim;
        B200_CHECK(cudaSetDevice(s.device));
        B200_CHECK(cudaStreamSynchronize(s.stream));
        memset(out, 0, sizeof(*out));
        out->n = s.n;
        out->steps = s.steps;
        out->num_cells = s.num_cells;
        out->grid_dim = s.grid_dim;
        out->key_bits = s.key_bits;
        out->cell_size = s.cell_size;
        out->grid_offset = s.grid_offset;
        unsigned long long pairs = 0;
        B200_CHECK(cudaMemcpy(&pairs, s.d_pairs, sizeof(pairs), cudaMemcpyDeviceToHost));
        out->neighbor_pairs = (int64_t)pairs;
        out->bytes_allocated = (int64_t)s.bytes_allocated;
        out->launches = s.launches;
        out->timed_steps = s.timer.count;
        for (int i = 0; i < B200_BOIDS_PHASES; ++i) out->phase_ms[i] = s.timer.ms[i];
    })
}

B200_API int b200_boids_reset_stats(b200_boids* h)
{
    B200_ARG(h, "handle is null");
    B200_TRY({
        b200::BoidsSim& s = h->sim;
        B200_CHECK(cudaSetDevice(s.device));
        B200_CHECK(cudaStreamSynchronize(s.stream));
        B200_CHECK(cudaMemset(s.d_pairs, 0, sizeof(unsigned long long)));
        s.timer.reset();
    })
}

B200_API int b200_boids_set_profiling(b200_boids* h, int enabled)
{
    B200_ARG(h, "handle is null");
    h->sim.timer.enabled = enabled != 0;
    return B200_OK;
}

B200_API int b200_boids_timed_steps(b200_boids* h, double dt, int nsteps, float* elapsed_ms)
{
    B200_ARG(h && elapsed_ms, "null argument");
    B200_TRY({
        b200::BoidsSim& s = h->sim;
        B200_CHECK(cudaSetDevice(s.device));
        cudaEvent_t e0, e1;
        B200_CHECK(cudaEventCreate(&e0));
        B200_CHECK(cudaEventCreate(&e1));
        B200_CHECK(cudaEventRecord(e0, s.stream));
        for (int i = 0; i < nsteps; ++i) b200::boids_step(s, dt);
        B200_CHECK(cudaEventRecord(e1, s.stream));
        B200_CHECK(cudaEventSynchronize(e1));
        B200_CHECK(cudaEventElapsedTime(elapsed_ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    })
}

B200_API int b200_boids_sync(b200_boids* h)
{
    B200_ARG(h, "handle is null");
    B200_TRY({
        B200_CHECK(cudaSetDevice(h->sim.device));
        B200_CHECK(cudaStreamSynchronize(h->sim.stream));
    })
}
