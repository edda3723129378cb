// common.cuh -- shared helpers for the sm_100a kernels of libb200sim.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace b200 {

void set_error(const std::string& msg);   // defined in capi.cu; read back through b200_last_error()

struct CudaError { std::string msg; };
// a device-side error flag (traversal stack / record pool overflow) or a misuse of the handle: B200_ERR_STATE
struct StateError { std::string msg; };

#define B200_CHECK(expr)                                                                      \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            char _b[512];                                                                     \
            snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                     __FILE__, __LINE__);                                                     \
            throw ::b200::CudaError{std::string(_b)};                                         \
        }                                                                                     \
    } while (0)

#define B200_REQUIRE(cond, text)                                                              \
    do {                                                                                      \
        if (!(cond)) throw ::b200::CudaError{std::string(text)};                              \
    } while (0)

template <typename T>
inline T* dev_alloc(size_t count)
{
    T* p = nullptr;
    if (count == 0) count = 1;
    B200_CHECK(cudaMalloc(&p, count * sizeof(T)));
    return p;
}

inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// gpu-scope relaxed accesses for flag+value status words (ld.volatile compiles to a system-scope
// strong load, which is slower than the look-back needs)
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned* p, unsigned v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Phase timer: CUDA events on the launching stream, accumulated per phase when enabled.
struct PhaseTimer {
    static constexpr int MAX = 16;
    cudaEvent_t ev[MAX + 1] = {};
    double ms[MAX] = {};
    int64_t count = 0;
    bool enabled = false;
    bool created = false;
    int cur = 0;
    void init()
    {
        if (created) return;
        for (int i = 0; i <= MAX; ++i) B200_CHECK(cudaEventCreate(&ev[i]));
        created = true;
    }
    void destroy()
    {
        if (!created) return;
        for (int i = 0; i <= MAX; ++i) cudaEventDestroy(ev[i]);
        created = false;
    }
    void begin(cudaStream_t s)
    {
        if (!enabled) return;
        cur = 0;
        cudaEventRecord(ev[0], s);
    }
    void mark(cudaStream_t s)   // ends phase `cur`
    {
        if (!enabled) return;
        ++cur;
        cudaEventRecord(ev[cur], s);
    }
    void collect()   // call after the stream is synchronised
    {
        if (!enabled) return;
        for (int i = 0; i < cur; ++i) {
            float t = 0.f;
            cudaEventElapsedTime(&t, ev[i], ev[i + 1]);
            ms[i] += t;
        }
        ++count;
    }
    void reset()
    {
        for (int i = 0; i < MAX; ++i) ms[i] = 0.0;
        count = 0;
    }
};

}  // namespace b200
