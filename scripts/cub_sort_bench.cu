// cub_sort_bench.cu -- speed comparator for the hand-written onesweep sort (NOT shipped, not linked into
// libb200sim.so): cub::DeviceRadixSort::SortPairs on 63-bit keys + 32-bit values, random and nearly sorted.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cub_sort_bench scripts/cub_sort_bench.cu
#include <cstdio>
#include <cstdint>
#include <cub/cub.cuh>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void fill(uint64_t* k, uint32_t* v, int n, int mode)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t x = (uint64_t)i * 0x9E3779B97F4A7C15ull;
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    if (mode == 0) k[i] = x >> 1;                                                  // random 63-bit keys
    else k[i] = ((uint64_t)i << 30) + (x & ((1ull << 33) - 1));                    // nearly sorted: local disorder over ~8 neighbours
    v[i] = (uint32_t)i;
}

int main(int argc, char** argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 50000000;
    uint64_t *k0, *k1; uint32_t *v0, *v1;
    CHECK(cudaMalloc(&k0, 8ull * n)); CHECK(cudaMalloc(&k1, 8ull * n));
    CHECK(cudaMalloc(&v0, 4ull * n)); CHECK(cudaMalloc(&v1, 4ull * n));
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k0, k1, v0, v1, n, 0, 63);
    void* tmp; CHECK(cudaMalloc(&tmp, tmp_bytes));
    cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            fill<<<(n + 255) / 256, 256>>>(k0, v0, n, mode);
            CHECK(cudaEventRecord(e0));
            CHECK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, n, 0, 63));
            CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
            float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        printf("cub::DeviceRadixSort::SortPairs n=%d u64(63 bits)+u32 %s: %.3f ms  (%.0f GB/s on 200 B/body)\n", n,
               mode ? "nearly sorted" : "random", best, 200.0 * n / (best * 1e-3) / 1e9);
    }
    return 0;
}
