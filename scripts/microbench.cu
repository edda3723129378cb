// microbench.cu -- instruction-rate probes on B200 that decide the traversal kernel's inner loop:
// scalar FFMA vs packed FFMA2 / FADD2 / FMUL2, MUFU.RSQ, shared-memory broadcast LDS.128, VOTE.
// Prints warp-instructions per clock per SM sub-partition (SMSP) and the implied TFLOP/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench scripts/microbench.cu && /tmp/microbench
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, float a, float b)
{
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(a, b, a, b);
    __syncthreads();
    float2 x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = make_float2(threadIdx.x + c, threadIdx.x - c);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    unsigned v = 0;
    unsigned w[CHAINS];
    float y[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { w[c] = threadIdx.x * c; y[c] = 1.0f + threadIdx.x + c; }
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (MODE == 0) { x[c].x = fmaf(x[c].x, a, b); }                                  // 1 FFMA
            if (MODE == 1) { x[c] = __ffma2_rn(x[c], a2, b2); }                              // 1 FFMA2
            if (MODE == 2) { x[c] = __fadd2_rn(x[c], b2); }                                  // 1 FADD2
            if (MODE == 3) { x[c] = __fmul2_rn(x[c], a2); }                                  // 1 FMUL2
            if (MODE == 4) { x[c] = __ffma2_rn(x[c], a2, b2); x[c].x = fmaf(x[c].x, a, b); } // FFMA2 + FFMA
            if (MODE == 5) { asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[c].x)); }  // 1 MUFU.RSQ
            if (MODE == 6) { const float4 t = sm[(i + c) & 63]; x[c].x += t.x; x[c].y += t.w; }   // LDS.128 broadcast + 2 FADD
            if (MODE == 7) { v += __ballot_sync(0xffffffffu, x[c].x > b); x[c].x += a; }     // VOTE + FADD (+IADD)
            if (MODE == 8) { x[c].x = x[c].x + b; }                                          // 1 FADD
            if (MODE == 10) { x[c] = __ffma2_rn(x[c], a2, b2); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[c]) : "r"(i), "r"(c + 1)); }   // FFMA2 + 1 LOP3 (ALU pipe)
            if (MODE == 16) { x[c] = __ffma2_rn(x[c], a2, b2); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[c]) : "r"(i), "r"(c + 1)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[c]) : "r"(i), "r"(c + 2)); }   // FFMA2 + 2 LOP3
            if (MODE == 17) { x[c] = __ffma2_rn(x[c], a2, b2); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[c]) : "r"(i), "r"(c + 1)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[c]) : "r"(i), "r"(c + 2)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[c]) : "r"(i), "r"(c + 3)); }   // FFMA2 + 3 LOP3
            if (MODE == 11) { x[c] = __ffma2_rn(x[c], a2, b2); w[c] = (w[c] ^ (unsigned)i) * 3u + v; w[c] = (w[c] >> 3) ^ w[c]; }   // FFMA2 + ~3 integer ops
            if (MODE == 12) { x[c] = __ffma2_rn(x[c], a2, b2); asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(y[c])); }   // FFMA2 + MUFU
            if (MODE == 13) { x[c] = __ffma2_rn(x[c], a2, b2); v += __ballot_sync(0xffffffffu, x[c].x > b); }   // FFMA2 + FSETP + VOTE + IADD
            if (MODE == 14) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[c]) : "r"(i), "r"(c + 1)); }   // 1 LOP3 only
            if (MODE == 15) { x[c] = __ffma2_rn(x[c], a2, b2); const float t = reinterpret_cast<const float*>(sm)[(i + c) & 63]; y[c] += t; }   // FFMA2 + LDS.32 + FADD
            if (MODE == 9) { x[c] = __ffma2_rn(x[c], a2, b2); x[c].x = x[c].x + b; x[c].y = x[c].y + b; }  // FFMA2 + 2 FADD
        }
    }
    float r = (float)v;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) r += x[c].x + x[c].y + (float)w[c] + y[c];
    if (r == 12345.678f) out[0] = r;
}

template <int MODE>
int run(const char* name, double instr_per_chain_iter, double flop_per_chain_iter, float* d, int sms, double mhz)
{
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    const int blocks = sms * 8;
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CHECK(cudaEventRecord(e0));
        probe<MODE><<<blocks, 256>>>(d, 1.0000001f, 1e-7f);
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        float ms;
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    const double warp_instr = instr_per_chain_iter * CHAINS * ITERS * (256.0 / 32) * blocks;
    const double cycles = best * 1e-3 * mhz * 1e6;
    const double per_smsp_clk = warp_instr / (cycles * sms * 4);
    const double tflops = flop_per_chain_iter * CHAINS * ITERS * 256.0 * blocks / (best * 1e-3) / 1e12;
    const double clk_per_iter = cycles * sms * 4 / (double(CHAINS) * ITERS * (256.0 / 32) * blocks);   // SMSP cycles per chain-iteration of one warp
    printf("%-28s %8.3f ms  %6.3f warp-instr/clk/SMSP (at %.0f MHz)  %7.2f TFLOP/s  %6.2f clk/iter\n", name, best, per_smsp_clk, mhz, tflops, clk_per_iter);
    return 0;
}

int main()
{
    cudaDeviceProp p;
    CHECK(cudaGetDeviceProperties(&p, 0));
    int khz = 0;
    CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double mhz = khz / 1000.0;
    printf("%s, %d SMs, clock attr %.0f MHz\n", p.name, p.multiProcessorCount, mhz);
    float* d;
    CHECK(cudaMalloc(&d, 4));
    const int sms = p.multiProcessorCount;
    run<0>("FFMA (scalar)", 1, 2, d, sms, mhz);
    run<1>("FFMA2 (packed)", 1, 4, d, sms, mhz);
    run<2>("FADD2", 1, 2, d, sms, mhz);
    run<3>("FMUL2", 1, 2, d, sms, mhz);
    run<4>("FFMA2 + FFMA", 2, 6, d, sms, mhz);
    run<5>("MUFU.RSQ", 1, 1, d, sms, mhz);
    run<6>("LDS.128 bcast + 2 FADD", 3, 2, d, sms, mhz);
    run<7>("VOTE + FADD + IADD", 3, 1, d, sms, mhz);
    run<8>("FADD (scalar)", 1, 1, d, sms, mhz);
    run<9>("FFMA2 + 2 FADD", 3, 6, d, sms, mhz);
    run<10>("FFMA2 + 1 LOP3", 2, 4, d, sms, mhz);
    run<16>("FFMA2 + 2 LOP3", 3, 4, d, sms, mhz);
    run<17>("FFMA2 + 3 LOP3", 4, 4, d, sms, mhz);
    run<11>("FFMA2 + ~5 int (count 6)", 6, 4, d, sms, mhz);
    run<12>("FFMA2 + MUFU", 2, 4, d, sms, mhz);
    run<13>("FFMA2 + FSETP+VOTE+IADD", 4, 4, d, sms, mhz);
    run<14>("LOP3 only", 1, 0, d, sms, mhz);
    run<15>("FFMA2 + LDS.32 + FADD", 3, 5, d, sms, mhz);
    cudaFree(d);
    return 0;
}
