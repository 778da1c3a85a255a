// pipes.cu -- issue-rate microbenchmark for the instruction classes the generator leans on (B200, sm_100a):
// FFMA, FFMA2 (packed fp32x2), IMAD.WIDE.U32, LOP3, MUFU.EX2, and FFMA2 + IMAD.WIDE mixed,
// and MUFU mixed with each of them (does the SFU pipe run beside the FMA-heavy pipe, or do they share a dispatch port?).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a pipes.cu -o pipes && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 4096, ILP = 8;

template <int KIND>
__global__ void __launch_bounds__(256) k(float* out, uint32_t seed) {
    float a[ILP];
    float2 b[ILP];
    uint32_t c[ILP];
    unsigned long long w[ILP];
    for (int i = 0; i < ILP; ++i) { a[i] = threadIdx.x * 1e-3f + i; b[i] = make_float2(a[i], a[i] + 1.f); c[i] = threadIdx.x * 2654435761u + i + seed; w[i] = c[i]; }
    const float m = 1.0000001f, d = 1e-7f;
    const float2 m2 = make_float2(m, m), d2 = make_float2(d, d);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0) a[i] = fmaf(a[i], m, d);
            if (KIND == 1) b[i] = __ffma2_rn(b[i], m2, d2);
            if (KIND == 2) { w[i] = (unsigned long long)0xD2511F53u * (uint32_t)w[i] + (w[i] >> 32); }
            if (KIND == 3) c[i] = (c[i] ^ (c[i] >> 3)) ^ seed;
            if (KIND == 4) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }
            if (KIND == 5) { b[i] = __ffma2_rn(b[i], m2, d2); w[i] = (unsigned long long)0xD2511F53u * (uint32_t)w[i] + (w[i] >> 32); }
            if (KIND == 6) { a[i] = fmaf(a[i], m, d); w[i] = (unsigned long long)0xD2511F53u * (uint32_t)w[i] + (w[i] >> 32); }
            if (KIND == 7) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); w[i] = (unsigned long long)0xD2511F53u * (uint32_t)w[i] + (w[i] >> 32); }
            if (KIND == 8) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); b[i] = __ffma2_rn(b[i], m2, d2); }
            if (KIND == 9) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); c[i] = (c[i] ^ (c[i] >> 3)) ^ seed; }
            if (KIND == 10) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); w[i] = (unsigned long long)0xD2511F53u * (uint32_t)w[i] + (w[i] >> 32);
                              w[(i + 1) % ILP] = (unsigned long long)0xCD9E8D57u * (uint32_t)w[(i + 1) % ILP] + (w[(i + 1) % ILP] >> 32); }
            if (KIND == 11) { a[i] = (float)c[i] * m; c[i] += seed; }
        }
    }
    float s = 0.f;
    for (int i = 0; i < ILP; ++i) s += a[i] + b[i].x + b[i].y + (float)c[i] + (float)(uint32_t)w[i] + (float)(uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
void run(const char* name, double ops_per_iter) {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND><<<148 * 8, 256>>>(out, 1);
    cudaEventRecord(e0);
    k<KIND><<<148 * 8, 256>>>(out, 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double warp_instr = (double)148 * 8 * 8 * ITER * ILP * ops_per_iter;  // warps x iterations x instructions
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s %8.3f ms  %.2f warp-instr / clk / SM  (%.2f per SMSP)\n", name, ms, warp_instr / cycles / 148, warp_instr / cycles / 148 / 4);
    cudaFree(out);
}

int main() {
    run<0>("FFMA", 1);
    run<1>("FFMA2 (packed fp32x2)", 1);
    run<2>("IMAD.WIDE.U32", 1);
    run<3>("LOP3 x2", 2);
    run<4>("MUFU.EX2", 1);
    run<5>("FFMA2 + IMAD.WIDE", 2);
    run<6>("FFMA + IMAD.WIDE", 2);
    run<7>("MUFU.EX2 + IMAD.WIDE", 2);
    run<8>("MUFU.EX2 + FFMA2", 2);
    run<9>("MUFU.EX2 + LOP3 x2", 3);
    run<10>("MUFU.EX2 + 2 IMAD.WIDE", 3);
    run<11>("I2FP.U32 + FMUL + IADD", 3);
    return 0;
}
