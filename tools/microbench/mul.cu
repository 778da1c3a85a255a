// mul.cu -- which 32x32 -> 64 multiply is cheapest on sm_100a?  (Philox4x32 needs hi and lo of two products per round.)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int ITER = 4096, ILP = 8;
template <int KIND>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
    uint32_t x[ILP], y[ILP];
    for (int i = 0; i < ILP; ++i) { x[i] = threadIdx.x * 2654435761u + i + seed; y[i] = x[i] ^ 0x9E3779B9u; }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0) { const unsigned long long p = (unsigned long long)0xD2511F53u * x[i]; x[i] = (uint32_t)(p >> 32) ^ y[i]; y[i] = (uint32_t)p; }   // IMAD.WIDE
            if (KIND == 1) { const uint32_t hi = __umulhi(0xD2511F53u, x[i]), lo = 0xD2511F53u * x[i]; x[i] = hi ^ y[i]; y[i] = lo; }                  // IMAD.HI + IMAD
            if (KIND == 2) { x[i] = 0xD2511F53u * x[i] + y[i]; }                                                                                        // IMAD (lo)
            if (KIND == 3) { x[i] = __umulhi(0xD2511F53u, x[i]) ^ y[i]; }                                                                               // IMAD.HI
            if (KIND == 4) {  // 16-bit pieces: hi/lo of M * x with M = Mh:Ml, x = xh:xl via four 16x16 products (IMAD.U16? -> plain IMAD on 16-bit values)
                const uint32_t Mh = 0xD251u, Ml = 0x1F53u, xh = x[i] >> 16, xl = x[i] & 0xffffu;
                const uint32_t ll = Ml * xl, lh = Ml * xh, hl = Mh * xl, hh = Mh * xh;
                const uint32_t mid = (ll >> 16) + (lh & 0xffffu) + (hl & 0xffffu);
                const uint32_t lo = (ll & 0xffffu) | (mid << 16);
                const uint32_t hi = hh + (lh >> 16) + (hl >> 16) + (mid >> 16);
                x[i] = hi ^ y[i]; y[i] = lo;
            }
        }
    }
    uint32_t s = 0;
    for (int i = 0; i < ILP; ++i) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND>
void run(const char* name) {
    uint32_t* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND><<<148 * 8, 256>>>(out, 1);
    cudaEventRecord(e0); k<KIND><<<148 * 8, 256>>>(out, 2); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double prods = (double)148 * 8 * 8 * ITER * ILP;  // warp-level products
    printf("%-34s %8.3f ms  %.2f clk per warp-product per SMSP\n", name, ms, ms * 1e-3 * clk * 1e3 * 148 * 4 / prods);
    cudaFree(out);
}
int main() {
    run<0>("IMAD.WIDE (hi & lo)");
    run<1>("IMAD.HI + IMAD (hi & lo)");
    run<2>("IMAD lo only");
    run<3>("IMAD.HI only");
    run<4>("four 16x16 pieces (hi & lo)");
    return 0;
}
