// philox.cuh -- counter-based normal generation on the device: Philox4x32-10 + Box-Muller, no cuRAND.
//
// The reference draws from std::mt19937 re-seeded from std::random_device three times per path
// (RoughVolatility.cpp:238-262): sequential, stateful and non-reproducible.  A counter-based generator has
// no state to carry, so any (path, step) cell can be produced by any thread and results do not depend on
// the launch geometry or the number of GPUs.  Stream layout (must match oracle/port/mcp_oracle.c):
//   rough-vol Z: ctr = (g_lo, g_hi, k>>1, 0) -> (x0,x1) => (Zre_k, Zim_k) for even k, (x2,x3) for odd k
//   rough-vol W: ctr = (g_lo, g_hi, k>>2, 2) -> four real normals W_{4q..4q+3} (the single N(0,1) that the reference
//                builds as rho W1 + sqrt(1-rho^2) W2; 3 normals per path-step instead of 4, same law)
//   gbm        : ctr = (g_lo, g_hi, q, 1)    -> normals of steps 4q .. 4q+3
//   rough-vol, pair stream (native mode of the 256-point generator, gen_rbergomi_pair.cuh; restated in tests/pair_stream.py):
//                ctr = (f_lo, f_hi, m>>1, 4) -> spectral normals G_m of transform f = 32 (g >> 6) + (g & 31), shared by paths
//                g and g ^ 32;  ctr = (g_lo, g_hi, 4 (k & 15) + (k >> 6), 6) -> W_k, lane (k >> 4) & 3
// with g the GLOBAL path id and key = 64-bit seed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MCP_PHILOX_M0 0xD2511F53u
#define MCP_PHILOX_M1 0xCD9E8D57u
#define MCP_PHILOX_W0 0x9E3779B9u
#define MCP_PHILOX_W1 0xBB67AE85u

// The ten round keys depend only on the seed: computed once on the host and passed by value in the kernel
// parameter block, so every LOP3 of the rounds takes its key straight from the constant bank.
struct PhiloxKeys {
    uint32_t k0[10];
    uint32_t k1[10];
};

static inline PhiloxKeys philox_make_keys(uint64_t seed) {
    PhiloxKeys K;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        K.k0[r] = a;
        K.k1[r] = b;
        a += MCP_PHILOX_W0;
        b += MCP_PHILOX_W1;
    }
    return K;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& K) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)MCP_PHILOX_M0 * c0;  // one IMAD.WIDE.U32 each
        const unsigned long long p1 = (unsigned long long)MCP_PHILOX_M1 * c2;
        c0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k0[r];
        c1 = (uint32_t)p1;
        c2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k1[r];
        c3 = (uint32_t)p0;
    }
    return make_uint4(c0, c1, c2, c3);
}

// Box-Muller in fp32: u1 = (a + 0.5) 2^-32 in (0,1], angle = 2 pi (b + 0.5) 2^-32.
// lg2.approx / sin.approx / cos.approx on the SFU pipe: |error| ~ 1e-6 absolute on a normal, immaterial
// next to Monte-Carlo noise and far inside the 1e-5 path tolerance (injected-draw parity never runs this).
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
    const float u1 = fmaf(__uint2float_rn(a), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float th = fmaf(__uint2float_rn(b), 1.4629180792671596e-09f, 7.314590396335798e-10f);  // 2pi * 2^-32 (b + 0.5)
    float lg, rad;  // sqrt(-2 ln u1), -2 ln u1 = -2 ln2 log2 u1; u1 >= 2^-33 is never sub-normal, so the .ftz forms are exact
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u1));  // one MUFU (__log2f adds a 3-instruction sub-normal guard)
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(-1.3862943611198906f * lg));
    float s, c;
    __sincosf(th, &s, &c);
    z0 = rad * c;
    z1 = rad * s;
}
