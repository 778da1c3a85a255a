// gen_rbergomi.cu -- rough-volatility (spectral "fGn") price paths, written time-major into HBM.
//
// Replaces RoughVolatility::GenerateStockPricePaths' hot loop (src/models/RoughVolatility.cpp:346-365):
//   Z_k complex normals                       :347 (genComplexGaussians :238-250)
//   A = phi (.) Z, zero-pad, DFT-, /M'        :348 (fractionalGaussian :264-292, fft :171-202)
//   X = sqrt(2H) eta Re(A)                    :284-291
//   v = xi exp(X - eta^2 t^{2H} / 2)          :349 (forwardVariance :294-309)
//   S_j = S_{j-1} exp((r - v/2) dt + sqrt(max(0,v)) sqrt(dt) (rho W1 + sqrt(1-rho^2) W2))   :354-364
//
// B200 design (not a translation of the serial per-path loop):
//   * one CTA owns a tile of TP consecutive paths; lane <-> path, so every shared-memory access of the
//     batched FFT is bank-conflict free, every twiddle / phi / compensator operand is warp-uniform, and every
//     global store is one full 128 B line of consecutive paths at one time index (time-major slab);
//   * normals come from Philox4x32-10 keyed by (global path id, step): one call = the four normals a
//     path-step consumes (Zre, Zim, W1, W2) -- nothing is carried between steps or paths;
//   * the M'-point complex DFT runs in shared memory as radix-16 (+ one radix-8/4/2) decimation-in-frequency
//     passes; output is left digit-reversed and read back through a position table, so there is no
//     reordering pass.  sqrt(2H) eta / M' and log2(e) are folded into the phi table on the host;
//   * the benchmark shape (128 < n <= 256 steps) has its own kernels with packed fp32x2 arithmetic: injected draws ->
//     two paths per thread, one transform per path (gen_rbergomi_x2.cuh); native Philox -> one transform per PAIR of
//     paths (gen_rbergomi_pair.cuh: same law, a third fewer normals, half the transforms); this file's kernel is the
//     generic one (any n <= 4096);
//   * the price recursion is a log-space prefix sum: per-thread serial chunk + one cross-chunk offset, then
//     S = S0 exp2(.).  fp32 throughout: measured |rel err| vs the fp64 oracle ~3e-7 (tolerance 1e-5).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"
#include "philox.cuh"
#include "transpose.cuh"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace {

constexpr int NT = 512;  // threads per CTA: 16 warps share one 32-path tile; 2 CTAs/SM -> 32 resident warps

struct RbParams {
    float S0, rd2, nkq, lsq, rho, rho_c;  // r dt log2e, -1/(2 log2e), log2(sqrt(dt) log2e), rho, sqrt(1-rho^2)
    int n;        // steps
    int Mp;       // DFT length = nextPow2(n)
    int lgMp;     // log2(Mp)
    int n_stage;  // DIF stages
    int lg_radix; // 3 bits per stage: log2(radix) of stage s at bits [3s, 3s+3)
    int64_t n_paths, ld;
    uint64_t path_offset;
    int64_t ld_draws;  // row stride of the slot-major draw tables
};

__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct RbParams;
__device__ __forceinline__ float log2_increment(float e, float w, const RbParams& P);

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// complex add / subtract as ONE packed fp32x2 instruction (Blackwell FADD2 / FFMA2): a float2 lives in an aligned
// register pair, so the two halves of a butterfly cost one issue slot instead of two
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

// forward (e^{-i theta}) 4-point DFT in place
__device__ __forceinline__ void dft4(float2& c0, float2& c1, float2& c2, float2& c3) {
    const float2 e0 = cadd(c0, c2), e1 = csub(c0, c2), o0 = cadd(c1, c3), d = csub(c1, c3);
    c0 = cadd(e0, o0);
    c2 = csub(e0, o0);
    c1 = make_float2(e1.x + d.y, e1.y - d.x);  // e1 + (-i) d : the rotation swaps the halves, scalar adds
    c3 = make_float2(e1.x - d.y, e1.y + d.x);  // e1 - (-i) d
}

// forward 8-point DFT: y[s] = sum_q x[q] w8^{qs}; result returned in natural order in x[]
__device__ __forceinline__ void dft8(float2 (&x)[8]) {
    const float h = 0.70710678118654752f;
    float2 a0 = cadd(x[0], x[4]), a1 = cadd(x[1], x[5]), a2 = cadd(x[2], x[6]), a3 = cadd(x[3], x[7]);
    float2 b0 = csub(x[0], x[4]), b1 = csub(x[1], x[5]), b2 = csub(x[2], x[6]), b3 = csub(x[3], x[7]);
    b1 = make_float2((b1.x + b1.y) * h, (b1.y - b1.x) * h);    // * w8
    b2 = mul_mi(b2);                                           // * w8^2
    b3 = make_float2((b3.y - b3.x) * h, -(b3.x + b3.y) * h);   // * w8^3
    dft4(a0, a1, a2, a3);  // even outputs 0,2,4,6
    dft4(b0, b1, b2, b3);  // odd outputs 1,3,5,7
    x[0] = a0; x[2] = a1; x[4] = a2; x[6] = a3;
    x[1] = b0; x[3] = b1; x[5] = b2; x[7] = b3;
}

// forward 16-point DFT as 4 x 4:  q = q1 + 4 q2,  s = 4 s1 + s2,
//   w16^{qs} = w4^{q2 s2} * w16^{q1 s2} * w4^{q1 s1}.  Output X[4 s1 + s2] is left in x[4 s2 + s1].
__device__ __forceinline__ void dft16_transposed(float2 (&x)[16]) {
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1) dft4(x[q1], x[q1 + 4], x[q1 + 8], x[q1 + 12]);  // x[q1 + 4 s2] = Y[q1][s2]
    // twiddles w16^{q1 s2}
    x[5] = cmul(x[5], make_float2(c1, -s1));                             // e = 1
    x[9] = make_float2((x[9].x + x[9].y) * h, (x[9].y - x[9].x) * h);    // e = 2
    x[13] = cmul(x[13], make_float2(s1, -c1));                           // e = 3
    x[6] = make_float2((x[6].x + x[6].y) * h, (x[6].y - x[6].x) * h);    // e = 2
    x[10] = mul_mi(x[10]);                                               // e = 4
    x[14] = make_float2((x[14].y - x[14].x) * h, -(x[14].x + x[14].y) * h);  // e = 6
    x[7] = cmul(x[7], make_float2(s1, -c1));                             // e = 3
    x[11] = make_float2((x[11].y - x[11].x) * h, -(x[11].x + x[11].y) * h);  // e = 6
    x[15] = cmul(x[15], make_float2(-c1, s1));                           // e = 9
#pragma unroll
    for (int s2 = 0; s2 < 4; ++s2) dft4(x[4 * s2], x[4 * s2 + 1], x[4 * s2 + 2], x[4 * s2 + 3]);
}

// index in x[] that holds output s after dftR
template <int R>
__device__ __forceinline__ constexpr int out_slot(int s) { return R == 16 ? 4 * (s & 3) + (s >> 2) : s; }

template <int R>
__device__ __forceinline__ void dftR(float2 (&x)[R]) {
    if constexpr (R == 16) {
        dft16_transposed(x);
    } else if constexpr (R == 8) {
        dft8(x);
    } else if constexpr (R == 4) {
        dft4(x[0], x[1], x[2], x[3]);
    } else {
        const float2 u = x[0];
        x[0] = cadd(u, x[1]);
        x[1] = csub(u, x[1]);
    }
}

// Log2-increment of one step from e = log2 v = X log2e + log2 xi - eta^2 t^{2H} log2e / 2 (tables carry the constants).
// One SFU op: u = sqrt(v) sqrt(dt) log2e = 2^{e/2 + lsq}  (v > 0 always, so the reference's max(0, v) is vacuous, :362);
// then v dt log2e / 2 = u^2 kq with kq = 1 / (2 log2e), and
//   d = ((r - v/2) dt + sqrt(v) sqrt(dt) dW) log2e = rd2 + u (w - kq u)                     RoughVolatility.cpp:356-363
__device__ __forceinline__ float log2_increment(float e, float w, const RbParams& P) {
    const float u = fast_ex2(fmaf(e, 0.5f, P.lsq));
    return fmaf(u, fmaf(u, P.nkq, w), P.rd2);
}

// One DIF pass of radix R = 2^LGR over sub-transforms of length L = 2^lgL, for the TP paths of the tile.
//   inputs  x_q = A[base + q*L/R],  outputs  y_s * w_L^{j s}  back to A[base + s*L/R].  All index math is shifts.
template <int LGR, int TP>
__device__ __forceinline__ void dif_pass(float2* __restrict__ A, const float2* __restrict__ tw, int lgMp, int lgL, int g, int p) {
    constexpr int R = 1 << LGR, G = NT / TP;
    const int lgS = lgL - LGR, stride = 1 << lgS, Mp = 1 << lgMp;
    for (int bf = g; bf < (Mp >> LGR); bf += G) {
        const int blk = bf >> lgS, j = bf & (stride - 1);
        float2* a = A + (size_t)((blk << lgL) + j) * TP + p;
        float2 x[R];
#pragma unroll
        for (int q = 0; q < R; ++q) x[q] = a[(size_t)(q << lgS) * TP];
        dftR<R>(x);
        if (lgS > 0) {
            const int jt = j << (lgMp - lgL);
#pragma unroll
            for (int s = 1; s < R; ++s) x[out_slot<R>(s)] = cmul(x[out_slot<R>(s)], tw[(jt * s) & (Mp - 1)]);
        }
#pragma unroll
        for (int s = 0; s < R; ++s) a[(size_t)(s << lgS) * TP] = x[out_slot<R>(s)];
    }
}

// LAST pass (sub-transform length == radix, no twiddles).  Only Re(X) is ever used (RoughVolatility.cpp:277-281),
// and this is the one place where every output X_m is in registers, so the variance and the log-increment
//   v_m = xi exp(X_m - eta^2 t_m^{2H}/2),  d_m = (r - v_m/2) dt + sqrt(max(0,v_m)) sqrt(dt) dW_m      (:294-309, :356-363)
// are formed right here and written over dW_m.  rev[] maps a storage position to its output index m.
template <int LGR, int TP>
__device__ __forceinline__ void last_pass(const float2* __restrict__ A, float* __restrict__ W, const int* __restrict__ rev,
                                          const float* __restrict__ comp2, const RbParams& P, int g, int p) {
    constexpr int R = 1 << LGR, G = NT / TP;
    for (int bf = g; bf < (P.Mp >> LGR); bf += G) {
        const float2* a = A + (size_t)(bf << LGR) * TP + p;
        float2 x[R];
#pragma unroll
        for (int q = 0; q < R; ++q) x[q] = a[(size_t)q * TP];
        dftR<R>(x);
#pragma unroll
        for (int s = 0; s < R; ++s) {
            const int m = rev[(bf << LGR) + s];
            if (m < P.n) {
                float* w = W + (size_t)m * TP + p;
                *w = log2_increment(x[out_slot<R>(s)].x + comp2[m], *w, P);
            }
        }
    }
}

// Normals of one (path, step) in the native stream -- generic (slow) form, used for ragged chunks only.
__device__ __forceinline__ void native_normals(uint32_t c0, uint32_t c1, int k, const PhiloxKeys& K, float& zr, float& zi, float& w) {
    const uint4 xz = philox4x32_10(c0, c1, (uint32_t)(k >> 1), 0u, K);
    if (k & 1) box_muller(xz.z, xz.w, zr, zi); else box_muller(xz.x, xz.y, zr, zi);
    const uint4 xw = philox4x32_10(c0, c1, (uint32_t)(k >> 2), 2u, K);
    float a, b;
    if (k & 2) box_muller(xw.z, xw.w, a, b); else box_muller(xw.x, xw.y, a, b);
    w = (k & 1) ? b : a;
}

// Dynamic shared memory carve-up (per CTA):
//   float2 A[Mp][TP] | float W[Mp][TP] | float tot[G][TP] | float2 phis[Mp] | float2 tw[Mp] | float comp2[Mp] | int rev[Mp]
template <int TP, bool INJECT, bool DUMP>
__device__ __forceinline__ void rbergomi_tiles(const RbParams& P, const PhiloxKeys& K, const float2* __restrict__ g_phis,
                                               const float2* __restrict__ g_tw, const float* __restrict__ g_comp2, const int* __restrict__ g_rev,
                                               const float* __restrict__ draws_in, float* __restrict__ draws_out, float* __restrict__ out,
                                               int64_t tile0, int64_t tile_stride) {
    constexpr int G = NT / TP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Mp = P.Mp, n = P.n;
    float2* A = reinterpret_cast<float2*>(smem_raw);
    float* W = reinterpret_cast<float*>(A + (size_t)Mp * TP);
    float* tot = W + (size_t)Mp * TP;
    float2* phis = reinterpret_cast<float2*>(tot + G * TP);
    float2* tw = phis + Mp;
    float* comp2 = reinterpret_cast<float*>(tw + Mp);
    int* rev = reinterpret_cast<int*>(comp2 + Mp);

    const int tid = threadIdx.x, p = tid % TP, g = tid / TP;
    for (int i = tid; i < Mp; i += NT) {
        phis[i] = i < n ? g_phis[i] : make_float2(0.f, 0.f);
        tw[i] = g_tw[i];
        comp2[i] = i < n ? g_comp2[i] : 0.f;
        rev[i] = g_rev[i];
    }
    const int CH = Mp >= G ? Mp / G : 1;  // contiguous time chunk owned by this thread
    const int k0 = g * CH, k1 = min(k0 + CH, Mp);
    const bool has_chunk = k0 < Mp;
    const int64_t n_tiles = (P.n_paths + TP - 1) / TP;
    __syncthreads();

    for (int64_t tile = tile0; tile < n_tiles; tile += tile_stride) {
        const int64_t path = tile * TP + p;  // local path index in this slab
        const bool live = path < P.n_paths;
        const uint64_t gid = P.path_offset + (uint64_t)path;
        const uint32_t c0 = (uint32_t)gid, c1 = (uint32_t)(gid >> 32);

        // ---- phase 1: normals -> A = phis (.) Z (zero padded), W = dW mix --------------------------------
        if (has_chunk) {
            if (!INJECT && (CH & 3) == 0) {
                // native stream, 4 steps at a time: 2 Philox calls -> 4 complex Z, 1 call -> 4 W
                for (int kq = k0; kq < k1; kq += 4) {
                    const uint4 xa = philox4x32_10(c0, c1, (uint32_t)(kq >> 1), 0u, K);
                    const uint4 xb = philox4x32_10(c0, c1, (uint32_t)(kq >> 1) + 1u, 0u, K);
                    const uint4 xw = philox4x32_10(c0, c1, (uint32_t)(kq >> 2), 2u, K);
                    float z[8], w[4];
                    box_muller(xa.x, xa.y, z[0], z[1]);
                    box_muller(xa.z, xa.w, z[2], z[3]);
                    box_muller(xb.x, xb.y, z[4], z[5]);
                    box_muller(xb.z, xb.w, z[6], z[7]);
                    box_muller(xw.x, xw.y, w[0], w[1]);
                    box_muller(xw.z, xw.w, w[2], w[3]);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int k = kq + t;
                        const bool in = k < n;
                        if (DUMP && live && in) {
                            draws_out[(int64_t)(2 * k) * P.ld_draws + path] = z[2 * t];
                            draws_out[(int64_t)(2 * k + 1) * P.ld_draws + path] = z[2 * t + 1];
                            draws_out[(int64_t)(2 * n + k) * P.ld_draws + path] = P.rho * w[t];
                            draws_out[(int64_t)(3 * n + k) * P.ld_draws + path] = P.rho_c * w[t];
                        }
                        A[(size_t)k * TP + p] = in ? cmul(phis[k], make_float2(z[2 * t], z[2 * t + 1])) : make_float2(0.f, 0.f);
                        W[(size_t)k * TP + p] = in ? w[t] : 0.f;
                    }
                }
            } else {
                for (int k = k0; k < k1; ++k) {
                    float2 a = make_float2(0.f, 0.f);
                    float w = 0.f;
                    if (k < n) {
                        float zr, zi;
                        if (INJECT) {
                            const int64_t col = live ? path : 0;
                            zr = draws_in[(int64_t)(2 * k) * P.ld_draws + col];
                            zi = draws_in[(int64_t)(2 * k + 1) * P.ld_draws + col];
                            w = P.rho * draws_in[(int64_t)(2 * n + k) * P.ld_draws + col] +
                                P.rho_c * draws_in[(int64_t)(3 * n + k) * P.ld_draws + col];  // RoughVolatility.cpp:356-358
                        } else {
                            native_normals(c0, c1, k, K, zr, zi, w);
                            if (DUMP && live) {
                                draws_out[(int64_t)(2 * k) * P.ld_draws + path] = zr;
                                draws_out[(int64_t)(2 * k + 1) * P.ld_draws + path] = zi;
                                draws_out[(int64_t)(2 * n + k) * P.ld_draws + path] = P.rho * w;
                                draws_out[(int64_t)(3 * n + k) * P.ld_draws + path] = P.rho_c * w;
                            }
                        }
                        a = cmul(phis[k], make_float2(zr, zi));
                    }
                    A[(size_t)k * TP + p] = a;
                    W[(size_t)k * TP + p] = w;
                }
            }
        }
        __syncthreads();

        // ---- phase 2: M'-point forward DFT in shared memory; the last pass also forms the log-increments ----
        {
            int lgL = P.lgMp;
            for (int s = 0; s < P.n_stage; ++s) {
                const int lgR = (P.lg_radix >> (3 * s)) & 7;
                if (s + 1 < P.n_stage) {
                    if (lgR == 4) dif_pass<4, TP>(A, tw, P.lgMp, lgL, g, p);
                    else if (lgR == 3) dif_pass<3, TP>(A, tw, P.lgMp, lgL, g, p);
                    else if (lgR == 2) dif_pass<2, TP>(A, tw, P.lgMp, lgL, g, p);
                    else dif_pass<1, TP>(A, tw, P.lgMp, lgL, g, p);
                } else {
                    if (lgR == 4) last_pass<4, TP>(A, W, rev, comp2, P, g, p);
                    else if (lgR == 3) last_pass<3, TP>(A, W, rev, comp2, P, g, p);
                    else if (lgR == 2) last_pass<2, TP>(A, W, rev, comp2, P, g, p);
                    else last_pass<1, TP>(A, W, rev, comp2, P, g, p);
                }
                lgL -= lgR;
                __syncthreads();
            }
            if (P.n_stage == 0) {  // n == 1: the transform is the identity
                if (tid < TP) {
                    W[p] = log2_increment(A[p].x + comp2[0], W[p], P);
                }
                __syncthreads();
            }
        }

        // ---- phase 3: log2-space prefix sum over time: chunk-local scan + cross-chunk offset, S = S0 2^(.) ----------
        if (CH == 16 && G * 16 == Mp) {
            // the common shape (252 steps: 16 warps x 16 steps): the chunk stays in registers between scan and store
            float c[16];
            const float* w = W + (size_t)k0 * TP + p;
#pragma unroll
            for (int t = 0; t < 16; ++t) c[t] = w[(size_t)t * TP];  // rows >= n hold 0
#pragma unroll
            for (int t = 1; t < 16; ++t) c[t] += c[t - 1];
            tot[g * TP + p] = c[15];
            __syncthreads();
            float off = 0.f;
            for (int gg = 0; gg < g; ++gg) off += tot[gg * TP + p];
            if (live) {
                if (g == 0) out[path] = P.S0;
                float* o = out + (int64_t)(k0 + 1) * P.ld + path;
#pragma unroll
                for (int t = 0; t < 16; ++t, o += P.ld)
                    if (k0 + t < n) *o = P.S0 * fast_ex2(off + c[t]);
            }
        } else {
            float run = 0.f;
            if (has_chunk) {
                float* w = W + (size_t)k0 * TP + p;
                const int kend = min(k1, n);
#pragma unroll 4
                for (int k = k0; k < kend; ++k, w += TP) {
                    run += *w;
                    *w = run;
                }
            }
            tot[g * TP + p] = run;
            __syncthreads();
            float off = 0.f;
            for (int gg = 0; gg < g; ++gg) off += tot[gg * TP + p];
            if (live) {
                if (g == 0) out[path] = P.S0;
                if (has_chunk) {
                    float* o = out + (int64_t)(k0 + 1) * P.ld + path;
                    const float* w = W + (size_t)k0 * TP + p;
                    const int kend = min(k1, n);
#pragma unroll 4
                    for (int k = k0; k < kend; ++k, o += P.ld, w += TP) *o = P.S0 * fast_ex2(off + *w);
                }
            }
        }
        __syncthreads();  // A / W / tot are rewritten by the next tile
    }
}

template <int TP, bool INJECT, bool DUMP>
__global__ void __launch_bounds__(NT, 2) rbergomi_paths_kernel(RbParams P, PhiloxKeys K, const float2* __restrict__ g_phis,
                                                              const float2* __restrict__ g_tw, const float* __restrict__ g_comp2,
                                                              const int* __restrict__ g_rev, const float* __restrict__ draws_in,
                                                              float* __restrict__ draws_out, float* __restrict__ out) {
    rbergomi_tiles<TP, INJECT, DUMP>(P, K, g_phis, g_tw, g_comp2, g_rev, draws_in, draws_out, out, blockIdx.x, gridDim.x);
}

// Batched rows (rows.cu): blockIdx.y = row, every row with its own parameters, tables (phis | tw | comp2 | rev, packed
// like the single-row scratch) and slab; blockIdx.x strides over the row's 32-path tiles.
struct RbRow {
    RbParams P;
    int64_t table_off;  // bytes into `tables`
    int64_t slab_off;   // floats into `slabs`
};

__global__ void __launch_bounds__(NT, 2) rbergomi_rows_kernel(const RbRow* __restrict__ rows, PhiloxKeys K, const unsigned char* __restrict__ tables,
                                                             float* __restrict__ slabs) {
    const RbRow& R = rows[blockIdx.y];
    const int Mp = R.P.Mp;
    const unsigned char* t = tables + R.table_off;
    const float2* phis = reinterpret_cast<const float2*>(t);
    const float2* tw = phis + Mp;
    const float* comp2 = reinterpret_cast<const float*>(tw + Mp);
    const int* rev = reinterpret_cast<const int*>(comp2 + Mp);
    rbergomi_tiles<32, false, false>(R.P, K, phis, tw, comp2, rev, nullptr, nullptr, slabs + R.slab_off, blockIdx.x, gridDim.x);
}

#include "gen_rbergomi_x2.cuh"
#include "gen_rbergomi_pair.cuh"

size_t smem_bytes(int Mp, int TP) {
    const int G = NT / TP;
    return (size_t)Mp * TP * 8 + (size_t)Mp * TP * 4 + (size_t)G * TP * 4 + (size_t)Mp * (8 + 8 + 4 + 4);
}

int next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

template <int TP>
int launch_tp(mcp_ctx* ctx, const RbParams& P, const PhiloxKeys& K, const float2* phis, const float2* tw, const float* comp2,
              const int* pos, const float* din, float* dout, float* out) {
    const size_t smem = smem_bytes(P.Mp, TP);
    const bool inject = din != nullptr, dump = dout != nullptr;
    void (*kern)(RbParams, PhiloxKeys, const float2*, const float2*, const float*, const int*, const float*, float*, float*);
    if (inject) kern = dump ? rbergomi_paths_kernel<TP, true, true> : rbergomi_paths_kernel<TP, true, false>;
    else kern = dump ? rbergomi_paths_kernel<TP, false, true> : rbergomi_paths_kernel<TP, false, false>;
    int occ = 0;
    MCP_TRY(mcp_kernel_config(ctx, (const void*)kern, NT, smem, &occ));
    if (occ < 1) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "rbergomi: kernel does not fit (smem %zu)", smem);
    const int64_t n_tiles = (P.n_paths + TP - 1) / TP;
    int64_t grid = (int64_t)ctx->sm_count * occ;
    if (grid > n_tiles) grid = n_tiles;
    kern<<<(unsigned)grid, NT, smem, ctx->stream>>>(P, K, phis, tw, comp2, pos, din, dout, out);
    MCP_LAUNCH_CHECK(ctx);
    return MCP_OK;
}

}  // namespace

// Process-wide constant tables for the host-side table builder (thread-safe lazy initialisation; engines on
// different host threads share them).
namespace {
struct UnitRoots {
    std::once_flag once;
    std::vector<double> c, s;  // cos / sin (2 pi q / 2^lg), q < 2^lg
};
UnitRoots g_unit_roots[16];
const UnitRoots& unit_roots(int lg) {
    UnitRoots& t = g_unit_roots[lg];
    std::call_once(t.once, [&] {
        const int M = 1 << lg;
        t.c.resize((size_t)M);
        t.s.resize((size_t)M);
        for (int q = 0; q < M; ++q) {
            const double ang = 2.0 * M_PI * (double)q / (double)M;
            t.c[q] = cos(ang);
            t.s[q] = sin(ang);
        }
    });
    return t;
}
const std::vector<double>& log_table() {  // ln i, i <= 4096 (entry 0 unused)
    static const std::vector<double> t = [] {
        std::vector<double> v(4097, 0.0);
        for (int i = 1; i <= 4096; ++i) v[i] = log((double)i);
        return v;
    }();
    return t;
}
}  // namespace

// Host-side tables, all in double then rounded once to fp32.
//   phi = DFT+(zero-pad(0.5 t^{2H}) to M = nextPow2(n+1))           RoughVolatility.cpp:212-236
//   phis_k = phi_k * sqrt(2H) eta / M' * log2(e)                    (:270 pads to M' = nextPow2(n); :284 scale; :198-200 1/M')
//   comp2_k = -0.5 eta^2 t_k^{2H} * log2(e) + log2(xi)              :304 (xi folded into the exponent)
int mcp_rbergomi_tables(int n, double H, double eta, double dt, double xi, std::vector<float>& phis, std::vector<float>& tw,
                        std::vector<float>& comp2, std::vector<int>& rev, int* Mp_out, int* lgMp_out, int* lg_radix_out,
                        int* n_stage) {
    const int M = next_pow2(n + 1), Mp = next_pow2(n);
    const double log2e = 1.4426950408889634074;
    const double log2_xi = xi > 0.0 ? log2(xi) : -INFINITY;  // xi = 0: v = 0 exactly, as xi exp(.) gives
    // t_i^{2H} = exp(2H (ln i + ln dt)), i >= 1: one exp per grid point, shared by lambda and the compensator
    // (the row driver builds these tables for thousands of rows per call: libm pow / cos / sin per element and an
    // O(n^2) DFT made it host-bound)
    const std::vector<double>& ln_i = log_table();
    const double ln_dt = log(dt);
    std::vector<double> t2h((size_t)n + 1);
    t2h[0] = pow(0.0, 2.0 * H);  // 1 for H = 0, else 0
    for (int i = 1; i <= n; ++i) t2h[i] = exp(2.0 * H * (ln_i[i] + ln_dt));
    const double scale = sqrt(2.0 * H) * eta / (double)Mp * log2e;
    phis.assign((size_t)2 * Mp, 0.f);
    int lgM = 0;
    while ((1 << lgM) < M) ++lgM;
    const UnitRoots& rt = unit_roots(lgM);  // e^{+2 pi i q / M}, q < M: every term of the DFT below is one of these
    if ((int64_t)n * (n + 1) <= (int64_t)6 * M * lgM) {
        // short rows: the direct sum is cheaper than a transform
        for (int k = 0; k < n; ++k) {
            double re = 0.0, im = 0.0;
            int idx = 0;  // (k * i) mod M, advanced incrementally (M is a power of two)
            for (int i = 0; i <= n; ++i) {
                re += 0.5 * t2h[i] * rt.c[idx];
                im += 0.5 * t2h[i] * rt.s[idx];
                idx = (idx + k) & (M - 1);
            }
            phis[2 * k] = (float)(re * scale);
            phis[2 * k + 1] = (float)(im * scale);
        }
    } else {
        // radix-2 decimation-in-time transform with the e^{+i theta} kernel (the reference's fft(., +1), RoughVolatility.cpp:171-202)
        std::vector<double> xr((size_t)M, 0.0), xi_((size_t)M, 0.0);
        for (int i = 0; i <= n; ++i) {
            int r = 0;
            for (int b = 0; b < lgM; ++b) r |= ((i >> b) & 1) << (lgM - 1 - b);
            xr[r] = 0.5 * t2h[i];
        }
        for (int len = 2; len <= M; len <<= 1) {
            const int half = len >> 1, stride = M / len;
            for (int base = 0; base < M; base += len)
                for (int j = 0; j < half; ++j) {
                    const double wr = rt.c[j * stride], wi = rt.s[j * stride];
                    const double ur = xr[base + j], ui = xi_[base + j];
                    const double vr = xr[base + j + half] * wr - xi_[base + j + half] * wi;
                    const double vi = xr[base + j + half] * wi + xi_[base + j + half] * wr;
                    xr[base + j] = ur + vr;
                    xi_[base + j] = ui + vi;
                    xr[base + j + half] = ur - vr;
                    xi_[base + j + half] = ui - vi;
                }
        }
        for (int k = 0; k < n; ++k) {
            phis[2 * k] = (float)(xr[k] * scale);
            phis[2 * k + 1] = (float)(xi_[k] * scale);
        }
    }
    int lgMp = 0;
    while ((1 << lgMp) < Mp) ++lgMp;
    const UnitRoots& rp = unit_roots(lgMp);
    tw.resize((size_t)2 * Mp);
    for (int q = 0; q < Mp; ++q) {  // e^{-2 pi i q / M'}
        tw[2 * q] = (float)rp.c[q];
        tw[2 * q + 1] = (float)(-rp.s[q]);
    }
    comp2.assign((size_t)Mp, 0.f);
    for (int k = 0; k < n; ++k) comp2[k] = (float)(-0.5 * eta * eta * t2h[k] * log2e + log2_xi);
    int lg = 0;
    while ((1 << lg) < Mp) ++lg;
    int radix[8], ns = 0, rem = lg, packed = 0;
    while (rem >= 4) { radix[ns++] = 16; rem -= 4; }
    if (rem == 3) radix[ns++] = 8;
    if (rem == 2) radix[ns++] = 4;
    if (rem == 1) radix[ns++] = 2;
    for (int s = 0; s < ns; ++s) packed |= (radix[s] == 16 ? 4 : radix[s] == 8 ? 3 : radix[s] == 4 ? 2 : 1) << (3 * s);
    *n_stage = ns;
    *lg_radix_out = packed;
    *lgMp_out = lg;
    // storage position of output m after the DIF passes: m = s1 + R1 (s2 + R2 (s3 ...)),  pos = s1 Mp/R1 + s2 Mp/(R1 R2) + ...
    // rev[pos] = m is what the last pass needs.
    rev.assign((size_t)Mp, 0);
    for (int m = 0; m < Mp; ++m) {
        int rest = m, L = Mp, q = 0;
        for (int s = 0; s < ns; ++s) {
            const int R = radix[s];
            L /= R;
            q += (rest % R) * L;
            rest /= R;
        }
        rev[q] = m;
    }
    *Mp_out = Mp;
    return 0;
}

// sqrt(w_m), w_m = (|phi_m|^2 + |phi_{M'-m}|^2) / 2 with phi_m = 0 for m >= n: the symmetrised spectrum of the pair
// stream (gen_rbergomi_pair.cuh)
static void pair_spectrum(const std::vector<float>& phis, int Mp, float* sw) {
    for (int m = 0; m < Mp; ++m) {
        const int mm = (Mp - m) & (Mp - 1);
        const double a = (double)phis[2 * m] * phis[2 * m] + (double)phis[2 * m + 1] * phis[2 * m + 1];
        const double b = (double)phis[2 * mm] * phis[2 * mm] + (double)phis[2 * mm + 1] * phis[2 * mm + 1];
        sw[m] = (float)sqrt(0.5 * (a + b));
    }
}

// The generator's host-side constant tables, exported for known-answer tests (pure host code: no device, no ctx).
extern "C" int mcp_rbergomi_host_tables(int n_steps, const mcp_rbergomi_params* prm, float* phis_out, float* comp2_out, float* sw_out) {
    if (!prm || n_steps < 1 || n_steps > 4096) return MCP_ERR_INVALID;
    if (!(prm->dt > 0.0) || !(prm->H >= 0.0) || !(prm->xi >= 0.0)) return MCP_ERR_DOMAIN;
    std::vector<float> phis, tw, comp2;
    std::vector<int> pos;
    int Mp = 0, lgMp = 0, lg_radix = 0, n_stage = 0;
    mcp_rbergomi_tables(n_steps, prm->H, prm->eta, prm->dt, prm->xi, phis, tw, comp2, pos, &Mp, &lgMp, &lg_radix, &n_stage);
    if (phis_out) memcpy(phis_out, phis.data(), (size_t)Mp * 8);
    if (comp2_out) memcpy(comp2_out, comp2.data(), (size_t)Mp * 4);
    if (sw_out) pair_spectrum(phis, Mp, sw_out);
    return Mp;
}

extern "C" int mcp_gen_rbergomi(mcp_ctx* ctx, mcp_pathset* ps, const mcp_rbergomi_params* prm, uint64_t seed, uint64_t path_offset,
                                const float* injected, float* dump) {
    if (!ctx || !ps || !prm) return MCP_ERR_INVALID;
    if (ps->ctx != ctx) return mcp_fail(ctx, MCP_ERR_INVALID, "rbergomi: pathset belongs to another ctx");
    if (ps->dtype != MCP_F32) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "rbergomi: generators write fp32 slabs");
    const int n = ps->n_steps;
    if (n < 1) return mcp_fail(ctx, MCP_ERR_INVALID, "rbergomi: n_steps must be >= 1");
    if (n > 4096) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "rbergomi: n_steps %d > 4096", n);
    if (!(prm->dt > 0.0) || !(prm->H >= 0.0) || !(fabs(prm->rho) <= 1.0) || !(prm->xi >= 0.0))
        return mcp_fail(ctx, MCP_ERR_DOMAIN, "rbergomi: need dt > 0, H >= 0, |rho| <= 1, xi >= 0");
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (injected && dump) {  // the normals used ARE the injected ones
        memcpy(dump, injected, (size_t)ps->n_paths * 4 * n * sizeof(float));
        dump = nullptr;
    }

    std::vector<float> phis, tw, comp2;
    std::vector<int> pos;  // rev[]: storage position -> output index
    RbParams P;
    memset(&P, 0, sizeof(P));
    mcp_rbergomi_tables(n, prm->H, prm->eta, prm->dt, prm->xi, phis, tw, comp2, pos, &P.Mp, &P.lgMp, &P.lg_radix, &P.n_stage);
    P.S0 = (float)prm->S0;
    const double log2e = 1.4426950408889634074;
    P.rd2 = (float)(prm->r * prm->dt * log2e);
    P.nkq = (float)(-0.5 / log2e);
    P.lsq = (float)log2(sqrt(prm->dt) * log2e);
    P.rho = (float)prm->rho;
    P.rho_c = (float)sqrt(1.0 - prm->rho * prm->rho);
    P.n = n;
    P.n_paths = ps->n_paths;
    P.ld = ps->ld;
    P.path_offset = path_offset;
    P.ld_draws = mcp_round_up(ps->n_paths, 32);

    // device tables (+ slot-major draw tables when injecting / dumping) live in the ctx scratch
    const int Mp = P.Mp;
    const size_t tab_bytes = (size_t)Mp * (8 + 8 + 4 + 4 + 4);
    const bool use_draws = injected || dump;
    const int64_t pc_max = use_draws ? (int64_t)((256u << 20) / ((size_t)4 * n * 4 * 2)) / 32 * 32 : 0;  // paths per draw chunk
    if (use_draws && pc_max < 32) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "rbergomi: draw staging too small for n=%d", n);
    const int64_t pc = use_draws ? (ps->n_paths < pc_max ? mcp_round_up(ps->n_paths, 32) : pc_max) : 0;
    const size_t draws_bytes = use_draws ? (size_t)4 * n * (size_t)pc * 4 : 0;
    MCP_TRY(mcp_scratch_reserve(ctx, mcp_round_up((int64_t)tab_bytes, 256) + 2 * draws_bytes));
    unsigned char* base = (unsigned char*)ctx->scratch;
    float2* d_phis = (float2*)base;
    float2* d_tw = d_phis + Mp;
    float* d_comp2 = (float*)(d_tw + Mp);
    int* d_pos = (int*)(d_comp2 + Mp);
    float* d_sw = (float*)(d_pos + Mp);  // symmetrised spectrum of the pair stream (gen_rbergomi_pair.cuh)
    float* d_slot = (float*)(base + mcp_round_up((int64_t)tab_bytes, 256));  // [4n][pc]  slot-major
    float* d_rows = d_slot + (size_t)4 * n * pc;                             // [pc][4n]  host order
    {   // the four tables are contiguous on the device: one copy through the pinned ring
        std::vector<unsigned char> pack(tab_bytes);
        memcpy(pack.data(), phis.data(), (size_t)Mp * 8);
        memcpy(pack.data() + (size_t)Mp * 8, tw.data(), (size_t)Mp * 8);
        memcpy(pack.data() + (size_t)Mp * 16, comp2.data(), (size_t)Mp * 4);
        memcpy(pack.data() + (size_t)Mp * 20, pos.data(), (size_t)Mp * 4);
        pair_spectrum(phis, Mp, (float*)(pack.data() + (size_t)Mp * 24));
        MCP_TRY(mcp_h2d(ctx, d_phis, pack.data(), tab_bytes));
    }

    const PhiloxKeys K = philox_make_keys(seed);
    int TP = 32;
    while (TP > 4 && smem_bytes(Mp, TP) > 200 * 1024) TP >>= 1;
    if (smem_bytes(Mp, TP) > 227 * 1024) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "rbergomi: n_steps %d needs too much shared memory", n);

    std::function<int(const RbParams&, const float*, float*, float*)> run = [&](const RbParams& Q, const float* din, float* dout, float* out) -> int {
        switch (TP) {
            case 32: return launch_tp<32>(ctx, Q, K, d_phis, d_tw, d_comp2, d_pos, din, dout, out);
            case 16: return launch_tp<16>(ctx, Q, K, d_phis, d_tw, d_comp2, d_pos, din, dout, out);
            case 8: return launch_tp<8>(ctx, Q, K, d_phis, d_tw, d_comp2, d_pos, din, dout, out);
            default: return launch_tp<4>(ctx, Q, K, d_phis, d_tw, d_comp2, d_pos, din, dout, out);
        }
    };

    // 256-point transforms (128 < n <= 256 steps) run the specialised kernels: injected draws -> two paths per thread, one
    // transform per path (gen_rbergomi_x2.cuh); native Philox -> one transform per PAIR of paths (gen_rbergomi_pair.cuh; its
    // own stream, same law).  MCP_GEN_IMPL=2 keeps the per-path stream on the x2 kernel, MCP_GEN_IMPL=0 forces the generic
    // kernel (per-path stream; used by tests that need bit-identical paths from the batched row driver).
    const char* impl_env = getenv("MCP_GEN_IMPL");
    const int impl = getenv("MCP_GEN_GENERIC") ? 0 : (impl_env && *impl_env ? atoi(impl_env) : 3);
    if (Mp == 256 && TP == 32 && impl >= 2) {
        run = [&, impl](const RbParams& Q, const float* din, float* dout, float* out) -> int {
            const bool inject = din != nullptr, dmp = dout != nullptr;
            if (!inject && impl >= 3) {
                void (*kern)(RbParams, PhiloxKeys, const float2*, const float*, const float2*, const float*, float*, float*) =
                    dmp ? rbergomi_paths_n256pair_kernel<true> : rbergomi_paths_n256pair_kernel<false>;
                int occ = 0;
                MCP_TRY(mcp_kernel_config(ctx, (const void*)kern, NT2, PAIR_SMEM, &occ));
                if (occ < 1) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "rbergomi: n256pair kernel does not fit");
                const int64_t n_tiles = (int64_t)(((Q.path_offset + (uint64_t)Q.n_paths + 63) >> 6) - (Q.path_offset >> 6));
                int64_t grid = (int64_t)ctx->sm_count * occ;
                if (grid > n_tiles) grid = n_tiles;
                kern<<<(unsigned)grid, NT2, PAIR_SMEM, ctx->stream>>>(Q, K, d_phis, d_sw, d_tw, d_comp2, dout, out);
                MCP_LAUNCH_CHECK(ctx);
                return MCP_OK;
            }
            void (*kern)(RbParams, PhiloxKeys, const float2*, const float2*, const float*, const float*, float*, float*);
            if (inject) kern = dmp ? rbergomi_paths_n256x2_kernel<true, true> : rbergomi_paths_n256x2_kernel<true, false>;
            else kern = dmp ? rbergomi_paths_n256x2_kernel<false, true> : rbergomi_paths_n256x2_kernel<false, false>;
            int occ = 0;
            MCP_TRY(mcp_kernel_config(ctx, (const void*)kern, NT2, X2_SMEM, &occ));
            if (occ < 1) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "rbergomi: n256x2 kernel does not fit");
            const int64_t n_tiles = (Q.n_paths + 31) / 32;
            int64_t grid = (int64_t)ctx->sm_count * occ;
            if (grid > n_tiles) grid = n_tiles;
            kern<<<(unsigned)grid, NT2, X2_SMEM, ctx->stream>>>(Q, K, d_phis, d_tw, d_comp2, din, dout, out);
            MCP_LAUNCH_CHECK(ctx);
            return MCP_OK;
        };
    }

    if (!use_draws) {
        if (ctx->profiling) cudaEventRecord(ctx->ev0, ctx->stream);
        MCP_TRY(run(P, nullptr, nullptr, (float*)ps->data));
        if (ctx->profiling) {
            cudaEventRecord(ctx->ev1, ctx->stream);
            MCP_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
            MCP_CUDA(ctx, cudaEventElapsedTime(&ctx->prof.gen_kernel_ms, ctx->ev0, ctx->ev1));
        }
        return MCP_OK;
    }

    // injected / dumped draws: stream the host [path][4n] table in chunks, transposing on the device
    const int slots = 4 * n;
    for (int64_t p0 = 0; p0 < ps->n_paths; p0 += pc) {
        const int64_t np = (ps->n_paths - p0 < pc) ? ps->n_paths - p0 : pc;
        RbParams Q = P;
        Q.n_paths = np;
        Q.path_offset = path_offset + (uint64_t)p0;
        Q.ld_draws = pc;
        if (injected) {
            MCP_CUDA(ctx, mcp_memcpy_async(ctx, d_rows, injected + (size_t)p0 * slots, (size_t)np * slots * 4, cudaMemcpyHostToDevice, ctx->stream));
            mcp_launch_transpose<float, float>(ctx->stream, d_rows, slots, np, slots, d_slot, pc);
            MCP_LAUNCH_CHECK(ctx);
        }
        // when both are requested the dump simply echoes the injected values (same table, in place)
        MCP_TRY(run(Q, injected ? d_slot : nullptr, dump ? d_slot : nullptr, (float*)ps->data + p0));
        if (dump) {
            mcp_launch_transpose<float, float>(ctx->stream, d_slot, pc, slots, np, d_rows, slots);
            MCP_LAUNCH_CHECK(ctx);
            MCP_CUDA(ctx, mcp_memcpy_async(ctx, dump + (size_t)p0 * slots, d_rows, (size_t)np * slots * 4, cudaMemcpyDeviceToHost, ctx->stream));
        }
        MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MCP_OK;
}

// Batched generation for the row driver (rows.cu): row r = its own model, step count and slab; normals of path i of
// row r are keyed by (seed, path_offset + r * n_paths + i) -- exactly what mcp_gen_rbergomi produces for that row
// when called with path_offset + r * n_paths.  One launch for all rows.
// bytes of staging (pinned host AND device, each) that mcp_rows_generate needs for a batch of n_rows rows of at most max_steps steps
size_t mcp_rows_stage_bytes(int n_rows, int max_steps) {
    return (size_t)mcp_round_up((int64_t)((size_t)n_rows * sizeof(RbRow)), 256) + (size_t)n_rows * (size_t)next_pow2(max_steps < 1 ? 1 : max_steps) * 24;
}

// host_stage / dev_stage (both of mcp_rows_stage_bytes, host one pinned): the row descriptors and tables are built straight into
// the pinned buffer and copied asynchronously -- the call returns with the copy and the kernel queued, nothing waits, so the
// caller can build the NEXT batch's tables while this one runs (rows.cu).  Both null: private buffers and a synchronous copy.
// ev_start (nullable) is recorded right before the first device operation of the batch.
int mcp_rows_generate(mcp_ctx* ctx, const mcp_rbergomi_params* models, size_t model_stride_bytes, const int* n_steps, int n_rows, int n_paths,
                      uint64_t seed, uint64_t path_offset, float* slabs, int64_t slab_stride, int64_t ld, void* host_stage, void* dev_stage,
                      size_t stage_bytes, cudaEvent_t ev_start) {
    std::vector<RbRow> rows_own;
    RbRow* rows = nullptr;
    if (host_stage) rows = (RbRow*)host_stage;
    else { rows_own.resize((size_t)n_rows); rows = rows_own.data(); }
    int max_Mp = 1, live_rows = 0;
    // pass 1 (serial, trivial): validate, place every row's tables in one staging buffer
    size_t tables_bytes = 0;
    for (int r = 0; r < n_rows; ++r) {
        const mcp_rbergomi_params* prm = (const mcp_rbergomi_params*)((const unsigned char*)models + (size_t)r * model_stride_bytes);
        RbRow& R = rows[(size_t)r];
        memset(&R, 0, sizeof(R));
        const int n = n_steps[r];
        if (n < 1) { R.P.n_paths = 0; R.P.Mp = 1; continue; }  // no tiles: the kernel's tile loop is empty
        if (n > 512) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "rows: n_steps %d > 512", n);
        // A degenerate model (the reference's unclamped DFA slope does go negative on real histories, RoughVolatility.cpp:120-149)
        // makes the reference write NaN paths for THAT row only (PredictionGen.cpp:753-777 rejects the row and goes on): such a
        // row gets no tiles here and the caller (mcp_price_rows) reports NaN for it; the other rows of the batch are priced.
        if (!(prm->dt > 0.0) || !(prm->H >= 0.0) || !(fabs(prm->rho) <= 1.0) || !(prm->xi >= 0.0)) { R.P.n_paths = 0; R.P.Mp = 1; continue; }
        const int Mp = next_pow2(n);
        R.P.Mp = Mp;
        R.table_off = (int64_t)tables_bytes;
        tables_bytes += (size_t)Mp * 24;
        if (Mp > max_Mp) max_Mp = Mp;
        ++live_rows;
    }
    const size_t rows_bytes = (size_t)n_rows * sizeof(RbRow);
    const size_t rows_span = (size_t)mcp_round_up((int64_t)rows_bytes, 256);
    std::vector<unsigned char> tables_own;
    unsigned char* tables = nullptr;
    if (host_stage) {
        if (rows_span + tables_bytes > stage_bytes) return mcp_fail(ctx, MCP_ERR_INVALID, "rows: staging buffer too small");
        tables = (unsigned char*)host_stage + rows_span;
    } else {
        tables_own.resize(tables_bytes);
        tables = tables_own.data();
    }
    // pass 2: the tables themselves (a few microseconds per row), split over a handful of host threads for large batches
    auto fill = [&](int r_begin, int r_end) {
        std::vector<float> phis, tw, comp2;
        std::vector<int> pos;
        for (int r = r_begin; r < r_end; ++r) {
            const int n = n_steps[r];
            if (n < 1) continue;
            const mcp_rbergomi_params* prm = (const mcp_rbergomi_params*)((const unsigned char*)models + (size_t)r * model_stride_bytes);
            if (!(prm->dt > 0.0) || !(prm->H >= 0.0) || !(fabs(prm->rho) <= 1.0) || !(prm->xi >= 0.0)) continue;  // degenerate: no tiles (pass 1)
            RbRow& R = rows[(size_t)r];
            mcp_rbergomi_tables(n, prm->H, prm->eta, prm->dt, prm->xi, phis, tw, comp2, pos, &R.P.Mp, &R.P.lgMp, &R.P.lg_radix, &R.P.n_stage);
            const int Mp = R.P.Mp;
            const double log2e = 1.4426950408889634074;
            R.P.S0 = (float)prm->S0;
            R.P.rd2 = (float)(prm->r * prm->dt * log2e);
            R.P.nkq = (float)(-0.5 / log2e);
            R.P.lsq = (float)log2(sqrt(prm->dt) * log2e);
            R.P.rho = (float)prm->rho;
            R.P.rho_c = (float)sqrt(1.0 - prm->rho * prm->rho);
            R.P.n = n;
            R.P.n_paths = n_paths;
            R.P.ld = ld;
            R.P.path_offset = path_offset + (uint64_t)r * (uint64_t)n_paths;
            R.P.ld_draws = 0;
            R.slab_off = (int64_t)r * slab_stride;
            unsigned char* o = tables + R.table_off;
            memcpy(o, phis.data(), (size_t)Mp * 8);
            memcpy(o + (size_t)Mp * 8, tw.data(), (size_t)Mp * 8);
            memcpy(o + (size_t)Mp * 16, comp2.data(), (size_t)Mp * 4);
            memcpy(o + (size_t)Mp * 20, pos.data(), (size_t)Mp * 4);
        }
    };
    {
        unsigned hw = std::thread::hardware_concurrency();
        int n_thr = (int)(hw ? hw : 1);
        if (n_thr > 8) n_thr = 8;
        if (n_thr > n_rows / 512) n_thr = n_rows / 512;  // small batches: not worth a thread
        if (n_thr <= 1) {
            fill(0, n_rows);
        } else {
            std::vector<std::thread> pool;
            const int per = (n_rows + n_thr - 1) / n_thr;
            int assigned = per < n_rows ? per : n_rows;  // [0, assigned) is this thread's share
            const int mine = assigned;
            try {
                for (int t = 1; t < n_thr && assigned < n_rows; ++t) {
                    const int end = assigned + per < n_rows ? assigned + per : n_rows;
                    pool.emplace_back(fill, assigned, end);
                    assigned = end;
                }
            } catch (...) {  // no thread to be had: the rest is done here (nothing may throw across the C ABI)
            }
            fill(0, mine);
            fill(assigned, n_rows);
            for (auto& th : pool) th.join();
        }
    }
    if (live_rows == 0) {
        if (ev_start) cudaEventRecord(ev_start, ctx->stream);
        return MCP_OK;
    }
    unsigned char* base = (unsigned char*)dev_stage;
    if (!base) {
        MCP_TRY(mcp_scratch_reserve(ctx, rows_span + tables_bytes));
        base = (unsigned char*)ctx->scratch;
    }
    if (ev_start) cudaEventRecord(ev_start, ctx->stream);
    if (host_stage) {  // one copy: descriptors and tables are contiguous in the pinned buffer
        MCP_CUDA(ctx, mcp_memcpy_async(ctx, base, host_stage, rows_span + tables_bytes, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        MCP_CUDA(ctx, mcp_memcpy_async(ctx, base, rows, rows_bytes, cudaMemcpyHostToDevice, ctx->stream));
        MCP_CUDA(ctx, mcp_memcpy_async(ctx, base + rows_span, tables, tables_bytes, cudaMemcpyHostToDevice, ctx->stream));
        MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // host vectors go out of scope
    }
    const size_t smem = smem_bytes(max_Mp, 32);
    if (smem > 227 * 1024) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "rows: transform of %d points does not fit", max_Mp);
    MCP_TRY(mcp_kernel_config(ctx, (const void*)rbergomi_rows_kernel, NT, smem, nullptr));
    const PhiloxKeys K = philox_make_keys(seed);
    const dim3 grid((unsigned)((n_paths + 31) / 32), (unsigned)n_rows);
    rbergomi_rows_kernel<<<grid, NT, smem, ctx->stream>>>((const RbRow*)base, K, base + rows_span, slabs);
    MCP_LAUNCH_CHECK(ctx);
    return MCP_OK;
}
