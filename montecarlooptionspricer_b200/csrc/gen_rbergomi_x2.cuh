// gen_rbergomi_x2.cuh -- the 256-point generator with TWO PATHS PER THREAD (included by gen_rbergomi.cu, inside its
// anonymous namespace).
//
// ncu on the one-path-per-thread kernel: 110 issued instructions per path-step, issue slots 62 % busy, FMA pipe only
// 42 % -- the kernel is bound by issue slots, not by arithmetic.  Blackwell's packed fp32x2 instructions (FFMA2 /
// FADD2 / FMUL2) do two fp32 operations per issue slot, so every thread now carries a PAIR of adjacent paths in the two
// halves of a float2: all floating-point work (Box-Muller scaling, phi (.) Z, the two radix-16 DFT passes, the
// log-increment, the scan, the final scaling) is issued once per pair; only Philox (integer) and the SFU calls stay
// per path.  Complex data is kept split (re[.][path], im[.][path]) so that a pair is one aligned float2 in shared
// memory and the -i rotations of the butterflies are register renames.
//
// CTA = 256 threads = 8 warps, tile = 32 paths = 16 pairs; lane l: pair l & 15, warp half l >> 4; time chunk
// g = 2 * warp + half (0..15) owns steps [16g, 16g+16).  Shared memory (112 KB, two CTAs per SM, 128 registers):
//   float re[256][32] | float im[256][32] | float W[256][32] | float tot[16][32] | Tw phis[256] | Tw tw2[16][16] | float2 comp[256]
#pragma once

constexpr int NT2 = 256;

struct C2 {  // one complex number for each of the two paths of a pair
    float2 re, im;
};
struct Tw {  // a complex constant shared by both paths, pre-splatted for the packed multiply
    float2 r, i, ni;  // (wr, wr), (wi, wi), (-wi, -wi)
};

constexpr int X2_SMEM = 3 * 256 * 32 * 4 + 16 * 32 * 4 + 2 * 256 * (int)sizeof(Tw) + 256 * 8;

__device__ __forceinline__ float2 f2splat(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

__device__ __forceinline__ C2 c2add(const C2& a, const C2& b) { return C2{f2add(a.re, b.re), f2add(a.im, b.im)}; }
__device__ __forceinline__ C2 c2sub(const C2& a, const C2& b) { return C2{f2sub(a.re, b.re), f2sub(a.im, b.im)}; }
// a * (wr + i wi):  re = a.re wr - a.im wi,  im = a.re wi + a.im wr
__device__ __forceinline__ C2 c2mul(const C2& a, const Tw& w) {
    return C2{f2fma(a.re, w.r, f2mul(a.im, w.ni)), f2fma(a.im, w.r, f2mul(a.re, w.i))};
}
__device__ __forceinline__ C2 c2mul_const(const C2& a, float wr, float wi) {
    const float2 r = f2splat(wr), i = f2splat(wi), ni = f2splat(-wi);
    return C2{f2fma(a.re, r, f2mul(a.im, ni)), f2fma(a.im, r, f2mul(a.re, i))};
}

// forward (e^{-i theta}) 4-point DFT in place
__device__ __forceinline__ void x2_dft4(C2& c0, C2& c1, C2& c2, C2& c3) {
    const C2 e0 = c2add(c0, c2), e1 = c2sub(c0, c2), o0 = c2add(c1, c3), d = c2sub(c1, c3);
    c0 = c2add(e0, o0);
    c2 = c2sub(e0, o0);
    c1 = C2{f2add(e1.re, d.im), f2sub(e1.im, d.re)};  // e1 + (-i) d
    c3 = C2{f2sub(e1.re, d.im), f2add(e1.im, d.re)};  // e1 - (-i) d
}

// forward 16-point DFT as 4 x 4 (q = q1 + 4 q2, s = 4 s1 + s2); output X[4 s1 + s2] is left in x[4 s2 + s1]
__device__ __forceinline__ void x2_dft16_transposed(C2 (&x)[16]) {
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1) x2_dft4(x[q1], x[q1 + 4], x[q1 + 8], x[q1 + 12]);
    // twiddles w16^{q1 s2} = cos(2 pi e / 16) - i sin(2 pi e / 16), e = q1 s2
    x[5] = c2mul_const(x[5], c1, -s1);                                                      // e = 1
    x[9] = C2{f2mul(f2add(x[9].re, x[9].im), f2splat(h)), f2mul(f2sub(x[9].im, x[9].re), f2splat(h))};      // e = 2: h (1 - i)
    x[13] = c2mul_const(x[13], s1, -c1);                                                    // e = 3
    x[6] = C2{f2mul(f2add(x[6].re, x[6].im), f2splat(h)), f2mul(f2sub(x[6].im, x[6].re), f2splat(h))};      // e = 2
    x[10] = C2{x[10].im, f2mul(x[10].re, f2splat(-1.f))};                                   // e = 4: -i
    x[14] = C2{f2mul(f2sub(x[14].im, x[14].re), f2splat(h)), f2mul(f2add(x[14].re, x[14].im), f2splat(-h))};  // e = 6: -h (1 + i)
    x[7] = c2mul_const(x[7], s1, -c1);                                                      // e = 3
    x[11] = C2{f2mul(f2sub(x[11].im, x[11].re), f2splat(h)), f2mul(f2add(x[11].re, x[11].im), f2splat(-h))};  // e = 6
    x[15] = c2mul_const(x[15], -c1, s1);                                                    // e = 9
#pragma unroll
    for (int s2 = 0; s2 < 4; ++s2) x2_dft4(x[4 * s2], x[4 * s2 + 1], x[4 * s2 + 2], x[4 * s2 + 3]);
}
__device__ __forceinline__ constexpr int x2_slot(int s) { return 4 * (s & 3) + (s >> 2); }

// Box-Muller for the two paths of a pair: (a, b) uniforms of path 0 and of path 1 -> z0 = (z0 of path 0, z0 of path 1), z1 likewise
__device__ __forceinline__ void box_muller_x2(uint32_t a0, uint32_t b0, uint32_t a1, uint32_t b1, float2& z0, float2& z1) {
    const float2 u1 = f2fma(make_float2(__uint2float_rn(a0), __uint2float_rn(a1)), f2splat(2.3283064365386963e-10f), f2splat(1.1641532182693481e-10f));
    const float2 th = f2fma(make_float2(__uint2float_rn(b0), __uint2float_rn(b1)), f2splat(1.4629180792671596e-09f), f2splat(7.314590396335798e-10f));
    float lx, ly, rx, ry;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lx) : "f"(u1.x));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(ly) : "f"(u1.y));
    const float2 t = f2mul(make_float2(lx, ly), f2splat(-1.3862943611198906f));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rx) : "f"(t.x));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(ry) : "f"(t.y));
    float sx, cx, sy, cy;
    __sincosf(th.x, &sx, &cx);
    __sincosf(th.y, &sy, &cy);
    const float2 rad = make_float2(rx, ry);
    z0 = f2mul(rad, make_float2(cx, cy));
    z1 = f2mul(rad, make_float2(sx, sy));
}

template <bool INJECT, bool DUMP>
__global__ void __launch_bounds__(NT2, 2) rbergomi_paths_n256x2_kernel(RbParams P, PhiloxKeys K, const float2* __restrict__ g_phis,
                                                                      const float2* __restrict__ g_tw, const float* __restrict__ g_comp2,
                                                                      const float* __restrict__ draws_in, float* __restrict__ draws_out,
                                                                      float* __restrict__ out) {
    constexpr int TP = 32, MP = 256;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Are = reinterpret_cast<float*>(smem_raw);
    float* Aim = Are + MP * TP;
    float* W = Aim + MP * TP;
    float* tot = W + MP * TP;
    Tw* phis = reinterpret_cast<Tw*>(tot + 16 * TP);
    Tw* tw2 = phis + MP;
    float2* comp = reinterpret_cast<float2*>(tw2 + MP);
    const int n = P.n;
    const int tid = threadIdx.x, pl = tid & 15, g = tid >> 4;  // g = 2 * warp + half
    {
        const float2 ph = tid < n ? g_phis[tid] : make_float2(0.f, 0.f);
        phis[tid] = Tw{f2splat(ph.x), f2splat(ph.y), f2splat(-ph.y)};
        const float2 w = g_tw[((tid >> 4) * (tid & 15)) & (MP - 1)];  // w256^{j s}, j = tid / 16, s = tid % 16
        tw2[tid] = Tw{f2splat(w.x), f2splat(w.y), f2splat(-w.y)};
        comp[tid] = f2splat(tid < n ? g_comp2[tid] : 0.f);
    }
    const int k0 = g * 16;
    const bool full = k0 + 16 <= n;  // only the last chunk can be ragged
    const int64_t n_tiles = (P.n_paths + TP - 1) / TP;
    const int col = 2 * pl;
    float2* const Rc = reinterpret_cast<float2*>(Are + k0 * TP + col);  // this thread's chunk, pair column (row stride TP floats = 16 float2)
    float2* const Ic = reinterpret_cast<float2*>(Aim + k0 * TP + col);
    float2* const Wc = reinterpret_cast<float2*>(W + k0 * TP + col);
    constexpr int RS = TP / 2;  // row stride in float2
    const float2 lsq2 = f2splat(P.lsq), nkq2 = f2splat(P.nkq), rd22 = f2splat(P.rd2), S02 = f2splat(P.S0), half2v = f2splat(0.5f);
    __syncthreads();

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t path = tile * TP + col;  // first path of the pair
        const bool live0 = path < P.n_paths, live1 = path + 1 < P.n_paths;
        const uint64_t gid0 = P.path_offset + (uint64_t)path, gid1 = gid0 + 1;
        const uint32_t a0 = (uint32_t)gid0, a1 = (uint32_t)(gid0 >> 32), b0 = (uint32_t)gid1, b1 = (uint32_t)(gid1 >> 32);

        // ---- phase 1: normals -> A = phis (.) Z, W = dW -----------------------------------------------------
#pragma unroll 1
        for (int kq = 0; kq < 16; kq += 4) {
            float2 zr[4], zi[4], w[4];
            if (INJECT) {
                const int64_t c0 = live0 ? path : 0, c1 = live1 ? path + 1 : 0;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int k = k0 + kq + t;
                    const bool in = k < n;
                    const float* d0 = draws_in + c0;
                    const float* d1 = draws_in + c1;
                    zr[t] = in ? make_float2(d0[(int64_t)(2 * k) * P.ld_draws], d1[(int64_t)(2 * k) * P.ld_draws]) : f2splat(0.f);
                    zi[t] = in ? make_float2(d0[(int64_t)(2 * k + 1) * P.ld_draws], d1[(int64_t)(2 * k + 1) * P.ld_draws]) : f2splat(0.f);
                    w[t] = in ? make_float2(P.rho * d0[(int64_t)(2 * n + k) * P.ld_draws] + P.rho_c * d0[(int64_t)(3 * n + k) * P.ld_draws],
                                            P.rho * d1[(int64_t)(2 * n + k) * P.ld_draws] + P.rho_c * d1[(int64_t)(3 * n + k) * P.ld_draws])
                              : f2splat(0.f);  // RoughVolatility.cpp:356-358
                }
            } else {
                const int kk = k0 + kq;
                const uint4 xa0 = philox4x32_10(a0, a1, (uint32_t)(kk >> 1), 0u, K), xa1 = philox4x32_10(b0, b1, (uint32_t)(kk >> 1), 0u, K);
                const uint4 xb0 = philox4x32_10(a0, a1, (uint32_t)(kk >> 1) + 1u, 0u, K), xb1 = philox4x32_10(b0, b1, (uint32_t)(kk >> 1) + 1u, 0u, K);
                const uint4 xw0 = philox4x32_10(a0, a1, (uint32_t)(kk >> 2), 2u, K), xw1 = philox4x32_10(b0, b1, (uint32_t)(kk >> 2), 2u, K);
                box_muller_x2(xa0.x, xa0.y, xa1.x, xa1.y, zr[0], zi[0]);
                box_muller_x2(xa0.z, xa0.w, xa1.z, xa1.w, zr[1], zi[1]);
                box_muller_x2(xb0.x, xb0.y, xb1.x, xb1.y, zr[2], zi[2]);
                box_muller_x2(xb0.z, xb0.w, xb1.z, xb1.w, zr[3], zi[3]);
                box_muller_x2(xw0.x, xw0.y, xw1.x, xw1.y, w[0], w[1]);
                box_muller_x2(xw0.z, xw0.w, xw1.z, xw1.w, w[2], w[3]);
                if (DUMP) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int k = kk + t;
                        if (k < n) {
                            if (live0) {
                                draws_out[(int64_t)(2 * k) * P.ld_draws + path] = zr[t].x;
                                draws_out[(int64_t)(2 * k + 1) * P.ld_draws + path] = zi[t].x;
                                draws_out[(int64_t)(2 * n + k) * P.ld_draws + path] = P.rho * w[t].x;
                                draws_out[(int64_t)(3 * n + k) * P.ld_draws + path] = P.rho_c * w[t].x;
                            }
                            if (live1) {
                                draws_out[(int64_t)(2 * k) * P.ld_draws + path + 1] = zr[t].y;
                                draws_out[(int64_t)(2 * k + 1) * P.ld_draws + path + 1] = zi[t].y;
                                draws_out[(int64_t)(2 * n + k) * P.ld_draws + path + 1] = P.rho * w[t].y;
                                draws_out[(int64_t)(3 * n + k) * P.ld_draws + path + 1] = P.rho_c * w[t].y;
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const bool in = full || (k0 + kq + t < n);
                const Tw ph = phis[k0 + kq + t];
                const C2 a = c2mul(C2{zr[t], zi[t]}, ph);  // phi (.) Z
                Rc[(kq + t) * RS] = in ? a.re : f2splat(0.f);
                Ic[(kq + t) * RS] = in ? a.im : f2splat(0.f);
                Wc[(kq + t) * RS] = in ? w[t] : f2splat(0.f);
            }
        }
        __syncthreads();

        // ---- phase 2a: DIF pass 1 on column g: elements g + 16 q, output s scaled by w256^{g s} -------------------
        {
            float2* ar = reinterpret_cast<float2*>(Are + g * TP + col);
            float2* ai = reinterpret_cast<float2*>(Aim + g * TP + col);
            C2 x[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) x[q] = C2{ar[q * 16 * RS], ai[q * 16 * RS]};
            x2_dft16_transposed(x);
            const Tw* t2 = tw2 + g * 16;
#pragma unroll
            for (int s = 1; s < 16; ++s) x[x2_slot(s)] = c2mul(x[x2_slot(s)], t2[s]);
#pragma unroll
            for (int s = 0; s < 16; ++s) {
                ar[s * 16 * RS] = x[x2_slot(s)].re;
                ai[s * 16 * RS] = x[x2_slot(s)].im;
            }
        }
        __syncthreads();

        // ---- phase 2b: DIF pass 2 on chunk g; output s is X_m, m = g + 16 s -> log2-increment over dW_m ---------
        {
            C2 x[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) x[q] = C2{Rc[q * RS], Ic[q * RS]};
            x2_dft16_transposed(x);
            float2* wm = reinterpret_cast<float2*>(W + g * TP + col);
#pragma unroll
            for (int s = 0; s < 16; ++s) {
                if (s < 15 || g + 240 < n) {
                    const float2 e = f2add(x[x2_slot(s)].re, comp[g + 16 * s]);
                    const float2 a = f2fma(e, half2v, lsq2);
                    const float2 u = make_float2(fast_ex2(a.x), fast_ex2(a.y));  // sqrt(v) sqrt(dt) log2e, see log2_increment()
                    wm[s * 16 * RS] = f2fma(u, f2fma(u, nkq2, wm[s * 16 * RS]), rd22);
                }
            }
        }
        __syncthreads();

        // ---- phase 3: log2-space prefix sum over time, S = S0 2^(.) -------------------------------------------
        {
            float2 c[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) c[t] = Wc[t * RS];  // rows >= n hold 0
#pragma unroll
            for (int t = 1; t < 16; ++t) c[t] = f2add(c[t], c[t - 1]);
            reinterpret_cast<float2*>(tot + g * TP + col)[0] = c[15];
            __syncthreads();
            float2 off = f2splat(0.f);
            for (int gg = 0; gg < g; ++gg) off = f2add(off, reinterpret_cast<const float2*>(tot + gg * TP + col)[0]);
            if (live0) {
                const bool both = live1 && ((P.ld & 1) == 0);  // 8-byte stores need an even row stride (ld is a multiple of 128)
                if (g == 0) {
                    out[path] = P.S0;
                    if (live1) out[path + 1] = P.S0;
                }
                float* o = out + (int64_t)(k0 + 1) * P.ld + path;
#pragma unroll
                for (int t = 0; t < 16; ++t, o += P.ld) {
                    if (full || k0 + t < n) {
                        const float2 a = f2add(off, c[t]);
                        const float2 sv = f2mul(S02, make_float2(fast_ex2(a.x), fast_ex2(a.y)));
                        if (both) *reinterpret_cast<float2*>(o) = sv;
                        else { o[0] = sv.x; if (live1) o[1] = sv.y; }
                    }
                }
            }
        }
        // no barrier needed here: the next tile's phase 1 writes only chunk g of re / im / W (read by this thread alone in
        // phase 3), and `tot` is rewritten only after the next three barriers
    }
}
