// gen_rbergomi_pair.cuh -- native-stream 256-point generator: ONE complex DFT drives TWO paths (included by
// gen_rbergomi.cu after gen_rbergomi_x2.cuh, inside its anonymous namespace; reuses its packed-pair helpers).
//
// The reference builds the fractional noise of one path as X_k = Re sum_{m<n} phi_m Z_m e^{-2 pi i k m / M'} with 2n
// real normals (RoughVolatility.cpp:264-292) and throws the imaginary half of the transform away.  X is a stationary
// Gaussian vector with Cov(X_k, X_l) = sum_m |phi_m|^2 cos(2 pi (k-l) m / M'): only |phi_m|^2 matters (Z_m is
// isotropic), and because cos is even in m the spectrum may be symmetrised,
//     w_m = ( |phi_m|^2 [m < n] + |phi_{M'-m}|^2 [M'-m < n] ) / 2            (indices mod M'),
// without changing the covariance.  For a SYMMETRIC spectrum the real and the imaginary part of
//     Y_k = sum_{m < M'} sqrt(w_m) G_m e^{-2 pi i k m / M'},     G_m iid standard complex normals,
// are uncorrelated (the cross-covariance is sum_m w_m sin(2 pi (k-l) m / M') = 0), hence independent, and each has
// exactly the reference's covariance.  So Re Y and Im Y are the X of two independent paths: M' complex normals per
// PAIR of paths instead of 2 n per path, half a transform per path, a real instead of a complex spectral multiply.
// With the single W normal per step that is ~2.03 normals per path-step instead of 3 (the reference draws 4) --
// Philox, the generator's floor, shrinks by a third.  Same law as the reference, different stream:
// this is the native (Philox) mode only; injected draws keep the per-path transform of gen_rbergomi_x2.cuh.
// tests/test_stream_law.py checks the covariance identity numerically; the DUMP mode writes, for every path, draws in
// the reference's order that reproduce this path through the reference's own formula (Z_m = u_m / phi_m with the
// m >= n frequencies folded onto their mirrors), so the native paths replay through the CPU oracle.
//
// Stream ("pair" stream; g = GLOBAL path id, T = g >> 6 the global 64-path tile, transform f = 32 T + (g & 31),
// Re -> path 64 T + (g & 31), Im -> path 64 T + 32 + (g & 31)):
//   G_m  : Philox ctr = (f_lo, f_hi, m >> 1, 4): (x0, x1) -> G_m for even m, (x2, x3) for odd m (Box-Muller pair = re, im)
//   W_k  : Philox ctr = (g_lo, g_hi, 4 (k & 15) + (k >> 6), 6): the four normals are steps k with (k >> 4) & 3 = 0..3
// Paths depend on (seed, g) only: any shard [path_offset, path_offset + n_paths) is a slice of the whole.
//
// CTA = 256 threads, tile = 64 paths = 32 transforms = 16 packed pairs of transforms; lane l: pair l & 15, chunk
// g = tid >> 4 (0..15).  Shared memory 78 KB: float re[256][32] | im[256][32] | tot[16][64] | float2 sw[256] | Tw tw2[256] |
// float2 comp[256].  The increments overwrite re / im in place (re row k = paths 0..31 of the tile, im row k = paths
// 32..63), so the W normals never touch shared memory: they are drawn in the second DFT pass, next to the X they meet.
#pragma once

constexpr int PAIR_SMEM = 2 * 256 * 32 * 4 + 16 * 64 * 4 + 256 * 8 + 256 * (int)sizeof(Tw) + 256 * 8;

// predicated 8-byte store: the guard becomes a predicate on the STG, never a branch (a branch per time index splits phase 3
// into sixteen basic blocks and keeps the scheduler from overlapping the SFU work of one row with the stores of the last)
__device__ __forceinline__ void st2_if(float* p, float2 v, bool ok) {
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %3, 0; @q st.global.v2.f32 [%0], {%1, %2}; }" ::"l"(p), "f"(v.x), "f"(v.y), "r"((int)ok) : "memory");
}

// Scheduling (r02; same stream and bits as the first version, 49.8 -> 45.5 ms at 2^26 x 252): the kernel alternates between
// integer-multiply-bound stretches (Philox: IMAD.WIDE on the FMA-heavy pipe) and SFU-bound stretches (Box-Muller, ex2),
// and the two pipes run side by side when fed (tools/microbench/pipes.cu: MUFU + IMAD.WIDE cost max, not sum).  So every
// phase is ONE straight-line block in which a Philox batch is drawn one step ahead of the Box-Muller work that consumes
// the previous one; the first spectral batch of the NEXT tile is drawn under the ex2 / store work of phase 3; the
// `m < n` guards of the second DFT pass are gone (rows >= n are never stored and only reach chunk totals that no stored
// row uses) and full tiles store through predicated STGs instead of a branch per time index.
template <bool DUMP>
__global__ void __launch_bounds__(NT2, 2) rbergomi_paths_n256pair_kernel(RbParams P, PhiloxKeys K, const float2* __restrict__ g_phis,
                                                                        const float* __restrict__ g_sw, const float2* __restrict__ g_tw,
                                                                        const float* __restrict__ g_comp2, float* __restrict__ draws_out,
                                                                        float* __restrict__ out) {
    constexpr int TC = 32, MP = 256;  // transforms (columns) per tile
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Are = reinterpret_cast<float*>(smem_raw);
    float* Aim = Are + MP * TC;
    float* tot = Aim + MP * TC;                               // [16][64]
    float2* sw = reinterpret_cast<float2*>(tot + 16 * 64);    // sqrt(w_m), splatted
    Tw* tw2 = reinterpret_cast<Tw*>(sw + MP);
    float2* comp = reinterpret_cast<float2*>(tw2 + MP);
    const int n = P.n;
    const int tid = threadIdx.x, pl = tid & 15, g = tid >> 4;
    {
        sw[tid] = f2splat(g_sw[tid]);
        const float2 w = g_tw[((tid >> 4) * (tid & 15)) & (MP - 1)];  // w256^{j s}, j = tid / 16, s = tid % 16
        tw2[tid] = Tw{f2splat(w.x), f2splat(w.y), f2splat(-w.y)};
        comp[tid] = f2splat(tid < n ? g_comp2[tid] : 0.f);
    }
    const int k0 = g * 16;
    const int col = 2 * pl;
    constexpr int RS = TC / 2;  // row stride in float2
    float2* const Rc = reinterpret_cast<float2*>(Are + k0 * TC + col);  // chunk g, this thread's pair of columns
    float2* const Ic = reinterpret_cast<float2*>(Aim + k0 * TC + col);
    const float2 lsq2 = f2splat(P.lsq), nkq2 = f2splat(P.nkq), rd22 = f2splat(P.rd2), S02 = f2splat(P.S0), half2v = f2splat(0.5f);
    const uint64_t T0 = P.path_offset >> 6;
    const int64_t n_tiles = (int64_t)(((P.path_offset + (uint64_t)P.n_paths + 63) >> 6) - T0);
    const bool even_rows = ((P.ld & 1) == 0) && ((P.path_offset & 1) == 0);  // 8-byte stores need even local ids and row stride
    __syncthreads();

    // the first spectral Philox batch of a tile is drawn one tile ahead, under the SFU / store work of phase 3
    uint4 cur[4];
    auto draw_first = [&](uint64_t Tn) {
        const uint64_t fn = (Tn << 5) + (uint64_t)col;
        const uint32_t n0 = (uint32_t)fn, n1 = (uint32_t)(fn >> 32), c2 = (uint32_t)(k0 >> 1);
        cur[0] = philox4x32_10(n0, n1, c2, 4u, K); cur[1] = philox4x32_10(n0 | 1u, n1, c2, 4u, K);
        cur[2] = philox4x32_10(n0, n1, c2 + 1u, 4u, K); cur[3] = philox4x32_10(n0 | 1u, n1, c2 + 1u, 4u, K);
    };
    draw_first(T0 + (uint64_t)blockIdx.x);

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t T = T0 + (uint64_t)tile;
        const uint64_t gidA = (T << 6) + (uint64_t)col, gidB = gidA + 32;   // first path of the Re pair / of the Im pair
        const int64_t locA = (int64_t)(gidA - P.path_offset), locB = locA + 32;  // may be negative in the first tile of a shard
        const bool liveA0 = locA >= 0 && locA < P.n_paths, liveA1 = locA + 1 >= 0 && locA + 1 < P.n_paths;
        const bool liveB0 = locB >= 0 && locB < P.n_paths, liveB1 = locB + 1 >= 0 && locB + 1 < P.n_paths;
        const uint64_t fid = (T << 5) + (uint64_t)col;  // transforms fid, fid + 1 (fid is even)
        const uint32_t f0 = (uint32_t)fid, f1 = (uint32_t)(fid >> 32);

        // ---- phase 1: spectral normals, in_m = sqrt(w_m) G_m ------------------------------------------------
        {
            // software pipeline: the integer rounds of batch q + 1 have no dependence on the SFU chain of batch q, so inside
            // one straight-line block the scheduler runs IMAD.WIDE / LOP3 under MUFU latency instead of after it
#pragma unroll
            for (int kq = 0; kq < 16; kq += 4) {
                const int m0 = k0 + kq;
                uint4 nxt[4];
                if (kq + 4 < 16) {
                    const uint32_t c2 = (uint32_t)((m0 + 4) >> 1);
                    nxt[0] = philox4x32_10(f0, f1, c2, 4u, K); nxt[1] = philox4x32_10(f0 | 1u, f1, c2, 4u, K);
                    nxt[2] = philox4x32_10(f0, f1, c2 + 1u, 4u, K); nxt[3] = philox4x32_10(f0 | 1u, f1, c2 + 1u, 4u, K);
                }
                float2 zr[4], zi[4];
                box_muller_x2(cur[0].x, cur[0].y, cur[1].x, cur[1].y, zr[0], zi[0]);
                box_muller_x2(cur[0].z, cur[0].w, cur[1].z, cur[1].w, zr[1], zi[1]);
                box_muller_x2(cur[2].x, cur[2].y, cur[3].x, cur[3].y, zr[2], zi[2]);
                box_muller_x2(cur[2].z, cur[2].w, cur[3].z, cur[3].w, zr[3], zi[3]);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float2 s = sw[m0 + t];
                    Rc[(kq + t) * RS] = f2mul(zr[t], s);
                    Ic[(kq + t) * RS] = f2mul(zi[t], s);
                }
                if (kq + 4 < 16) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) cur[t] = nxt[t];
                }
            }
        }
        __syncthreads();

        if (DUMP) {  // reference-order draws that reproduce these paths through X = Re DFT(phi (.) Z): Z_m = u_m / phi_m
            for (int t = 0; t < 16; ++t) {
                const int m = k0 + t;
                if (m < n) {
                    const float2 ar = Rc[t * RS], ai = Ic[t * RS];
                    float2 ur = ar, ui = ai;                                   // Re half: u = in_m
                    float2 vr = ai, vi = make_float2(-ar.x, -ar.y);            // Im half: u = -i in_m
                    if (m >= 1 && MP - m >= n) {                               // frequency M' - m has no slot in the reference: fold it
                        const float2 mr = *reinterpret_cast<const float2*>(Are + (MP - m) * TC + col);
                        const float2 mi = *reinterpret_cast<const float2*>(Aim + (MP - m) * TC + col);
                        ur = f2add(ur, mr);  ui = f2sub(ui, mi);               // + conj(in_{M'-m})
                        vr = f2add(vr, mi);  vi = f2add(vi, mr);               // + i conj(in_{M'-m})
                    }
                    const float2 ph = g_phis[m];
                    const float inv = 1.f / (ph.x * ph.x + ph.y * ph.y);
                    float* dre = draws_out + (int64_t)(2 * m) * P.ld_draws;
                    float* dim = draws_out + (int64_t)(2 * m + 1) * P.ld_draws;
                    if (liveA0) { dre[locA] = (ur.x * ph.x + ui.x * ph.y) * inv; dim[locA] = (ui.x * ph.x - ur.x * ph.y) * inv; }
                    if (liveA1) { dre[locA + 1] = (ur.y * ph.x + ui.y * ph.y) * inv; dim[locA + 1] = (ui.y * ph.x - ur.y * ph.y) * inv; }
                    if (liveB0) { dre[locB] = (vr.x * ph.x + vi.x * ph.y) * inv; dim[locB] = (vi.x * ph.x - vr.x * ph.y) * inv; }
                    if (liveB1) { dre[locB + 1] = (vr.y * ph.x + vi.y * ph.y) * inv; dim[locB + 1] = (vi.y * ph.x - vr.y * ph.y) * inv; }
                }
            }
            __syncthreads();
        }

        // ---- phase 2a: DIF pass 1 on column g: elements g + 16 q, output s scaled by w256^{g s} -------------------
        const uint32_t a0 = (uint32_t)gidA, a1 = (uint32_t)(gidA >> 32), b0 = (uint32_t)gidB, b1 = (uint32_t)(gidB >> 32);
        uint4 wq[4];  // Brownian Philox batch q of phase 2b
        auto draw_w = [&](int q) {
            const uint32_t ctr = (uint32_t)(4 * g + q);
            wq[0] = philox4x32_10(a0, a1, ctr, 6u, K); wq[1] = philox4x32_10(a0 | 1u, a1, ctr, 6u, K);
            wq[2] = philox4x32_10(b0, b1, ctr, 6u, K); wq[3] = philox4x32_10(b0 | 1u, b1, ctr, 6u, K);
        };
        {
            float2* ar = reinterpret_cast<float2*>(Are + g * TC + col);
            float2* ai = reinterpret_cast<float2*>(Aim + g * TC + col);
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

        // ---- phase 2b: DIF pass 2 on chunk g; output s is Y_m, m = g + 16 s; Re -> X of the first 32 paths of the
        //      tile, Im -> X of the other 32; the log2-increment replaces Y_m in place (after everyone has read) -----
        {
            C2 x[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) x[q] = C2{Rc[q * RS], Ic[q * RS]};
            __syncthreads();
            x2_dft16_transposed(x);
            float2* oa = reinterpret_cast<float2*>(Are + g * TC + col);
            float2* ob = reinterpret_cast<float2*>(Aim + g * TC + col);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                draw_w(q);
                const uint4 xA0 = wq[0], xA1 = wq[1], xB0 = wq[2], xB1 = wq[3];
                float2 wA[4], wB[4];
                box_muller_x2(xA0.x, xA0.y, xA1.x, xA1.y, wA[0], wA[1]);
                box_muller_x2(xA0.z, xA0.w, xA1.z, xA1.w, wA[2], wA[3]);
                box_muller_x2(xB0.x, xB0.y, xB1.x, xB1.y, wB[0], wB[1]);
                box_muller_x2(xB0.z, xB0.w, xB1.z, xB1.w, wB[2], wB[3]);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int s = 4 * q + t, m = g + 16 * s;
                    const bool in = m < n;
                    const float2 cm = comp[m];
                    const float2 uA0 = f2fma(f2add(x[x2_slot(s)].re, cm), half2v, lsq2), uB0 = f2fma(f2add(x[x2_slot(s)].im, cm), half2v, lsq2);
                    const float2 uA = make_float2(fast_ex2(uA0.x), fast_ex2(uA0.y));  // sqrt(v) sqrt(dt) log2e, see log2_increment()
                    const float2 uB = make_float2(fast_ex2(uB0.x), fast_ex2(uB0.y));
                    // rows >= n: finite, never stored, and they only reach chunk totals that no stored row uses
                    oa[s * 16 * RS] = f2fma(uA, f2fma(uA, nkq2, wA[t]), rd22);
                    ob[s * 16 * RS] = f2fma(uB, f2fma(uB, nkq2, wB[t]), rd22);
                    if (DUMP && in) {  // the reference's W1 / W2 slots: rho W1 + sqrt(1 - rho^2) W2 = w
                        float* d1 = draws_out + (int64_t)(2 * n + m) * P.ld_draws;
                        float* d2 = draws_out + (int64_t)(3 * n + m) * P.ld_draws;
                        if (liveA0) { d1[locA] = P.rho * wA[t].x; d2[locA] = P.rho_c * wA[t].x; }
                        if (liveA1) { d1[locA + 1] = P.rho * wA[t].y; d2[locA + 1] = P.rho_c * wA[t].y; }
                        if (liveB0) { d1[locB] = P.rho * wB[t].x; d2[locB] = P.rho_c * wB[t].x; }
                        if (liveB1) { d1[locB + 1] = P.rho * wB[t].y; d2[locB + 1] = P.rho_c * wB[t].y; }
                    }
                }
            }
        }
        __syncthreads();

        // ---- phase 3: log2-space prefix sum over time, S = S0 2^(.) -------------------------------------------
        {
            float2 ca[16], cb[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) { ca[t] = Rc[t * RS]; cb[t] = Ic[t * RS]; }  // rows >= n hold 0
#pragma unroll
            for (int t = 1; t < 16; ++t) { ca[t] = f2add(ca[t], ca[t - 1]); cb[t] = f2add(cb[t], cb[t - 1]); }
            reinterpret_cast<float2*>(tot + g * 64 + col)[0] = ca[15];
            reinterpret_cast<float2*>(tot + g * 64 + 32 + col)[0] = cb[15];
            __syncthreads();
            float2 offa = f2splat(0.f), offb = f2splat(0.f);
            for (int gg = 0; gg < g; ++gg) {
                offa = f2add(offa, reinterpret_cast<const float2*>(tot + gg * 64 + col)[0]);
                offb = f2add(offb, reinterpret_cast<const float2*>(tot + gg * 64 + 32 + col)[0]);
            }
            if (g == 0) {
                if (liveA0) out[locA] = P.S0;
                if (liveA1) out[locA + 1] = P.S0;
                if (liveB0) out[locB] = P.S0;
                if (liveB1) out[locB + 1] = P.S0;
            }
            const bool bothA = even_rows && liveA0 && liveA1, bothB = even_rows && liveB0 && liveB1;
            float* o = out + (int64_t)(k0 + 1) * P.ld + locA;
            if (bothA && bothB) {  // the whole tile is inside the shard (all but the edge tiles): no branch per row
                const int rows = n - k0;       // rows of this chunk that exist (>= 16 for all chunks but the last)
                draw_first(T + (uint64_t)gridDim.x);
#pragma unroll
                for (int t = 0; t < 16; ++t, o += P.ld) {
                    const float2 a = f2add(offa, ca[t]), b = f2add(offb, cb[t]);
                    const float2 sa = f2mul(S02, make_float2(fast_ex2(a.x), fast_ex2(a.y)));
                    const float2 sb = f2mul(S02, make_float2(fast_ex2(b.x), fast_ex2(b.y)));
                    st2_if(o, sa, t < rows);
                    st2_if(o + 32, sb, t < rows);
                }
            } else {
            draw_first(T + (uint64_t)gridDim.x);
#pragma unroll
            for (int t = 0; t < 16; ++t, o += P.ld) {
                if (k0 + t < n) {
                    const float2 a = f2add(offa, ca[t]), b = f2add(offb, cb[t]);
                    const float2 sa = f2mul(S02, make_float2(fast_ex2(a.x), fast_ex2(a.y)));
                    const float2 sb = f2mul(S02, make_float2(fast_ex2(b.x), fast_ex2(b.y)));
                    if (bothA) *reinterpret_cast<float2*>(o) = sa;
                    else { if (liveA0) o[0] = sa.x; if (liveA1) o[1] = sa.y; }
                    if (bothB) *reinterpret_cast<float2*>(o + 32) = sb;
                    else { if (liveB0) o[32] = sb.x; if (liveB1) o[33] = sb.y; }
                }
            }
            }
        }
        // no barrier needed here: the next tile's phase 1 writes only chunk g of re / im (read by this thread alone in
        // phase 3), and `tot` is rewritten only after the next four barriers
    }
}
