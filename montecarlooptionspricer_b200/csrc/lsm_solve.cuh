// lsm_solve.cuh -- pieces shared by every pricer kernel: payoff, exact float->double widening, warp sums and the
// tiny normal-equation solver (equilibrated Cholesky with a rank-revealing Jacobi fallback).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int MAXP = 6;
constexpr int COEF_LD = 8;  // coefficient row stride (p+1 <= 7)

__device__ __forceinline__ double payoff_fn(int is_call, double S, double K) {  // include/core/common.h:8-14
    const double x = is_call ? S - K : K - S;
    return x > 0.0 ? x : 0.0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Exact float -> double widening on the integer pipe (the F2F conversion unit is quarter rate and was the top
// stall of the first sweep kernel: 44% of samples).  Sub-normals flush to zero (prices and option values never
// are); zero, inf and nan keep their meaning.
__device__ __forceinline__ double f2d(float f) {
    const uint32_t b = __float_as_uint(f), e = b & 0x7f800000u;
    uint32_t hi = (b & 0x80000000u) | (((b & 0x7fffffffu) >> 3) + 0x38000000u);
    uint32_t lo = b << 29;
    if (e == 0u) { hi = b & 0x80000000u; lo = 0u; }
    if (e == 0x7f800000u) hi |= 0x7ff00000u;
    return __hiloint2double((int)hi, (int)lo);
}

// Rank-revealing fallback: cyclic Jacobi on the equilibrated Gram matrix, pseudo-inverse with a relative cut
// (projection of y onto the realised column space == what the reference's min-norm SVD solve evaluates to at the
// regression points, LSMPricer.cpp:76-85).  Rare (j = 0 in the money, fewer ITM paths than basis functions).
static __device__ __noinline__ void solve_fallback_jacobi(double* G /*[n][MAXP+1], destroyed*/, const double* rhs, int n, double* z) {
    double Q[MAXP + 1][MAXP + 1];
    auto g = [&](int a, int b) -> double& { return G[a * (MAXP + 1) + b]; };
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) Q[a][b] = a == b ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0.0;
        for (int a = 0; a < n; ++a)
            for (int b = a + 1; b < n; ++b) off += g(a, b) * g(a, b);
        if (off < 1e-60) break;
        for (int pp = 0; pp < n - 1; ++pp)
            for (int q = pp + 1; q < n; ++q) {
                const double apq = g(pp, q);
                if (apq == 0.0) continue;
                const double theta = (g(q, q) - g(pp, pp)) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < n; ++k) {
                    const double gkp = g(k, pp), gkq = g(k, q);
                    g(k, pp) = cs * gkp - sn * gkq;
                    g(k, q) = sn * gkp + cs * gkq;
                }
                for (int k = 0; k < n; ++k) {
                    const double gpk = g(pp, k), gqk = g(q, k);
                    g(pp, k) = cs * gpk - sn * gqk;
                    g(q, k) = sn * gpk + cs * gqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double qkp = Q[k][pp], qkq = Q[k][q];
                    Q[k][pp] = cs * qkp - sn * qkq;
                    Q[k][q] = sn * qkp + cs * qkq;
                }
            }
    }
    double lmax = 0.0;
    for (int a = 0; a < n; ++a) lmax = fmax(lmax, g(a, a));
    const double thr = lmax * (double)n * 64.0 * 2.220446049250313e-16;
    for (int a = 0; a < n; ++a) z[a] = 0.0;
    for (int e = 0; e < n; ++e) {
        if (g(e, e) > thr) {
            double proj = 0.0;
            for (int a = 0; a < n; ++a) proj += Q[a][e] * rhs[a];
            proj /= g(e, e);
            for (int a = 0; a < n; ++a) z[a] += Q[a][e] * proj;
        }
    }
}

// Solve the normal equations of one step from the (globally reduced) moments -> coef row (one thread).
//   G[a][b] = s[a+b], rhs[a] = t[a].  Diagonal equilibration; Cholesky when safely positive definite (unrolled,
//   in registers), else the Jacobi fallback above.
template <int P>
__device__ __forceinline__ void solve_normal_equations(const double* mom, double* __restrict__ coef_row) {
    constexpr int n = P + 1;
    double G[n][n], rhs[n], d[n], z[n], L[n][n];
#pragma unroll
    for (int k = 0; k < COEF_LD; ++k) coef_row[k] = 0.0;
    if (!(mom[0] > 0.0)) return;  // no in-the-money path at this step (LSMPricer.cpp:60)
#pragma unroll
    for (int a = 0; a < n; ++a) d[a] = mom[2 * a] > 0.0 ? rsqrt(mom[2 * a]) : 0.0;
#pragma unroll
    for (int a = 0; a < n; ++a) {
#pragma unroll
        for (int b = 0; b < n; ++b) G[a][b] = mom[a + b] * d[a] * d[b];
        rhs[a] = mom[2 * P + 1 + a] * d[a];
    }
    bool ok = true;
#pragma unroll
    for (int k = 0; k < n; ++k) {
        double piv = G[k][k];
#pragma unroll
        for (int m = 0; m < k; ++m) piv -= L[k][m] * L[k][m];
        ok = ok && (piv > 1e-10);
        const double inv = rsqrt(ok ? piv : 1.0);
        L[k][k] = inv;  // store 1/l_kk
#pragma unroll
        for (int i = k + 1; i < n; ++i) {
            double sacc = G[i][k];
#pragma unroll
            for (int m = 0; m < k; ++m) sacc -= L[i][m] * L[k][m];
            L[i][k] = sacc * inv;
        }
    }
    if (ok) {
#pragma unroll
        for (int i = 0; i < n; ++i) {
            double sacc = rhs[i];
#pragma unroll
            for (int m = 0; m < i; ++m) sacc -= L[i][m] * z[m];
            z[i] = sacc * L[i][i];
        }
#pragma unroll
        for (int i = n - 1; i >= 0; --i) {
            double sacc = z[i];
#pragma unroll
            for (int m = i + 1; m < n; ++m) sacc -= L[m][i] * z[m];
            z[i] = sacc * L[i][i];
        }
    } else {
        double Gf[(MAXP + 1) * (MAXP + 1)], rf[MAXP + 1], zf[MAXP + 1];
        for (int a = 0; a < n; ++a) {
            for (int b = 0; b < n; ++b) Gf[a * (MAXP + 1) + b] = G[a][b];
            rf[a] = rhs[a];
        }
        solve_fallback_jacobi(Gf, rf, n, zf);
        for (int a = 0; a < n; ++a) z[a] = zf[a];
    }
#pragma unroll
    for (int a = 0; a < n; ++a) coef_row[a] = z[a] * d[a];
}

static __device__ __noinline__ void solve_dispatch(const double* mom, int p, double* coef_row) {
    switch (p) {
        case 0: solve_normal_equations<0>(mom, coef_row); break;
        case 1: solve_normal_equations<1>(mom, coef_row); break;
        case 2: solve_normal_equations<2>(mom, coef_row); break;
        case 3: solve_normal_equations<3>(mom, coef_row); break;
        case 4: solve_normal_equations<4>(mom, coef_row); break;
        case 5: solve_normal_equations<5>(mom, coef_row); break;
        default: solve_normal_equations<6>(mom, coef_row); break;
    }
}

