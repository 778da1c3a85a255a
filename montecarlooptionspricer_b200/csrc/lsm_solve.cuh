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

// The reference solves  [1, S, .., S^p] c = y  with Eigen's bdcSvd().solve (LSMPricer.cpp:61-76): a min-norm solve that DROPS
// the singular directions of the RAW monomial design with sigma_i < sigma_max * min(rows, cols) * eps (SVDBase::rank()).
// For p >= 5 and S ~ 100 that cut really bites (sigma_min / sigma_max ~ 4e-15 at p = 5, 6e-18 at p = 6), so a full-rank
// fit in the standardised basis is NOT what the reference returns.  The cut can be reproduced without ever forming the
// raw design:  with G = L L^T the Cholesky factor of the standardised Gram matrix and  r(S) = M u(x)  the (exact,
// lower-triangular) change of basis from the standardised monomials u to the raw ones,
//     A_raw = Q B,   Q = U L^{-T} (orthonormal columns),   B = L^T M^T  (upper triangular, column-graded like Eigen's R),
// so A_raw and the small matrix B share their singular values and  c = L^{-T} P z,  z = L^{-1} X^T y,  P = projector onto the
// left singular vectors of B that survive the reference's cut.  One-sided (Hestenes) Jacobi on the columns of B is
// accurate for column-graded matrices -- it is what the CPU oracle (and Eigen, below 16 columns) runs on R.  Full rank
// (every realistic case for p <= 4) leaves z untouched, bit for bit.
struct RefRank {
    double mu, inv_s;  // standardisation of the step being regressed: x = (S - mu) * inv_s
};

template <int P>
static __device__ __noinline__ void project_onto_reference_rank(const double (&L)[P + 1][P + 1], const double (&d)[P + 1], const RefRank& rr, double count,
                                                                double (&z)[P + 1]) {
    constexpr int n = P + 1;
    const double s = 1.0 / rr.inv_s;
    // M[k][j] = C(k, j) mu^(k-j) s^j  (S^k = (mu + s x)^k), lower triangular
    double M[n][n];
    for (int k = 0; k < n; ++k) {
        double binom = 1.0;
        for (int j = 0; j <= k; ++j) {
            double v = binom;
            for (int q = 0; q < k - j; ++q) v *= rr.mu;
            for (int q = 0; q < j; ++q) v *= s;
            M[k][j] = v;
            binom = binom * (double)(k - j) / (double)(j + 1);
        }
        for (int j = k + 1; j < n; ++j) M[k][j] = 0.0;
    }
    // B = L^T D^{-1} M^T with the stored factor of the EQUILIBRATED Gram matrix: L_full[k][a] = (k == a ? 1 / L[k][k] : L[k][a]), D = diag(d)
    double W[n][n];
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) {
            double acc = 0.0;
            for (int k = a; k <= b; ++k) acc += (k == a ? 1.0 / L[k][k] : L[k][a]) / d[k] * M[b][k];
            W[a][b] = acc;
        }
    // one-sided Jacobi: rotate column pairs until mutually orthogonal; the columns end up as sigma_i w_i
    bool rotated = true;
    for (int sweep = 0; sweep < 60 && rotated; ++sweep) {
        rotated = false;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                double a = 0.0, bb = 0.0, c = 0.0;
                for (int i = 0; i < n; ++i) { a += W[i][p] * W[i][p]; bb += W[i][q] * W[i][q]; c += W[i][p] * W[i][q]; }
                if (c == 0.0 || fabs(c) <= 2.220446049250313e-16 * sqrt(a * bb)) continue;
                rotated = true;
                const double zeta = (bb - a) / (2.0 * c);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                for (int i = 0; i < n; ++i) {
                    const double wp = W[i][p], wq = W[i][q];
                    W[i][p] = cs * wp - sn * wq;
                    W[i][q] = sn * wp + cs * wq;
                }
            }
    }
    double sig[n], smax = 0.0;
    for (int j = 0; j < n; ++j) {
        double t = 0.0;
        for (int i = 0; i < n; ++i) t += W[i][j] * W[i][j];
        sig[j] = sqrt(t);
        smax = fmax(smax, sig[j]);
    }
    const double diag = count < (double)n ? count : (double)n;                // Eigen: min(rows, cols)
    double thr = smax * diag * 2.220446049250313e-16;                         // SVDBase::rank(): keep sigma_i >= this
    if (thr < 2.2250738585072014e-308) thr = 2.2250738585072014e-308;
    bool all = true;
    for (int j = 0; j < n; ++j) all = all && (sig[j] >= thr && sig[j] > 0.0);
    if (all) return;                                                           // full rank: z stays as it is, bit for bit
    double zp[n];
    for (int i = 0; i < n; ++i) zp[i] = 0.0;
    for (int j = 0; j < n; ++j) {
        if (!(sig[j] >= thr && sig[j] > 0.0)) continue;
        double proj = 0.0;
        for (int i = 0; i < n; ++i) proj += W[i][j] * z[i];
        proj /= sig[j] * sig[j];
        for (int i = 0; i < n; ++i) zp[i] += W[i][j] * proj;
    }
    for (int i = 0; i < n; ++i) z[i] = zp[i];
}

// Solve the normal equations of one step from the (globally reduced) moments -> coef row (one thread).
//   G[a][b] = s[a+b], rhs[a] = t[a].  Diagonal equilibration; Cholesky when safely positive definite (unrolled,
//   in registers), else the Jacobi fallback above.  rr != nullptr: apply the reference's rank cut (see above).
template <int P>
__device__ __forceinline__ void solve_normal_equations(const double* mom, double* __restrict__ coef_row, const RefRank* rr = nullptr) {
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
        if (rr != nullptr && P >= 1) project_onto_reference_rank<P>(L, d, *rr, mom[0], z);
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

static __device__ __noinline__ void solve_dispatch(const double* mom, int p, double* coef_row, const RefRank* rr = nullptr) {
    switch (p) {
        case 0: solve_normal_equations<0>(mom, coef_row, rr); break;
        case 1: solve_normal_equations<1>(mom, coef_row, rr); break;
        case 2: solve_normal_equations<2>(mom, coef_row, rr); break;
        case 3: solve_normal_equations<3>(mom, coef_row, rr); break;
        case 4: solve_normal_equations<4>(mom, coef_row, rr); break;
        case 5: solve_normal_equations<5>(mom, coef_row, rr); break;
        default: solve_normal_equations<6>(mom, coef_row, rr); break;
    }
}
