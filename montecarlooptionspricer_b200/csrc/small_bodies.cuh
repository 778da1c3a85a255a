// small_bodies.cuh -- the four pricers as SINGLE-CTA device routines for small path sets (<= 4096 paths; the
// reference's production rows are 250 paths x floor(dte/365*252) steps, PredictionGen.cpp:718-719).  One CTA walks the
// whole time axis with block barriers, so a pricer is one launch -- and a batch of rows is one launch with one CTA per
// row (rows.cu).  All decisions in fp64 on the stored path values, predicates literal:
//   LSM          src/models/LSMPricer.cpp:19-102
//   Branching    src/models/BranchingProcessPricer.cpp:41-134
//   Asymptotic   src/models/AsymptoticAnalysisPricer.cpp:8-113
//   Martingale   src/models/MartingaleOptimizationPricer.cpp:21-188
// Every routine must be called by ALL threads of the CTA (any block size that is a multiple of 32, at most SB_NT: the
// loops stride by blockDim.x, so the batched row driver can run two 256-thread CTAs per SM where the per-row API uses 512).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lsm_solve.cuh"
#include "philox.cuh"

constexpr int SB_NT = 512;
constexpr int SB_MAX_PATHS = 4096;

template <typename ST>
__device__ __forceinline__ double sb_ld(const ST* p);
template <>
__device__ __forceinline__ double sb_ld<float>(const float* p) { return f2d(*p); }
template <>
__device__ __forceinline__ double sb_ld<double>(const double* p) { return *p; }

// deterministic CTA-wide sum of NV doubles per thread -> out[0..NV) (shared), visible to all threads on return
template <int NV>
__device__ __forceinline__ void sb_sum(double (&acc)[NV], double* __restrict__ out) {
    __shared__ double red[SB_NT / 32][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const double s = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

// mean / 1/std of the in-the-money prices (standardisation of the regression variable), fallbacks as on the big path
__device__ __forceinline__ void sb_scale(double cnt, double s1, double s2, double K, double* mu, double* inv_s) {
    double m = K, sd = fabs(K) > 0.0 ? fabs(K) : 1.0;
    if (cnt >= 2.0) {
        m = s1 / cnt;
        const double var = (s2 - cnt * m * m) / (cnt - 1.0);
        if (var > 1e-12 * m * m) sd = sqrt(var);
        else sd = fabs(m) > 0.0 ? fabs(m) : 1.0;
    }
    *mu = m;
    *inv_s = 1.0 / sd;
}

// ------------------------------------------------------------------------------------------------------- LSM
// sV: shared [n] (the carry).  Optional global outputs: tau[n], coef_out[M][COEF_LD] (standardised basis), mu_out[M],
// is_out[M], v_out[n].  fin[0..3) = {sum V0, sum (V0 - mean)^2, n} (written by thread 0; shared or global).
template <typename ST, int P>
__device__ void sb_lsm(const ST* __restrict__ S, int64_t ld, int n, int M, double K, int is_call, double disc, double dt, double maturity,
                       double* __restrict__ sV, int32_t* __restrict__ tau, double* __restrict__ coef_out, double* __restrict__ mu_out,
                       double* __restrict__ is_out, double* __restrict__ v_out, double* __restrict__ fin) {
    constexpr int NM = 3 * P + 2;
    constexpr int NV = NM > 3 ? NM : 3;
    __shared__ double mom[NV], cf[COEF_LD], sc[2];
    const int tid = threadIdx.x;
    // disc = exp(-r dt), evaluated by the caller on the host exactly like LSMPricer.cpp:46,69,92
    for (int i = tid; i < n; i += (int)blockDim.x) sV[i] = payoff_fn(is_call, sb_ld<ST>(S + (int64_t)(M - 1) * ld + i), K);  // :37-40
    __syncthreads();
    for (int j = M - 2; j >= 0; --j) {                                                        // :42
        const ST* Sj = S + (int64_t)j * ld;
        if ((double)j * dt > maturity) {                                                      // :43-49
            for (int i = tid; i < n; i += (int)blockDim.x) sV[i] *= disc;
            if (tid == 0 && coef_out)
                for (int k = 0; k < COEF_LD; ++k) coef_out[(int64_t)j * COEF_LD + k] = 0.0;
            __syncthreads();
            continue;
        }
        // standardisation of step j from its in-the-money prices
        double st[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) st[k] = 0.0;
        for (int i = tid; i < n; i += (int)blockDim.x) {
            const double s = sb_ld<ST>(Sj + i);
            if (payoff_fn(is_call, s, K) > 1e-14) { st[0] += 1.0; st[1] += s; st[2] = fma(s, s, st[2]); }
        }
        sb_sum<NV>(st, mom);
        if (tid == 0) {
            sb_scale(mom[0], mom[1], mom[2], K, &sc[0], &sc[1]);
            if (mu_out) mu_out[j] = sc[0];
            if (is_out) is_out[j] = sc[1];
        }
        __syncthreads();
        const double mu = sc[0], inv_s = sc[1];
        // normal-equation moments over the in-the-money paths                                  :51-74
        double acc[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] = 0.0;
        for (int i = tid; i < n; i += (int)blockDim.x) {
            const double s = sb_ld<ST>(Sj + i);
            if (payoff_fn(is_call, s, K) > 1e-14) {
                const double x = (s - mu) * inv_s, y = sV[i] * disc;                          // :69
                double xp = x;
                acc[0] += 1.0;
                acc[2 * P + 1] += y;
#pragma unroll
                for (int k = 1; k <= 2 * P; ++k) {
                    acc[k] += xp;
                    if (k <= P) acc[2 * P + 1 + k] = fma(xp, y, acc[2 * P + 1 + k]);
                    if (k < 2 * P) xp *= x;
                }
            }
        }
        sb_sum<NV>(acc, mom);
        if (tid == 0) {
            const RefRank rr{mu, inv_s};
            solve_normal_equations<P>(mom, cf, &rr);                                          // :76 (with Eigen's rank cut, lsm_solve.cuh)
            if (coef_out)
                for (int k = 0; k < COEF_LD; ++k) coef_out[(int64_t)j * COEF_LD + k] = cf[k];
        }
        __syncthreads();
        double c[P + 1];
#pragma unroll
        for (int k = 0; k <= P; ++k) c[k] = cf[k];
        for (int i = tid; i < n; i += (int)blockDim.x) {
            const double s = sb_ld<ST>(Sj + i), pay = payoff_fn(is_call, s, K);
            const double x = (s - mu) * inv_s;
            double cont = c[P];
#pragma unroll
            for (int k = P - 1; k >= 0; --k) cont = fma(cont, x, c[k]);
            const bool itm = pay > 1e-14, ex = !(pay < cont);                                 // :55, :85
            const double carried = pay < 1e-14 ? sV[i] * disc : 0.0;                          // :89-94; == 1e-14 keeps the initial 0 (:35)
            sV[i] = itm ? (ex ? pay : cont) : carried;
            if (tau && itm && ex) tau[i] = j;
        }
        __syncthreads();
    }
    // payoff averaging (:97-101) + two-pass standard error
    double t[3] = {0.0, 0.0, 0.0};
    for (int i = tid; i < n; i += (int)blockDim.x) t[0] += sV[i];
    sb_sum<3>(t, mom);
    const double total = mom[0], mean = total / (double)n;
    __syncthreads();
    double q[3] = {0.0, 0.0, 0.0};
    for (int i = tid; i < n; i += (int)blockDim.x) {
        const double dlt = sV[i] - mean;
        q[0] = fma(dlt, dlt, q[0]);
        if (v_out) v_out[i] = sV[i];
    }
    sb_sum<3>(q, mom);
    if (tid == 0) { fin[0] = total; fin[1] = mom[0]; fin[2] = (double)n; }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------- Branching
// F, best: shared [n] each.  ex[0..n_ex): visited exercise dates (strictly increasing); is_ex[j] = e + 1 for a visited
// date, else 0 (either a table, or nullptr meaning ex = 0 .. n_ex-1, the reference's own choice PredictionGen.cpp:780-783).
// out[0] = sum of lower bounds, out[1] = sum of upper bounds (thread 0 / 1 write).
template <typename ST>
__device__ void sb_branching(const ST* __restrict__ S, int64_t ld, int n, int j_hi, int j_lo, int kend, int ex_back, const int* __restrict__ is_ex,
                             const int* __restrict__ ex, int n_ex, double r, double dt, double K, int is_call, int n_br, const PhiloxKeys& keys,
                             uint64_t path_offset, const int32_t* __restrict__ inj, double* __restrict__ F, double* __restrict__ best,
                             double* __restrict__ out) {
    __shared__ double red2[2];
    const int tid = threadIdx.x;
    double lo[2] = {0.0, 0.0};
    for (int i = tid; i < n; i += (int)blockDim.x) {
        F[i] = 0.0;
        best[i] = 0.0;
        double b = 0.0;  // lower bound: first listed date with a positive discounted payoff (:55-70)
        for (int e = 0; e < n_ex; ++e) {
            const int j = ex ? ex[e] : e;
            const double d = exp(-r * ((double)j * dt)) * payoff_fn(is_call, sb_ld<ST>(S + (int64_t)j * ld + i), K);
            if (d > b) { b = d; break; }
        }
        lo[0] += b;
    }
    __syncthreads();
    for (int j = j_hi; j >= j_lo; --j) {
        const int e = is_ex ? is_ex[j] : (j < n_ex ? j + 1 : 0);
        const bool has_cont = j < ex_back, j_valid = j < kend;
        const double dj = exp(-r * ((double)j * dt));
        if (e) {
            for (int i = tid; i < n; i += (int)blockDim.x) {
                const double d = dj * payoff_fn(is_call, sb_ld<ST>(S + (int64_t)j * ld + i), K);
                double cont = 0.0;
                if (has_cont) {                                                               // :103
                    double sum = 0.0;
                    const uint64_t gid = path_offset + (uint64_t)i;
                    const int32_t* row = inj ? inj + ((int64_t)(e - 1) * n + i) * n_br : nullptr;
                    for (int b0 = 0; b0 < n_br; b0 += 4) {
                        uint4 u = make_uint4(0u, 0u, 0u, 0u);
                        if (!inj) u = philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)j, 0x10000u + (uint32_t)(b0 >> 2), keys);
                        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (b0 + q < n_br) sum += F[inj ? (int)row[b0 + q] : (int)(((uint64_t)uu[q] * (uint64_t)n) >> 32)];
                    }
                    cont = sum / (double)n_br;                                                // :123 (e^{-r t} is inside F)
                }
                const double better = d < cont ? cont : d;                                    // :126
                if (better > best[i]) best[i] = better;                                       // :127-129
            }
            __syncthreads();  // every gather of date j is done before F takes index j in
        }
        if (j_valid)
            for (int i = tid; i < n; i += (int)blockDim.x) {
                const double d = dj * payoff_fn(is_call, sb_ld<ST>(S + (int64_t)j * ld + i), K);
                if (d > F[i]) F[i] = d;
            }
        __syncthreads();
    }
    for (int i = tid; i < n; i += (int)blockDim.x) lo[1] += best[i];
    sb_sum<2>(lo, red2);
    if (tid < 2) out[tid] = red2[tid];
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------ Asymptotic
// tab: shared [2 * M] (boundary | discount).  out[0] = sum of per-path bests, out[1] = number of finite ones.
template <typename ST>
__device__ void sb_asymptotic(const ST* __restrict__ S, int64_t ld, int n, int M, double K, int is_call, double r, double dt, double maturity,
                              double sigma, double dividend, double* __restrict__ tab, double* __restrict__ out) {
    __shared__ double red2[2];
    const int tid = threadIdx.x;
    int jend = M;
    for (int j = 0; j < M; ++j)
        if ((double)j * dt > maturity) { jend = j; break; }                                   // :71
    for (int j = tid; j < M; j += (int)blockDim.x) {
        const double t = (double)j * dt, eps = maturity - t;
        double b = K;
        if (!(eps < 1e-10)) {                                                                 // :10-11, :25-26
            const double c0 = 0.5 * sigma * sqrt(eps * log(1.0 / eps));                       // NaN for eps > 1, as in the reference
            if (is_call) { b = K - c0; if (eps < 0.01) b += 0.5 * (dividend - r) * eps; }     // :28-34
            else         { b = K + c0; if (eps < 0.01) b -= 0.5 * (r - dividend) * eps; }     // :13-19
        }
        tab[j] = b;
        tab[M + j] = exp(-r * t);
    }
    __syncthreads();
    double acc[2] = {0.0, 0.0};
    for (int i = tid; i < n; i += (int)blockDim.x) {
        double best = 0.0;
        for (int j = 0; j < jend; ++j) {
            const double s = sb_ld<ST>(S + (int64_t)j * ld + i);
            if (isnan(s) || isinf(s)) continue;                                               // :74
            const double b = tab[j];
            const bool in = is_call ? (s > b) : (s < b);                                      // :80-85
            if (in) {
                const double pay = payoff_fn(is_call, s, K);
                if (isnan(pay) || isinf(pay)) continue;                                       // :89
                const double d = tab[M + j] * pay;                                            // :90
                if (d > best) best = d;
            }
        }
        if (!isnan(best) && !isinf(best)) { acc[0] += best; acc[1] += 1.0; }                  // :101-106
    }
    sb_sum<2>(acc, red2);
    if (tid < 2) out[tid] = red2[tid];
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------ Martingale
// smp: shared [4 * n] (S_stop | target | S_other | target2);  DF: shared [M].  out[0] = sum primal, out[1] = sum dual
// (== primal when max_iterations == 1).  The fitted martingale does not depend on the iteration count (pricers.cu).
template <typename ST, int P>
__device__ void sb_martingale(const ST* __restrict__ S, int64_t ld, int n, int M, double K, int is_call, double r, double dt, double maturity,
                              int max_iterations, double* __restrict__ smp, double* __restrict__ DF, double* __restrict__ out) {
    constexpr int NM = 3 * P + 2;
    constexpr int NV = NM > 3 ? NM : 3;
    __shared__ double mom[NV], cf[COEF_LD], sc[3];
    const int tid = threadIdx.x;
    int jend = M;
    for (int j = 0; j < M; ++j)
        if ((double)j * dt > maturity) { jend = j; break; }
    for (int j = tid; j < M; j += (int)blockDim.x) {                                                    // PathDiscountFactor (.h:44-49)
        double t = (double)j * dt;
        if (t > maturity) t = maturity;
        DF[j] = exp(-r * t);
    }
    __syncthreads();
    double pr[3] = {0.0, 0.0, 0.0};
    for (int i = tid; i < n; i += (int)blockDim.x) {                                                    // :72-94, :130-150
        double best = 0.0, s_stop = sb_ld<ST>(S + i);
        int idx = 0;
        for (int j = 0; j < jend; ++j) {
            const double s = sb_ld<ST>(S + (int64_t)j * ld + i);
            const double dp = payoff_fn(is_call, s, K) * DF[j];
            if (dp > best) { best = dp; idx = j; s_stop = s; }
        }
        pr[0] += best;
        const int jo = (idx + M / 2) % M;
        const double s_other = sb_ld<ST>(S + (int64_t)jo * ld + i);
        smp[i] = s_stop;
        smp[n + i] = 0.5 * (payoff_fn(is_call, s_stop, K) * DF[idx]);
        smp[2 * n + i] = s_other;
        smp[3 * n + i] = 0.2 * (payoff_fn(is_call, s_other, K) * DF[jo]);
        pr[1] += 2.0;
        pr[2] += s_stop + s_other;
    }
    sb_sum<3>(pr, sc);
    const double primal_sum = sc[0], cnt = sc[1], s1 = sc[2];
    __syncthreads();
    if (max_iterations < 2) {
        if (tid == 0) { out[0] = primal_sum; out[1] = primal_sum; }
        __syncthreads();
        return;
    }
    // standardisation over all 2n samples, then the normal-equation moments (:152-166)
    double sq[3] = {0.0, 0.0, 0.0};
    for (int i = tid; i < n; i += (int)blockDim.x) sq[0] += smp[i] * smp[i] + smp[2 * n + i] * smp[2 * n + i];
    sb_sum<3>(sq, sc);
    double mu, inv_s;
    sb_scale(cnt, s1, sc[0], K, &mu, &inv_s);
    __syncthreads();
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
    for (int i = tid; i < n; i += (int)blockDim.x) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double x = (smp[(2 * h) * n + i] - mu) * inv_s, y = smp[(2 * h + 1) * n + i];
            double xp = 1.0;
#pragma unroll
            for (int k = 0; k <= 2 * P; ++k) {
                acc[k] += xp;
                if (k <= P) acc[2 * P + 1 + k] = fma(xp, y, acc[2 * P + 1 + k]);
                xp *= x;
            }
        }
    }
    sb_sum<NV>(acc, mom);
    if (tid == 0) {
        for (int k = 0; k < COEF_LD; ++k) cf[k] = 0.0;
        if (2.0 * (double)n >= (double)(P + 1)) solve_normal_equations<P>(mom, cf);           // :152-154, :166
    }
    __syncthreads();
    double c[P + 1];
#pragma unroll
    for (int k = 0; k <= P; ++k) c[k] = cf[k];
    auto poly = [&](double s) {
        const double x = (s - mu) * inv_s;
        double v = c[P];
#pragma unroll
        for (int k = P - 1; k >= 0; --k) v = fma(v, x, c[k]);
        return v;
    };
    double of[3] = {0.0, 0.0, 0.0};
    for (int i = tid; i < n; i += (int)blockDim.x) of[0] += poly(sb_ld<ST>(S + i));                     // :172-177
    sb_sum<3>(of, sc);
    const double offset = sc[0] / (double)n;
    __syncthreads();
    double du[3] = {0.0, 0.0, 0.0};
    for (int i = tid; i < n; i += (int)blockDim.x) {                                                    // :96-117
        double best = 0.0;
        for (int j = 0; j < jend; ++j) {
            const double s = sb_ld<ST>(S + (int64_t)j * ld + i);
            const double cand = payoff_fn(is_call, s, K) * DF[j] - (poly(s) - offset);
            if (cand > best) best = cand;
        }
        du[0] += best;
    }
    sb_sum<3>(du, sc);
    if (tid == 0) { out[0] = primal_sum; out[1] = sc[0]; }
    __syncthreads();
}
