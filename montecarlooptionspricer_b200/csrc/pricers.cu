// pricers.cu -- the other three pricing-method plugins of the reference as streaming kernels over the same
// time-major slab:
//   AsymptoticAnalysis::PredictOptionPrice      src/models/AsymptoticAnalysisPricer.cpp:38-113
//   MartingaleOptimization::PredictOptionPrice  src/models/MartingaleOptimizationPricer.cpp:21-188
//   BranchingProcesses::PredictOptionPrice      src/models/BranchingProcessPricer.cpp:13-134
// All three are per-path scans over time (thread <-> path, so every load is a coalesced run of consecutive paths
// at one time index) plus, at most, one tiny regression (Martingale) or random cross-path gathers (Branching).
// Every comparison is made in fp64 on the stored path values, with the reference's predicates kept literally.
// Per-step scalars (discount factors, exercise boundary, maturity cut) are evaluated on the host in double exactly
// as the reference writes them and handed to the kernels as tables.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "lsm_solve.cuh"
#include "philox.cuh"
#include "small_bodies.cuh"

namespace {

constexpr int PR_NT = 256;
constexpr int PR_LD = 24;  // partial-row stride (>= 3 * MAXP + 2)

template <typename ST>
__device__ __forceinline__ double ldS(const ST* p);
template <>
__device__ __forceinline__ double ldS<float>(const float* p) { return f2d(__ldg(p)); }
template <>
__device__ __forceinline__ double ldS<double>(const double* p) { return __ldg(p); }

// Block-wide deterministic sum of NV doubles per thread -> row `blockIdx.x` of `partial`.
template <int NV>
__device__ __forceinline__ void block_sum_to_row(double (&acc)[NV], double* __restrict__ row) {
    __shared__ double red[PR_NT / 32][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const double s = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < PR_NT / 32; ++w) s += red[w][threadIdx.x];
        row[threadIdx.x] = s;
    }
}

// Fold the per-CTA rows in a fixed order (bitwise reproducible) -> out[0..nv).
__global__ void __launch_bounds__(256) fold_rows_kernel(const double* __restrict__ partial, int nblocks, int nv, double* __restrict__ out) {
    __shared__ double red[8][32];
    const int k = threadIdx.x & 31, grp = threadIdx.x >> 5;
    double s = 0.0;
    if (k < nv)
        for (int b = grp; b < nblocks; b += 8) s += partial[(int64_t)b * PR_LD + k];
    red[grp][k] = s;
    __syncthreads();
    if (threadIdx.x < nv) {
        double t = 0.0;
#pragma unroll
        for (int g = 0; g < 8; ++g) t += red[g][threadIdx.x];
        out[threadIdx.x] = t;
    }
}

// ------------------------------------------------------------------------------------------- Asymptotic
// Per path: best = max over j (j dt <= maturity) of e^{-r t_j} payoff(S_j) restricted to the exercise region
// S < b_j (put) / S > b_j (call); non-finite S are skipped (AsymptoticAnalysisPricer.cpp:67-97).  Mean over the
// paths whose best is finite (:99-108).
template <typename ST>
__global__ void __launch_bounds__(PR_NT) asym_kernel(const ST* __restrict__ S, int64_t ld, int64_t n, int jend, const double* __restrict__ bnd,
                                                    const double* __restrict__ disc, double K, int is_call, double* __restrict__ partial) {
    double acc[2] = {0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * PR_NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PR_NT) {
        double best = 0.0;
#pragma unroll 4
        for (int j = 0; j < jend; ++j) {
            const double s = ldS<ST>(S + (int64_t)j * ld + i);
            if (isnan(s) || isinf(s)) continue;                   // :74
            const double b = bnd[j];
            const bool in = is_call ? (s > b) : (s < b);          // :80-85 (NaN boundary => never in the region)
            if (in) {
                const double pay = payoff_fn(is_call, s, K);
                if (isnan(pay) || isinf(pay)) continue;           // :89
                const double d = disc[j] * pay;                   // :90
                if (d > best) best = d;
            }
        }
        if (!isnan(best) && !isinf(best)) { acc[0] += best; acc[1] += 1.0; }  // :101-106
    }
    block_sum_to_row<2>(acc, partial + (int64_t)blockIdx.x * PR_LD);
}

// ------------------------------------------------------------------------------------------- Martingale
// Pass 1 (MartingaleOptimizationPricer.cpp:72-94 and :130-150): per path the best discounted payoff, its (first)
// index, and the two regression samples (S_stop, 0.5 dp_stop), (S_other, 0.2 dp_other), jOther = (jStop + M/2) % M.
template <typename ST>
__global__ void __launch_bounds__(PR_NT) mart_primal_kernel(const ST* __restrict__ S, int64_t ld, int64_t n, int M, int jend,
                                                           const double* __restrict__ DF, double K, int is_call, double* __restrict__ smp /*[4][ld]*/,
                                                           double* __restrict__ partial) {
    double acc[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * PR_NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PR_NT) {
        double best = 0.0, s_stop = ldS<ST>(S + i);
        int idx = 0;
#pragma unroll 4
        for (int j = 0; j < jend; ++j) {
            const double s = ldS<ST>(S + (int64_t)j * ld + i);
            const double dp = payoff_fn(is_call, s, K) * DF[j];
            if (dp > best) { best = dp; idx = j; s_stop = s; }
        }
        acc[0] += best;
        const int jo = (idx + M / 2) % M;
        const double s_other = ldS<ST>(S + (int64_t)jo * ld + i);
        smp[i] = s_stop;
        smp[ld + i] = 0.5 * (payoff_fn(is_call, s_stop, K) * DF[idx]);
        smp[2 * ld + i] = s_other;
        smp[3 * ld + i] = 0.2 * (payoff_fn(is_call, s_other, K) * DF[jo]);
    }
    block_sum_to_row<1>(acc, partial + (int64_t)blockIdx.x * PR_LD);
}

// cnt / sum x / sum x^2 over the samples of the first `ns` paths (standardisation of the regression variable)
__global__ void __launch_bounds__(PR_NT) mart_stats_kernel(const double* __restrict__ smp, int64_t ld, int ns, double* __restrict__ out) {
    double acc[3] = {0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < ns; i += PR_NT) {
        const double a = smp[i], b = smp[2 * ld + i];
        acc[0] += 2.0;
        acc[1] += a + b;
        acc[2] += a * a + b * b;
    }
    block_sum_to_row<3>(acc, out);
}

// Normal-equation moments of the 2N-sample regression (:152-166) in x = (X - mu) inv_s.
template <int P>
__global__ void __launch_bounds__(PR_NT) mart_moments_kernel(const double* __restrict__ smp, int64_t ld, int64_t n, const double* __restrict__ stats,
                                                            double K, double* __restrict__ partial) {
    constexpr int NV = 3 * P + 2;
    const double cnt = stats[0];
    double mu = K, sd = fabs(K) > 0.0 ? fabs(K) : 1.0;
    if (cnt >= 2.0) {
        mu = stats[1] / cnt;
        const double var = (stats[2] - cnt * mu * mu) / (cnt - 1.0);
        sd = var > 1e-12 * mu * mu ? sqrt(var) : (fabs(mu) > 0.0 ? fabs(mu) : 1.0);
    }
    const double inv_s = 1.0 / sd;
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * PR_NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PR_NT) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double x = (smp[(2 * h) * ld + i] - mu) * inv_s, y = smp[(2 * h + 1) * ld + i];
            double xp = 1.0;
#pragma unroll
            for (int k = 0; k <= 2 * P; ++k) {
                acc[k] += xp;
                if (k <= P) acc[2 * P + 1 + k] = fma(xp, y, acc[2 * P + 1 + k]);
                xp *= x;
            }
        }
    }
    block_sum_to_row<NV>(acc, partial + (int64_t)blockIdx.x * PR_LD);
    if (blockIdx.x == 0 && threadIdx.x == 0) { partial[(int64_t)gridDim.x * PR_LD] = mu; partial[(int64_t)gridDim.x * PR_LD + 1] = inv_s; }
}

__global__ void mart_solve_kernel(const double* __restrict__ mom, int p, double n_samples_global, double* __restrict__ coef) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int k = 0; k < COEF_LD; ++k) coef[k] = 0.0;
    if (n_samples_global < (double)(p + 1)) return;  // :152-154: too few samples, the martingale stays zero
    solve_dispatch(mom, p, coef);
}

__device__ __forceinline__ double poly_std(const double* __restrict__ c, int p, double x) {
    double v = c[p];
    for (int k = p - 1; k >= 0; --k) v = fma(v, x, c[k]);
    return v;
}

// offset = mean_i M(S_i0) (:172-177): per-CTA sums of the fitted polynomial at column 0
template <typename ST>
__global__ void __launch_bounds__(PR_NT) mart_offset_kernel(const ST* __restrict__ S, int64_t n, const double* __restrict__ coef, int p,
                                                           const double* __restrict__ musig, double* __restrict__ partial) {
    const double mu = musig[0], inv_s = musig[1];
    double acc[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * PR_NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PR_NT)
        acc[0] += poly_std(coef, p, (ldS<ST>(S + i) - mu) * inv_s);
    block_sum_to_row<1>(acc, partial + (int64_t)blockIdx.x * PR_LD);
}

// Pass 2 (:96-117): dual_i = max(0, max_j [dp_j - (M(S_j) - offset)])
template <typename ST>
__global__ void __launch_bounds__(PR_NT) mart_dual_kernel(const ST* __restrict__ S, int64_t ld, int64_t n, int jend, const double* __restrict__ DF,
                                                         double K, int is_call, const double* __restrict__ coef, int p,
                                                         const double* __restrict__ musig, const double* __restrict__ offs /*[sum, n]*/,
                                                         double* __restrict__ partial) {
    const double mu = musig[0], inv_s = musig[1], offset = offs[0] / offs[1];
    double c[MAXP + 1];
    for (int k = 0; k <= MAXP; ++k) c[k] = k <= p ? coef[k] : 0.0;
    double acc[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * PR_NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PR_NT) {
        double best = 0.0;
#pragma unroll 2
        for (int j = 0; j < jend; ++j) {
            const double s = ldS<ST>(S + (int64_t)j * ld + i);
            const double dp = payoff_fn(is_call, s, K) * DF[j];
            const double cand = dp - (poly_std(c, p, (s - mu) * inv_s) - offset);
            if (cand > best) best = cand;
        }
        acc[0] += best;
    }
    block_sum_to_row<1>(acc, partial + (int64_t)blockIdx.x * PR_LD);
}

// ------------------------------------------------------------------------------------------- Branching
// Lower bound (BranchingProcessPricer.cpp:41-71): exercise at the FIRST listed date whose discounted payoff is positive.
template <typename ST>
__global__ void __launch_bounds__(PR_NT) branch_lower_kernel(const ST* __restrict__ S, int64_t ld, int64_t n, const int* __restrict__ ex, int n_ex,
                                                            const double* __restrict__ disc, double K, int is_call, double* __restrict__ partial) {
    double acc[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * PR_NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PR_NT) {
        double best = 0.0;
        for (int e = 0; e < n_ex; ++e) {
            const int j = ex[e];
            const double d = disc[j] * payoff_fn(is_call, ldS<ST>(S + (int64_t)j * ld + i), K);
            if (d > best) { best = d; break; }
        }
        acc[0] += best;
    }
    block_sum_to_row<1>(acc, partial + (int64_t)blockIdx.x * PR_LD);
}

// Upper bound (:73-134), one launch per time index j, descending.  F_old[i] = max_{k > j, t_k <= T} e^{-r t_k} payoff(S_k[i])
// (the reference's inner loop :110-121, times e^{-r t}); F_new adds index j.  At an exercise date each path draws
// numBranches random OTHER paths of this shard and averages their F_old (:104-124).  Ping-pong buffers keep the
// gathers of step j apart from the update of step j.
template <typename ST>
__global__ void __launch_bounds__(PR_NT) branch_upper_kernel(const ST* __restrict__ Sj, int64_t n, int j, int j_valid, int is_ex, int has_cont,
                                                            double disc_j, double K, int is_call, int n_br, const double* __restrict__ F_old,
                                                            double* __restrict__ F_new, double* __restrict__ best, PhiloxKeys keys,
                                                            uint64_t path_offset, const int32_t* __restrict__ inj /*[n][n_br] or null*/) {
    for (int64_t i = (int64_t)blockIdx.x * PR_NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PR_NT) {
        const double d = disc_j * payoff_fn(is_call, ldS<ST>(Sj + i), K);
        const double fo = F_old[i];
        F_new[i] = (j_valid && d > fo) ? d : fo;
        if (is_ex) {
            double cont = 0.0;
            if (has_cont) {                                             // :103 tIdx < exerciseTimes.back()
                double sum = 0.0;
                const uint64_t gid = path_offset + (uint64_t)i;
                for (int b0 = 0; b0 < n_br; b0 += 4) {
                    uint4 u = make_uint4(0u, 0u, 0u, 0u);
                    if (!inj) u = philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)j, 0x10000u + (uint32_t)(b0 >> 2), keys);
                    const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (b0 + q < n_br) {
                            const int64_t rp = inj ? (int64_t)inj[i * n_br + b0 + q] : (int64_t)(((uint64_t)uu[q] * (uint64_t)n) >> 32);
                            sum += F_old[rp];
                        }
                    }
                }
                cont = sum / (double)n_br;                               // :123 (e^{-r t} is already inside F)
            }
            const double better = d < cont ? cont : d;                    // std::max(discNow, continuation) :126
            if (better > best[i]) best[i] = better;                      // :127-129
        }
    }
}

// Single-launch forms for small path sets (<= 4096 paths: the reference's rows are 250): one CTA runs the whole
// pricer out of shared memory (small_bodies.cuh).  Same arithmetic and, for Branching, the same resampling stream as
// the streaming kernels above.
constexpr int BR_SMALL_NT = SB_NT;
constexpr int BR_SMALL_MAX = SB_MAX_PATHS;

template <typename ST>
__global__ void __launch_bounds__(SB_NT, 1) branch_small_kernel(const ST* __restrict__ S, int64_t ld, int n, int j_hi, int j_lo, int kend, int ex_back,
                                                               const int* __restrict__ is_ex, const int* __restrict__ ex, int n_ex, double r, double dt,
                                                               double K, int is_call, int n_br, PhiloxKeys keys, uint64_t path_offset,
                                                               const int32_t* __restrict__ inj, double* __restrict__ out) {
    extern __shared__ double sm[];  // F[n] | best[n]
    sb_branching<ST>(S, ld, n, j_hi, j_lo, kend, ex_back, is_ex, ex, n_ex, r, dt, K, is_call, n_br, keys, path_offset, inj, sm, sm + n, out);
}

template <typename ST>
__global__ void __launch_bounds__(SB_NT, 1) asym_small_kernel(const ST* __restrict__ S, int64_t ld, int n, int M, double K, int is_call, double r, double dt,
                                                             double maturity, double sigma, double dividend, double* __restrict__ out) {
    extern __shared__ double sm[];  // boundary[M] | discount[M]
    sb_asymptotic<ST>(S, ld, n, M, K, is_call, r, dt, maturity, sigma, dividend, sm, out);
}

template <typename ST, int P>
__global__ void __launch_bounds__(SB_NT, 1) mart_small_kernel(const ST* __restrict__ S, int64_t ld, int n, int M, double K, int is_call, double r, double dt,
                                                             double maturity, int max_iterations, double* __restrict__ out) {
    extern __shared__ double sm[];  // samples[4 n] | DF[M]
    sb_martingale<ST, P>(S, ld, n, M, K, is_call, r, dt, maturity, max_iterations, sm, sm + 4 * (size_t)n, out);
}

typedef void (*MartSmallFn)(const void*, int64_t, int, int, double, int, double, double, double, int, double*);
template <typename ST>
MartSmallFn pick_mart_small(int p) {
    switch (p) {
        case 0: return (MartSmallFn)mart_small_kernel<ST, 0>;
        case 1: return (MartSmallFn)mart_small_kernel<ST, 1>;
        case 2: return (MartSmallFn)mart_small_kernel<ST, 2>;
        case 3: return (MartSmallFn)mart_small_kernel<ST, 3>;
        case 4: return (MartSmallFn)mart_small_kernel<ST, 4>;
        case 5: return (MartSmallFn)mart_small_kernel<ST, 5>;
        default: return (MartSmallFn)mart_small_kernel<ST, 6>;
    }
}

bool use_small(const mcp_ctx* ctx, int64_t N, int M) {
    return N <= SB_MAX_PATHS && M <= 4096 && !(ctx->nranks > 1 && ctx->comm) && !getenv("MCP_PRICERS_STREAMING");
}

__global__ void __launch_bounds__(PR_NT) sum_vector_kernel(const double* __restrict__ v, int64_t n, double* __restrict__ partial) {
    double acc[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * PR_NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PR_NT) acc[0] += v[i];
    block_sum_to_row<1>(acc, partial + (int64_t)blockIdx.x * PR_LD);
}

int grid_for(const mcp_ctx* ctx, int64_t n) {
    int64_t g = (n + PR_NT - 1) / PR_NT;
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

// first j with j*dt > maturity, evaluated in double exactly like the reference's `if (t > maturity) break;`
int cut_index(int M, double dt, double maturity) {
    for (int j = 0; j < M; ++j)
        if ((double)j * dt > maturity) return j;
    return M;
}

template <typename F>
int fold_sum(mcp_ctx* ctx, double* d_partial, int grid, int nv, double* d_out, F&& launch) {
    launch();
    MCP_LAUNCH_CHECK(ctx);
    fold_rows_kernel<<<1, 256, 0, ctx->stream>>>(d_partial, grid, nv, d_out);
    MCP_LAUNCH_CHECK(ctx);
    return MCP_OK;
}

}  // namespace

// ===================================================================================== Asymptotic (C ABI)
extern "C" int mcp_asymptotic_price(mcp_ctx* ctx, const mcp_pathset* ps, double r, double strike, double maturity, double dt, int is_call,
                                    double sigma, double dividend, double* price) {
    if (!ctx || !price) return MCP_ERR_INVALID;
    *price = 0.0;
    if (!ps || ps->n_paths <= 0) return MCP_OK;  // the reference returns 0.0 on empty input (:48-50)
    if (ps->ctx != ctx) return mcp_fail(ctx, MCP_ERR_INVALID, "asymptotic: pathset belongs to another ctx");
    if (!(sigma > 0.0)) return mcp_fail(ctx, MCP_ERR_DOMAIN, "AsymptoticAnalysis: Volatility must be positive.");  // :51-53
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));
    const int M = ps->n_steps + 1;
    const int64_t N = ps->n_paths;
    const int jend = cut_index(M, dt, maturity);
    std::vector<double> tab(2 * (size_t)M);
    for (int j = 0; j < M; ++j) {
        const double t = (double)j * dt, eps = maturity - t;
        double b = strike;
        if (!(eps < 1e-10)) {                                           // :10-11 / :25-26
            const double c0 = 0.5 * sigma * sqrt(eps * log(1.0 / eps));  // NaN for eps > 1, as in the reference
            if (is_call) { b = strike - c0; if (eps < 0.01) b += 0.5 * (dividend - r) * eps; }   // :28-34
            else         { b = strike + c0; if (eps < 0.01) b -= 0.5 * (r - dividend) * eps; }   // :13-19
        }
        tab[j] = b;
        tab[M + j] = exp(-r * t);
    }
    const int grid = grid_for(ctx, N);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_tab = take(2 * (size_t)M * 8), o_part = take((size_t)grid * PR_LD * 8), o_out = take(4 * 8);
    MCP_TRY(mcp_scratch_reserve(ctx, off));
    unsigned char* sb = (unsigned char*)ctx->scratch;
    double *d_tab = (double*)(sb + o_tab), *d_part = (double*)(sb + o_part), *d_out = (double*)(sb + o_out);
    cudaStream_t st = ctx->stream;
    if (use_small(ctx, N, M)) {
        const size_t smem = 2 * (size_t)M * sizeof(double);
        if (ps->dtype == MCP_F32) {
            MCP_TRY(mcp_kernel_config(ctx, (const void*)asym_small_kernel<float>, SB_NT, 2 * 4096 * sizeof(double), nullptr));
            asym_small_kernel<float><<<1, SB_NT, smem, st>>>((const float*)ps->data, ps->ld, (int)N, M, strike, is_call, r, dt, maturity, sigma, dividend, d_out);
        } else {
            MCP_TRY(mcp_kernel_config(ctx, (const void*)asym_small_kernel<double>, SB_NT, 2 * 4096 * sizeof(double), nullptr));
            asym_small_kernel<double><<<1, SB_NT, smem, st>>>((const double*)ps->data, ps->ld, (int)N, M, strike, is_call, r, dt, maturity, sigma, dividend, d_out);
        }
        MCP_LAUNCH_CHECK(ctx);
    } else {
    MCP_TRY(mcp_h2d(ctx, d_tab, tab.data(), 2 * (size_t)M * 8));
    MCP_TRY(fold_sum(ctx, d_part, grid, 2, d_out, [&] {
        if (ps->dtype == MCP_F32) asym_kernel<float><<<grid, PR_NT, 0, st>>>((const float*)ps->data, ps->ld, N, jend, d_tab, d_tab + M, strike, is_call, d_part);
        else asym_kernel<double><<<grid, PR_NT, 0, st>>>((const double*)ps->data, ps->ld, N, jend, d_tab, d_tab + M, strike, is_call, d_part);
    }));
    }
    MCP_TRY(mcp_allreduce_f64(ctx, d_out, 2));
    double h[2] = {0, 0};
    double* hp = (double*)mcp_stage_alloc(ctx, 16);
    MCP_CUDA(ctx, mcp_memcpy_async(ctx, hp ? hp : h, d_out, 16, cudaMemcpyDeviceToHost, st));
    MCP_CUDA(ctx, cudaStreamSynchronize(st));
    if (hp) memcpy(h, hp, 16);
    *price = h[1] > 0.0 ? h[0] / h[1] : 0.0;  // :108
    return MCP_OK;
}

// ===================================================================================== Martingale (C ABI)
// The reference runs maxIterations rounds of {primal, dual with the PREVIOUS martingale, refit} (:54-61).  The stopping
// indices, the primal value and the regression samples depend only on the paths (:72-94, :130-150), so the fitted
// martingale is the same after every round: the returned 0.5 (primal + dual) uses M = 0 when maxIterations == 1 and
// the single fitted polynomial otherwise.  Two passes over the slab instead of 2 x maxIterations.
extern "C" int mcp_martingale_price(mcp_ctx* ctx, const mcp_pathset* ps, double r, double strike, double maturity, double dt, int is_call,
                                    int poly_order, int max_iterations, double* price, double* primal_out, double* dual_out) {
    if (!ctx || !price) return MCP_ERR_INVALID;
    if (!ps || ps->n_paths <= 0) return mcp_fail(ctx, MCP_ERR_EMPTY_PATHS, "MartingaleOptimization: Empty pricePaths.");  // :31-33
    if (max_iterations <= 0) return mcp_fail(ctx, MCP_ERR_DOMAIN, "MartingaleOptimization: maxIterations must be positive.");  // :34-36
    if (ps->ctx != ctx) return mcp_fail(ctx, MCP_ERR_INVALID, "martingale: pathset belongs to another ctx");
    const int p = poly_order;
    if (p < 0) return mcp_fail(ctx, MCP_ERR_INVALID, "martingale: poly_order %d < 0", p);
    if (p > MAXP) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "martingale: poly_order %d > %d", p, MAXP);
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));
    const int M = ps->n_steps + 1;
    const int64_t N = ps->n_paths;
    const int jend = cut_index(M, dt, maturity);
    std::vector<double> DF(M);
    for (int j = 0; j < M; ++j) {  // PathDiscountFactor, MartingaleOptimizationPricer.h:44-49
        double t = (double)j * dt;
        if (t > maturity) t = maturity;
        DF[j] = exp(-r * t);
    }
    const int grid = grid_for(ctx, N);
    const int nm = 3 * p + 2;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_df = take((size_t)M * 8), o_part = take(((size_t)grid + 1) * PR_LD * 8), o_stats = take(PR_LD * 8), o_mom = take(PR_LD * 8);
    const size_t o_coef = take(COEF_LD * 8), o_fin = take(8 * 8);
    MCP_TRY(mcp_scratch_reserve(ctx, off));
    MCP_TRY(mcp_carry_reserve(ctx, (size_t)4 * ps->ld * 8));
    unsigned char* sb = (unsigned char*)ctx->scratch;
    double *d_df = (double*)(sb + o_df), *d_part = (double*)(sb + o_part), *d_stats = (double*)(sb + o_stats), *d_mom = (double*)(sb + o_mom);
    double *d_coef = (double*)(sb + o_coef), *d_fin = (double*)(sb + o_fin);  // fin: [0] primal sum, [1] n, [2] offset sum, [3] n, [4] dual sum
    double* d_smp = (double*)ctx->carry;
    double* d_musig = d_part + (size_t)grid * PR_LD;
    cudaStream_t st = ctx->stream;
    const bool f32 = ps->dtype == MCP_F32;
    const double nloc = (double)N;
    MCP_TRY(mcp_h2d(ctx, d_df, DF.data(), (size_t)M * 8));
    MCP_CUDA(ctx, cudaMemsetAsync(d_fin, 0, 8 * 8, st));
    MCP_TRY(mcp_h2d(ctx, d_fin + 1, &nloc, 8));
    MCP_TRY(mcp_h2d(ctx, d_fin + 3, &nloc, 8));

    if (use_small(ctx, N, M)) {
        // one launch: d_fin[0] = primal sum, d_fin[4] = dual sum (laid out like the streaming path's results)
        const size_t smem = (4 * (size_t)N + (size_t)M) * sizeof(double);
        MartSmallFn fn = f32 ? pick_mart_small<float>(p) : pick_mart_small<double>(p);
        MCP_TRY(mcp_kernel_config(ctx, (const void*)fn, SB_NT, (4 * (size_t)SB_MAX_PATHS + 4096) * sizeof(double), nullptr));
        fn<<<1, SB_NT, smem, st>>>(ps->data, ps->ld, (int)N, M, strike, is_call, r, dt, maturity, max_iterations, d_fin + 5);
        MCP_LAUNCH_CHECK(ctx);
        double h2[8];
        double* hp2 = (double*)mcp_stage_alloc(ctx, 64);
        MCP_CUDA(ctx, mcp_memcpy_async(ctx, hp2 ? hp2 : h2, d_fin, 8 * 8, cudaMemcpyDeviceToHost, st));
        MCP_CUDA(ctx, cudaStreamSynchronize(st));
        if (hp2) memcpy(h2, hp2, 64);
        const double primal = h2[5] / (double)N, dual = h2[6] / (double)N;
        if (primal_out) *primal_out = primal;
        if (dual_out) *dual_out = dual;
        *price = 0.5 * (primal + dual);  // :63
        return MCP_OK;
    }
    // pass 1: primal + regression samples
    MCP_TRY(fold_sum(ctx, d_part, grid, 1, d_fin, [&] {
        if (f32) mart_primal_kernel<float><<<grid, PR_NT, 0, st>>>((const float*)ps->data, ps->ld, N, M, jend, d_df, strike, is_call, d_smp, d_part);
        else mart_primal_kernel<double><<<grid, PR_NT, 0, st>>>((const double*)ps->data, ps->ld, N, M, jend, d_df, strike, is_call, d_smp, d_part);
    }));
    MCP_TRY(mcp_allreduce_f64(ctx, d_fin, 2));  // primal sum, N
    if (max_iterations >= 2) {
        // one regression over the 2N samples
        const int ns = (int)(N < 16384 ? N : 16384);
        mart_stats_kernel<<<1, PR_NT, 0, st>>>(d_smp, ps->ld, ns, d_stats);
        MCP_LAUNCH_CHECK(ctx);
        MCP_TRY(mcp_allreduce_f64(ctx, d_stats, 3));
        MCP_TRY(fold_sum(ctx, d_part, grid, nm, d_mom, [&] {
            switch (p) {
                case 0: mart_moments_kernel<0><<<grid, PR_NT, 0, st>>>(d_smp, ps->ld, N, d_stats, strike, d_part); break;
                case 1: mart_moments_kernel<1><<<grid, PR_NT, 0, st>>>(d_smp, ps->ld, N, d_stats, strike, d_part); break;
                case 2: mart_moments_kernel<2><<<grid, PR_NT, 0, st>>>(d_smp, ps->ld, N, d_stats, strike, d_part); break;
                case 3: mart_moments_kernel<3><<<grid, PR_NT, 0, st>>>(d_smp, ps->ld, N, d_stats, strike, d_part); break;
                case 4: mart_moments_kernel<4><<<grid, PR_NT, 0, st>>>(d_smp, ps->ld, N, d_stats, strike, d_part); break;
                case 5: mart_moments_kernel<5><<<grid, PR_NT, 0, st>>>(d_smp, ps->ld, N, d_stats, strike, d_part); break;
                default: mart_moments_kernel<6><<<grid, PR_NT, 0, st>>>(d_smp, ps->ld, N, d_stats, strike, d_part); break;
            }
        }));
        MCP_TRY(mcp_allreduce_f64(ctx, d_mom, nm));
        const double n_samples = 2.0 * (double)N * (double)(ctx->comm ? ctx->nranks : 1);
        mart_solve_kernel<<<1, 32, 0, st>>>(d_mom, p, n_samples, d_coef);
        MCP_LAUNCH_CHECK(ctx);
        MCP_TRY(fold_sum(ctx, d_part, grid, 1, d_fin + 2, [&] {
            if (f32) mart_offset_kernel<float><<<grid, PR_NT, 0, st>>>((const float*)ps->data, N, d_coef, p, d_musig, d_part);
            else mart_offset_kernel<double><<<grid, PR_NT, 0, st>>>((const double*)ps->data, N, d_coef, p, d_musig, d_part);
        }));
        MCP_TRY(mcp_allreduce_f64(ctx, d_fin + 2, 2));  // offset sum, N
        // pass 2: dual
        MCP_TRY(fold_sum(ctx, d_part, grid, 1, d_fin + 4, [&] {
            if (f32) mart_dual_kernel<float><<<grid, PR_NT, 0, st>>>((const float*)ps->data, ps->ld, N, jend, d_df, strike, is_call, d_coef, p, d_musig, d_fin + 2, d_part);
            else mart_dual_kernel<double><<<grid, PR_NT, 0, st>>>((const double*)ps->data, ps->ld, N, jend, d_df, strike, is_call, d_coef, p, d_musig, d_fin + 2, d_part);
        }));
        MCP_TRY(mcp_allreduce_f64(ctx, d_fin + 4, 1));
    }
    double h[8];
    double* hp = (double*)mcp_stage_alloc(ctx, 64);
    MCP_CUDA(ctx, mcp_memcpy_async(ctx, hp ? hp : h, d_fin, 8 * 8, cudaMemcpyDeviceToHost, st));
    MCP_CUDA(ctx, cudaStreamSynchronize(st));
    if (hp) memcpy(h, hp, 64);
    const double primal = h[0] / h[1];
    const double dual = max_iterations >= 2 ? h[4] / h[1] : primal;  // round 1 uses M = 0, offset = 0: dual == primal
    if (primal_out) *primal_out = primal;
    if (dual_out) *dual_out = dual;
    *price = 0.5 * (primal + dual);  // :63
    return MCP_OK;
}

// ====================================================================================== Branching (C ABI)
extern "C" int mcp_branching_price(mcp_ctx* ctx, const mcp_pathset* ps, double r, double strike, double maturity, double dt, int is_call,
                                   int num_branches, const int* exercise_times, int n_exercise, uint64_t seed, uint64_t path_offset,
                                   const int32_t* injected_rp, double* price, double* lower_out, double* upper_out) {
    if (!ctx || !price) return MCP_ERR_INVALID;
    if (!ps || ps->n_paths <= 0) return mcp_fail(ctx, MCP_ERR_EMPTY_PATHS, "BranchingProcesses: Empty pricePaths.");  // :22-24
    if (!exercise_times || n_exercise <= 0) return mcp_fail(ctx, MCP_ERR_DOMAIN, "BranchingProcesses: No exercise times.");  // :25-27
    if (!(strike > 0.0)) return mcp_fail(ctx, MCP_ERR_DOMAIN, "BranchingProcesses: Strike must be positive.");  // :28-30
    if (ps->ctx != ctx) return mcp_fail(ctx, MCP_ERR_INVALID, "branching: pathset belongs to another ctx");
    if (num_branches <= 0) return mcp_fail(ctx, MCP_ERR_INVALID, "branching: numBranches must be positive");
    const int M = ps->n_steps + 1;
    const int64_t N = ps->n_paths;
    if (N >= ((int64_t)1 << 31)) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "branching: more than 2^31 paths per rank");
    for (int e = 0; e < n_exercise; ++e) {
        if (exercise_times[e] < 0 || exercise_times[e] >= M) return mcp_fail(ctx, MCP_ERR_INVALID, "branching: exercise index %d outside [0,%d)", exercise_times[e], M);
        if (e > 0 && exercise_times[e] <= exercise_times[e - 1])
            return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "branching: exercise times must be strictly increasing (as PredictionGen.cpp:780-783 builds them)");
    }
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));
    // dates the loops actually visit: the prefix before the first t > maturity (`break`, :58-61 / :95-98)
    int n_ex = 0;
    while (n_ex < n_exercise && !((double)exercise_times[n_ex] * dt > maturity)) ++n_ex;
    const int ex_back = exercise_times[n_exercise - 1];
    const int kend = cut_index(M, dt, maturity);  // inner loop :110-114 stops at the first t_k > maturity
    std::vector<double> disc(M);
    for (int j = 0; j < M; ++j) disc[j] = exp(-r * ((double)j * dt));
    std::vector<int> is_ex(M, 0);
    for (int e = 0; e < n_ex; ++e) is_ex[exercise_times[e]] = e + 1;

    const int grid = grid_for(ctx, N);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_disc = take((size_t)M * 8), o_ex = take((size_t)(n_ex > 0 ? n_ex : 1) * 4), o_part = take((size_t)grid * PR_LD * 8), o_fin = take(4 * 8);
    const size_t o_isex = take((size_t)M * 4);
    const size_t o_inj = take(injected_rp ? (size_t)N * num_branches * 4 * (size_t)(N <= BR_SMALL_MAX && n_ex > 0 ? n_ex : 1) : 0);
    MCP_TRY(mcp_scratch_reserve(ctx, off));
    MCP_TRY(mcp_carry_reserve(ctx, (size_t)3 * ps->ld * 8));
    unsigned char* sb = (unsigned char*)ctx->scratch;
    double *d_disc = (double*)(sb + o_disc), *d_part = (double*)(sb + o_part), *d_fin = (double*)(sb + o_fin);
    int* d_ex = (int*)(sb + o_ex);
    int* d_isex = (int*)(sb + o_isex);
    int32_t* d_inj = injected_rp ? (int32_t*)(sb + o_inj) : nullptr;
    double* F0 = (double*)ctx->carry;
    double* F1 = F0 + ps->ld;
    double* d_best = F1 + ps->ld;
    cudaStream_t st = ctx->stream;
    const bool f32 = ps->dtype == MCP_F32;
    const double nloc = (double)N;
    MCP_TRY(mcp_h2d(ctx, d_disc, disc.data(), (size_t)M * 8));
    if (n_ex > 0) MCP_TRY(mcp_h2d(ctx, d_ex, exercise_times, (size_t)n_ex * 4));
    MCP_CUDA(ctx, cudaMemsetAsync(d_fin, 0, 4 * 8, st));
    MCP_TRY(mcp_h2d(ctx, d_fin + 2, &nloc, 8));
    MCP_CUDA(ctx, cudaMemsetAsync(F0, 0, (size_t)3 * ps->ld * 8, st));

    const PhiloxKeys keys = philox_make_keys(seed);
    const bool small = use_small(ctx, N, M) && !getenv("MCP_BRANCH_PER_DATE") && n_ex > 0;
    if (small) {
        // one launch: both bounds
        MCP_TRY(mcp_h2d(ctx, d_isex, is_ex.data(), (size_t)M * 4));
        if (d_inj) MCP_CUDA(ctx, mcp_memcpy_async(ctx, d_inj, injected_rp, (size_t)n_ex * N * num_branches * 4, cudaMemcpyHostToDevice, st));
        const int j_hi = kend - 1 > exercise_times[n_ex - 1] ? kend - 1 : exercise_times[n_ex - 1];
        const size_t smem = (size_t)2 * N * sizeof(double);
        if (f32) {
            MCP_TRY(mcp_kernel_config(ctx, (const void*)branch_small_kernel<float>, BR_SMALL_NT, (size_t)2 * BR_SMALL_MAX * sizeof(double), nullptr));
            branch_small_kernel<float><<<1, BR_SMALL_NT, smem, st>>>((const float*)ps->data, ps->ld, (int)N, j_hi, exercise_times[0], kend, ex_back, d_isex, d_ex, n_ex,
                                                                   r, dt, strike, is_call, num_branches, keys, path_offset, d_inj, d_fin);
        } else {
            MCP_TRY(mcp_kernel_config(ctx, (const void*)branch_small_kernel<double>, BR_SMALL_NT, (size_t)2 * BR_SMALL_MAX * sizeof(double), nullptr));
            branch_small_kernel<double><<<1, BR_SMALL_NT, smem, st>>>((const double*)ps->data, ps->ld, (int)N, j_hi, exercise_times[0], kend, ex_back, d_isex, d_ex, n_ex,
                                                                    r, dt, strike, is_call, num_branches, keys, path_offset, d_inj, d_fin);
        }
        MCP_LAUNCH_CHECK(ctx);
    } else {
    // lower bound
    MCP_TRY(fold_sum(ctx, d_part, grid, 1, d_fin, [&] {
        if (f32) branch_lower_kernel<float><<<grid, PR_NT, 0, st>>>((const float*)ps->data, ps->ld, N, d_ex, n_ex, d_disc, strike, is_call, d_part);
        else branch_lower_kernel<double><<<grid, PR_NT, 0, st>>>((const double*)ps->data, ps->ld, N, d_ex, n_ex, d_disc, strike, is_call, d_part);
    }));
    // upper bound: descending sweep from the last index any loop can touch down to the first exercise date
    if (n_ex > 0) {
        const int j_hi = kend - 1 > exercise_times[n_ex - 1] ? kend - 1 : exercise_times[n_ex - 1];
        double *Fo = F0, *Fn = F1;
        for (int j = j_hi; j >= exercise_times[0]; --j) {
            const int e = is_ex[j];
            if (e && d_inj) {
                MCP_CUDA(ctx, mcp_memcpy_async(ctx, d_inj, injected_rp + (size_t)(e - 1) * N * num_branches, (size_t)N * num_branches * 4, cudaMemcpyHostToDevice, st));
            }
            const int j_valid = j < kend ? 1 : 0;  // index j enters the future maxima of earlier dates only if t_j <= maturity
            const int has_cont = j < ex_back ? 1 : 0;
            if (f32)
                branch_upper_kernel<float><<<grid, PR_NT, 0, st>>>((const float*)ps->data + (int64_t)j * ps->ld, N, j, j_valid, e ? 1 : 0, has_cont, disc[j], strike,
                                                                  is_call, num_branches, Fo, Fn, d_best, keys, path_offset, d_inj);
            else
                branch_upper_kernel<double><<<grid, PR_NT, 0, st>>>((const double*)ps->data + (int64_t)j * ps->ld, N, j, j_valid, e ? 1 : 0, has_cont, disc[j], strike,
                                                                   is_call, num_branches, Fo, Fn, d_best, keys, path_offset, d_inj);
            MCP_LAUNCH_CHECK(ctx);
            if (e && d_inj) MCP_CUDA(ctx, cudaStreamSynchronize(st));  // the staging table is reused by the next date
            double* t = Fo; Fo = Fn; Fn = t;
        }
    }
    MCP_TRY(fold_sum(ctx, d_part, grid, 1, d_fin + 1, [&] { sum_vector_kernel<<<grid, PR_NT, 0, st>>>(d_best, N, d_part); }));
    }
    MCP_TRY(mcp_allreduce_f64(ctx, d_fin, 3));
    double h[3];
    double* hp = (double*)mcp_stage_alloc(ctx, 24);
    MCP_CUDA(ctx, mcp_memcpy_async(ctx, hp ? hp : h, d_fin, 24, cudaMemcpyDeviceToHost, st));
    MCP_CUDA(ctx, cudaStreamSynchronize(st));
    if (hp) memcpy(h, hp, 24);
    const double lower = h[0] / h[2], upper = h[1] / h[2];
    if (lower_out) *lower_out = lower;
    if (upper_out) *upper_out = upper;
    *price = 0.5 * (lower + upper);  // :38
    return MCP_OK;
}
