// dual.cu -- nested-simulation duality (Andersen-Broadie) under GBM: BASELINE config 4 [new].
//
// There is NO reference algorithm for this: the reference's "MartingaleOptimization" pricer is a polynomial fit to
// heuristic targets (src/models/MartingaleOptimizationPricer.cpp:122-178), and its rough-volatility driver
// X = Re DFT(phi (.) Z) (RoughVolatility.cpp:264-292) is not adapted -- every X_k depends on all Z -- so "inner paths
// conditional on the outer path up to t_j" has no meaning for that model.  The algorithm is therefore defined here
// (see include/mcp_b200.h) for the GBM model of config 1 and checked against an independent restatement in the tests.
//
// Kernels: the policy comes from the LSM sweep (lsm.cu) on an independent path set; `dual_nested_kernel` is the hot
// one: work item (date j, outer path i) -- consecutive threads share j, so a warp's inner paths have the same horizon --
// simulates n_inner GBM paths from S_j[i] under the policy until it exercises (one Philox call per four inner steps,
// Box-Muller on the SFU, fp32) and writes the time-0 discounted continuation value Q[j][i]; `dual_combine_kernel`
// walks every outer path once: policy value, martingale, max_j (h_j - M_j), and the fp64 sums.
#include <math.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "lsm_solve.cuh"
#include "philox.cuh"

namespace {

constexpr int DU_NT = 256;
constexpr int DU_LD = 24;
constexpr int DU_MAXSTEPS = 1024;

struct DualStep {  // policy at one date, evaluated in fp32: exercise iff payoff > 1e-14 and !(payoff < cont(x))
    float c[MAXP + 1];
    float mu, inv_s, disc;  // disc = e^{-r j dt}
};

struct DualArgs {
    const float* S;  // outer slab [n+1][ld]
    int64_t ld, n_outer;
    int n, p, K, is_call;
    float strike, drift2, vol2;  // GBM step: s *= 2^(drift2 + vol2 z)
    uint64_t path_offset;
    const DualStep* steps;  // [n+1]
    float* Q;               // [n][ld]
};

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float payoff32(int is_call, float s, float K) { return fmaxf(is_call ? s - K : K - s, 0.f); }

__device__ __forceinline__ bool policy_exercises(const DualStep& st, int p, float s, float pay) {
    if (!(pay > 1e-14f)) return false;
    const float x = (s - st.mu) * st.inv_s;
    float cont = st.c[p];
    for (int k = p - 1; k >= 0; --k) cont = fmaf(cont, x, st.c[k]);
    return !(pay < cont);
}

__global__ void __launch_bounds__(DU_NT) dual_nested_kernel(DualArgs a, PhiloxKeys keys) {
    extern __shared__ DualStep sst[];  // [n+1]
    for (int j = threadIdx.x; j <= a.n; j += DU_NT) sst[j] = a.steps[j];
    __syncthreads();
    const int64_t total = (int64_t)a.n * a.n_outer;
    for (int64_t t = (int64_t)blockIdx.x * DU_NT + threadIdx.x; t < total; t += (int64_t)gridDim.x * DU_NT) {
        const int j = (int)(t / a.n_outer);
        const int64_t i = t - (int64_t)j * a.n_outer;
        const float s0 = a.S[(int64_t)j * a.ld + i];
        const uint64_t gid = a.path_offset + (uint64_t)i;
        float sum = 0.f;
        for (int k = 0; k < a.K; ++k) {
            float s = s0, val = 0.f;
            const uint32_t c2 = (uint32_t)j * (uint32_t)a.K + (uint32_t)k;
            for (int m = j + 1; m <= a.n; m += 4) {
                const uint4 u = philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), c2, 0x30000u + (uint32_t)((m - j - 1) >> 2), keys);
                float z[4];
                box_muller(u.x, u.y, z[0], z[1]);
                box_muller(u.z, u.w, z[2], z[3]);
                bool done = false;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int mm = m + q;
                    if (!done && mm <= a.n) {
                        s *= ex2f(fmaf(a.vol2, z[q], a.drift2));
                        const float pay = payoff32(a.is_call, s, a.strike);
                        const bool ex = mm == a.n ? (pay > 0.f) : policy_exercises(sst[mm], a.p, s, pay);
                        if (ex || mm == a.n) { val = ex ? pay * sst[mm].disc : 0.f; done = true; }
                    }
                }
                if (done) break;
            }
            sum += val;
        }
        a.Q[(int64_t)j * a.ld + i] = sum / (float)a.K;
    }
}

// per outer path: lower-bound payoff (first policy exercise), martingale and dual maximum; fp64 sums per CTA
__global__ void __launch_bounds__(DU_NT) dual_combine_kernel(DualArgs a, double* __restrict__ partial) {
    extern __shared__ DualStep sst[];
    for (int j = threadIdx.x; j <= a.n; j += DU_NT) sst[j] = a.steps[j];
    __syncthreads();
    double acc[4] = {0.0, 0.0, 0.0, 0.0};  // sum lower, sum lower^2, sum dual, sum dual^2
    for (int64_t i = (int64_t)blockIdx.x * DU_NT + threadIdx.x; i < a.n_outer; i += (int64_t)gridDim.x * DU_NT) {
        double M = 0.0, best = -1e300, lower = 0.0, q_prev = 0.0;
        bool stopped = false;
        for (int j = 0; j <= a.n; ++j) {
            const float s = a.S[(int64_t)j * a.ld + i];
            const float pay = payoff32(a.is_call, s, a.strike);
            const double h = (double)pay * (double)sst[j].disc;
            const bool ex = j == a.n ? (pay > 0.f) : policy_exercises(sst[j], a.p, s, pay);
            const double q = j < a.n ? (double)a.Q[(int64_t)j * a.ld + i] : 0.0;
            const double L = (ex || j == a.n) ? (ex ? h : 0.0) : q;   // value of the policy at date j
            if (j > 0) M += L - q_prev;                                 // M_j = M_{j-1} + L_j - Q_{j-1}
            if (h - M > best) best = h - M;
            if (!stopped && (ex || j == a.n)) { lower = ex ? h : 0.0; stopped = true; }
            q_prev = q;
        }
        acc[0] += lower; acc[1] = fma(lower, lower, acc[1]);
        acc[2] += best;  acc[3] = fma(best, best, acc[3]);
    }
    __shared__ double red[DU_NT / 32][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < 4; ++k) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double v = 0.0;
        for (int w = 0; w < DU_NT / 32; ++w) v += red[w][threadIdx.x];
        partial[(int64_t)blockIdx.x * DU_LD + threadIdx.x] = v;
    }
}

__global__ void __launch_bounds__(256) dual_fold_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ out) {
    __shared__ double red[8][4];
    const int k = threadIdx.x & 3, grp = (threadIdx.x >> 2) & 7;
    double s = 0.0;
    if (threadIdx.x < 32)
        for (int b = grp; b < nblocks; b += 8) s += partial[(int64_t)b * DU_LD + k];
    if (threadIdx.x < 32) red[grp][k] = s;
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int g = 0; g < 8; ++g) t += red[g][threadIdx.x];
        out[threadIdx.x] = t;
    }
}

}  // namespace

extern "C" int mcp_gbm_nested_dual(mcp_ctx* ctx, const mcp_gbm_params* model, double strike, int is_call, int n_steps, int poly_order,
                                   int64_t n_policy_paths, int64_t n_outer, int n_inner, uint64_t seed, uint64_t path_offset,
                                   mcp_dual_result* out) {
    if (!ctx || !model || !out) return MCP_ERR_INVALID;
    if (n_steps < 1 || n_steps > DU_MAXSTEPS) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "dual: n_steps %d outside [1, %d]", n_steps, DU_MAXSTEPS);
    if (poly_order < 0 || poly_order > MAXP) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "dual: poly_order %d outside [0, %d]", poly_order, MAXP);
    if (n_policy_paths < 1 || n_outer < 1 || n_inner < 1) return mcp_fail(ctx, MCP_ERR_INVALID, "dual: path counts must be positive");
    if ((int64_t)n_steps * n_inner >= ((int64_t)1 << 32)) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "dual: n_steps * n_inner must stay below 2^32");
    if (!(model->dt > 0.0) || !(model->sigma >= 0.0) || !(model->S0 > 0.0)) return mcp_fail(ctx, MCP_ERR_DOMAIN, "dual: need dt > 0, sigma >= 0, S0 > 0");
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));
    memset(out, 0, sizeof(*out));
    const int n = n_steps, p = poly_order;
    const double maturity = (double)n * model->dt;
    cudaEvent_t ev[4];
    for (auto& e : ev) MCP_CUDA(ctx, cudaEventCreate(&e));
    cudaStream_t st = ctx->stream;

    // ---- 1. policy: LSM regression on an independent path set (every rank fits the same policy) ----
    cudaEventRecord(ev[0], st);
    std::vector<double> coef((size_t)n * (p + 3), 0.0);
    {
        mcp_pathset* pp = nullptr;
        MCP_TRY(mcp_pathset_create(ctx, n_policy_paths, n, MCP_F32, &pp));
        int rc = mcp_gen_gbm(ctx, pp, model, seed ^ 1ull, 0, nullptr, nullptr);
        if (rc == MCP_OK) {
            mcp_lsm_params lp;
            lp.r = model->r; lp.strike = strike; lp.maturity = maturity + model->dt; lp.dt = model->dt;
            lp.is_call = is_call; lp.poly_order = p; lp.basis = MCP_BASIS_STANDARDISED; lp.carry = MCP_F64;
            mcp_lsm_result lr;
            void* comm = ctx->comm;  // the policy sample is replicated, not sharded: fit it locally
            ctx->comm = nullptr;
            rc = mcp_lsm_price(ctx, pp, &lp, &lr, coef.data(), nullptr, nullptr);
            ctx->comm = comm;
        }
        mcp_pathset_destroy(pp);
        if (rc != MCP_OK) return rc;
    }
    std::vector<DualStep> steps((size_t)n + 1);
    for (int j = 0; j <= n; ++j) {
        DualStep& d = steps[(size_t)j];
        memset(&d, 0, sizeof(d));
        d.disc = (float)exp(-model->r * (double)j * model->dt);
        if (j < n) {
            for (int k = 0; k <= p; ++k) d.c[k] = (float)coef[(size_t)j * (p + 3) + k];
            d.mu = (float)coef[(size_t)j * (p + 3) + p + 1];
            d.inv_s = (float)coef[(size_t)j * (p + 3) + p + 2];
        }
    }

    // ---- 2. outer paths ----
    cudaEventRecord(ev[1], st);
    mcp_pathset* po = nullptr;
    MCP_TRY(mcp_pathset_create(ctx, n_outer, n, MCP_F32, &po));
    int rc = mcp_gen_gbm(ctx, po, model, seed, path_offset, nullptr, nullptr);
    const int64_t ld = po->ld;
    const int grid_c = (int)((n_outer + DU_NT - 1) / DU_NT < (int64_t)ctx->sm_count * 8 ? (n_outer + DU_NT - 1) / DU_NT : (int64_t)ctx->sm_count * 8);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_steps = take(steps.size() * sizeof(DualStep)), o_part = take((size_t)grid_c * DU_LD * 8), o_fin = take(8 * 8);
    if (rc == MCP_OK) rc = mcp_scratch_reserve(ctx, off);
    if (rc == MCP_OK) rc = mcp_carry_reserve(ctx, (size_t)n * ld * sizeof(float));
    if (rc != MCP_OK) { mcp_pathset_destroy(po); return rc; }
    unsigned char* sb = (unsigned char*)ctx->scratch;
    DualArgs a;
    memset(&a, 0, sizeof(a));
    a.S = (const float*)po->data; a.ld = ld; a.n_outer = n_outer; a.n = n; a.p = p; a.K = n_inner; a.is_call = is_call;
    a.strike = (float)strike;
    const double log2e = 1.4426950408889634074;
    a.drift2 = (float)((model->r - 0.5 * model->sigma * model->sigma) * model->dt * log2e);
    a.vol2 = (float)(model->sigma * sqrt(model->dt) * log2e);
    a.path_offset = path_offset;
    a.steps = (const DualStep*)(sb + o_steps);
    a.Q = (float*)ctx->carry;
    double* d_part = (double*)(sb + o_part);
    double* d_fin = (double*)(sb + o_fin);
    rc = mcp_h2d(ctx, sb + o_steps, steps.data(), steps.size() * sizeof(DualStep));
    const size_t smem = steps.size() * sizeof(DualStep);
    if (rc == MCP_OK && smem > 48 * 1024) {
        rc = mcp_kernel_config(ctx, (const void*)dual_nested_kernel, DU_NT, smem, nullptr);
        if (rc == MCP_OK) rc = mcp_kernel_config(ctx, (const void*)dual_combine_kernel, DU_NT, smem, nullptr);
    }
    if (rc != MCP_OK) { mcp_pathset_destroy(po); return rc; }

    // ---- 3. nested simulation + combination ----
    cudaEventRecord(ev[2], st);
    const PhiloxKeys keys = philox_make_keys(seed ^ 0x9E3779B97F4A7C15ull);  // inner-path stream
    const int64_t items = (int64_t)n * n_outer;
    const int grid_n = (int)((items + DU_NT - 1) / DU_NT < (int64_t)ctx->sm_count * 16 ? (items + DU_NT - 1) / DU_NT : (int64_t)ctx->sm_count * 16);
    dual_nested_kernel<<<grid_n, DU_NT, smem, st>>>(a, keys);
    ctx->launches++;
    dual_combine_kernel<<<grid_c, DU_NT, smem, st>>>(a, d_part);
    ctx->launches++;
    dual_fold_kernel<<<1, 256, 0, st>>>(d_part, grid_c, d_fin);
    ctx->launches++;
    const double nloc = (double)n_outer;
    rc = mcp_h2d(ctx, d_fin + 4, &nloc, 8);
    if (rc == MCP_OK) rc = mcp_allreduce_f64(ctx, d_fin, 5);
    cudaEventRecord(ev[3], st);
    double h[5] = {0, 0, 0, 0, 0};
    double* hp = (double*)mcp_stage_alloc(ctx, 40);
    if (rc == MCP_OK && mcp_memcpy_async(ctx, hp ? hp : h, d_fin, 40, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = mcp_fail(ctx, MCP_ERR_CUDA, "dual: D2H failed");
    if (rc == MCP_OK && (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess))
        rc = mcp_fail(ctx, MCP_ERR_CUDA, "dual: kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
    mcp_pathset_destroy(po);
    if (rc != MCP_OK) return rc;
    if (hp) memcpy(h, hp, 40);
    const double N = h[4];
    out->n_outer_global = (int64_t)llround(N);
    out->lower = h[0] / N;
    out->upper = h[2] / N;
    const double vl = N > 1 ? (h[1] - N * out->lower * out->lower) / (N - 1) : 0.0, vu = N > 1 ? (h[3] - N * out->upper * out->upper) / (N - 1) : 0.0;
    out->lower_se = vl > 0 ? sqrt(vl / N) : 0.0;
    out->upper_se = vu > 0 ? sqrt(vu / N) : 0.0;
    cudaEventElapsedTime(&out->policy_ms, ev[0], ev[1]);
    cudaEventElapsedTime(&out->outer_ms, ev[1], ev[2]);
    cudaEventElapsedTime(&out->nested_ms, ev[2], ev[3]);
    for (auto& e : ev) cudaEventDestroy(e);
    return MCP_OK;
}
