// gen_gbm.cu -- geometric-Brownian-motion price paths [new: BASELINE config 1; the reference has no GBM
// generator].  It is the constant-variance special case of the reference recursion
// (src/models/RoughVolatility.cpp:354-364) with v = sigma^2 and one driving normal per step:
//     S_j = S_{j-1} exp((r - sigma^2/2) dt + sigma sqrt(dt) z_j).
// Thread <-> path (consecutive lanes = consecutive paths => every store is a full 128 B line of the
// time-major slab); one Philox4x32-10 call feeds four steps; log-space running sum, S = S0 exp2(.).
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "philox.cuh"
#include "transpose.cuh"

namespace {

struct GbmParams {
    float S0, drift2, vol2;  // (r - sigma^2/2) dt log2e,  sigma sqrt(dt) log2e
    int n;
    int64_t n_paths, ld, ld_draws;
    uint64_t path_offset;
};

__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool INJECT, bool DUMP>
__global__ void __launch_bounds__(256) gbm_paths_kernel(GbmParams P, PhiloxKeys K, const float* __restrict__ draws_in,
                                                       float* __restrict__ draws_out, float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t path = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; path < P.n_paths; path += stride) {
        const uint64_t gid = P.path_offset + (uint64_t)path;
        const uint32_t c0 = (uint32_t)gid, c1 = (uint32_t)(gid >> 32);
        float cum = 0.f;
        out[path] = P.S0;
        for (int q = 0; 4 * q < P.n; ++q) {
            float z[4];
            if (INJECT) {
#pragma unroll
                for (int i = 0; i < 4; ++i) z[i] = (4 * q + i < P.n) ? draws_in[(int64_t)(4 * q + i) * P.ld_draws + path] : 0.f;
            } else {
                const uint4 x = philox4x32_10(c0, c1, (uint32_t)q, 1u, K);
                box_muller(x.x, x.y, z[0], z[1]);
                box_muller(x.z, x.w, z[2], z[3]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = 4 * q + i;
                if (j < P.n) {
                    if (DUMP) draws_out[(int64_t)j * P.ld_draws + path] = z[i];
                    cum += fmaf(P.vol2, z[i], P.drift2);
                    out[(int64_t)(j + 1) * P.ld + path] = P.S0 * fast_ex2(cum);
                }
            }
        }
    }
}

}  // namespace

extern "C" int mcp_gen_gbm(mcp_ctx* ctx, mcp_pathset* ps, const mcp_gbm_params* prm, uint64_t seed, uint64_t path_offset,
                           const float* injected, float* dump) {
    if (!ctx || !ps || !prm) return MCP_ERR_INVALID;
    if (ps->ctx != ctx) return mcp_fail(ctx, MCP_ERR_INVALID, "gbm: pathset belongs to another ctx");
    if (ps->dtype != MCP_F32) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "gbm: generators write fp32 slabs");
    const int n = ps->n_steps;
    if (n < 1) return mcp_fail(ctx, MCP_ERR_INVALID, "gbm: n_steps must be >= 1");
    if (!(prm->dt > 0.0)) return mcp_fail(ctx, MCP_ERR_DOMAIN, "gbm: need dt > 0");
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (injected && dump) {
        memcpy(dump, injected, (size_t)ps->n_paths * n * sizeof(float));
        dump = nullptr;
    }
    const double log2e = 1.4426950408889634074;
    GbmParams P;
    memset(&P, 0, sizeof(P));
    P.S0 = (float)prm->S0;
    P.drift2 = (float)((prm->r - 0.5 * prm->sigma * prm->sigma) * prm->dt * log2e);
    P.vol2 = (float)(prm->sigma * sqrt(prm->dt) * log2e);
    P.n = n;
    P.n_paths = ps->n_paths;
    P.ld = ps->ld;
    P.path_offset = path_offset;
    const PhiloxKeys K = philox_make_keys(seed);

    auto launch = [&](const GbmParams& Q, const float* din, float* dout, float* out) -> int {
        int64_t blocks = (Q.n_paths + 255) / 256;
        const int64_t cap = (int64_t)ctx->sm_count * 8;
        if (blocks > cap) blocks = cap;
        if (din) {
            if (dout) gbm_paths_kernel<true, true><<<(unsigned)blocks, 256, 0, ctx->stream>>>(Q, K, din, dout, out);
            else gbm_paths_kernel<true, false><<<(unsigned)blocks, 256, 0, ctx->stream>>>(Q, K, din, dout, out);
        } else {
            if (dout) gbm_paths_kernel<false, true><<<(unsigned)blocks, 256, 0, ctx->stream>>>(Q, K, din, dout, out);
            else gbm_paths_kernel<false, false><<<(unsigned)blocks, 256, 0, ctx->stream>>>(Q, K, din, dout, out);
        }
        MCP_LAUNCH_CHECK(ctx);
        return MCP_OK;
    };

    if (!injected && !dump) {
        if (ctx->profiling) cudaEventRecord(ctx->ev0, ctx->stream);
        MCP_TRY(launch(P, nullptr, nullptr, (float*)ps->data));
        if (ctx->profiling) {
            cudaEventRecord(ctx->ev1, ctx->stream);
            MCP_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
            MCP_CUDA(ctx, cudaEventElapsedTime(&ctx->prof.gen_kernel_ms, ctx->ev0, ctx->ev1));
        }
        return MCP_OK;
    }

    int64_t pc = (int64_t)((256u << 20) / ((size_t)n * 4 * 2)) / 32 * 32;
    if (pc < 32) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "gbm: draw staging too small for n=%d", n);
    if (ps->n_paths < pc) pc = mcp_round_up(ps->n_paths, 32);
    MCP_TRY(mcp_scratch_reserve(ctx, (size_t)2 * n * pc * 4));
    float* d_slot = (float*)ctx->scratch;       // [n][pc]
    float* d_rows = d_slot + (size_t)n * pc;    // [pc][n]
    for (int64_t p0 = 0; p0 < ps->n_paths; p0 += pc) {
        const int64_t np = (ps->n_paths - p0 < pc) ? ps->n_paths - p0 : pc;
        GbmParams Q = P;
        Q.n_paths = np;
        Q.path_offset = path_offset + (uint64_t)p0;
        Q.ld_draws = pc;
        if (injected) {
            MCP_CUDA(ctx, mcp_memcpy_async(ctx, d_rows, injected + (size_t)p0 * n, (size_t)np * n * 4, cudaMemcpyHostToDevice, ctx->stream));
            mcp_launch_transpose<float, float>(ctx->stream, d_rows, n, np, n, d_slot, pc);
            MCP_LAUNCH_CHECK(ctx);
        }
        MCP_TRY(launch(Q, injected ? d_slot : nullptr, dump ? d_slot : nullptr, (float*)ps->data + p0));
        if (dump) {
            mcp_launch_transpose<float, float>(ctx->stream, d_slot, pc, n, np, d_rows, n);
            MCP_LAUNCH_CHECK(ctx);
            MCP_CUDA(ctx, mcp_memcpy_async(ctx, dump + (size_t)p0 * n, d_rows, (size_t)np * n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        }
        MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MCP_OK;
}

// ---- raw Philox words for known-answer tests ------------------------------------------------------------
namespace {
__global__ void philox_raw_kernel(PhiloxKeys K, uint64_t first, int64_t count, uint32_t c2, uint32_t c3, uint4* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        const uint64_t g = first + (uint64_t)i;
        out[i] = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), c2, c3, K);
    }
}
}  // namespace

extern "C" int mcp_philox_raw(mcp_ctx* ctx, uint64_t seed, uint64_t first, int64_t count, uint32_t c2, uint32_t c3, uint32_t* out_host) {
    if (!ctx || !out_host || count <= 0) return MCP_ERR_INVALID;
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));
    MCP_TRY(mcp_scratch_reserve(ctx, (size_t)count * 16));
    philox_raw_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(philox_make_keys(seed), first, count, c2, c3, (uint4*)ctx->scratch);
    MCP_LAUNCH_CHECK(ctx);
    MCP_CUDA(ctx, mcp_memcpy_async(ctx, out_host, ctx->scratch, (size_t)count * 16, cudaMemcpyDeviceToHost, ctx->stream));
    MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MCP_OK;
}
