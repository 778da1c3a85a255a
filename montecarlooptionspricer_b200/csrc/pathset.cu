// pathset.cu -- the device-resident path slab and its host<->device movers.
//
// Layout in HBM: TIME-MAJOR  S[j][i], j = 0..n_steps (row), i = path (contiguous), row stride ld = n_paths
// rounded up to 128 elements, so every row starts 512 B aligned and a warp touching 32 consecutive paths of
// one time index moves exactly one 128 B line.  The reference holds the transposed, path-major
// std::vector<std::vector<double>> (RoughVolatility.cpp:344); LSM walks it column-wise (LSMPricer.cpp:51-94),
// one cache line per element.  The movers below transpose through a padded shared-memory tile so both the
// host-layout side and the slab side are accessed coalesced.
#include <string.h>

#include "common.cuh"
#include "transpose.cuh"

extern "C" {

int mcp_pathset_create(mcp_ctx* ctx, int64_t n_paths, int n_steps, int dtype, mcp_pathset** out) {
    if (!ctx || !out) return MCP_ERR_INVALID;
    *out = nullptr;
    if (n_paths <= 0 || n_steps < 0) return mcp_fail(ctx, MCP_ERR_EMPTY_PATHS, "pathset: empty (n_paths=%lld, n_steps=%d)", (long long)n_paths, n_steps);
    if (dtype != MCP_F32 && dtype != MCP_F64) return mcp_fail(ctx, MCP_ERR_INVALID, "pathset: bad dtype %d", dtype);
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));
    mcp_pathset* ps = new mcp_pathset();
    ps->ctx = ctx;
    ps->n_paths = n_paths;
    ps->n_steps = n_steps;
    ps->ld = mcp_round_up(n_paths, 128);
    ps->dtype = dtype;
    ps->bytes = (size_t)ps->ld * (size_t)(n_steps + 1) * (dtype == MCP_F32 ? 4 : 8);
    // best fit from the ctx pool (stream order makes reuse safe: every user of the old slab ran on ctx->stream)
    int best = -1;
    for (int i = 0; i < (int)ctx->slab_pool.size(); ++i)
        if (ctx->slab_pool[i].second >= ps->bytes && (best < 0 || ctx->slab_pool[i].second < ctx->slab_pool[best].second)) best = i;
    if (best >= 0 && ctx->slab_pool[best].second <= 4 * ps->bytes + (1u << 20)) {
        ps->data = ctx->slab_pool[best].first;
        ps->capacity = ctx->slab_pool[best].second;
        ctx->slab_pool_bytes -= ps->capacity;
        ctx->slab_pool.erase(ctx->slab_pool.begin() + best);
    } else {
        ps->capacity = ps->bytes;
        if (cudaMalloc(&ps->data, ps->bytes) != cudaSuccess) {
            cudaGetLastError();
            delete ps;
            return mcp_fail(ctx, MCP_ERR_NOMEM, "pathset: cudaMalloc of %zu bytes failed", (size_t)ps->bytes);
        }
    }
    // generators and uploads write live paths only: the pad columns [n_paths, ld) of every row start as zeros, so the
    // vector / bulk-copy readers that touch them never see recycled NaN or Inf bit patterns
    if (ps->ld > n_paths) {
        const size_t esz = dtype == MCP_F32 ? 4 : 8;
        cudaError_t e = cudaMemset2DAsync((char*)ps->data + (size_t)n_paths * esz, (size_t)ps->ld * esz, 0, (size_t)(ps->ld - n_paths) * esz,
                                          (size_t)(n_steps + 1), ctx->stream);
        if (e != cudaSuccess) {
            mcp_pathset_destroy(ps);
            return mcp_fail(ctx, MCP_ERR_CUDA, "pathset: clearing the pad columns failed: %s", cudaGetErrorString(e));
        }
    }
    ctx->live_ps.push_back(ps);
    *out = ps;
    return MCP_OK;
}

int mcp_pathset_destroy(mcp_pathset* ps) {
    if (!ps) return MCP_OK;
    cudaSetDevice(ps->ctx->device);
    mcp_ctx* ctx = ps->ctx;
    for (size_t i = 0; i < ctx->live_ps.size(); ++i)
        if (ctx->live_ps[i] == ps) { ctx->live_ps.erase(ctx->live_ps.begin() + (long)i); break; }
    if (ps->capacity > ((size_t)256 << 20)) cudaStreamSynchronize(ctx->stream);  // pooled blocks are reused in stream order instead
    if (ctx->cached_ps == ps) ctx->cached_ps = nullptr;
    if (ctx->cached_surface_ps == ps) ctx->cached_surface_ps = nullptr;
    constexpr size_t POOL_BLOCK_MAX = (size_t)256 << 20, POOL_TOTAL_MAX = (size_t)1 << 30;
    if (ps->data && ps->capacity <= POOL_BLOCK_MAX && ctx->slab_pool_bytes + ps->capacity <= POOL_TOTAL_MAX && ctx->slab_pool.size() < 16) {
        ctx->slab_pool.push_back(std::make_pair(ps->data, ps->capacity));
        ctx->slab_pool_bytes += ps->capacity;
    } else if (ps->data) {
        cudaFree(ps->data);
    }
    delete ps;
    return MCP_OK;
}

int mcp_pathset_info(const mcp_pathset* ps, int64_t* n_paths, int* n_steps, int64_t* ld, int* dtype, void** dptr) {
    if (!ps) return MCP_ERR_INVALID;
    if (n_paths) *n_paths = ps->n_paths;
    if (n_steps) *n_steps = ps->n_steps;
    if (ld) *ld = ps->ld;
    if (dtype) *dtype = ps->dtype;
    if (dptr) *dptr = ps->data;
    return MCP_OK;
}

}  // extern "C"

// Upload a chunk that already sits in device staging as [pc][cols] doubles.
static int scatter_chunk(mcp_pathset* ps, const double* stage, int64_t p0, int64_t pc) {
    mcp_ctx* ctx = ps->ctx;
    const int cols = ps->n_steps + 1;
    if (ps->dtype == MCP_F32)
        mcp_launch_transpose<double, float>(ctx->stream, stage, cols, pc, cols, (float*)ps->data + p0, ps->ld);
    else
        mcp_launch_transpose<double, double>(ctx->stream, stage, cols, pc, cols, (double*)ps->data + p0, ps->ld);
    MCP_LAUNCH_CHECK(ctx);
    return MCP_OK;
}

static int64_t chunk_paths(const mcp_pathset* ps) {
    const int cols = ps->n_steps + 1;
    int64_t pc = (int64_t)(64u << 20) / ((int64_t)cols * 8);  // ~64 MiB of doubles per chunk
    if (pc < 32) pc = 32;
    pc = pc / 32 * 32;
    return pc < ps->n_paths ? pc : ps->n_paths;
}

extern "C" {

int mcp_pathset_upload_f64(mcp_pathset* ps, const double* host, int64_t ld_host) {
    if (!ps || !host) return MCP_ERR_INVALID;
    mcp_ctx* ctx = ps->ctx;
    const int cols = ps->n_steps + 1;
    if (ld_host < cols) return mcp_fail(ctx, MCP_ERR_INVALID, "upload: ld_host %lld < %d columns", (long long)ld_host, cols);
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t pc = chunk_paths(ps);
    MCP_TRY(mcp_scratch_reserve(ctx, (size_t)pc * cols * 8));
    for (int64_t p0 = 0; p0 < ps->n_paths; p0 += pc) {
        const int64_t n = (ps->n_paths - p0 < pc) ? ps->n_paths - p0 : pc;
        ctx->h2d_bytes += (size_t)cols * 8 * (size_t)n;
        MCP_CUDA(ctx, cudaMemcpy2DAsync(ctx->scratch, (size_t)cols * 8, host + p0 * ld_host, (size_t)ld_host * 8,
                                        (size_t)cols * 8, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        MCP_TRY(scatter_chunk(ps, (const double*)ctx->scratch, p0, n));
        MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // staging is reused by the next chunk
    }
    return MCP_OK;
}

int mcp_pathset_upload_rows_f64(mcp_pathset* ps, const double* const* rows) {
    if (!ps || !rows) return MCP_ERR_INVALID;
    mcp_ctx* ctx = ps->ctx;
    const int cols = ps->n_steps + 1;
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t pc = chunk_paths(ps);
    MCP_TRY(mcp_scratch_reserve(ctx, (size_t)pc * cols * 8));
    MCP_TRY(mcp_pinned_reserve(ctx, (size_t)pc * cols * 8));
    for (int64_t p0 = 0; p0 < ps->n_paths; p0 += pc) {
        const int64_t n = (ps->n_paths - p0 < pc) ? ps->n_paths - p0 : pc;
        double* pin = (double*)ctx->pinned;
        for (int64_t i = 0; i < n; ++i) memcpy(pin + i * cols, rows[p0 + i], (size_t)cols * 8);
        MCP_CUDA(ctx, mcp_memcpy_async(ctx, ctx->scratch, pin, (size_t)n * cols * 8, cudaMemcpyHostToDevice, ctx->stream));
        MCP_TRY(scatter_chunk(ps, (const double*)ctx->scratch, p0, n));
        MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MCP_OK;
}

int mcp_pathset_download_f64(const mcp_pathset* ps, double* host, int64_t ld_host) {
    if (!ps || !host) return MCP_ERR_INVALID;
    mcp_ctx* ctx = ps->ctx;
    const int cols = ps->n_steps + 1;
    if (ld_host < cols) return mcp_fail(ctx, MCP_ERR_INVALID, "download: ld_host %lld < %d columns", (long long)ld_host, cols);
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t pc = chunk_paths(ps);
    MCP_TRY(mcp_scratch_reserve(ctx, (size_t)pc * cols * 8));
    for (int64_t p0 = 0; p0 < ps->n_paths; p0 += pc) {
        const int64_t n = (ps->n_paths - p0 < pc) ? ps->n_paths - p0 : pc;
        double* stage = (double*)ctx->scratch;
        if (ps->dtype == MCP_F32)
            mcp_launch_transpose<float, double>(ctx->stream, (const float*)ps->data + p0, ps->ld, cols, n, stage, cols);
        else
            mcp_launch_transpose<double, double>(ctx->stream, (const double*)ps->data + p0, ps->ld, cols, n, stage, cols);
        MCP_LAUNCH_CHECK(ctx);
        ctx->d2h_bytes += (size_t)cols * 8 * (size_t)n;
        MCP_CUDA(ctx, cudaMemcpy2DAsync(host + p0 * ld_host, (size_t)ld_host * 8, stage, (size_t)cols * 8, (size_t)cols * 8,
                                        (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MCP_OK;
}

int mcp_pathset_download_rows_f64(const mcp_pathset* ps, double* const* rows) {
    if (!ps || !rows) return MCP_ERR_INVALID;
    mcp_ctx* ctx = ps->ctx;
    const int cols = ps->n_steps + 1;
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t pc = chunk_paths(ps);
    MCP_TRY(mcp_scratch_reserve(ctx, (size_t)pc * cols * 8));
    MCP_TRY(mcp_pinned_reserve(ctx, (size_t)pc * cols * 8));
    for (int64_t p0 = 0; p0 < ps->n_paths; p0 += pc) {
        const int64_t n = (ps->n_paths - p0 < pc) ? ps->n_paths - p0 : pc;
        double* stage = (double*)ctx->scratch;
        if (ps->dtype == MCP_F32)
            mcp_launch_transpose<float, double>(ctx->stream, (const float*)ps->data + p0, ps->ld, cols, n, stage, cols);
        else
            mcp_launch_transpose<double, double>(ctx->stream, (const double*)ps->data + p0, ps->ld, cols, n, stage, cols);
        MCP_LAUNCH_CHECK(ctx);
        double* pin = (double*)ctx->pinned;
        MCP_CUDA(ctx, mcp_memcpy_async(ctx, pin, stage, (size_t)n * cols * 8, cudaMemcpyDeviceToHost, ctx->stream));
        MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int64_t i = 0; i < n; ++i) memcpy(rows[p0 + i], pin + i * cols, (size_t)cols * 8);
    }
    return MCP_OK;
}

int mcp_pathset_download_timemajor_f32(const mcp_pathset* ps, float* host, int64_t ld_host) {
    if (!ps || !host) return MCP_ERR_INVALID;
    mcp_ctx* ctx = ps->ctx;
    if (ps->dtype != MCP_F32) return mcp_fail(ctx, MCP_ERR_INVALID, "download_timemajor_f32: slab is not fp32");
    if (ld_host < ps->n_paths) return mcp_fail(ctx, MCP_ERR_INVALID, "download: ld_host too small");
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->d2h_bytes += (size_t)ps->n_paths * 4 * (size_t)(ps->n_steps + 1);
    MCP_CUDA(ctx, cudaMemcpy2DAsync(host, (size_t)ld_host * 4, ps->data, (size_t)ps->ld * 4, (size_t)ps->n_paths * 4,
                                    (size_t)(ps->n_steps + 1), cudaMemcpyDeviceToHost, ctx->stream));
    MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MCP_OK;
}

}  // extern "C"
