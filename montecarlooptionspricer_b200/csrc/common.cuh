// common.cuh -- shared host-side plumbing for libmcp_b200.so (engine handle, pathset, error handling).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mcp_b200.h"

struct McpNccl;  // dlopen'ed NCCL entry points (ctx.cu)

// Mailbox layout (legacy per-launch exchange): [parity 0|1][source rank][MCP_XROW 8-byte words]; word 2k / 2k+1 = {low / high
// 32 bits of value k, 32-bit sequence tag}: payload and tag travel in ONE 8-byte store (atomic over NVLink), so no fence
// and no separate flag.
constexpr int MCP_XROW = 64;
constexpr int MCP_XMAX_RANKS = 16;
struct McpXchg {
    int nranks = 1, rank = 0;
    int enabled = 0;
    double* const* peer = nullptr;  // device array [nranks]: mailbox base of every rank (own entry = local mailbox)
    int* err = nullptr;             // device flag: set when a wait timed out
};

// Persistent-sweep exchange (lsm_persist.cuh).  Same tagged-word format.  Two regions:
//   * the PEER region, part of the IPC-shared mailbox block behind the legacy rows: [parity][source rank][MCP_PX_BIGW words],
//     written by the reducer CTA of every rank into every rank (one NVLink hop), read locally;
//   * the LOCAL region (plain cudaMalloc, zeroed once): worker rows [parity][MCP_PX_MAXW][MCP_PX_ROWW] (worker CTA -> reducer CTA)
//     and broadcast slots [parity][MCP_PX_MAXW][MCP_PX_BCW] (reducer CTA -> worker CTAs; one slot per worker, so that
//     a hundred CTAs never spin on one L2 line).
constexpr int MCP_PX_BIGW = 4096;
constexpr int MCP_PX_MAXW = 192;
constexpr int MCP_PX_ROWW = 48;
constexpr int MCP_PX_BCW = 16;
constexpr size_t MCP_XLEGACY_WORDS = (size_t)2 * MCP_XMAX_RANKS * MCP_XROW;
constexpr size_t MCP_XBOX_WORDS = MCP_XLEGACY_WORDS + (size_t)2 * MCP_XMAX_RANKS * MCP_PX_BIGW;  // whole IPC-shared block
constexpr size_t MCP_PX_LOCAL_WORDS = (size_t)2 * MCP_PX_MAXW * MCP_PX_ROWW + (size_t)2 * MCP_PX_MAXW * MCP_PX_BCW;
struct McpPx {
    unsigned long long* const* peer = nullptr;  // device array [nranks]: base of every rank's IPC block (own entry = local block)
    unsigned long long* local = nullptr;        // local region
    int* err = nullptr;
    int nranks = 1, rank = 0;
};

// One engine handle per host thread / per GPU rank.
struct mcp_ctx {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;  // == own_stream unless mcp_set_stream installed a caller stream
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    uint64_t launches = 0;

    // multi-GPU
    void* comm = nullptr;  // ncclComm_t
    int rank = 0, nranks = 1;
    // peer-memory exchange (NVLink / NVSwitch P2P): every rank owns a small mailbox that all peers can store into,
    // so the per-step moment all-reduce runs INSIDE the sweep kernel instead of as a separate NCCL launch
    McpXchg xchg;
    void* xchg_local = nullptr;            // this rank's mailbox (cudaMalloc, exported through CUDA IPC)
    void* xchg_peer_ptrs_dev = nullptr;    // device copy of xchg.peer[]
    std::vector<void*> xchg_opened;        // peers' mailboxes opened with cudaIpcOpenMemHandle
    unsigned long long xchg_seq = 0;       // exchange counter (all ranks advance in lock step)
    void* px_local = nullptr;              // local region of the persistent-sweep exchange (see McpPx)
    void* px_trace = nullptr;              // MCP_PX_TRACE=1: time stamps of the last persistent sweep
    size_t px_trace_bytes = 0;
    int px_trace_rows = 0, px_trace_cols = 0;

    // grow-only device scratch (regression partials, coefficient tables, transposition staging ...)
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // pinned parameter ring (1 MiB): small host->device tables and device->host results go through it, so no copy of
    // the per-row calls touches pageable memory (pageable copies serialise all host threads inside the driver)
    unsigned char* stage = nullptr;
    size_t stage_off = 0;
    // grow-only pinned host staging
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    // grow-only LSM carry buffer (kept across calls so repeated pricing does not re-allocate)
    void* carry = nullptr;
    size_t carry_bytes = 0;
    // small-slab pool: device blocks of destroyed pathsets (<= 256 MiB each) are kept for the next mcp_pathset_create on
    // this ctx -- the reference's row loop creates and drops one path matrix per row (PredictionGen.cpp:736-737), and
    // cudaMalloc / cudaFree serialise every host thread of the process
    std::vector<std::pair<void*, size_t>> slab_pool;
    size_t slab_pool_bytes = 0;
    // cached pathset for mcp_price_rbergomi_lsm
    mcp_pathset* cached_ps = nullptr;
    // cached slab of mcp_price_surface_rbergomi_lsm (a 4 GB cudaMalloc / cudaFree pair per call costs more than the generation)
    mcp_pathset* cached_surface_ps = nullptr;
    // every pathset created on this ctx and not destroyed yet: mcp_destroy releases what the caller forgot
    std::vector<mcp_pathset*> live_ps;
    // optional per-kernel timing
    bool profiling = false;
    std::vector<cudaEvent_t> prof_ev;  // grow-only pool
    mcp_profile prof = {0.f, 0.f, 0, 0.f, 0};
    // host<->device bytes moved by this ctx (every copy the library issues is counted where it is issued)
    uint64_t h2d_bytes = 0, d2h_bytes = 0;
};

struct mcp_pathset {
    mcp_ctx* ctx = nullptr;
    int64_t n_paths = 0;
    int n_steps = 0;   // slab has n_steps+1 rows
    int64_t ld = 0;    // row stride in elements (n_paths rounded up to 128)
    int dtype = MCP_F32;
    void* data = nullptr;
    size_t bytes = 0;
    size_t capacity = 0;  // size of the device block behind `data` (>= bytes when it came from the pool)
};

int mcp_fail(mcp_ctx* ctx, int code, const char* fmt, ...);
int mcp_scratch_reserve(mcp_ctx* ctx, size_t bytes);
int mcp_pinned_reserve(mcp_ctx* ctx, size_t bytes);
int mcp_carry_reserve(mcp_ctx* ctx, size_t bytes);
// pinned scratch of `bytes` from the ctx ring (valid until ~1 MiB more has been handed out); nullptr when too large
void* mcp_stage_alloc(mcp_ctx* ctx, size_t bytes);
// host -> device through the pinned ring (falls back to a direct copy for large sources)
int mcp_h2d(mcp_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
// once per (device, kernel, dynamic smem): raise the kernel's dynamic shared-memory limit and query its occupancy
int mcp_kernel_config(mcp_ctx* ctx, const void* kernel, int block, size_t smem, int* occ_out);
// all-reduce (sum, fp64) of a device buffer on ctx->stream; no-op without a communicator
int mcp_allreduce_f64(mcp_ctx* ctx, double* dev, int count);
// exchange plumbing of the persistent sweep: allocates the local region (and, without a communicator, a one-rank mailbox block)
int mcp_px_get(mcp_ctx* ctx, McpPx* out);
// event `i` of the profiling pool (created on demand)
cudaEvent_t mcp_prof_event(mcp_ctx* ctx, size_t i);

#define MCP_CUDA(ctx, call)                                                                            \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return mcp_fail((ctx), MCP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                            __FILE__, __LINE__);                                                       \
    } while (0)

#define MCP_LAUNCH_CHECK(ctx)                                                                          \
    do {                                                                                               \
        (ctx)->launches++;                                                                             \
        cudaError_t e__ = cudaGetLastError();                                                          \
        if (e__ != cudaSuccess)                                                                        \
            return mcp_fail((ctx), MCP_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                   \
                            cudaGetErrorString(e__), __FILE__, __LINE__);                              \
    } while (0)

#define MCP_TRY(expr)                  \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != MCP_OK) return rc__; \
    } while (0)

// every host<->device copy of the product goes through here (or counts itself): bench.py reports the bytes it COUNTED
static inline cudaError_t mcp_memcpy_async(mcp_ctx* ctx, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st) {
    if (kind == cudaMemcpyHostToDevice) ctx->h2d_bytes += bytes;
    else if (kind == cudaMemcpyDeviceToHost) ctx->d2h_bytes += bytes;
    return cudaMemcpyAsync(dst, src, bytes, kind, st);
}

static inline int64_t mcp_round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// batched row generation (gen_rbergomi.cu), used by rows.cu
int mcp_rows_generate(mcp_ctx* ctx, const mcp_rbergomi_params* models, size_t model_stride_bytes, const int* n_steps, int n_rows, int n_paths,
                      uint64_t seed, uint64_t path_offset, float* slabs, int64_t slab_stride, int64_t ld, void* host_stage = nullptr,
                      void* dev_stage = nullptr, size_t stage_bytes = 0, cudaEvent_t ev_start = nullptr);
size_t mcp_rows_stage_bytes(int n_rows, int max_steps);
