// ctx.cu -- engine handle, error reporting, workspaces, and the NCCL doorway (dlopen'ed so that the same
// libmcp_b200.so works inside a torch process -- which already maps its own libnccl.so.2 -- and from plain
// C++ hosts against the system NCCL).
#include <dlfcn.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <map>
#include <mutex>

#include "common.cuh"

static thread_local std::string g_create_err;

int mcp_fail(mcp_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_err = buf;
    return code;
}

// ---------------------------------------------------------------------------------------------- NCCL
typedef struct { char internal[128]; } mcp_nccl_uid;
struct McpNccl {
    void* lib = nullptr;
    int (*GetUniqueId)(mcp_nccl_uid*) = nullptr;
    int (*CommInitRank)(void**, int, mcp_nccl_uid, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
static McpNccl g_nccl;
static std::mutex g_nccl_mu;  // several per-thread engines may call mcp_comm_init / mcp_comm_unique_id at once

static bool nccl_load(std::string* why) {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.ok) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) { *why = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return false; }
    g_nccl.GetUniqueId = (int (*)(mcp_nccl_uid*))dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, mcp_nccl_uid, int))dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclAllReduce");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclAllGather");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce) {
        *why = "NCCL symbols missing";
        return false;
    }
    g_nccl.ok = true;
    return true;
}

int mcp_allreduce_f64(mcp_ctx* ctx, double* dev, int count) {
    if (!ctx->comm || ctx->nranks <= 1) return MCP_OK;
    const int ncclFloat64 = 8, ncclSum = 0;
    int rc = g_nccl.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclSum, ctx->comm, ctx->stream);
    if (rc != 0)
        return mcp_fail(ctx, MCP_ERR_NCCL, "ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    return MCP_OK;
}

// ------------------------------------------------------------------------- peer-memory mailboxes (P2P)
// Every rank cudaMalloc's one mailbox, the CUDA IPC handles travel through one ncclAllGather, every rank opens the
// others' mailboxes (one process per GPU; NVLink peer access is enabled lazily by cudaIpcOpenMemHandle).  Any
// failure simply leaves the NCCL all-reduce path in charge.
struct XchgCard {
    cudaIpcMemHandle_t handle;
    long long pid;
    int device;
    int ok;
};

static void xchg_teardown(mcp_ctx* ctx) {
    for (void* p : ctx->xchg_opened) cudaIpcCloseMemHandle(p);
    ctx->xchg_opened.clear();
    if (ctx->xchg_peer_ptrs_dev) cudaFree(ctx->xchg_peer_ptrs_dev);
    if (ctx->xchg.err) cudaFree(ctx->xchg.err);
    if (ctx->xchg_local) cudaFree(ctx->xchg_local);
    ctx->xchg_peer_ptrs_dev = nullptr;
    ctx->xchg_local = nullptr;
    ctx->xchg = McpXchg();
}

// Returns with ctx->xchg.enabled = 1 on EVERY rank or on none: every allocation this rank needs is made before the
// verdict all-reduce, so a local failure is part of the agreed verdict.
static void xchg_setup(mcp_ctx* ctx) {
    const char* impl = getenv("MCP_COMM_IMPL");
    if (impl && strcmp(impl, "nccl") == 0) return;
    const int n = ctx->nranks;
    if (n < 2 || n > MCP_XMAX_RANKS || !g_nccl.AllGather) return;
    if (ctx->xchg_local) xchg_teardown(ctx);  // a one-rank block from an earlier single-GPU price on this ctx
    const size_t box_bytes = MCP_XBOX_WORDS * sizeof(unsigned long long);
    XchgCard mine;
    memset(&mine, 0, sizeof(mine));
    mine.pid = (long long)getpid();
    mine.device = ctx->device;
    mine.ok = cudaMalloc(&ctx->xchg_local, box_bytes) == cudaSuccess && cudaMemset(ctx->xchg_local, 0, box_bytes) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.handle, ctx->xchg_local) == cudaSuccess &&
              cudaMalloc(&ctx->xchg_peer_ptrs_dev, sizeof(double*) * (size_t)n) == cudaSuccess &&
              cudaMalloc((void**)&ctx->xchg.err, sizeof(int)) == cudaSuccess && cudaMemset(ctx->xchg.err, 0, sizeof(int)) == cudaSuccess;
    cudaGetLastError();
    // all-gather the cards (device staging; the cards are plain bytes)
    XchgCard* d_cards = nullptr;
    std::vector<XchgCard> cards((size_t)n);
    bool ok = cudaMalloc(&d_cards, sizeof(XchgCard) * (size_t)(n + 1)) == cudaSuccess;
    if (ok) {
        cudaMemcpyAsync(d_cards + n, &mine, sizeof(XchgCard), cudaMemcpyHostToDevice, ctx->stream);
        ok = g_nccl.AllGather(d_cards + n, d_cards, sizeof(XchgCard), /*ncclInt8*/ 0, ctx->comm, ctx->stream) == 0;
        ok = ok && cudaMemcpyAsync(cards.data(), d_cards, sizeof(XchgCard) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess;
        ok = ok && cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    }
    if (d_cards) cudaFree(d_cards);
    std::vector<double*> ptrs((size_t)n, nullptr);
    for (int r = 0; ok && r < n; ++r) {
        if (!cards[(size_t)r].ok) { ok = false; break; }
        if (r == ctx->rank) { ptrs[(size_t)r] = (double*)ctx->xchg_local; continue; }
        if (cards[(size_t)r].pid == mine.pid) { ok = false; break; }  // same process: IPC handles cannot be opened
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, cards[(size_t)r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
        ctx->xchg_opened.push_back(p);
        ptrs[(size_t)r] = (double*)p;
    }
    ok = ok && cudaMemcpy(ctx->xchg_peer_ptrs_dev, ptrs.data(), sizeof(double*) * (size_t)n, cudaMemcpyHostToDevice) == cudaSuccess;
    cudaGetLastError();
    // every rank must agree, otherwise some would wait on mailboxes nobody writes: all-reduce the verdict
    double* d_flag = nullptr;
    double verdict = ok ? 0.0 : 1.0;
    if (cudaMalloc(&d_flag, sizeof(double)) == cudaSuccess) {
        cudaMemcpyAsync(d_flag, &verdict, sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
        if (g_nccl.AllReduce(d_flag, d_flag, 1, /*ncclFloat64*/ 8, /*ncclSum*/ 0, ctx->comm, ctx->stream) != 0) verdict = 1.0;
        else cudaMemcpyAsync(&verdict, d_flag, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        cudaFree(d_flag);
    } else {
        verdict = 1.0;  // the peers block in their all-reduce until the communicator is torn down: fail loudly upstream
    }
    if (verdict != 0.0) { xchg_teardown(ctx); cudaGetLastError(); return; }
    ctx->xchg.nranks = n;
    ctx->xchg.rank = ctx->rank;
    ctx->xchg.peer = (double* const*)ctx->xchg_peer_ptrs_dev;
    ctx->xchg.enabled = 1;
    ctx->xchg_seq = 0;
}

// Exchange plumbing of the persistent sweep.  With mailboxes up (multi-GPU) it rides on the IPC block; on one GPU (or on
// the NCCL fallback, where the persistent sweep is not used across ranks) a one-rank block is allocated on first use.
int mcp_px_get(mcp_ctx* ctx, McpPx* out) {
    if (!ctx->px_local) {
        const size_t bytes = MCP_PX_LOCAL_WORDS * sizeof(unsigned long long);
        if (cudaMalloc(&ctx->px_local, bytes) != cudaSuccess || cudaMemset(ctx->px_local, 0, bytes) != cudaSuccess) {
            cudaGetLastError();
            if (ctx->px_local) cudaFree(ctx->px_local);
    if (ctx->px_trace) cudaFree(ctx->px_trace);
            ctx->px_local = nullptr;
            return mcp_fail(ctx, MCP_ERR_NOMEM, "persistent sweep: local exchange region allocation failed");
        }
    }
    if (!ctx->xchg_local) {  // single rank: own block, own pointer table
        const size_t box_bytes = MCP_XBOX_WORDS * sizeof(unsigned long long);
        bool ok = cudaMalloc(&ctx->xchg_local, box_bytes) == cudaSuccess && cudaMemset(ctx->xchg_local, 0, box_bytes) == cudaSuccess &&
                  cudaMalloc(&ctx->xchg_peer_ptrs_dev, sizeof(void*)) == cudaSuccess &&
                  cudaMemcpy(ctx->xchg_peer_ptrs_dev, &ctx->xchg_local, sizeof(void*), cudaMemcpyHostToDevice) == cudaSuccess &&
                  cudaMalloc((void**)&ctx->xchg.err, sizeof(int)) == cudaSuccess && cudaMemset(ctx->xchg.err, 0, sizeof(int)) == cudaSuccess;
        if (!ok) {
            cudaGetLastError();
            xchg_teardown(ctx);
            return mcp_fail(ctx, MCP_ERR_NOMEM, "persistent sweep: mailbox allocation failed");
        }
    }
    out->peer = (unsigned long long* const*)ctx->xchg_peer_ptrs_dev;
    out->local = (unsigned long long*)ctx->px_local;
    out->err = ctx->xchg.err;
    out->nranks = ctx->xchg.enabled ? ctx->xchg.nranks : 1;
    out->rank = ctx->xchg.enabled ? ctx->xchg.rank : 0;
    return MCP_OK;
}

extern "C" {

int mcp_abi_version(void) { return MCP_B200_ABI_VERSION; }

int mcp_create(int device, mcp_ctx** out) {
    if (!out) return mcp_fail(nullptr, MCP_ERR_INVALID, "mcp_create: out is NULL");
    *out = nullptr;
    if (const char* bs = getenv("MCP_SYNC_MODE")) {  // host-thread behaviour while waiting: spin (default) | yield | block
        const unsigned f = strcmp(bs, "block") == 0 ? cudaDeviceScheduleBlockingSync : strcmp(bs, "yield") == 0 ? cudaDeviceScheduleYield : cudaDeviceScheduleSpin;
        cudaSetDeviceFlags(f);
        cudaGetLastError();
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return mcp_fail(nullptr, MCP_ERR_CUDA, "mcp_create: no CUDA device (%s); this library has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev) return mcp_fail(nullptr, MCP_ERR_INVALID, "mcp_create: device %d out of range [0,%d)", device, ndev);
    mcp_ctx* ctx = new mcp_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        delete ctx;
        return mcp_fail(nullptr, MCP_ERR_CUDA, "mcp_create: %s", cudaGetErrorString(e));
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    if (prop.major != 10) {
        delete ctx;
        return mcp_fail(nullptr, MCP_ERR_CUDA, "mcp_create: device is sm_%d%d; this library ships sm_100a code only", prop.major, prop.minor);
    }
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) {
        delete ctx;
        return mcp_fail(nullptr, MCP_ERR_CUDA, "mcp_create: %s", cudaGetErrorString(e));
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return MCP_OK;
}

int mcp_destroy(mcp_ctx* ctx) {
    if (!ctx) return MCP_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->cached_ps) mcp_pathset_destroy(ctx->cached_ps);
    if (ctx->cached_surface_ps) mcp_pathset_destroy(ctx->cached_surface_ps);
    while (!ctx->live_ps.empty()) mcp_pathset_destroy(ctx->live_ps.back());  // handles the caller never destroyed (they are dead after this)
    for (auto& blk : ctx->slab_pool) cudaFree(blk.first);
    ctx->slab_pool.clear();
    xchg_teardown(ctx);
    if (ctx->px_local) cudaFree(ctx->px_local);
    if (ctx->px_trace) cudaFree(ctx->px_trace);
    if (ctx->comm && g_nccl.ok) g_nccl.CommDestroy(ctx->comm);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->carry) cudaFree(ctx->carry);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->stage) cudaFreeHost(ctx->stage);
    for (cudaEvent_t e : ctx->prof_ev)
        if (e) cudaEventDestroy(e);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return MCP_OK;
}

const char* mcp_last_error(const mcp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int mcp_set_stream(mcp_ctx* ctx, void* s) {
    if (!ctx) return MCP_ERR_INVALID;
    cudaStreamSynchronize(ctx->stream);  // pooled slabs and workspaces are recycled in stream order
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return MCP_OK;
}

int mcp_synchronize(mcp_ctx* ctx) {
    if (!ctx) return MCP_ERR_INVALID;
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));
    MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MCP_OK;
}

int mcp_device_info(mcp_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* free_b, size_t* total_b) {
    if (!ctx) return MCP_ERR_INVALID;
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    size_t f = 0, t = 0;
    MCP_CUDA(ctx, cudaMemGetInfo(&f, &t));
    if (free_b) *free_b = f;
    if (total_b) *total_b = t;
    return MCP_OK;
}

uint64_t mcp_launch_count(const mcp_ctx* ctx) { return ctx ? ctx->launches : 0; }

int mcp_copy_counters(const mcp_ctx* ctx, uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
    if (!ctx) return MCP_ERR_INVALID;
    if (h2d_bytes) *h2d_bytes = ctx->h2d_bytes;
    if (d2h_bytes) *d2h_bytes = ctx->d2h_bytes;
    return MCP_OK;
}

int mcp_set_profiling(mcp_ctx* ctx, int on) {
    if (!ctx) return MCP_ERR_INVALID;
    ctx->profiling = on != 0;
    return MCP_OK;
}

int mcp_get_profile(const mcp_ctx* ctx, mcp_profile* out) {
    if (!ctx || !out) return MCP_ERR_INVALID;
    *out = ctx->prof;
    return MCP_OK;
}

int mcp_comm_unique_id(void* id128) {
    std::string why;
    if (!id128) return MCP_ERR_INVALID;
    if (!nccl_load(&why)) return mcp_fail(nullptr, MCP_ERR_NCCL, "%s", why.c_str());
    mcp_nccl_uid uid;
    int rc = g_nccl.GetUniqueId(&uid);
    if (rc != 0) return mcp_fail(nullptr, MCP_ERR_NCCL, "ncclGetUniqueId failed (%d)", rc);
    memcpy(id128, &uid, 128);
    return MCP_OK;
}

int mcp_comm_init(mcp_ctx* ctx, int rank, int nranks, const void* id128) {
    if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return mcp_fail(ctx, MCP_ERR_INVALID, "mcp_comm_init: bad arguments");
    if (ctx->comm) return mcp_fail(ctx, MCP_ERR_INVALID, "mcp_comm_init: this ctx already has a communicator (one mcp_comm_init per ctx)");
    std::string why;
    if (!nccl_load(&why)) return mcp_fail(ctx, MCP_ERR_NCCL, "%s", why.c_str());
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));
    MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    mcp_nccl_uid uid;
    memcpy(&uid, id128, 128);
    void* comm = nullptr;
    int rc = g_nccl.CommInitRank(&comm, nranks, uid, rank);
    if (rc != 0) return mcp_fail(ctx, MCP_ERR_NCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    ctx->comm = comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
    xchg_setup(ctx);  // optional: mailboxes over NVLink peer memory for the in-kernel moment exchange
    return MCP_OK;
}

int mcp_comm_uses_peer_memory(const mcp_ctx* ctx) { return ctx && ctx->xchg.enabled ? 1 : 0; }

int mcp_comm_info(const mcp_ctx* ctx, int* rank, int* nranks) {
    if (!ctx) return MCP_ERR_INVALID;
    if (rank) *rank = ctx->rank;
    if (nranks) *nranks = ctx->nranks;
    return MCP_OK;
}

}  // extern "C"

cudaEvent_t mcp_prof_event(mcp_ctx* ctx, size_t i) {
    if (ctx->prof_ev.size() <= i) ctx->prof_ev.resize(i + 1, nullptr);  // slots are created on first use
    if (!ctx->prof_ev[i]) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        ctx->prof_ev[i] = e;
    }
    return ctx->prof_ev[i];
}

constexpr size_t MCP_STAGE_BYTES = (size_t)1 << 20;

void* mcp_stage_alloc(mcp_ctx* ctx, size_t bytes) {
    bytes = (bytes + 255) / 256 * 256;
    if (bytes > MCP_STAGE_BYTES / 4) return nullptr;
    if (!ctx->stage) {
        void* p = nullptr;
        if (cudaMallocHost(&p, MCP_STAGE_BYTES) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        ctx->stage = (unsigned char*)p;
        ctx->stage_off = 0;
    }
    if (ctx->stage_off + bytes > MCP_STAGE_BYTES) {  // wrap: everything handed out before must have been consumed
        cudaStreamSynchronize(ctx->stream);
        ctx->stage_off = 0;
    }
    void* out = ctx->stage + ctx->stage_off;
    ctx->stage_off += bytes;
    return out;
}

int mcp_h2d(mcp_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
    if (bytes == 0) return MCP_OK;
    ctx->h2d_bytes += bytes;
    void* pin = mcp_stage_alloc(ctx, bytes);
    if (pin) {
        memcpy(pin, src_host, bytes);
        MCP_CUDA(ctx, cudaMemcpyAsync(dst_dev, pin, bytes, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        MCP_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
        MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the caller's buffer may go away
    }
    return MCP_OK;
}

namespace {
struct KernelKey {
    int device;
    const void* fn;
    size_t smem;
    int block;
    bool operator<(const KernelKey& o) const {
        if (device != o.device) return device < o.device;
        if (fn != o.fn) return fn < o.fn;
        if (smem != o.smem) return smem < o.smem;
        return block < o.block;
    }
};
std::mutex g_kernel_mu;
std::map<KernelKey, int> g_kernel_occ;
std::map<std::pair<int, const void*>, size_t> g_kernel_smem_limit;  // dynamic shared-memory limit already granted (only ever raised)
}  // namespace

int mcp_kernel_config(mcp_ctx* ctx, const void* kernel, int block, size_t smem, int* occ_out) {
    const KernelKey key{ctx->device, kernel, smem, block};
    {
        std::lock_guard<std::mutex> lk(g_kernel_mu);
        auto it = g_kernel_occ.find(key);
        if (it != g_kernel_occ.end()) {
            if (occ_out) *occ_out = it->second;
            return MCP_OK;
        }
    }
    if (smem > 48 * 1024) {
        std::lock_guard<std::mutex> lk(g_kernel_mu);
        size_t& granted = g_kernel_smem_limit[std::make_pair(ctx->device, kernel)];
        if (smem > granted) {
            MCP_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            granted = smem;
        }
    }
    int occ = 0;
    MCP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, smem));
    {
        std::lock_guard<std::mutex> lk(g_kernel_mu);
        g_kernel_occ[key] = occ;
    }
    if (occ_out) *occ_out = occ;
    return MCP_OK;
}

int mcp_scratch_reserve(mcp_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return MCP_OK;
    // grow geometrically and in whole MiB: a caller that walks through growing sizes (the maturities of a surface) would
    // otherwise pay a stream synchronisation + cudaFree + cudaMalloc at every step
    if (bytes < 2 * ctx->scratch_bytes) bytes = 2 * ctx->scratch_bytes;
    bytes = (bytes + ((size_t)1 << 20) - 1) >> 20 << 20;
    MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    if (cudaMalloc(&ctx->scratch, bytes) != cudaSuccess) {
        cudaGetLastError();
        return mcp_fail(ctx, MCP_ERR_NOMEM, "device scratch allocation of %zu bytes failed", bytes);
    }
    ctx->scratch_bytes = bytes;
    return MCP_OK;
}

int mcp_carry_reserve(mcp_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->carry_bytes) return MCP_OK;
    MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->carry) cudaFree(ctx->carry);
    ctx->carry = nullptr;
    ctx->carry_bytes = 0;
    if (cudaMalloc(&ctx->carry, bytes) != cudaSuccess) {
        cudaGetLastError();
        return mcp_fail(ctx, MCP_ERR_NOMEM, "device carry allocation of %zu bytes failed", bytes);
    }
    ctx->carry_bytes = bytes;
    return MCP_OK;
}

int mcp_pinned_reserve(mcp_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->pinned_bytes) return MCP_OK;
    MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr;
    ctx->pinned_bytes = 0;
    if (cudaMallocHost(&ctx->pinned, bytes) != cudaSuccess) {
        cudaGetLastError();
        return mcp_fail(ctx, MCP_ERR_NOMEM, "pinned host allocation of %zu bytes failed", bytes);
    }
    ctx->pinned_bytes = bytes;
    return MCP_OK;
}
