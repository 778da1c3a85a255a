// rows.cu -- the reference's row loop as ONE batch: src/core/PredictionGen.cpp:542-866 prices every CSV row with
// 250 fresh paths (:719) through four pricers (:788-791), one row per OpenMP thread at a time.  Per-row calls from many
// host threads serialise inside the CUDA driver; here a batch of rows is three launches in total:
//   1. rbergomi_rows_kernel   all rows' paths (blockIdx.y = row, per-row model / step count / tables)
//   2. rows_price_kernel      one CTA per row runs Asymptotic, Branching, LSM and Martingale out of shared memory
//                             (small_bodies.cuh -- the same device routines the per-row API uses for small path sets)
// Rows are independent: a multi-GPU run gives each rank a slice of the rows, no collective.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "small_bodies.cuh"

namespace {

struct RowDev {
    int64_t slab_off;
    int n_steps, is_call;
    double K, r, dt, maturity, disc, sigma, dividend;
};

// 250 paths per row (PredictionGen.cpp:719) keep at most 250 threads busy whatever the CTA size, and the four pricers are chains of
// short loops separated by block-wide sums: latency, not throughput.  Small CTAs, several per SM, let one row's barriers hide
// behind another row's work: one 512-thread CTA per SM issued 19 % of the time with 23 % of its stalls at barriers (ncu r02a,
// 70 ms per 16384 rows); two of 256 threads 38 ms; four of 128 threads 27 ms; eight of 64 threads the same 27 ms.
constexpr int ROWS_NT = 128;
template <int P>
__global__ void __launch_bounds__(ROWS_NT, 4) rows_price_kernel(const RowDev* __restrict__ rows, const float* __restrict__ slabs, int64_t ld, int n, int n_br,
                                                             int max_iter, PhiloxKeys keys, uint64_t path_offset, double* __restrict__ out /*[rows][8]*/) {
    extern __shared__ double sm[];
    __shared__ double res[4];
    const RowDev R = rows[blockIdx.x];
    double* o = out + (size_t)blockIdx.x * 8;
    const int M = R.n_steps + 1;
    if (R.n_steps < 1) {  // PredictionGen.cpp:720-733: such rows are skipped and written as zeros
        if (threadIdx.x < 8) o[threadIdx.x] = 0.0;
        return;
    }
    const float* S = slabs + R.slab_off;
    // Asymptotic (PredictionGen.cpp:788)
    sb_asymptotic<float>(S, ld, n, M, R.K, R.is_call, R.r, R.dt, R.maturity, R.sigma, R.dividend, sm, res);
    if (threadIdx.x == 0) o[0] = res[1] > 0.0 ? res[0] / res[1] : 0.0;
    __syncthreads();
    // Branching, exercise dates 0 .. steps-1 (:780-783, :789)
    {
        int kend = M;
        for (int j = 0; j < M; ++j)
            if ((double)j * R.dt > R.maturity) { kend = j; break; }
        const int n_ex = R.n_steps < kend ? R.n_steps : kend;  // dates visited before the first t > maturity
        if (n_ex > 0) {
            const int j_hi = kend - 1 > n_ex - 1 ? kend - 1 : n_ex - 1;
            sb_branching<float>(S, ld, n, j_hi, 0, kend, R.n_steps - 1, nullptr, nullptr, n_ex, R.r, R.dt, R.K, R.is_call, n_br, keys,
                                path_offset + (uint64_t)blockIdx.x * (uint64_t)n, nullptr, sm, sm + n, res);
            if (threadIdx.x == 0) o[1] = 0.5 * (res[0] / (double)n + res[1] / (double)n);
        } else if (threadIdx.x == 0) {
            o[1] = 0.0;
        }
        __syncthreads();
    }
    // LSM (:790)
    sb_lsm<float, P>(S, ld, n, M, R.K, R.is_call, R.disc, R.dt, R.maturity, sm, nullptr, nullptr, nullptr, nullptr, nullptr, res);
    if (threadIdx.x == 0) {
        o[2] = res[0] / res[2];
        const double var = res[2] > 1.0 ? res[1] / (res[2] - 1.0) : 0.0;
        o[4] = var > 0.0 ? sqrt(var / res[2]) : 0.0;
    }
    __syncthreads();
    // Martingale (:791)
    sb_martingale<float, P>(S, ld, n, M, R.K, R.is_call, R.r, R.dt, R.maturity, max_iter, sm, sm + 4 * (size_t)n, res);
    if (threadIdx.x == 0) o[3] = 0.5 * (res[0] / (double)n + res[1] / (double)n);
}

typedef void (*RowsFn)(const RowDev*, const float*, int64_t, int, int, int, PhiloxKeys, uint64_t, double*);
RowsFn pick_rows(int p) {
    switch (p) {
        case 0: return rows_price_kernel<0>;
        case 1: return rows_price_kernel<1>;
        case 2: return rows_price_kernel<2>;
        case 3: return rows_price_kernel<3>;
        case 4: return rows_price_kernel<4>;
        case 5: return rows_price_kernel<5>;
        default: return rows_price_kernel<6>;
    }
}

}  // namespace

extern "C" int mcp_price_rows(mcp_ctx* ctx, const mcp_row* rows, int n_rows, int n_paths, int poly_order, int num_branches, int max_iterations,
                              uint64_t seed, uint64_t path_offset, mcp_row_result* out, float* gen_ms, float* price_ms) {
    if (!ctx || !rows || !out) return MCP_ERR_INVALID;
    if (n_rows <= 0) return MCP_OK;
    if (n_paths < 1 || n_paths > SB_MAX_PATHS) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "rows: n_paths %d outside [1, %d]", n_paths, SB_MAX_PATHS);
    if (poly_order < 0 || poly_order > MAXP) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "rows: poly_order %d outside [0, %d]", poly_order, MAXP);
    if (num_branches <= 0) return mcp_fail(ctx, MCP_ERR_INVALID, "rows: numBranches must be positive");
    if (max_iterations <= 0) return mcp_fail(ctx, MCP_ERR_DOMAIN, "MartingaleOptimization: maxIterations must be positive.");
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));
    int max_steps = 1;
    for (int r = 0; r < n_rows; ++r) {
        if (rows[r].n_steps > 512) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "rows: row %d has %d steps (> 512)", r, rows[r].n_steps);
        if (rows[r].n_steps >= 1 && !(rows[r].sigma > 0.0)) return mcp_fail(ctx, MCP_ERR_DOMAIN, "AsymptoticAnalysis: Volatility must be positive. (row %d)", r);
        if (rows[r].n_steps >= 1 && !(rows[r].strike > 0.0)) return mcp_fail(ctx, MCP_ERR_DOMAIN, "BranchingProcesses: Strike must be positive. (row %d)", r);
        if (rows[r].n_steps > max_steps) max_steps = rows[r].n_steps;
    }
    const int64_t ld = mcp_round_up(n_paths, 128);
    const int64_t slab_stride = (int64_t)(max_steps + 1) * ld;
    // Rows go through in chunks, TWO IN FLIGHT: while the device generates and prices chunk c, the host builds the transform
    // tables of chunk c + 1 straight into pinned memory (one asynchronous copy per chunk; the pageable, synchronous copy of the
    // first version cost more than the generation kernel).  Every per-chunk buffer exists twice (parity of the chunk index);
    // nothing in the loop waits for the device except the hand-over of a chunk's results, one chunk late.
    int64_t chunk = ((int64_t)2 << 30) / (slab_stride * 4);  // at most ~2 GiB of slabs per chunk (two chunks are resident)
    {
        // rows per chunk: measured at 16384 rows, 2048 / 4096 / 8192 / 16384 rows per chunk -> 44.2 / 40.0 / 38.5 / 38.3 ms (many CTA
        // waves per launch keep the tail of the pricing kernel small); MCP_ROWS_CHUNK overrides
        const char* ce = getenv("MCP_ROWS_CHUNK");
        const int64_t cap = ce && *ce ? atoll(ce) : 8192;
        if (cap >= 1 && chunk > cap) chunk = cap;
    }
    if (chunk < 1) chunk = 1;
    if (chunk > n_rows) chunk = n_rows;
    const size_t smem = (4 * (size_t)n_paths + (size_t)(max_steps + 1) * 2 + 8) * sizeof(double);
    if (smem > 200 * 1024) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "rows: %d paths x %d steps does not fit one CTA", n_paths, max_steps);
    RowsFn fn = pick_rows(poly_order);
    MCP_TRY(mcp_kernel_config(ctx, (const void*)fn, ROWS_NT, smem, nullptr));
    const size_t slab_bytes = (size_t)chunk * slab_stride * 4;
    const size_t rowdev_bytes = (size_t)mcp_round_up((int64_t)(chunk * sizeof(RowDev)), 256), out_bytes = (size_t)mcp_round_up(chunk * 8 * 8, 256);
    const size_t stage_bytes = (size_t)mcp_round_up((int64_t)mcp_rows_stage_bytes((int)chunk, max_steps), 256);
    MCP_TRY(mcp_carry_reserve(ctx, 2 * (slab_bytes + rowdev_bytes + out_bytes) + 4096));
    MCP_TRY(mcp_scratch_reserve(ctx, 2 * stage_bytes));
    MCP_TRY(mcp_pinned_reserve(ctx, 2 * (stage_bytes + rowdev_bytes + out_bytes)));
    struct Slot {
        float* d_slabs; RowDev* d_rows; double* d_out; unsigned char* d_stage;
        unsigned char* h_stage; RowDev* h_rows; double* h_out;
        cudaEvent_t e0, e1, e2, done;
        int64_t r0; int nr;
    } slot[2];
    for (int b = 0; b < 2; ++b) {
        unsigned char* dc = (unsigned char*)ctx->carry + (size_t)b * (slab_bytes + rowdev_bytes + out_bytes);
        slot[b].d_slabs = (float*)dc; slot[b].d_rows = (RowDev*)(dc + slab_bytes); slot[b].d_out = (double*)(dc + slab_bytes + rowdev_bytes);
        slot[b].d_stage = (unsigned char*)ctx->scratch + (size_t)b * stage_bytes;
        unsigned char* hp = (unsigned char*)ctx->pinned + (size_t)b * (stage_bytes + rowdev_bytes + out_bytes);
        slot[b].h_stage = hp; slot[b].h_rows = (RowDev*)(hp + stage_bytes); slot[b].h_out = (double*)(hp + stage_bytes + rowdev_bytes);
        slot[b].e0 = mcp_prof_event(ctx, 4080 + 4 * (size_t)b); slot[b].e1 = mcp_prof_event(ctx, 4081 + 4 * (size_t)b);
        slot[b].e2 = mcp_prof_event(ctx, 4082 + 4 * (size_t)b); slot[b].done = mcp_prof_event(ctx, 4083 + 4 * (size_t)b);
        if (!slot[b].e0 || !slot[b].e1 || !slot[b].e2 || !slot[b].done) return mcp_fail(ctx, MCP_ERR_CUDA, "cudaEventCreate failed");
        slot[b].r0 = 0; slot[b].nr = 0;
    }
    const PhiloxKeys keys = philox_make_keys(seed ^ 0x5bd1e995ull);  // resampling stream of the Branching pricer
    float g_total = 0.f, p_total = 0.f;
    int rc = MCP_OK;
    std::vector<int> steps((size_t)chunk);

    auto issue = [&](Slot& s, int64_t r0, int nr) -> int {
        s.r0 = r0; s.nr = nr;
        for (int k = 0; k < nr; ++k) {
            const mcp_row& R = rows[r0 + k];
            RowDev& D = s.h_rows[k];
            D.slab_off = (int64_t)k * slab_stride;
            // a degenerate model (H < 0, |rho| > 1, xi < 0, NaN ...) spoils only its own row: no paths, NaN results (see mcp_rows_generate)
            const bool degenerate = !(R.model.dt > 0.0) || !(R.model.H >= 0.0) || !(fabs(R.model.rho) <= 1.0) || !(R.model.xi >= 0.0);
            D.n_steps = degenerate ? 0 : R.n_steps;
            D.is_call = R.is_call;
            D.K = R.strike; D.r = R.r; D.dt = R.dt; D.maturity = R.maturity; D.sigma = R.sigma; D.dividend = R.dividend;
            D.disc = exp(-R.r * R.dt);
            steps[(size_t)k] = D.n_steps;
        }
        MCP_TRY(mcp_rows_generate(ctx, &rows[r0].model, sizeof(mcp_row), steps.data(), nr, n_paths, seed, path_offset + (uint64_t)r0 * (uint64_t)n_paths, s.d_slabs,
                                  slab_stride, ld, s.h_stage, s.d_stage, stage_bytes, s.e0));
        MCP_CUDA(ctx, cudaEventRecord(s.e1, ctx->stream));
        MCP_CUDA(ctx, mcp_memcpy_async(ctx, s.d_rows, s.h_rows, (size_t)nr * sizeof(RowDev), cudaMemcpyHostToDevice, ctx->stream));
        fn<<<nr, ROWS_NT, smem, ctx->stream>>>(s.d_rows, s.d_slabs, ld, n_paths, num_branches, max_iterations, keys, path_offset + (uint64_t)r0 * (uint64_t)n_paths, s.d_out);
        MCP_LAUNCH_CHECK(ctx);
        MCP_CUDA(ctx, cudaEventRecord(s.e2, ctx->stream));
        MCP_CUDA(ctx, mcp_memcpy_async(ctx, s.h_out, s.d_out, (size_t)nr * 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MCP_CUDA(ctx, cudaEventRecord(s.done, ctx->stream));
        return MCP_OK;
    };
    auto finish = [&](Slot& s) -> int {
        if (s.nr == 0) return MCP_OK;
        if (cudaEventSynchronize(s.done) != cudaSuccess || cudaGetLastError() != cudaSuccess)
            return mcp_fail(ctx, MCP_ERR_CUDA, "rows: pricing kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
        for (int k = 0; k < s.nr; ++k) {
            mcp_row_result& o = out[s.r0 + k];
            o.asymptotic = s.h_out[(size_t)k * 8 + 0];
            o.branching = s.h_out[(size_t)k * 8 + 1];
            o.lsm = s.h_out[(size_t)k * 8 + 2];
            o.martingale = s.h_out[(size_t)k * 8 + 3];
            o.lsm_std_error = s.h_out[(size_t)k * 8 + 4];
            if (s.h_rows[k].n_steps != rows[s.r0 + k].n_steps) {  // degenerate model: what the reference's NaN paths give
                const double qnan = nan("");
                o.asymptotic = o.branching = o.lsm = o.martingale = o.lsm_std_error = qnan;
            }
        }
        float a = 0.f, b = 0.f;
        if (cudaEventElapsedTime(&a, s.e0, s.e1) == cudaSuccess) g_total += a;
        if (cudaEventElapsedTime(&b, s.e1, s.e2) == cudaSuccess) p_total += b;
        s.nr = 0;
        return MCP_OK;
    };
    int c = 0;
    for (int64_t r0 = 0; r0 < n_rows && rc == MCP_OK; r0 += chunk, ++c) {
        const int nr = (int)(n_rows - r0 < chunk ? n_rows - r0 : chunk);
        Slot& s = slot[c & 1];
        rc = finish(s);                      // the chunk that used this slot two issues ago
        if (rc == MCP_OK) rc = issue(s, r0, nr);
    }
    for (int b = 0; b < 2; ++b) {
        const int r2 = finish(slot[(c + b) & 1]);  // oldest first
        if (rc == MCP_OK) rc = r2;
    }
    if (rc != MCP_OK) cudaStreamSynchronize(ctx->stream);
    if (gen_ms) *gen_ms = g_total;
    if (price_ms) *price_ms = p_total;
    return rc;
}
