// surface.cu -- batched strike x maturity sweep under rough-volatility LSM (BASELINE config 5).
//
// The reference prices one contract per CSV row, regenerating paths for every row (src/core/PredictionGen.cpp:718-737,
// steps = floor(maturity * 252), 250 paths).  A surface of C = n_maturities x n_strikes contracts on the same
// underlying shares its paths per maturity: one slab of n_paths x floor(T_m * steps_per_year) steps is generated per
// maturity and every strike is priced on it by the fused LSM sweep, so generation is paid n_maturities times, not C.
// Contracts are independent, hence the multi-GPU split needs NO collective: rank g prices maturities
// g, g + G, g + 2G, ... (mat_first / mat_stride) with its own un-sharded path set (pass a ctx WITHOUT a communicator;
// with one attached, paths are sharded instead and every contract's regression is global).
#include <math.h>
#include <stdlib.h>

#include <chrono>

#include <vector>

#include "common.cuh"

extern "C" int mcp_price_surface_rbergomi_lsm(mcp_ctx* ctx, const mcp_rbergomi_params* model, const mcp_lsm_params* lsm_tmpl, const double* strikes,
                                              int n_strikes, const double* maturities, int n_maturities, int steps_per_year, int64_t n_paths,
                                              uint64_t seed, uint64_t path_offset, int mat_first, int mat_stride, double* prices,
                                              double* std_errors, float* gen_ms_total, float* lsm_ms_total) {
    if (!ctx || !model || !lsm_tmpl || !strikes || !maturities || !prices) return MCP_ERR_INVALID;
    if (n_strikes <= 0 || n_maturities <= 0 || steps_per_year <= 0 || n_paths <= 0 || mat_first < 0 || mat_stride <= 0)
        return mcp_fail(ctx, MCP_ERR_INVALID, "surface: bad sizes");
    float gen_total = 0.f, lsm_total = 0.f;
    cudaEvent_t e0, e1;
    MCP_CUDA(ctx, cudaEventCreate(&e0));
    MCP_CUDA(ctx, cudaEventCreate(&e1));
    int rc = MCP_OK;
    // ONE slab, sized for the longest maturity this call owns, serves every maturity (row j of the time-major slab is the
    // same memory whatever the number of rows in use): no 4 GB cudaMalloc / cudaFree pair per maturity
    int n_steps_max = 0;
    for (int mi = mat_first; mi < n_maturities; mi += mat_stride) {
        const int n_steps = (int)floor(maturities[mi] * (double)steps_per_year);
        if (n_steps > n_steps_max) n_steps_max = n_steps;
    }
    // ... and the slab survives the call: the next surface with the same path count and no more rows reuses it
    mcp_pathset* ps = ctx->cached_surface_ps;
    int n_rows_have = ps ? ps->n_steps : 0;
    if (ps && (ps->n_paths != n_paths || ps->dtype != MCP_F32 || n_rows_have < n_steps_max)) {
        mcp_pathset_destroy(ps);
        ps = nullptr;
    }
    if (!ps && n_steps_max >= 1) {
        rc = mcp_pathset_create(ctx, n_paths, n_steps_max, MCP_F32, &ps);
        n_rows_have = n_steps_max;
        if (rc == MCP_OK) ctx->cached_surface_ps = ps;
    }
    // longest maturity first: every grow-only workspace (scratch, carry) is sized once, by the first ladder
    std::vector<int> owned;
    for (int mi = mat_first; mi < n_maturities; mi += mat_stride) owned.push_back(mi);
    for (size_t oi = owned.size(); oi-- > 0 && rc == MCP_OK;) {
        const int mi = owned[oi];
        const double T = maturities[mi];
        const int n_steps = (int)floor(T * (double)steps_per_year);  // PredictionGen.cpp:718
        if (n_steps < 1) {  // PredictionGen.cpp:720-733 skips such rows and writes zeros
            for (int k = 0; k < n_strikes; ++k) {
                prices[(size_t)mi * n_strikes + k] = 0.0;
                if (std_errors) std_errors[(size_t)mi * n_strikes + k] = 0.0;
            }
            continue;
        }
        ps->n_steps = n_steps;  // a view of the first n_steps + 1 rows
        const bool trace = getenv("MCP_SURFACE_TRACE") != nullptr;
        const auto t0 = std::chrono::steady_clock::now();
        cudaEventRecord(e0, ctx->stream);
        rc = mcp_gen_rbergomi(ctx, ps, model, seed + 0x9E3779B97F4A7C15ull * (uint64_t)(mi + 1), path_offset, nullptr, nullptr);
        cudaEventRecord(e1, ctx->stream);
        const auto t1 = std::chrono::steady_clock::now();
        if (rc == MCP_OK) {  // the whole strike ladder of this maturity: one sweep per 16 strikes in throughput mode
            mcp_lsm_params prm = *lsm_tmpl;
            prm.maturity = T;
            prm.dt = model->dt;
            std::vector<mcp_lsm_result> res((size_t)n_strikes);
            rc = mcp_lsm_price_multi(ctx, ps, &prm, strikes, n_strikes, res.data());
            for (int k = 0; k < n_strikes && rc == MCP_OK; ++k) {
                prices[(size_t)mi * n_strikes + k] = res[(size_t)k].price;
                if (std_errors) std_errors[(size_t)mi * n_strikes + k] = res[(size_t)k].std_error;
                lsm_total += res[(size_t)k].elapsed_ms;
            }
        }
        if (rc == MCP_OK) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) gen_total += ms;
            if (trace) {
                const auto t2 = std::chrono::steady_clock::now();
                fprintf(stderr, "surface: maturity %d (%d steps): generator call %.2f ms host (%.2f ms device), ladder call %.2f ms host\n", mi, n_steps,
                        std::chrono::duration<double, std::milli>(t1 - t0).count(), ms, std::chrono::duration<double, std::milli>(t2 - t1).count());
            }
        }
    }
    if (ps) ps->n_steps = n_rows_have;  // the cached slab keeps its full height
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (gen_ms_total) *gen_ms_total = gen_total;
    if (lsm_ms_total) *lsm_ms_total = lsm_total;
    return rc;
}
