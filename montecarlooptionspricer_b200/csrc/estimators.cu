// estimators.cu -- host-side glue of the exact-signature generator call:
//     RoughVolatility::GenerateStockPricePaths(historical_prices, forward_steps, path_num)
//     (include/models/RoughVolatility.h:15-19, src/models/RoughVolatility.cpp:312-368).
// The reference cannot be told its model parameters: it estimates them from the price history on every call
// (:324-331).  Those estimators are O(history) scalar fp64 work (<= 1826 prices, PredictionGen.cpp:247-258) and
// stay on the host; the path generation they parameterise runs on the device (gen_rbergomi.cu).
//   xi  = var(logret) / dt                         :141-145   (sample variance, n-1)
//   H   = slope of log F(w) vs log w, DFA-1         :72-122    (windows 4, 8, .. <= n/4; NOT clamped)
//   eta = 2 std(logret)                             :151-155
//   rho = corr(ret, ret^2), replaced by -0.3 if > 0 :157-169
//   r = 0.04, dt = 1/252, S0 = history.back()       :321-331
#include <math.h>

#include <limits>
#include <vector>

#include "common.cuh"

namespace {

double mean_of(const std::vector<double>& v) {  // RoughVolatility.cpp:20-23
    double s = 0.0;
    for (double x : v) s += x;
    return v.empty() ? 0.0 : s / (double)v.size();
}

double sample_var(const std::vector<double>& v) {  // :25-33
    if (v.size() < 2) return 0.0;
    const double m = mean_of(v);
    double acc = 0.0;
    for (double x : v) acc += (x - m) * (x - m);
    return acc / (double)(v.size() - 1);
}

double sample_cov(const std::vector<double>& x, const std::vector<double>& y) {  // :35-42
    if (x.size() != y.size() || x.size() < 2) return 0.0;
    const double mx = mean_of(x), my = mean_of(y);
    double acc = 0.0;
    for (size_t i = 0; i < x.size(); ++i) acc += (x[i] - mx) * (y[i] - my);
    return acc / (double)(x.size() - 1);
}

// Detrended fluctuation analysis, order 1 (:44-122): integrate the demeaned series, split into windows of w = 4, 8, ...
// <= n/4, remove a least-squares line (abscissa 1..w) from each window, F(w) = mean over windows of the residual
// rms, H = OLS slope of log F on log w.
double dfa_hurst(const std::vector<double>& series) {
    const size_t n = series.size();
    if (n < 2) return 0.5;
    std::vector<double> prof(series);
    const double m = mean_of(prof);
    for (double& x : prof) x -= m;
    for (size_t i = 1; i < n; ++i) prof[i] += prof[i - 1];

    std::vector<double> lw, lf;
    for (size_t w = 4; w <= n / 4; w *= 2) {
        std::vector<double> t(w), rms_list;
        for (size_t i = 0; i < w; ++i) t[i] = (double)(i + 1);
        const double tm = mean_of(t);
        for (size_t start = 0; start + w <= n; start += w) {
            std::vector<double> seg(prof.begin() + start, prof.begin() + start + w);
            const double ym = mean_of(seg);
            double num = 0.0, den = 0.0;
            for (size_t i = 0; i < w; ++i) {
                num += (t[i] - tm) * (seg[i] - ym);
                den += (t[i] - tm) * (t[i] - tm);
            }
            if (!(fabs(den) < 1e-14)) {  // :58
                const double slope = num / den, icpt = ym - slope * tm;
                for (size_t i = 0; i < w; ++i) seg[i] -= slope * t[i] + icpt;
            }
            double ss = 0.0;
            for (double e : seg) ss += e * e;
            rms_list.push_back(sqrt(ss / (double)w));
        }
        const double f = mean_of(rms_list);
        if (f > 0.0) {
            lw.push_back(log((double)w));
            lf.push_back(log(f));
        }
    }
    const size_t k = lw.size();
    if (k < 2) return 0.5;
    double sx = 0.0, sy = 0.0, sxx = 0.0, sxy = 0.0;
    for (size_t i = 0; i < k; ++i) {
        sx += lw[i];
        sy += lf[i];
        sxx += lw[i] * lw[i];
        sxy += lw[i] * lf[i];
    }
    return ((double)k * sxy - sx * sy) / ((double)k * sxx - sx * sx);
}

}  // namespace

extern "C" int mcp_estimate_rbergomi_params(const double* hist, int64_t n_hist, mcp_rbergomi_params* out) {
    if (!hist || !out) return MCP_ERR_INVALID;
    if (n_hist < 2) return mcp_fail(nullptr, MCP_ERR_DOMAIN, "Historical prices vector too small.");  // :317-319
    std::vector<double> ret;
    ret.reserve((size_t)n_hist - 1);
    for (int64_t i = 1; i < n_hist; ++i) ret.push_back(log(hist[i] / hist[i - 1]));  // :126-133
    const double dt = 1.0 / 252.0;
    std::vector<double> sq(ret.size());
    for (size_t i = 0; i < ret.size(); ++i) sq[i] = ret[i] * ret[i];
    const double var = sample_var(ret);
    double rho = sample_cov(ret, sq) / sqrt(var * sample_var(sq));
    if (rho > 0.0) rho = -0.3;  // :165-167 (a NaN correlation stays NaN, as in the reference)
    out->S0 = hist[n_hist - 1];
    out->r = 0.04;
    out->xi = var / dt;
    out->H = dfa_hurst(ret);
    out->eta = 2.0 * sqrt(var);
    out->rho = rho;
    out->dt = dt;
    return MCP_OK;
}

// One call = RoughVolatility().GenerateStockPricePaths(hist, forward_steps, path_num): estimate on the host, generate
// on the device with native Philox streams, hand back host rows [path][0..forward_steps] of doubles.
// Where the reference's own arithmetic yields NaN (H < 0 => sqrt(2H), :284; NaN rho) every column after S0 is NaN
// here too -- PredictionGen.cpp:753-777 rejects such rows on the caller's side.
extern "C" int mcp_generate_stock_price_paths(mcp_ctx* ctx, const double* hist, int64_t n_hist, int forward_steps, int path_num,
                                              uint64_t seed, uint64_t path_offset, double* const* rows) {
    if (!ctx || !rows) return MCP_ERR_INVALID;
    if (!hist || n_hist < 2) return mcp_fail(ctx, MCP_ERR_DOMAIN, "Historical prices vector too small.");
    if (path_num <= 0) return MCP_OK;  // the reference returns an empty vector
    mcp_rbergomi_params prm;
    MCP_TRY(mcp_estimate_rbergomi_params(hist, n_hist, &prm));
    if (forward_steps <= 0) {
        for (int i = 0; i < path_num; ++i) rows[i][0] = prm.S0;
        return MCP_OK;
    }
    const bool degenerate = !(prm.H >= 0.0) || !(fabs(prm.rho) <= 1.0) || !isfinite(prm.xi) || !isfinite(prm.eta) || !isfinite(prm.S0);
    if (degenerate) {
        const double qnan = std::numeric_limits<double>::quiet_NaN();
        for (int i = 0; i < path_num; ++i) {
            rows[i][0] = prm.S0;
            for (int j = 1; j <= forward_steps; ++j) rows[i][j] = qnan;
        }
        return MCP_OK;
    }
    mcp_pathset* ps = nullptr;
    MCP_TRY(mcp_pathset_create(ctx, path_num, forward_steps, MCP_F32, &ps));
    int rc = mcp_gen_rbergomi(ctx, ps, &prm, seed, path_offset, nullptr, nullptr);
    if (rc == MCP_OK) rc = mcp_pathset_download_rows_f64(ps, rows);
    if (rc == MCP_OK)
        for (int i = 0; i < path_num; ++i) rows[i][0] = prm.S0;  // column 0 is the caller's double, not its fp32 image
    mcp_pathset_destroy(ps);
    return rc;
}
