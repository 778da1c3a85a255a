// oracle/ref_api.cpp -- TEST INFRASTRUCTURE (CPU oracle), never part of the product path.
//
// extern "C" doorway into the UNMODIFIED reference translation units, which oracle/Makefile compiles from
// where they lie under /root/reference into oracle/_ref/libmcp_ref.so:
//   src/models/RoughVolatility.cpp            (force-including oracle/shim_normal.h: draw injection)
//   src/models/LSMPricer.cpp                  (against oracle/eigen_shim/Eigen/Dense)
//   src/models/MartingaleOptimizationPricer.cpp (same shim)
//   src/models/BranchingProcessPricer.cpp, src/models/AsymptoticAnalysisPricer.cpp (as they are)
// Nothing from the reference is copied into this repository; this file only CALLS it.
//
// The reference API can only *estimate* (xi, H, eta, rho) from a price history
// (RoughVolatility.cpp:326-331), so for explicit-parameter runs (BASELINE config 2) we call its private
// members rbergomiLambda / rbergomiPhi / fractionalGaussian / forwardVariance directly and restate only
// the price recursion of RoughVolatility.cpp:354-364.
#include <chrono>
#include <complex>
#include <cstddef>
#include <cstring>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#define private public
#include "models/RoughVolatility.h"
#undef private
#include "models/LSMPricer.h"
#include "models/MartingaleOptimizationPricer.h"
#include "models/BranchingProcessPricer.h"
#include "models/AsymptoticAnalysisPricer.h"

extern "C" {
thread_local const double* orc_inject_ptr = nullptr;
thread_local std::size_t orc_inject_left = 0;
thread_local std::size_t orc_inject_used = 0;
// resampling indices for BranchingProcessPricer.cpp (oracle/shim_uniform.h)
thread_local const int* orc_index_ptr = nullptr;
thread_local std::size_t orc_index_left = 0;
thread_local std::size_t orc_index_used = 0;
}

namespace {
thread_local std::string g_err;

struct InjectGuard {
    InjectGuard(const double* p, std::size_t n) {
        orc_inject_ptr = p;
        orc_inject_left = p ? n : 0;
        orc_inject_used = 0;
    }
    ~InjectGuard() { orc_inject_ptr = nullptr; orc_inject_left = 0; }
};

std::vector<std::vector<double>> to_vv(const double* flat, long N, long M) {
    std::vector<std::vector<double>> vv(static_cast<size_t>(N));
    for (long i = 0; i < N; ++i) vv[static_cast<size_t>(i)].assign(flat + i * M, flat + (i + 1) * M);
    return vv;
}

template <class F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    } catch (...) {
        g_err = "unknown exception";
        return -2;
    }
}
}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

int ref_omp_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// RoughVolatility::GenerateStockPricePaths (RoughVolatility.cpp:312-368) exactly as shipped; `draws`
// (nullable) replaces the RNG in reference order.  out is [n_paths][steps+1] row-major.
int ref_generate_paths(const double* hist, int n_hist, int steps, int n_paths, const double* draws,
                       std::size_t n_draws, double* out, std::size_t* used) {
    return guarded([&] {
        InjectGuard g(draws, n_draws);
        RoughVolatility rv;
        std::vector<double> h(hist, hist + n_hist);
        auto p = rv.GenerateStockPricePaths(h, steps, n_paths);
        if (used) *used = orc_inject_used;
        for (size_t i = 0; i < p.size(); ++i) std::memcpy(out + i * (steps + 1), p[i].data(), sizeof(double) * (steps + 1));
    });
}

// The estimators the reference applies to the history (RoughVolatility.cpp:324-331). out = {xi,H,eta,rho,S0}.
int ref_estimate_params(const double* hist, int n_hist, double* out5) {
    return guarded([&] {
        RoughVolatility rv;
        std::vector<double> h(hist, hist + n_hist);
        std::vector<double> rets = rv.logReturns(h);
        out5[0] = rv.estimateXi(rets, 1.0 / 252.0);
        out5[1] = rv.estimateH(rets);
        out5[2] = rv.estimateEta(rets, out5[1]);
        out5[3] = rv.estimateRho(rets);
        out5[4] = h.back();
    });
}

// phi = fft(+1)(zero-padded 0.5 t^{2H}) (RoughVolatility.cpp:212-236).  out holds 2*M doubles (re,im).
int ref_rbergomi_phi(int n_steps, double H, double dt, double* out, int cap, int* M_out) {
    return guarded([&] {
        RoughVolatility rv;
        std::vector<double> t(static_cast<size_t>(n_steps) + 1);
        for (size_t i = 0; i < t.size(); ++i) t[i] = i * dt;
        auto lam = rv.rbergomiLambda(t, H);
        auto phi = rv.rbergomiPhi(lam, H);
        *M_out = static_cast<int>(phi.size());
        if (static_cast<int>(phi.size()) > cap) throw std::runtime_error("phi buffer too small");
        for (size_t i = 0; i < phi.size(); ++i) { out[2 * i] = phi[i].real(); out[2 * i + 1] = phi[i].imag(); }
    });
}

// Explicit-parameter rough-vol paths through the reference's private members; draws [n_paths][4n] in
// reference order.  Optional X_out / v_out are [n_paths][n].
int ref_rbergomi_paths(double S0, double r, double xi, double H, double eta, double rho, double dt, int n,
                       long n_paths, const double* draws, double* out, double* X_out, double* v_out) {
    return guarded([&] {
        RoughVolatility rv;
        std::vector<double> t(static_cast<size_t>(n) + 1);
        for (size_t i = 0; i < t.size(); ++i) t[i] = i * dt;
        auto lam = rv.rbergomiLambda(t, H);
        auto phi = rv.rbergomiPhi(lam, H);
        const double sq_dt = std::sqrt(dt), rho_c = std::sqrt(1.0 - rho * rho);
        std::vector<std::complex<double>> Z(static_cast<size_t>(n));
        for (long p = 0; p < n_paths; ++p) {
            const double* d = draws + static_cast<size_t>(p) * 4 * n;
            for (int k = 0; k < n; ++k) Z[k] = std::complex<double>(d[2 * k], d[2 * k + 1]);
            std::vector<double> X = rv.fractionalGaussian(phi, Z, H, eta);
            std::vector<double> v = rv.forwardVariance(X, t, xi, H, eta);
            const double* W1 = d + 2 * n;
            const double* W2 = d + 3 * n;
            double* S = out + static_cast<size_t>(p) * (n + 1);
            S[0] = S0;
            for (int j = 1; j <= n; ++j) {  // restates RoughVolatility.cpp:354-364
                double dw1 = sq_dt * W1[j - 1], dw2 = sq_dt * W2[j - 1];
                double dW = rho * dw1 + rho_c * dw2;
                double vt = v[j - 1];
                S[j] = S[j - 1] * std::exp((r - 0.5 * vt) * dt + std::sqrt(std::max(0.0, vt)) * dW);
            }
            if (X_out) std::memcpy(X_out + static_cast<size_t>(p) * n, X.data(), sizeof(double) * n);
            if (v_out) std::memcpy(v_out + static_cast<size_t>(p) * n, v.data(), sizeof(double) * n);
        }
    });
}

// The four pricer plugins, unmodified.  paths is [N][M] row-major.
int ref_lsm_price(const double* paths, long N, long M, double r, double K, double T, double dt, int isCall,
                  int polyOrder, double* price) {
    return guarded([&] {
        auto vv = to_vv(paths, N, M);
        LSM lsm;
        *price = lsm.PredictOptionPrice(vv, r, K, T, dt, isCall != 0, polyOrder);
    });
}

int ref_martingale_price(const double* paths, long N, long M, double r, double K, double T, double dt, int isCall,
                         int polyOrder, int maxIterations, double* price) {
    return guarded([&] {
        auto vv = to_vv(paths, N, M);
        MartingaleOptimization mo;
        *price = mo.PredictOptionPrice(vv, r, K, T, dt, isCall != 0, polyOrder, maxIterations);
    });
}

// rp (nullable): resampled path indices in the reference's consumption order [path][date with continuation][branch]
// (BranchingProcessPricer.cpp:93-110, serial build); n_rp entries; *used (nullable) receives how many were drawn.
int ref_branching_price(const double* paths, long N, long M, double r, double K, double T, double dt, int isCall,
                        int numBranches, const int* ex, int n_ex, const int* rp, long n_rp, long* used, double* price) {
    return guarded([&] {
        auto vv = to_vv(paths, N, M);
        std::vector<int> e(ex, ex + n_ex);
        BranchingProcesses bp;
        orc_index_ptr = rp;
        orc_index_left = rp ? static_cast<std::size_t>(n_rp) : 0;
        orc_index_used = 0;
        struct Reset { ~Reset() { orc_index_ptr = nullptr; orc_index_left = 0; } } reset;
        *price = bp.PredictOptionPrice(vv, r, K, T, dt, isCall != 0, numBranches, e);
        if (used) *used = static_cast<long>(orc_index_used);
    });
}

int ref_asymptotic_price(const double* paths, long N, long M, double r, double K, double T, double dt, int isCall,
                         double sigma, double dividend, double* price) {
    return guarded([&] {
        auto vv = to_vv(paths, N, M);
        AsymptoticAnalysis aa;
        *price = aa.PredictOptionPrice(vv, r, K, T, dt, isCall != 0, sigma, dividend);
    });
}

// CPU baseline harness: the reference parallelises only over contracts/rows
// (`#pragma omp parallel` + `omp for schedule(dynamic)`, src/core/PredictionGen.cpp:542-546), each row
// calling GenerateStockPricePaths (:736-737) then the pricers (:788-791).  We run n_contracts
// independent (generate + LSM price) rows of n_paths x n_steps under the same pragma, RNG as shipped,
// and return wall seconds.  prices_out[n_contracts].  The history is synthetic (hist, n_hist).
int ref_bench_rows(const double* hist, int n_hist, int n_contracts, int n_paths, int n_steps, double r,
                   double strike, int isCall, int polyOrder, int threads, double* prices_out, double* seconds,
                   double* gen_seconds_sum, double* lsm_seconds_sum) {
    return guarded([&] {
        std::vector<double> h(hist, hist + n_hist);
        const double dt = 1.0 / 252.0, maturity = n_steps * dt + 1e-9;
        double gsum = 0.0, lsum = 0.0;
        bool failed = false;
        std::string msg;
#ifdef _OPENMP
        if (threads > 0) omp_set_num_threads(threads);
#endif
        auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel reduction(+ : gsum, lsum)
        {
#pragma omp for schedule(dynamic)
            for (int c = 0; c < n_contracts; ++c) {
                try {
                    RoughVolatility rv;
                    LSM lsm;
                    auto a = std::chrono::steady_clock::now();
                    auto paths = rv.GenerateStockPricePaths(h, n_steps, n_paths);
                    auto b = std::chrono::steady_clock::now();
                    double px = lsm.PredictOptionPrice(paths, r, strike, maturity, dt, isCall != 0, polyOrder);
                    auto d = std::chrono::steady_clock::now();
                    gsum += std::chrono::duration<double>(b - a).count();
                    lsum += std::chrono::duration<double>(d - b).count();
                    if (prices_out) prices_out[c] = px;
                } catch (const std::exception& e) {
#pragma omp critical
                    { failed = true; msg = e.what(); }
                }
            }
        }
        auto t1 = std::chrono::steady_clock::now();
        if (failed) throw std::runtime_error(msg);
        *seconds = std::chrono::duration<double>(t1 - t0).count();
        if (gen_seconds_sum) *gen_seconds_sum = gsum;
        if (lsm_seconds_sum) *lsm_seconds_sum = lsum;
    });
}

}  // extern "C"
