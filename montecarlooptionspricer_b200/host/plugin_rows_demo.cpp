// plugin_rows_demo.cpp -- a PredictionGen-shaped caller of the C++ plugin classes (src/core/PredictionGen.cpp:542-570,
// :700-719, :736-737, :780-791): per "row" one history -> GenerateStockPricePaths(hist, steps, 250) -> four pricers,
// rows spread over OpenMP threads with private pricer instances.  Used by tests/test_gpu_plugins_cpp.py: it writes
// the generated paths and the four prices of every row to a binary file that the test re-prices with the CPU oracle,
// and it checks the exception contract and thread re-entrancy itself (exit code != 0 on any failure).
//
//   plugin_rows_demo <out.bin> [rows=8] [paths=250] [dte=91]
#define MCP_B200_DROP_IN
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "mcp_plugins.hpp"

static std::vector<double> synthetic_history(int n, uint64_t seed) {
    // deterministic pseudo-random walk, ~20% annualised volatility (splitmix64 + Box-Muller on the host: input data only)
    auto next = [&seed]() {
        seed += 0x9E3779B97F4A7C15ull;
        uint64_t z = seed;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    std::vector<double> h(n);
    double s = 100.0;
    for (int i = 0; i < n; ++i) {
        const double u1 = ((double)(next() >> 11) + 0.5) / 9007199254740992.0, u2 = ((double)(next() >> 11) + 0.5) / 9007199254740992.0;
        s *= std::exp(0.0126 * std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2));
        h[i] = s;
    }
    return h;
}

template <typename F>
static int expect_throw(const char* what, const char* msg, F&& f) {
    try {
        f();
    } catch (const std::runtime_error& e) {
        if (std::strcmp(e.what(), msg) == 0) return 0;
        std::fprintf(stderr, "FAIL %s: message '%s' != '%s'\n", what, e.what(), msg);
        return 1;
    }
    std::fprintf(stderr, "FAIL %s: no exception\n", what);
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s out.bin [rows] [paths] [dte]\n", argv[0]); return 2; }
    const int n_rows = argc > 2 ? std::atoi(argv[2]) : 8, n_paths = argc > 3 ? std::atoi(argv[3]) : 250, dte = argc > 4 ? std::atoi(argv[4]) : 91;
    const double r = 0.04, dt = 1.0 / 252.0, maturity = dte / 365.0;          // PredictionGen.cpp:700-703
    const int steps = (int)std::floor(maturity * 252.0);                        // :718
    const int polyOrder = 2, numBranches = 10;                                  // :789-791
    int failures = 0;

    struct Row { std::vector<double> hist; mcp_b200::PathMatrix paths; double strike, sigma, aa, bp, lsm, mo; };
    std::vector<Row> rows(n_rows);
    const auto t_begin = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic) reduction(+ : failures)
    for (int idx = 0; idx < n_rows; ++idx) {
        try {
            RoughVolatility rv(1000 + idx);                                     // :566-570: private instances per thread
            AsymptoticAnalysis aa; BranchingProcesses bp(77 + idx); LSM lsm; MartingaleOptimization mo;
            Row& row = rows[idx];
            row.hist = synthetic_history(300, 42 + idx);
            row.strike = row.hist.back() * (1.0 - 0.02 * (idx % 3 - 1));        // :704-705 K = S (1 - distance)
            row.sigma = 0.2;
            row.paths = rv.GenerateStockPricePaths(row.hist, steps, n_paths);   // :736-737
            std::vector<int> ex(steps);
            for (int j = 0; j < steps; ++j) ex[j] = j;                          // :780-783
            const bool isCall = (idx % 2) == 1;
            row.aa = aa.PredictOptionPrice(row.paths, r, row.strike, maturity, dt, isCall, row.sigma, 0.0);
            row.bp = bp.PredictOptionPrice(row.paths, r, row.strike, maturity, dt, isCall, numBranches, ex);
            row.lsm = lsm.PredictOptionPrice(row.paths, r, row.strike, maturity, dt, isCall, polyOrder);
            row.mo = mo.PredictOptionPrice(row.paths, r, row.strike, maturity, dt, isCall, polyOrder);
        } catch (const std::exception& e) {
            std::fprintf(stderr, "row %d: %s\n", idx, e.what());
            failures += 1;
        }
    }
    const double row_loop_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    if (failures) return 1;

    // determinism + re-entrancy: the same seed on another thread reproduces row 0 bit for bit
    {
        RoughVolatility rv(1000);
        auto again = rv.GenerateStockPricePaths(rows[0].hist, steps, n_paths);
        if (again != rows[0].paths) { std::fprintf(stderr, "FAIL: seeded generation is not reproducible\n"); ++failures; }
        auto fresh = rv.GenerateStockPricePaths(rows[0].hist, steps, n_paths);
        if (fresh == again) { std::fprintf(stderr, "FAIL: a second call returned the same paths\n"); ++failures; }
    }
    // exception contract (messages are the reference's)
    const mcp_b200::PathMatrix empty;
    const std::vector<int> one_ex{0};
    failures += expect_throw("rv", "Historical prices vector too small.", [] { RoughVolatility().GenerateStockPricePaths({100.0}, 5, 5); });
    failures += expect_throw("lsm", "LSM::PredictOptionPrice: Empty pricePaths.", [&] { LSM().PredictOptionPrice(empty, r, 100, 1, dt, false, 2); });
    failures += expect_throw("mo", "MartingaleOptimization: Empty pricePaths.", [&] { MartingaleOptimization().PredictOptionPrice(empty, r, 100, 1, dt, false, 2); });
    failures += expect_throw("mo-iter", "MartingaleOptimization: maxIterations must be positive.",
                             [&] { MartingaleOptimization().PredictOptionPrice(rows[0].paths, r, 100, 1, dt, false, 2, 0); });
    failures += expect_throw("bp", "BranchingProcesses: Empty pricePaths.", [&] { BranchingProcesses().PredictOptionPrice(empty, r, 100, 1, dt, false, 10, one_ex); });
    failures += expect_throw("bp-ex", "BranchingProcesses: No exercise times.", [&] { BranchingProcesses().PredictOptionPrice(rows[0].paths, r, 100, 1, dt, false, 10, {}); });
    failures += expect_throw("bp-k", "BranchingProcesses: Strike must be positive.", [&] { BranchingProcesses().PredictOptionPrice(rows[0].paths, r, 0.0, 1, dt, false, 10, one_ex); });
    failures += expect_throw("aa", "AsymptoticAnalysis: Volatility must be positive.", [&] { AsymptoticAnalysis().PredictOptionPrice(rows[0].paths, r, 100, 1, dt, false, 0.0, 0.0); });
    if (AsymptoticAnalysis().PredictOptionPrice(empty, r, 100, 1, dt, false, 0.2, 0.0) != 0.0) { std::fprintf(stderr, "FAIL: aa(empty) != 0\n"); ++failures; }

    FILE* f = std::fopen(argv[1], "wb");
    if (!f) { std::perror("fopen"); return 2; }
    const int32_t hdr[6] = {n_rows, n_paths, steps + 1, 300, polyOrder, numBranches};
    std::fwrite(hdr, sizeof(hdr), 1, f);
    const double scal[3] = {r, dt, maturity};
    std::fwrite(scal, sizeof(scal), 1, f);
    for (const Row& row : rows) {
        const double v[6] = {row.strike, row.sigma, row.aa, row.bp, row.lsm, row.mo};
        std::fwrite(v, sizeof(v), 1, f);
        std::fwrite(row.hist.data(), 8, row.hist.size(), f);
        for (const auto& p : row.paths) std::fwrite(p.data(), 8, p.size(), f);
    }
    std::fclose(f);
    int threads = 1;
#ifdef _OPENMP
    threads = omp_get_max_threads();
#endif
    std::printf("plugin_rows_demo: %d rows x %d paths x %d steps on %d host threads, %d failures; row loop %.3f s = %.0f rows/s (4 pricers + generation per row)\n",
                n_rows, n_paths, steps, threads, failures, row_loop_s, n_rows / row_loop_s);
    return failures ? 1 : 0;
}
