// plugin_latency.cpp -- per-call latency of the C++ plugin classes on ONE PredictionGen-sized row (250 paths,
// floor(dte/365*252) steps, PredictionGen.cpp:718-719), single host thread: where a row's time goes.
//   plugin_latency [dte=91] [paths=250] [repeats=20]
#define MCP_B200_DROP_IN
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "mcp_plugins.hpp"

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    const int dte = argc > 1 ? std::atoi(argv[1]) : 91, n_paths = argc > 2 ? std::atoi(argv[2]) : 250, reps = argc > 3 ? std::atoi(argv[3]) : 20;
    const double r = 0.04, dt = 1.0 / 252.0, maturity = dte / 365.0;
    const int steps = (int)std::floor(maturity * 252.0);
    std::vector<double> hist(300);
    double s = 100.0;
    unsigned long long z = 88172645463325252ull;
    for (int i = 0; i < 300; ++i) {
        z ^= z << 13; z ^= z >> 7; z ^= z << 17;
        s *= std::exp(0.0126 * (((double)(z >> 11) / 9007199254740992.0) - 0.5) * 3.4641);
        hist[i] = s;
    }
    std::vector<int> ex(steps);
    for (int j = 0; j < steps; ++j) ex[j] = j;
    RoughVolatility rv(1);
    AsymptoticAnalysis aa; BranchingProcesses bp(2); LSM lsm; MartingaleOptimization mo;
    double t[5] = {0, 0, 0, 0, 0}, sink = 0.0;
    for (int it = -2; it < reps; ++it) {  // two warm-up rounds
        double t0 = now_ms();
        auto paths = rv.GenerateStockPricePaths(hist, steps, n_paths);
        double t1 = now_ms();
        sink += aa.PredictOptionPrice(paths, r, hist.back(), maturity, dt, false, 0.2, 0.0);
        double t2 = now_ms();
        sink += bp.PredictOptionPrice(paths, r, hist.back(), maturity, dt, false, 10, ex);
        double t3 = now_ms();
        sink += lsm.PredictOptionPrice(paths, r, hist.back(), maturity, dt, false, 2);
        double t4 = now_ms();
        sink += mo.PredictOptionPrice(paths, r, hist.back(), maturity, dt, false, 2);
        double t5 = now_ms();
        if (it >= 0) { t[0] += t1 - t0; t[1] += t2 - t1; t[2] += t3 - t2; t[3] += t4 - t3; t[4] += t5 - t4; }
    }
    std::printf("row %d paths x %d steps, ms per call: generate %.3f | asymptotic %.3f | branching %.3f | lsm %.3f | martingale %.3f | row total %.3f (sink %.3f)\n",
                n_paths, steps, t[0] / reps, t[1] / reps, t[2] / reps, t[3] / reps, t[4] / reps, (t[0] + t[1] + t[2] + t[3] + t[4]) / reps, sink);
    return 0;
}
