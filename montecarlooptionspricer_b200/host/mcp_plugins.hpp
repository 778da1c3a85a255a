// mcp_plugins.hpp -- C++ host mirror of the reference's pricing-method plugin interface on top of the C ABI
// (include/mcp_b200.h -> libmcp_b200.so).  Class names, method names, argument order and the std::runtime_error
// messages are the reference's, so a call site such as src/core/PredictionGen.cpp:566-570 / :736-737 / :788-791
// compiles against this header unchanged (define MCP_B200_DROP_IN to get the names in the global namespace):
//
//   RoughVolatility::GenerateStockPricePaths   include/models/RoughVolatility.h:15-19
//   LSM::PredictOptionPrice                    include/models/LSMPricer.h:8-14
//   MartingaleOptimization::PredictOptionPrice include/models/MartingaleOptimizationPricer.h:10-18
//   BranchingProcesses::PredictOptionPrice     include/models/BranchingProcessPricer.h:8-16
//   AsymptoticAnalysis::PredictOptionPrice     include/models/AsymptoticAnalysisPricer.h:8-15
//
// Every method uploads the caller's rows (kept in fp64 on the device), runs the CUDA path and returns by value;
// nothing is retained past the call.  The reference creates one private instance of each class per OpenMP thread
// (PredictionGen.cpp:542-570): here every host thread lazily owns one Engine (CUDA stream + workspaces), so the
// classes stay default-constructible and re-entrant across threads.  There is no CPU fallback: without a usable
// sm_100 device every call throws std::runtime_error carrying the C ABI's message.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "mcp_b200.h"

namespace mcp_b200 {

using PathMatrix = std::vector<std::vector<double>>;  // [path][step], the reference's layout

// RAII owner of one mcp_ctx.  Not copyable; one per host thread or per GPU rank.
class Engine {
public:
    explicit Engine(int device = 0);
    ~Engine();
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;

    mcp_ctx* handle() const { return ctx_; }
    int device() const { return device_; }
    // The calling thread's lazily created engine on `device` (env MCP_B200_DEVICE overrides the default 0).
    static Engine& thread_default();
    // Attach an NCCL communicator (rank 0 creates the 128-byte id with unique_id() and hands it to the others).
    static std::vector<unsigned char> unique_id();
    void comm_init(int rank, int nranks, const std::vector<unsigned char>& id128);
    // status -> the reference's exception contract
    void check(int status) const;

private:
    friend class DevicePaths;
    mcp_ctx* ctx_ = nullptr;
    int device_ = 0;
    // last uploaded path matrix (<= 64 MiB), keyed by dimensions + content hash: see DevicePaths
    mcp_pathset* cache_ps_ = nullptr;
    size_t cache_n_ = 0, cache_m_ = 0;
    uint64_t cache_hash_ = 0;
    std::vector<double> cache_copy_;  // host copy of the cached matrix: a hash match is CONFIRMED by comparing contents
};

// Device-resident copy of a caller's path matrix (fp64 slab, time-major).  Matrices up to 64 MiB are kept in the
// engine's one-entry cache keyed by a content hash, so consecutive pricers called on the same matrix upload it once.
class DevicePaths {
public:
    DevicePaths(Engine& eng, const PathMatrix& paths, int dtype = MCP_F64);
    ~DevicePaths();
    DevicePaths(const DevicePaths&) = delete;
    DevicePaths& operator=(const DevicePaths&) = delete;
    mcp_pathset* handle() const { return ps_; }

private:
    mcp_pathset* ps_ = nullptr;
    bool owned_ = true;  // false when the slab lives in the engine's upload cache
};

class RoughVolatility {
public:
    RoughVolatility();                                      // seed from std::random_device, like the reference's RNGs
    explicit RoughVolatility(uint64_t seed, Engine* engine = nullptr);
    PathMatrix GenerateStockPricePaths(const std::vector<double>& historical_prices, int forward_steps, int path_num);
    // [new] the reference can only estimate its parameters; explicit-parameter generation for benchmarks and parity
    PathMatrix GenerateWithParams(const mcp_rbergomi_params& prm, int forward_steps, int path_num);
    static mcp_rbergomi_params EstimateParams(const std::vector<double>& historical_prices);

private:
    uint64_t seed_;
    uint64_t next_path_ = 0;
    Engine* engine_ = nullptr;
};

class LSM {
public:
    LSM() = default;
    explicit LSM(Engine* engine) : engine_(engine) {}
    double PredictOptionPrice(const PathMatrix& pricePaths, double r, double strike, double maturity, double dt, bool isCall,
                              int polyOrder);

private:
    Engine* engine_ = nullptr;
};

class MartingaleOptimization {
public:
    MartingaleOptimization() = default;
    explicit MartingaleOptimization(Engine* engine) : engine_(engine) {}
    double PredictOptionPrice(const PathMatrix& pricePaths, double r, double strike, double maturity, double dt, bool isCall,
                              int polyOrder, int maxIterations = 5);

private:
    Engine* engine_ = nullptr;
};

class BranchingProcesses {
public:
    BranchingProcesses();                                   // resampling seed from std::random_device (reference: :84-85)
    explicit BranchingProcesses(uint64_t seed, Engine* engine = nullptr);
    double PredictOptionPrice(const PathMatrix& pricePaths, double r, double strike, double maturity, double dt, bool isCall,
                              int numBranches, const std::vector<int>& exerciseTimes);

private:
    uint64_t seed_;
    uint64_t calls_ = 0;
    Engine* engine_ = nullptr;
};

class AsymptoticAnalysis {
public:
    AsymptoticAnalysis() = default;
    explicit AsymptoticAnalysis(Engine* engine) : engine_(engine) {}
    double PredictOptionPrice(const PathMatrix& pricePaths, double r, double strike, double maturity, double dt, bool isCall,
                              double sigma, double dividend);

private:
    Engine* engine_ = nullptr;
};

}  // namespace mcp_b200

#ifdef MCP_B200_DROP_IN
using mcp_b200::AsymptoticAnalysis;
using mcp_b200::BranchingProcesses;
using mcp_b200::LSM;
using mcp_b200::MartingaleOptimization;
using mcp_b200::RoughVolatility;
#endif
