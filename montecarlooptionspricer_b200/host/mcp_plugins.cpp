// mcp_plugins.cpp -- thin C++ plugin classes over the C ABI (see mcp_plugins.hpp).  Marshalling only: row pointers
// in, a double out; all arithmetic happens in libmcp_b200.so on the GPU.
#include "mcp_plugins.hpp"

#include <cstdlib>
#include <cstring>
#include <memory>
#include <random>

namespace mcp_b200 {

namespace {

uint64_t entropy_seed() {
    std::random_device rd;
    return ((uint64_t)rd() << 32) ^ (uint64_t)rd();
}

// rows of a rectangular matrix; empty vector when pricePaths is empty / has an empty first row / is ragged
bool row_pointers(const PathMatrix& paths, std::vector<const double*>& rows, bool* ragged) {
    *ragged = false;
    if (paths.empty() || paths[0].empty()) return false;
    const size_t m = paths[0].size();
    rows.resize(paths.size());
    for (size_t i = 0; i < paths.size(); ++i) {
        if (paths[i].size() != m) { *ragged = true; return false; }
        rows[i] = paths[i].data();
    }
    return true;
}

Engine& pick(Engine* e) { return e ? *e : Engine::thread_default(); }

}  // namespace

// ------------------------------------------------------------------------------------------- upload cache
// The reference's row loop hands the SAME path matrix to four pricers in a row (PredictionGen.cpp:788-791).  Each
// Engine therefore keeps the last uploaded matrix on the device, keyed by its dimensions and a 64-bit hash of its
// contents: the second to fourth pricer of a row find their slab already resident.  A hash match is only a candidate:
// the hit is confirmed against a host copy of the cached matrix (memcmp, row by row), so a collision can never price
// the previous matrix.  Matrices above 64 MiB are not cached (hashing them would cost as much as the copy).
namespace {

uint64_t hash_rows(const PathMatrix& paths) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ ((uint64_t)paths.size() << 32) ^ (uint64_t)paths[0].size();
    for (const auto& row : paths) {
        const double* p = row.data();
        for (size_t j = 0; j < row.size(); ++j) {
            uint64_t w;
            std::memcpy(&w, p + j, 8);
            h = (h ^ w) * 0x100000001B3ull;
            h ^= h >> 29;
        }
    }
    return h;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ Engine
Engine::Engine(int device) : device_(device) {
    const int rc = mcp_create(device, &ctx_);
    if (rc != MCP_OK) throw std::runtime_error(std::string("mcp_b200: ") + mcp_last_error(nullptr));
}

Engine::~Engine() {
    if (cache_ps_) mcp_pathset_destroy(cache_ps_);
    mcp_destroy(ctx_);
}

Engine& Engine::thread_default() {
    thread_local std::unique_ptr<Engine> eng;
    if (!eng) {
        const char* d = std::getenv("MCP_B200_DEVICE");
        eng.reset(new Engine(d ? std::atoi(d) : 0));
    }
    return *eng;
}

std::vector<unsigned char> Engine::unique_id() {
    std::vector<unsigned char> id(128);
    if (mcp_comm_unique_id(id.data()) != MCP_OK) throw std::runtime_error(std::string("mcp_b200: ") + mcp_last_error(nullptr));
    return id;
}

void Engine::comm_init(int rank, int nranks, const std::vector<unsigned char>& id128) {
    if (id128.size() != 128) throw std::runtime_error("mcp_b200: the NCCL unique id is 128 bytes");
    check(mcp_comm_init(ctx_, rank, nranks, id128.data()));
}

void Engine::check(int status) const {
    if (status == MCP_OK) return;
    const std::string msg = mcp_last_error(ctx_);
    // MCP_ERR_EMPTY_PATHS / MCP_ERR_DOMAIN carry the reference's own std::runtime_error text verbatim
    if (status == MCP_ERR_EMPTY_PATHS || status == MCP_ERR_DOMAIN) throw std::runtime_error(msg);
    throw std::runtime_error("mcp_b200 [" + std::to_string(status) + "]: " + msg);
}

// ------------------------------------------------------------------------------------------- DevicePaths
DevicePaths::DevicePaths(Engine& eng, const PathMatrix& paths, int dtype) {
    std::vector<const double*> rows;
    bool ragged = false;
    if (!row_pointers(paths, rows, &ragged))
        throw std::runtime_error(ragged ? "mcp_b200: ragged pricePaths (rows of different length)" : "mcp_b200: Empty pricePaths.");
    const size_t n = paths.size(), m = paths[0].size();
    const bool cacheable = dtype == MCP_F64 && n * m * sizeof(double) <= ((size_t)64 << 20);
    uint64_t h = 0;
    if (cacheable) {
        h = hash_rows(paths);
        if (eng.cache_ps_ && eng.cache_n_ == n && eng.cache_m_ == m && eng.cache_hash_ == h && eng.cache_copy_.size() == n * m) {
            bool same = true;
            for (size_t i = 0; i < n && same; ++i) same = std::memcmp(eng.cache_copy_.data() + i * m, rows[i], m * sizeof(double)) == 0;
            if (same) {
                ps_ = eng.cache_ps_;  // resident already
                owned_ = false;
                return;
            }
        }
    }
    eng.check(mcp_pathset_create(eng.handle(), (int64_t)n, (int)m - 1, dtype, &ps_));
    const int rc = mcp_pathset_upload_rows_f64(ps_, rows.data());
    if (rc != MCP_OK) {
        mcp_pathset_destroy(ps_);
        ps_ = nullptr;
        eng.check(rc);
    }
    if (cacheable) {
        if (eng.cache_ps_) mcp_pathset_destroy(eng.cache_ps_);
        eng.cache_ps_ = ps_; eng.cache_n_ = n; eng.cache_m_ = m; eng.cache_hash_ = h;
        eng.cache_copy_.resize(n * m);
        for (size_t i = 0; i < n; ++i) std::memcpy(eng.cache_copy_.data() + i * m, rows[i], m * sizeof(double));
        owned_ = false;  // the engine's cache owns it now
    }
}

DevicePaths::~DevicePaths() {
    if (owned_) mcp_pathset_destroy(ps_);
}

// ------------------------------------------------------------------------------------------ RoughVolatility
RoughVolatility::RoughVolatility() : seed_(entropy_seed()) {}
RoughVolatility::RoughVolatility(uint64_t seed, Engine* engine) : seed_(seed), engine_(engine) {}

mcp_rbergomi_params RoughVolatility::EstimateParams(const std::vector<double>& hist) {
    if (hist.size() < 2) throw std::runtime_error("Historical prices vector too small.");  // RoughVolatility.cpp:317-319
    mcp_rbergomi_params prm;
    if (mcp_estimate_rbergomi_params(hist.data(), (int64_t)hist.size(), &prm) != MCP_OK) throw std::runtime_error(mcp_last_error(nullptr));
    return prm;
}

PathMatrix RoughVolatility::GenerateStockPricePaths(const std::vector<double>& hist, int forward_steps, int path_num) {
    if (hist.size() < 2) throw std::runtime_error("Historical prices vector too small.");
    const int np = path_num > 0 ? path_num : 0, ns = forward_steps > 0 ? forward_steps : 0;
    PathMatrix out((size_t)np, std::vector<double>((size_t)ns + 1, 0.0));
    if (np == 0) return out;
    std::vector<double*> rows((size_t)np);
    for (int i = 0; i < np; ++i) rows[(size_t)i] = out[(size_t)i].data();
    Engine& eng = pick(engine_);
    eng.check(mcp_generate_stock_price_paths(eng.handle(), hist.data(), (int64_t)hist.size(), forward_steps, path_num, seed_, next_path_,
                                             rows.data()));
    next_path_ += (uint64_t)np;  // successive calls draw fresh Philox streams
    return out;
}

PathMatrix RoughVolatility::GenerateWithParams(const mcp_rbergomi_params& prm, int forward_steps, int path_num) {
    if (path_num <= 0 || forward_steps <= 0) throw std::runtime_error("mcp_b200: forward_steps and path_num must be positive");
    Engine& eng = pick(engine_);
    mcp_pathset* ps = nullptr;
    eng.check(mcp_pathset_create(eng.handle(), path_num, forward_steps, MCP_F32, &ps));
    PathMatrix out((size_t)path_num, std::vector<double>((size_t)forward_steps + 1, 0.0));
    std::vector<double*> rows((size_t)path_num);
    for (int i = 0; i < path_num; ++i) rows[(size_t)i] = out[(size_t)i].data();
    int rc = mcp_gen_rbergomi(eng.handle(), ps, &prm, seed_, next_path_, nullptr, nullptr);
    if (rc == MCP_OK) rc = mcp_pathset_download_rows_f64(ps, rows.data());
    mcp_pathset_destroy(ps);
    eng.check(rc);
    next_path_ += (uint64_t)path_num;
    return out;
}

// ------------------------------------------------------------------------------------------------- LSM
double LSM::PredictOptionPrice(const PathMatrix& paths, double r, double strike, double maturity, double dt, bool isCall, int polyOrder) {
    if (paths.empty() || paths[0].empty()) throw std::runtime_error("LSM::PredictOptionPrice: Empty pricePaths.");  // LSMPricer.cpp:28-30
    Engine& eng = pick(engine_);
    DevicePaths dp(eng, paths);  // fp64 slab: every decision on the caller's doubles
    mcp_lsm_params prm;
    prm.r = r; prm.strike = strike; prm.maturity = maturity; prm.dt = dt;
    prm.is_call = isCall ? 1 : 0; prm.poly_order = polyOrder; prm.basis = MCP_BASIS_MONOMIAL; prm.carry = MCP_F64;
    mcp_lsm_result res;
    eng.check(mcp_lsm_price(eng.handle(), dp.handle(), &prm, &res, nullptr, nullptr, nullptr));
    return res.price;
}

// ---------------------------------------------------------------------------------- MartingaleOptimization
double MartingaleOptimization::PredictOptionPrice(const PathMatrix& paths, double r, double strike, double maturity, double dt, bool isCall,
                                                  int polyOrder, int maxIterations) {
    if (paths.empty() || paths[0].empty()) throw std::runtime_error("MartingaleOptimization: Empty pricePaths.");  // :31-33
    if (maxIterations <= 0) throw std::runtime_error("MartingaleOptimization: maxIterations must be positive.");  // :34-36
    Engine& eng = pick(engine_);
    DevicePaths dp(eng, paths);
    double price = 0.0;
    eng.check(mcp_martingale_price(eng.handle(), dp.handle(), r, strike, maturity, dt, isCall ? 1 : 0, polyOrder, maxIterations, &price,
                                   nullptr, nullptr));
    return price;
}

// -------------------------------------------------------------------------------------- BranchingProcesses
BranchingProcesses::BranchingProcesses() : seed_(entropy_seed()) {}
BranchingProcesses::BranchingProcesses(uint64_t seed, Engine* engine) : seed_(seed), engine_(engine) {}

double BranchingProcesses::PredictOptionPrice(const PathMatrix& paths, double r, double strike, double maturity, double dt, bool isCall,
                                              int numBranches, const std::vector<int>& exerciseTimes) {
    if (paths.empty() || paths[0].empty()) throw std::runtime_error("BranchingProcesses: Empty pricePaths.");  // :22-24
    if (exerciseTimes.empty()) throw std::runtime_error("BranchingProcesses: No exercise times.");            // :25-27
    if (strike <= 0.0) throw std::runtime_error("BranchingProcesses: Strike must be positive.");              // :28-30
    Engine& eng = pick(engine_);
    DevicePaths dp(eng, paths);
    double price = 0.0;
    eng.check(mcp_branching_price(eng.handle(), dp.handle(), r, strike, maturity, dt, isCall ? 1 : 0, numBranches, exerciseTimes.data(),
                                  (int)exerciseTimes.size(), seed_ + calls_++, 0, nullptr, &price, nullptr, nullptr));
    return price;
}

// -------------------------------------------------------------------------------------- AsymptoticAnalysis
double AsymptoticAnalysis::PredictOptionPrice(const PathMatrix& paths, double r, double strike, double maturity, double dt, bool isCall,
                                              double sigma, double dividend) {
    if (paths.empty() || paths[0].empty()) return 0.0;                                                         // :48-50
    if (sigma <= 0.0) throw std::runtime_error("AsymptoticAnalysis: Volatility must be positive.");           // :51-53
    for (const auto& row : paths)
        if (row.size() != paths[0].size()) return 0.0;                                                         // :58-62
    // The reference also swallows every exception of its scan loop (:110-112); nothing in that loop can throw here,
    // and a missing device must stay loud, so engine errors propagate.
    Engine& eng = pick(engine_);
    DevicePaths dp(eng, paths);
    double price = 0.0;
    eng.check(mcp_asymptotic_price(eng.handle(), dp.handle(), r, strike, maturity, dt, isCall ? 1 : 0, sigma, dividend, &price));
    return price;
}

}  // namespace mcp_b200
