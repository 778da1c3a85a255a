/*
 * mcp_b200.h -- C ABI of the B200-native Monte-Carlo hot path (libmcp_b200.so).
 *
 * This is the drop-in boundary for ONE path of bcosm/MonteCarloOptionsPricer: normal generation ->
 * GBM / rough-volatility path simulation -> Longstaff-Schwartz backward induction + payoff averaging.
 * Everything behind these entry points is hand-written CUDA for sm_100a; there is no CPU fallback: every
 * call fails with MCP_ERR_CUDA when no device is usable.
 *
 * Reference interfaces each entry point replaces (paths relative to the reference repository):
 *   mcp_gen_rbergomi            RoughVolatility::GenerateStockPricePaths   include/models/RoughVolatility.h:15-19
 *                               (body src/models/RoughVolatility.cpp:312-368, fGn :212-309)
 *   mcp_gen_gbm                 [new] constant-variance special case of the same recursion (:354-364)
 *   mcp_pathset_upload_*        the `const std::vector<std::vector<double>>& pricePaths` argument that
 *                               every pricer takes (include/models/LSMPricer.h:8-14 and siblings)
 *   mcp_lsm_price               LSM::PredictOptionPrice                     include/models/LSMPricer.h:8-14
 *                               (body src/models/LSMPricer.cpp:19-102)
 *   mcp_lsm_price_host_rows     the same call, bound directly to host rows  src/core/PredictionGen.cpp:790
 *   mcp_estimate_rbergomi_params  RoughVolatility::estimateXi/H/Eta/Rho     src/models/RoughVolatility.cpp:72-169, :324-331
 *   mcp_generate_stock_price_paths  GenerateStockPricePaths, exact call shape  src/core/PredictionGen.cpp:736-737
 *   mcp_price_surface_rbergomi_lsm  [new] the row loop of src/core/PredictionGen.cpp:542-866 for a strike x maturity grid
 *   mcp_price_rows              the whole per-row block of src/core/PredictionGen.cpp:700-791, batched over rows
 *   mcp_gbm_nested_dual         [new] BASELINE config 4 (nested-simulation duality); no reference counterpart
 *   mcp_asymptotic_price        AsymptoticAnalysis::PredictOptionPrice      include/models/AsymptoticAnalysisPricer.h:8-15
 *   mcp_martingale_price        MartingaleOptimization::PredictOptionPrice  include/models/MartingaleOptimizationPricer.h:10-18
 *   mcp_branching_price         BranchingProcesses::PredictOptionPrice      include/models/BranchingProcessPricer.h:8-16
 *
 * Conventions: plain pointers and sizes only; caller allocates every output; no ownership transfer; no
 * exceptions cross the boundary; every function returns an int status (0 = ok, negative = error class) and
 * the message is available from mcp_last_error(ctx).  A ctx is NOT thread-safe: use one per host thread
 * (the reference instantiates its pricers per OpenMP thread, src/core/PredictionGen.cpp:566-570).
 */
#ifndef MCP_B200_H
#define MCP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCP_B200_ABI_VERSION 1

typedef struct mcp_ctx mcp_ctx;         /* engine handle: device, stream, workspaces, optional NCCL communicator */
typedef struct mcp_pathset mcp_pathset; /* device-resident TIME-MAJOR slab  S[(n_steps+1)][ld]  (ld >= n_paths) */

enum mcp_status {
    MCP_OK = 0,
    MCP_ERR_INVALID = -1,     /* bad argument */
    MCP_ERR_CUDA = -2,        /* CUDA runtime / no device / kernel failure */
    MCP_ERR_NCCL = -3,        /* NCCL missing or failed */
    MCP_ERR_NOMEM = -4,       /* device or host allocation failed */
    MCP_ERR_EMPTY_PATHS = -5, /* the reference's "Empty pricePaths." std::runtime_error (LSMPricer.cpp:28-30) */
    MCP_ERR_UNSUPPORTED = -6, /* valid request outside the implemented envelope (e.g. poly_order > 6) */
    MCP_ERR_DOMAIN = -7       /* the reference would throw a domain error (e.g. sigma <= 0, strike <= 0) */
};

enum mcp_dtype { MCP_F32 = 0, MCP_F64 = 1 };       /* storage type of a path slab / of the LSM carry */
enum mcp_basis { MCP_BASIS_MONOMIAL = 0,           /* 1, S, S^2 ... (LSMPricer.cpp:9-17) */
                 MCP_BASIS_LAGUERRE = 1,           /* L_0..L_p(S/K), unweighted: spans the same space */
                 MCP_BASIS_STANDARDISED = 2 };     /* the device's own basis: rows [c_0..c_p, mu, 1/s], x = (S - mu) / s */

/* ------------------------------------------------------------------------------------------- engine */
int mcp_abi_version(void);
int mcp_create(int device, mcp_ctx **out);
int mcp_destroy(mcp_ctx *ctx); /* also releases every pathset of this ctx that is still alive: their handles die with it */
const char *mcp_last_error(const mcp_ctx *ctx); /* ctx may be NULL: error of the last failed mcp_create */
int mcp_set_stream(mcp_ctx *ctx, void *cuda_stream); /* run on a caller-owned cudaStream_t (NULL = own stream) */
int mcp_synchronize(mcp_ctx *ctx);
int mcp_device_info(mcp_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *free_bytes,
                    size_t *total_bytes);
uint64_t mcp_launch_count(const mcp_ctx *ctx);  /* kernels launched by this ctx since creation */
/* bytes this ctx has copied host->device and device->host so far (counted where each copy is issued) */
int mcp_copy_counters(const mcp_ctx *ctx, uint64_t *h2d_bytes, uint64_t *d2h_bytes);

/* Optional per-kernel timing (CUDA events on the ctx stream around every launch of the two hot kernels).
 * Off by default: the extra event records sit between launches and are not wanted in a throughput run. */
typedef struct mcp_profile {
    float gen_kernel_ms;    /* last mcp_gen_*: the path kernel alone */
    float sweep_kernels_ms; /* last mcp_lsm_price: sum over its sweep launches */
    int n_sweep_launches;
    float lsm_total_ms;     /* last mcp_lsm_price: whole backward induction incl. solves / collectives */
    int n_sweep_steps;      /* time steps those launches swept (the persistent sweep runs all of them in one launch) */
} mcp_profile;
int mcp_set_profiling(mcp_ctx *ctx, int on);
int mcp_get_profile(const mcp_ctx *ctx, mcp_profile *out);

/* ---------------------------------------------------------------------------------------- multi-GPU
 * Paths shard across ranks (global path id = path_offset + local id); the only exchanged data are the
 * per-step regression moments and the final sums (NCCL all-reduce, fp64).  One ctx per rank/GPU. */
int mcp_comm_unique_id(void *id128);                                  /* rank 0: 128-byte ncclUniqueId */
int mcp_comm_init(mcp_ctx *ctx, int rank, int nranks, const void *id128);
int mcp_comm_info(const mcp_ctx *ctx, int *rank, int *nranks);
/* 1 when the per-step moment exchange runs inside the sweep kernel over NVLink peer memory (one process per GPU,
 * CUDA IPC mailboxes), 0 when it goes through ncclAllReduce (MCP_COMM_IMPL=nccl forces that) */
int mcp_comm_uses_peer_memory(const mcp_ctx *ctx);

/* ----------------------------------------------------------------------------------------- pathsets */
int mcp_pathset_create(mcp_ctx *ctx, int64_t n_paths, int n_steps, int dtype, mcp_pathset **out);
int mcp_pathset_destroy(mcp_pathset *ps);
int mcp_pathset_info(const mcp_pathset *ps, int64_t *n_paths, int *n_steps, int64_t *ld, int *dtype,
                     void **device_ptr);
/* host [path][step] (the reference's layout), row stride `ld_host` elements, n_steps+1 columns used */
int mcp_pathset_upload_f64(mcp_pathset *ps, const double *host, int64_t ld_host);
int mcp_pathset_upload_rows_f64(mcp_pathset *ps, const double *const *rows); /* vector<vector<double>> rows */
int mcp_pathset_download_f64(const mcp_pathset *ps, double *host, int64_t ld_host);
int mcp_pathset_download_rows_f64(const mcp_pathset *ps, double *const *rows); /* into vector<vector<double>> rows */
/* exact device values, time-major [step][path] */
int mcp_pathset_download_timemajor_f32(const mcp_pathset *ps, float *host, int64_t ld_host);

/* --------------------------------------------------------------------------------------- generators */
typedef struct mcp_rbergomi_params {
    double S0, r, xi, H, eta, rho, dt;
} mcp_rbergomi_params;

typedef struct mcp_gbm_params {
    double S0, r, sigma, dt;
} mcp_gbm_params;

/* Fills ps (n_steps, n_paths from the pathset).  Normals: native Philox4x32-10 streams keyed by `seed`
 * and the GLOBAL path id (path_offset + i), or, when `injected` != NULL, host floats in the reference's
 * consumption order: rbergomi [n_paths][4n] = Zre0,Zim0,..,Zre(n-1),Zim(n-1),W1[0..n),W2[0..n)
 * (RoughVolatility.cpp:346-352); gbm [n_paths][n].  `dump` (nullable, same layout) receives draws that
 * reproduce the generated paths through the reference's own formulas, so a native-Philox run can be replayed
 * through the CPU oracle: the normals actually used, except for rbergomi with 128 < n_steps <= 256, whose native
 * stream drives TWO paths with one complex transform (same law, 2 normals per path-step instead of 3;
 * csrc/gen_rbergomi_pair.cuh) -- there the Z slots hold the per-path draws equivalent to that shared transform.
 * Either way path i depends on (seed, path_offset + i) only: shards are slices of the whole. */
int mcp_gen_rbergomi(mcp_ctx *ctx, mcp_pathset *ps, const mcp_rbergomi_params *p, uint64_t seed,
                     uint64_t path_offset, const float *injected, float *dump);
int mcp_gen_gbm(mcp_ctx *ctx, mcp_pathset *ps, const mcp_gbm_params *p, uint64_t seed, uint64_t path_offset,
                const float *injected, float *dump);
/* Host-side constant tables of the rough-vol generator, for known-answer tests (pure host code, no device):
 * phis[2 M'] = phi_k sqrt(2H) eta / M' log2(e) for k < n, 0 beyond (RoughVolatility.cpp:212-236, :270, :284, :198-200);
 * comp2[M'] = -eta^2 t_k^{2H} log2(e) / 2 + log2(xi) (:304); sw[M'] = symmetrised spectrum of the pair stream.
 * M' = nextPow2(n_steps) is returned (negative = error); outputs are nullable. */
int mcp_rbergomi_host_tables(int n_steps, const mcp_rbergomi_params *p, float *phis, float *comp2, float *sw);
/* Raw generator words, for known-answer tests: out[4*i..4*i+3] = Philox4x32-10(ctr=(i_lo,i_hi,c2,c3), key=seed) */
int mcp_philox_raw(mcp_ctx *ctx, uint64_t seed, uint64_t first, int64_t count, uint32_t c2, uint32_t c3,
                   uint32_t *out_host);

/* ---------------------------------------------------------------------------------------------- LSM */
typedef struct mcp_lsm_params {
    double r, strike, maturity, dt;
    int is_call;
    int poly_order; /* 0..6.  The reference accepts any order, but its raw-monomial design is rank deficient by Eigen's own rule from
                     * order 5 on (S ~ 100): bdcSvd().solve() then drops directions (LSMPricer.cpp:76).  That cut is reproduced on the
                     * device (csrc/lsm_solve.cuh: exercise indices and price match the reference for orders 5 and 6 as they do for
                     * 1..4); orders above 6 would only add directions the reference discards and return MCP_ERR_UNSUPPORTED. */
    int basis;      /* mcp_basis: affects only the coefficient table that is returned */
    int carry;      /* mcp_dtype of the value carry V: MCP_F64 = parity mode, MCP_F32 = throughput mode */
} mcp_lsm_params;

typedef struct mcp_lsm_result {
    double price;      /* mean_i V[i][0]                              (LSMPricer.cpp:97-101) */
    double std_error;  /* [new] sample std of V[:,0] / sqrt(N): the error of the final average GIVEN the fitted regressions;
                        * independent runs of this value-iteration estimator scatter 2-4x wider */
    double sum_v0, sum_sq_dev; /* sum_i V0_i and sum_i (V0_i - mean)^2 over ALL ranks */
    int64_t n_paths_global;
    float elapsed_ms;  /* device time of the sweep (CUDA events on the ctx stream) */
    int n_kernel_launches;
} mcp_lsm_result;

/* coeffs (nullable): host [n_steps][poly_order+1] ([poly_order+3] for MCP_BASIS_STANDARDISED), row j = regression at step j in the requested basis,
 *                    zeros where no path was in the money / past maturity;
 * first_exercise (nullable): host int32 [n_paths], tau_i = min{ j : exercised } else n_steps;
 * v0 (nullable): host double [n_paths], V[i][0]. */
int mcp_lsm_price(mcp_ctx *ctx, const mcp_pathset *ps, const mcp_lsm_params *p, mcp_lsm_result *res,
                  double *coeffs, int32_t *first_exercise, double *v0);

/* n_strikes contracts that differ only in their strike, priced on the SAME path set (prm->strike is ignored).  In
 * throughput mode (fp32 slab, fp32 carry, one GPU, more than 4096 paths) up to 16 strikes share one sweep: the slab is
 * read once per step for the whole ladder, one warp per contract.  res: caller's array [n_strikes]. */
int mcp_lsm_price_multi(mcp_ctx *ctx, const mcp_pathset *ps, const mcp_lsm_params *prm, const double *strikes, int n_strikes,
                        mcp_lsm_result *res);

/* [new -- the reference computes no error estimate] Out-of-sample value of a fitted exercise policy.  coeffs_std: host
 * [n_steps][poly_order+3], the MCP_BASIS_STANDARDISED table returned by mcp_lsm_price on ANOTHER, independent path set with the
 * same contract and step grid.  Every path of `ps` is stopped at the first date where the reference's own rule exercises
 * (in the money by more than 1e-14 and !(immediate < fitted continuation), LSMPricer.cpp:55,85; the last column always pays
 * off, :37-40; no exercise past maturity, :43-49); res->price is the mean discounted realised payoff, res->std_error its
 * standard error (paths are independent given the coefficients, so this one is exact, unlike mcp_lsm_result::std_error of
 * the in-sample run), and in expectation a LOWER bound of the true price.  mean_stop_index (nullable): average stopping
 * column.  With a communicator attached the sums run over all ranks' shards. */
int mcp_lsm_policy_value(mcp_ctx *ctx, const mcp_pathset *ps, const mcp_lsm_params *prm, const double *coeffs_std,
                         mcp_lsm_result *res, double *mean_stop_index);

/* One call = LSM::PredictOptionPrice(pricePaths, r, strike, maturity, dt, isCall, polyOrder): uploads
 * n_paths host rows of n_cols doubles (kept in fp64 on the device), prices, returns the mean. */
int mcp_lsm_price_host_rows(mcp_ctx *ctx, const double *const *rows, int64_t n_paths, int n_cols, double r,
                            double strike, double maturity, double dt, int is_call, int poly_order,
                            double *price);

/* One call = generate (native Philox) + LSM on the device, nothing but parameters in and a result out.
 * n_paths is THIS rank's share; with a communicator the regression and the mean are global. */
int mcp_price_rbergomi_lsm(mcp_ctx *ctx, const mcp_rbergomi_params *model, const mcp_lsm_params *lsm,
                           int64_t n_paths, int n_steps, uint64_t seed, uint64_t path_offset,
                           mcp_lsm_result *res, float *gen_ms);

/* ------------------------------------------------- nested-simulation duality under GBM (config 4) [new]
 * Andersen-Broadie upper bound for the Bermudan option with exercise dates j dt, j = 0..n_steps, under GBM:
 *   1. exercise policy = the LSM regression fitted on n_policy_paths independent paths (seed ^ 1);
 *   2. n_outer fresh paths (seed, path_offset + i): lower bound = mean discounted payoff of following the policy;
 *   3. at every date of every outer path, n_inner inner paths continue from the outer state under the policy
 *      (normals keyed by (outer id, date, inner id, step block)) -> continuation value Q_j; martingale
 *      M_0 = 0, M_{j+1} = M_j + L_{j+1} - Q_j with L_j = policy value at j; upper = mean max_j (h_j - M_j), all discounted.
 * The reference has no such algorithm (its MartingaleOptimization pricer is a polynomial fit), and its rough-vol
 * driver is not adapted, so conditional inner simulation is only defined for the GBM model here: parity is unpinned by
 * the reference; the test oracle is an independent restatement of this description. */
typedef struct mcp_dual_result {
    double lower, lower_se, upper, upper_se;
    int64_t n_outer_global;
    float policy_ms, outer_ms, nested_ms;
} mcp_dual_result;
int mcp_gbm_nested_dual(mcp_ctx *ctx, const mcp_gbm_params *model, double strike, int is_call, int n_steps,
                        int poly_order, int64_t n_policy_paths, int64_t n_outer, int n_inner, uint64_t seed,
                        uint64_t path_offset, mcp_dual_result *out);

/* ---------------------------------------------------------------- batched strike x maturity surface (config 5)
 * prices[m][k] (row-major [n_maturities][n_strikes]) = LSM price of strike k at maturity m under the rough-vol model:
 * one path slab of n_paths x floor(T_m * steps_per_year) steps per maturity (PredictionGen.cpp:718), shared by all
 * strikes.  Only maturities mat_first, mat_first + mat_stride, ... are priced (the others' entries are left alone):
 * contracts are independent, so ranks split maturities with no collective.  lsm_tmpl supplies r, is_call,
 * poly_order, basis, carry; its strike / maturity / dt are ignored (dt = model->dt). */
int mcp_price_surface_rbergomi_lsm(mcp_ctx *ctx, const mcp_rbergomi_params *model, const mcp_lsm_params *lsm_tmpl,
                                   const double *strikes, int n_strikes, const double *maturities, int n_maturities,
                                   int steps_per_year, int64_t n_paths, uint64_t seed, uint64_t path_offset,
                                   int mat_first, int mat_stride, double *prices, double *std_errors /*nullable*/,
                                   float *gen_ms_total /*nullable*/, float *lsm_ms_total /*nullable*/);

/* ------------------------------------------------------------------------ batched row driver
 * The reference's row loop (src/core/PredictionGen.cpp:542-866) as one call: for every row, n_paths (<= 4096; the
 * reference uses 250, :719) rough-vol paths of row.n_steps steps (<= 512) are generated and priced by all four
 * pricers with the reference's per-row settings (exercise dates 0 .. n_steps-1, :780-783).  Three kernel launches per
 * batch instead of several per row and pricer.  Paths of row k are keyed by (seed, path_offset + k * n_paths + i), i.e.
 * the paths mcp_gen_rbergomi(.., seed, path_offset + k * n_paths) generates.  A row whose model is degenerate (H < 0 -- the
 * reference's DFA slope is unclamped --, |rho| > 1, xi < 0, dt <= 0, NaN) gets NaN in all five outputs, which is what the
 * reference's NaN paths lead to for that row (PredictionGen.cpp:753-777); the rest of the batch is priced normally. */
typedef struct mcp_row {
    mcp_rbergomi_params model; /* from mcp_estimate_rbergomi_params(history) or explicit */
    int n_steps;               /* floor(maturity * 252) in the reference (:718); rows with n_steps < 1 yield zeros (:720-733) */
    int is_call;
    double r, strike, maturity, dt, sigma, dividend;
} mcp_row;
typedef struct mcp_row_result {
    double asymptotic, branching, lsm, martingale; /* the four columns PredictionGen appends (:809-815) */
    double lsm_std_error;
} mcp_row_result;
int mcp_price_rows(mcp_ctx *ctx, const mcp_row *rows, int n_rows, int n_paths, int poly_order, int num_branches,
                   int max_iterations, uint64_t seed, uint64_t path_offset, mcp_row_result *out,
                   float *gen_ms /*nullable*/, float *price_ms /*nullable*/);

/* ------------------------------------------------------------- exact-signature generator
 * Host estimators of the reference (pure host arithmetic, no device needed): xi = var(logret)/dt, H = DFA slope
 * (unclamped), eta = 2 std(logret), rho = corr(ret, ret^2) or -0.3 when positive, r = 0.04, dt = 1/252,
 * S0 = hist[n-1].  MCP_ERR_DOMAIN + "Historical prices vector too small." for n_hist < 2 (RoughVolatility.cpp:317-319). */
int mcp_estimate_rbergomi_params(const double *hist, int64_t n_hist, mcp_rbergomi_params *out);
/* rows[path][0..forward_steps] (caller-allocated) = GenerateStockPricePaths(hist, forward_steps, path_num) with native
 * Philox normals keyed by (seed, path_offset + path). */
int mcp_generate_stock_price_paths(mcp_ctx *ctx, const double *hist, int64_t n_hist, int forward_steps, int path_num,
                                   uint64_t seed, uint64_t path_offset, double *const *rows);

/* ------------------------------------------------------------------- the other three plugins
 * Same arguments as the reference methods after `pricePaths`; with a communicator the sums are global. */
int mcp_asymptotic_price(mcp_ctx *ctx, const mcp_pathset *ps, double r, double strike, double maturity, double dt,
                         int is_call, double sigma, double dividend, double *price);
int mcp_martingale_price(mcp_ctx *ctx, const mcp_pathset *ps, double r, double strike, double maturity, double dt,
                         int is_call, int poly_order, int max_iterations, double *price, double *primal /*nullable*/,
                         double *dual /*nullable*/);
/* exercise_times must be strictly increasing.  The upper bound resamples paths with Philox (seed, path_offset) or,
 * when injected_rp != NULL, with the caller's indices [visited exercise date][path][branch] (parity runs). */
int mcp_branching_price(mcp_ctx *ctx, const mcp_pathset *ps, double r, double strike, double maturity, double dt,
                        int is_call, int num_branches, const int *exercise_times, int n_exercise, uint64_t seed,
                        uint64_t path_offset, const int32_t *injected_rp, double *price, double *lower /*nullable*/,
                        double *upper /*nullable*/);

#ifdef __cplusplus
}
#endif
#endif /* MCP_B200_H */
