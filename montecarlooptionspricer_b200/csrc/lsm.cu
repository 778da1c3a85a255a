// lsm.cu -- Longstaff-Schwartz backward induction (the reference's value-iteration variant) on the device.
//
// Replaces LSM::PredictOptionPrice (src/models/LSMPricer.cpp:19-102):
//   V_{M-1} = payoff(S_{M-1})                                              :37-40
//   for j = M-2 .. 0:                                                      :42
//     j dt > maturity      -> V_j = e^{-r dt} V_{j+1}                      :43-49
//     ITM = { payoff(S_j) > 1e-14 }                                        :51-58
//     c_j = argmin || [1,S,..,S^p] c - e^{-r dt} V_{j+1} ||  over ITM      :61-76  (Eigen bdcSvd, min-norm)
//     ITM: V_j = max(payoff, basis(S_j) . c_j)                             :78-86
//     payoff < 1e-14: V_j = e^{-r dt} V_{j+1}                              :89-94
//   price = mean_i V_0[i]                                                  :97-101
//
// B200 design.  The reference materialises Values[N][M] doubles and, per step, an index list, a design
// matrix and a thin SVD -- all strided over a path-major vector<vector>.  Here only a carry vector V lives
// next to the time-major slab, and each time step is ONE streaming kernel:
//   sweep(j):  read S_j, S_{j-1}, V;  apply the step-j decision with the already solved c_j;  write V;  and in
//              the same pass accumulate the normal-equation moments of step j-1 from the V_j just produced
//              (fused continuation/exercise update + discounted carry + next regression's X^T X, X^T y).
//   solve(j-1): the LAST CTA of the same launch to finish (atomic ticket) folds the per-CTA fp64 partials in a
//              fixed order (bitwise reproducible), all-reduces them across GPUs when a communicator is attached
//              (the ONLY data-path collective: 3p+2 doubles, exchanged in-kernel through NVLink peer mailboxes
//              or, as the fallback, by ncclAllReduce between two launches), and solves the (p+1)x(p+1) system.
//              So a time step is exactly one launch; programmatic dependent launch overlaps step j-1's
//              prologue with step j's tail.
// Kernels in this file: the fp64-decision parity kernel (lsm_sweep_kernel), the throughput kernels
// (lsm_sweep_fast2_kernel: direct 256-bit loads; lsm_sweep_tma_kernel: persistent CTAs fed by a TMA bulk-copy
// ring), the single-CTA small-problem kernel (lsm_small_kernel) and the multi-contract kernel (lsm_multi_kernel:
// one pass over the slab prices up to 16 strikes).
// Regression numerics.  X^T X in raw monomials of S~100 is singular in fp64 (cond ~ 4e17 at config 1), so
// moments are accumulated in the standardised variable x = (S - mu_j) / s_j (mu_j, s_j = mean / std of the
// in-the-money prices of a fixed leading sample of paths).  Any basis of the same polynomial space yields the
// same fitted values at the regression points, which is all the reference evaluates (LSMPricer.cpp:82-85); the
// Hankel structure needs only sum x^k (k <= 2p) and sum x^k y (k <= p).  Rank-deficient steps (j = 0: all paths
// at S0; fewer ITM paths than basis functions) reproduce the min-norm / projection behaviour through a Jacobi
// eigen-decomposition with a relative cut, the normal-equation image of Eigen's singular-value threshold.
// All decisions are taken in fp64 on the stored path values, exactly like the reference.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>
#include <vector>

#include "common.cuh"
#include "lsm_solve.cuh"
#include "small_bodies.cuh"

// ---------------------------------------------------------------------------------------------------------
// MCP_DEBUG_BOUNDS build (python -m montecarlooptionspricer_b200.build --debug -> libmcp_b200_dbg.so): compute-sanitizer is
// closed on the development pool, so the asynchronous parts -- bulk-copy ring, cross-step cursors, exchange rows --
// carry their own index checks.  Every check that fails bumps a counter (class below) instead of corrupting memory
// silently; tests/test_gpu_debug_bounds.py runs ragged sizes through every sweep kernel and requires all counters zero.
// In a normal build the macros vanish.
// ---------------------------------------------------------------------------------------------------------
enum McpDbgClass { DBG_CARRY_STORE = 0, DBG_TAU_STORE = 1, DBG_RING_ISSUE = 2, DBG_RING_STAGE = 3, DBG_RING_CANARY = 4, DBG_XCHG_ROW = 5, DBG_NCLASS = 8 };
#ifdef MCP_DEBUG_BOUNDS
__device__ unsigned int g_dbg_violations[DBG_NCLASS];
#define MCP_DBG_CHECK(cond, cls)                                  \
    do {                                                          \
        if (!(cond)) atomicAdd(&g_dbg_violations[(cls)], 1u);     \
    } while (0)
constexpr int MCP_DBG_CANARY_BYTES = 128;  // between the last ring stage and the barriers
constexpr unsigned int MCP_DBG_CANARY = 0xC0FFEE11u;
#else
#define MCP_DBG_CHECK(cond, cls) \
    do {                         \
    } while (0)
constexpr int MCP_DBG_CANARY_BYTES = 0;
#endif

extern "C" int mcp_debug_violations(unsigned int* out, int n) {
#ifdef MCP_DEBUG_BOUNDS
    unsigned int v[DBG_NCLASS];
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpyFromSymbol(v, g_dbg_violations, sizeof(v)) != cudaSuccess) return -2;
    for (int i = 0; i < n && i < DBG_NCLASS; ++i) out[i] = v[i];
    return DBG_NCLASS;
#else
    (void)out; (void)n;
    return 0;  // not a debug build
#endif
}

namespace {

constexpr int LSM_NT = 256;
constexpr int MOM_LD = 24;       // moment row stride (3p+2 <= 20)
constexpr int SAMPLE_MAX = 16384;

enum StepKind { STEP_NORMAL = 0, STEP_DISCOUNT = 1 };

struct LsmDev {  // pointers into ctx scratch
    double* coef;     // [M][COEF_LD]
    double* mu;       // [M]
    double* inv_s;    // [M]
    double* ssum;     // [M][4]   sample sums: cnt, sum S, sum S^2
    double* partial;  // [max_blocks][MOM_LD]
    double* moments;  // [MOM_LD]
    double* fin;      // [4]      sum V0, sum V0^2, n
    int* kind;        // [M]
    unsigned int* counter;  // CTA completion ticket for the last-block epilogue (self-resetting)
};

// Constants of a contract (not of a step), computed ONCE on the host (or once per launch on the device): as kernel
// arguments they sit in the constant bank and cost no register -- and no per-tile re-derivation from the doubles.
struct StepK {
    float sg, nsK, nsKlo;  // payoff = max((sg * S + nsK) + nsKlo, 0): strike as a two-float sum     include/core/common.h:8-14
    float ne1;             // -(1 - e^{-r dt}): V e^{-r dt} = fma(V, ne1, V), ONE operation, correctly rounded for a factor that is
                           // within 1.2e-11 of the true one (1 - e^{-r dt} ~ 2e-4 keeps 24 bits of its own)
    float thrS;            // in the money (payoff > 1e-14, LSMPricer.cpp:55)  <=>  (S > thrS) != flip   (exactly, see make_stepk)
    int flip;              // 1 for puts
};

__host__ __device__ inline float stepk_payoff(float u /* = sg * S */, float nsK, float nsKlo) {
#ifdef __CUDA_ARCH__
    const float t = __fadd_rn(__fadd_rn(u, nsK), nsKlo);
#else
    volatile float t0 = u + nsK;  // the device's FFMA with sg = +-1 rounds exactly this sum once
    volatile float t1 = t0 + nsKlo;
    const float t = t1;
#endif
    return t > 0.f ? t : 0.f;
}

// The payoff as the kernels compute it is a monotone function of u = sg * S, so { payoff > 1e-14f } = { u > thr } for one
// float thr: bisection over the ordered bit patterns finds it exactly, and the regression's in-the-money test of a path
// (taken one pass BEFORE its exercise decision) is a single compare that agrees bit for bit with the decision's own test.
__host__ __device__ inline StepK make_stepk(double K, double disc, int is_call) {
    StepK g;
    const float sgn = is_call ? 1.f : -1.f;
    const float K_hi = (float)K, K_lo = (float)(K - (double)K_hi);
    g.sg = sgn;
    g.nsK = -sgn * K_hi;
    g.nsKlo = -sgn * K_lo;
    g.ne1 = -(float)(1.0 - disc);
    auto key2f = [](uint32_t key) -> float {  // order-preserving map of [0, 2^32) onto the floats
        const uint32_t b = (key & 0x80000000u) ? (key ^ 0x80000000u) : ~key;
#ifdef __CUDA_ARCH__
        return __uint_as_float(b);
#else
        float f;
        memcpy(&f, &b, 4);
        return f;
#endif
    };
    uint32_t lo = 0x00800000u, hi = 0xff7fffffu;  // keys of -FLT_MAX (never in the money) and +FLT_MAX (always)
    while (hi - lo > 1u) {
        const uint32_t mid = lo + (hi - lo) / 2u;
        if (stepk_payoff(key2f(mid), g.nsK, g.nsKlo) > 1e-14f) hi = mid; else lo = mid;
    }
    // lo = key of the largest u = sg * S that is NOT in the money.  Calls (u = S): ITM <=> S > u_lo.  Puts (u = -S): ITM <=>
    // -S > u_lo <=> S < -u_lo <=> not (S > pred(-u_lo)), and pred(-u) = -(succ(u)) = -key2f(hi).
    g.flip = is_call ? 0 : 1;
    g.thrS = is_call ? key2f(lo) : -key2f(hi);
    return g;
}

struct SweepArgs {
    const void* S;   // slab
    int64_t ld, n;   // row stride, paths
    void* V;         // carry
    int32_t* tau;    // first exercise index (nullable)
    LsmDev d;
    double K, disc;
    int is_call;
    int j;           // step being decided
    int terminal;    // j == M-1: V = payoff
    int do_moments;  // accumulate moments of step j-1
    int do_final;    // j == 0: accumulate sum V0
    int solve_here;  // single GPU: the last CTA to finish also solves step j-1 (no extra launches)
    McpXchg x;          // peer-memory mailboxes (multi-GPU): the moment all-reduce happens inside this kernel
    unsigned long long seq;  // sequence number of this launch's exchange
    int l2_resident;    // two slab rows + the carry fit in L2: keep them there instead of streaming
    StepK g;            // contract constants of the packed-fp32 kernels (make_stepk)
    int ref_rank;       // apply the reference's SVD rank cut in the solve (parity mode, and always for p >= 5; lsm_solve.cuh)
};

// Block-wide deterministic sum of NV doubles per thread -> row `blockIdx.x` of `partial`.
template <int NV, int NT = LSM_NT>
__device__ __forceinline__ void block_reduce_to_partial(double (&acc)[NV], double* __restrict__ partial_row) {
    __shared__ double red[NT / 32][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const double s = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) s += red[w][threadIdx.x];
        partial_row[threadIdx.x] = s;
    }
}

struct SweepArgs;
template <int NV, int P, int NT = 256>
__device__ __forceinline__ void sweep_epilogue(const SweepArgs& a, double (&acc)[NV]);

template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, double (&o)[4]) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        o[0] = f2d(v.x); o[1] = f2d(v.y); o[2] = f2d(v.z); o[3] = f2d(v.w);
    }
    static __device__ __forceinline__ void store_round(float* p, double (&io)[4]) {  // stores and returns the stored values
        float4 v = make_float4((float)io[0], (float)io[1], (float)io[2], (float)io[3]);
        *reinterpret_cast<float4*>(p) = v;
        io[0] = f2d(v.x); io[1] = f2d(v.y); io[2] = f2d(v.z); io[3] = f2d(v.w);
    }
};
template <>
struct Vec4<double> {
    static __device__ __forceinline__ void load(const double* p, double (&o)[4]) {
        const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
        o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
    }
    static __device__ __forceinline__ void store_round(double* p, double (&io)[4]) {
        *reinterpret_cast<double2*>(p) = make_double2(io[0], io[1]);
        *reinterpret_cast<double2*>(p + 2) = make_double2(io[2], io[3]);
    }
};

// Vector index -> first path of the 4-wide vector.  Odd steps walk the slab backwards so that what the
// previous launch touched last (still resident in the 126 MB L2: S_{j-1} and V) is what this launch reads first.
__device__ __forceinline__ int64_t vec_base(int64_t idx, int64_t nvec, int j) { return ((j & 1) ? (nvec - 1 - idx) : idx) * 4; }

// ---------------------------------------------------------------------------------------------------------
// sweep(j), PARITY kernel: every decision and every moment in fp64 on the stored values, exactly the
// reference's arithmetic.  ST = slab storage, CT = carry storage, P = poly order.
// ---------------------------------------------------------------------------------------------------------
template <typename ST, typename CT, int P>
__global__ void __launch_bounds__(LSM_NT, 3) lsm_sweep_kernel(SweepArgs a) {
    constexpr int NM = 3 * P + 2;  // s[0..2P], t[0..P]
    constexpr int NV = NM > 2 ? NM : 2;
    const ST* __restrict__ Sj = reinterpret_cast<const ST*>(a.S) + (int64_t)a.j * a.ld;
    const ST* __restrict__ Sp = reinterpret_cast<const ST*>(a.S) + (int64_t)(a.j > 0 ? a.j - 1 : 0) * a.ld;
    CT* __restrict__ V = reinterpret_cast<CT*>(a.V);

    const int mode = a.terminal ? 2 : a.d.kind[a.j];  // 0 regress/decide, 1 discount only, 2 terminal payoff
    double c[P + 1];
#pragma unroll
    for (int k = 0; k <= P; ++k) c[k] = a.d.coef[(int64_t)a.j * COEF_LD + k];
    const double mu = a.d.mu[a.j], inv_s = a.d.inv_s[a.j];
    const double mu_p = a.d.mu[a.j > 0 ? a.j - 1 : 0], inv_s_p = a.d.inv_s[a.j > 0 ? a.j - 1 : 0];

    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
    int cnt = 0;

    const int64_t nvec = (a.n + 3) >> 2, vstride = (int64_t)gridDim.x * LSM_NT;
    for (int64_t iv = (int64_t)blockIdx.x * LSM_NT + threadIdx.x; iv < nvec; iv += vstride) {
        const int64_t i = vec_base(iv, nvec, a.j);
        const int nvalid = (int)(a.n - i < 4 ? a.n - i : 4);
        double s[4], sp[4], v[4];
        Vec4<ST>::load(Sj + i, s);
        if (a.do_moments) Vec4<ST>::load(Sp + i, sp);
        if (mode != 2) Vec4<CT>::load(V + i, v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const double pay = payoff_fn(a.is_call, s[e], a.K);
            if (mode == 2) {
                v[e] = pay;  // LSMPricer.cpp:37-40
            } else if (mode == 1) {
                v[e] = v[e] * a.disc;  // LSMPricer.cpp:43-49
            } else {
                const double x = (s[e] - mu) * inv_s;
                double cont = c[P];
#pragma unroll
                for (int k = P - 1; k >= 0; --k) cont = fma(cont, x, c[k]);
                const bool itm = pay > 1e-14;   // LSMPricer.cpp:55
                const bool ex = !(pay < cont);  // std::max(immediate, cont) returns immediate (LSMPricer.cpp:85)
                const double carried = pay < 1e-14 ? v[e] * a.disc : 0.0;  // LSMPricer.cpp:89-94; == 1e-14 keeps the initial 0 (:35)
                v[e] = itm ? (ex ? pay : cont) : carried;
                if (a.tau && itm && ex && e < nvalid) a.tau[i + e] = a.j;
            }
        }
        Vec4<CT>::store_round(V + i, v);  // v[] now holds exactly what the next step will read
        if (a.do_moments) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (e < nvalid && payoff_fn(a.is_call, sp[e], a.K) > 1e-14) {  // LSMPricer.cpp:51-58 for step j-1
                    const double x = (sp[e] - mu_p) * inv_s_p, y = v[e] * a.disc;  // LSMPricer.cpp:69
                    double xp = x;
                    ++cnt;
                    acc[2 * P + 1] += y;
#pragma unroll
                    for (int k = 1; k <= 2 * P; ++k) {
                        acc[k] += xp;
                        if (k <= P) acc[2 * P + 1 + k] = fma(xp, y, acc[2 * P + 1 + k]);
                        if (k < 2 * P) xp *= x;
                    }
                }
            }
        }
        if (a.do_final) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (e < nvalid) acc[0] += v[e];
        }
    }
    if (a.do_moments) acc[0] = (double)cnt;
    if (a.do_moments || a.do_final) sweep_epilogue<NV, P>(a, acc);
}

// ---------------------------------------------------------------------------------------------------------
// sweep(j), THROUGHPUT arithmetic (fp32 slab + fp32 carry), shared by the direct-load kernel below and the TMA-ring
// kernel further down.  All per-path arithmetic in fp32, only the cross-path accumulation in fp64:
//   * packed fp32x2 math (FFMA2 / FADD2 / FMUL2, one issue slot per two paths): the first scalar version issued ~77
//     instructions per path and ran at 63% issue utilisation with DRAM only 56% busy -- as much instruction-bound as
//     bandwidth-bound;
//   * the in-the-money filter of the moments is a 0/1 multiplier on (x, y) instead of a divergent branch
//     (x = 0 kills every power, y = 0 every cross moment);
//   * fp32 partial sums of <= 64 paths per lane are folded into fp64 accumulators that live in shared memory;
//   * the discount is applied as a two-float product (d_hi + d_lo) so that 252 chained roundings stay unbiased; the
//     strike is split the same way.  Prices agree with the fp64 oracle to ~1e-7 relative (tolerance 1e-5); exercise
//     indices may differ from it at near-ties only (use the parity kernel when they must not).
// L2 pinning of a carry prefix (evict_last hints, and a persisting access-policy window) was measured and gave nothing;
// it is not in the code any more.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 splat2(float a) { return make_float2(a, a); }

// 256-bit global accesses with an L2 eviction priority (sm_100a: LDG/STG.E.256 with .EFL2 / .ENL2 / .ELL2).
struct F8 {
    float2 q[4];
};
#define MCP_LD8(NAME, MOD)                                                                                                     \
    __device__ __forceinline__ F8 NAME(const float* p) {                                                                       \
        uint32_t r[8];                                                                                                         \
        asm volatile("ld.global" MOD ".v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                                                \
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])          \
                     : "l"(p));                                                                                                \
        F8 o;                                                                                                                  \
        _Pragma("unroll") for (int i = 0; i < 4; ++i) o.q[i] = make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])); \
        return o;                                                                                                              \
    }
#define MCP_ST8(NAME, MOD)                                                                                                     \
    __device__ __forceinline__ void NAME(float* p, const F8& o) {                                                              \
        asm volatile("st.global" MOD ".v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(__float_as_uint(o.q[0].x)),      \
                     "r"(__float_as_uint(o.q[0].y)), "r"(__float_as_uint(o.q[1].x)), "r"(__float_as_uint(o.q[1].y)),           \
                     "r"(__float_as_uint(o.q[2].x)), "r"(__float_as_uint(o.q[2].y)), "r"(__float_as_uint(o.q[3].x)),           \
                     "r"(__float_as_uint(o.q[3].y))                                                                            \
                     : "memory");                                                                                              \
    }
MCP_LD8(ld8_stream, ".L1::no_allocate.L2::evict_first")  // touched once per sweep
MCP_LD8(ld8_keep, ".L1::no_allocate.L2::evict_normal")   // S_{j-1}: read again by the next sweep
MCP_ST8(st8_stream, ".L2::evict_first")

template <int P>
struct FastConsts {
    float2 c[P + 1];
    float2 is, c0, is_p, c0_p;  // x = S * is + c0 with is = 1/s, c0 = -mu/s, of step j and of step j-1
};

// Arithmetic of one group of 8 paths held in registers: s8 = S_j, p8 = S_{j-1}, v8 = carry (updated in place).
// Path e of the group is global path idx[e >> 2] + (e & 3), i.e. two runs of four consecutive paths (the direct-load
// kernel uses one run of eight: idx1 = idx0 + 4).  TAIL: the ragged last group, paths >= a.n are masked out.
// KIND 0: the common launch (regress-and-decide step that also accumulates the next regression, 251 of 253 launches
// at config 3) with every mode test resolved at compile time; KIND 1: any launch, flags read at run time.
// la[0] counts the in-the-money paths of a regression step and sums V0 on the last step (`cnt` is a spare integer counter).
// The FP32 pipe is what bounds these kernels once the data sits in L2 (ncu, profiles/r02b): per pair of paths 23
// packed operations (payoff 2, discount 2, standardise 1, Horner P, mask 1, standardise 1, discount 1, powers P - 1,
// sums 3P + 1); masks, selects and the count run on the ALU pipe (FSETP / FSEL / predicated add).
template <int P, bool TAU, bool TAIL, int KIND = 1>
__device__ __forceinline__ void fast2_compute(const SweepArgs& a, const StepK& g, const FastConsts<P>& k, const F8& s8, const F8& p8, F8& v8, int64_t idx0,
                                              int64_t idx1, int mode_rt, float2 (&la)[(3 * P + 2) > 2 ? (3 * P + 2) : 2], int& cnt) {
    const int mode = KIND == 0 ? 0 : mode_rt;
    const bool do_moments = KIND == 0 ? true : (a.do_moments != 0), do_final = KIND == 0 ? false : (a.do_final != 0);
    float2 s[4], sp[4];
    float2(&v)[4] = v8.q;
    bool ok[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) ok[e] = !TAIL || ((e < 4 ? idx0 : idx1) + (e & 3) < a.n);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        s[q] = s8.q[q];
        sp[q] = p8.q[q];
        if (TAIL) {
            // lanes past the last path hold whatever the padding / a stale ring stage held (possibly NaN or Inf bit
            // patterns, and NaN * 0 = NaN would poison the step's moments): replace them by zeros BEFORE any arithmetic
            if (!ok[2 * q]) { s[q].x = 0.f; sp[q].x = 0.f; v[q].x = 0.f; }
            if (!ok[2 * q + 1]) { s[q].y = 0.f; sp[q].y = 0.f; v[q].y = 0.f; }
        }
    }
    const float2 sg = splat2(g.sg), nsK = splat2(g.nsK), nsKlo = splat2(g.nsKlo), ne1 = splat2(g.ne1);
    const bool flip = g.flip != 0, klo = g.nsKlo != 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float2 pay = __ffma2_rn(s[q], sg, nsK);  // include/core/common.h:8-14
        if (klo) pay = __fadd2_rn(pay, nsKlo);   // (strikes that are exact in fp32 have no low word)
        pay.x = fmaxf(pay.x, 0.f);
        pay.y = fmaxf(pay.y, 0.f);
        if (mode == 2) {
            v[q] = pay;  // LSMPricer.cpp:37-40
        } else {
            const float2 vd = __ffma2_rn(v[q], ne1, v[q]);
            if (mode == 1) {
                v[q] = vd;  // LSMPricer.cpp:43-49
            } else {
                const float2 x = __ffma2_rn(s[q], k.is, k.c0);
                float2 cont = k.c[P];
#pragma unroll
                for (int m = P - 1; m >= 0; --m) cont = __ffma2_rn(cont, x, k.c[m]);
                // ITM: V = max(payoff, fitted)  (:78-86);  otherwise the discounted carry (:89-94).  (The reference leaves V = 0
                // when the payoff EQUALS 1e-14; the fp32 payoff is 0 or >= one ulp of the strike's low word, never that.)
                v[q].x = pay.x > 1e-14f ? fmaxf(pay.x, cont.x) : vd.x;
                v[q].y = pay.y > 1e-14f ? fmaxf(pay.y, cont.y) : vd.y;
                if (TAU) {
                    const int64_t ib = (q < 2 ? idx0 : idx1) + 2 * (q & 1);
                    MCP_DBG_CHECK(!ok[2 * q] || (ib >= 0 && ib < a.n), DBG_TAU_STORE);
                    MCP_DBG_CHECK(!ok[2 * q + 1] || (ib + 1 >= 0 && ib + 1 < a.n), DBG_TAU_STORE);
                    if (pay.x > 1e-14f && !(pay.x < cont.x) && ok[2 * q]) a.tau[ib] = a.j;
                    if (pay.y > 1e-14f && !(pay.y < cont.y) && ok[2 * q + 1]) a.tau[ib + 1] = a.j;
                }
            }
        }
    }
    if (do_moments) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            // LSMPricer.cpp:51-58 for step j-1: one compare per path on the ALU pipe (the FP32 pipe is the busy one)
            const bool in_x = ((sp[q].x > g.thrS) != flip) && ok[2 * q], in_y = ((sp[q].y > g.thrS) != flip) && ok[2 * q + 1];
            float2 x = __ffma2_rn(sp[q], k.is_p, k.c0_p);
            float2 y = __ffma2_rn(v[q], ne1, v[q]);  // LSMPricer.cpp:69
            x.x = in_x ? x.x : 0.f;  x.y = in_y ? x.y : 0.f;   // x = 0 kills every power, y = 0 every cross moment
            y.x = in_x ? y.x : 0.f;  y.y = in_y ? y.y : 0.f;
            la[0] = __fadd2_rn(la[0], make_float2(in_x ? 1.f : 0.f, in_y ? 1.f : 0.f));
            // power sums without forming the high powers: x^e = x^ceil(e/2) * x^floor(e/2) goes straight into the FMA that
            // accumulates it (P - 1 multiplies + 3P - 1 FMAs + 2 adds per pair)
            float2 xp[P + 1];
            xp[0] = x;
            if (P >= 1) xp[1] = x;
#pragma unroll
            for (int e = 2; e <= P; ++e) xp[e] = __fmul2_rn(xp[e - 1], x);
            la[2 * P + 1] = __fadd2_rn(la[2 * P + 1], y);
            if (P >= 1) la[1] = __fadd2_rn(la[1], x);
#pragma unroll
            for (int e = 2; e <= 2 * P; ++e) la[e] = __ffma2_rn(xp[(e + 1) / 2], xp[e / 2], la[e]);
#pragma unroll
            for (int e = 1; e <= P; ++e) la[2 * P + 1 + e] = __ffma2_rn(xp[e], y, la[2 * P + 1 + e]);
        }
    }
    if (do_final) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float2 w = v[q];
            if (TAIL) {
                if (!ok[2 * q]) w.x = 0.f;
                if (!ok[2 * q + 1]) w.y = 0.f;
            }
            la[0] = __fadd2_rn(la[0], w);
        }
    }
}

// Direct-load form: one group = 8 consecutive paths, 256-bit loads / stores with L2 eviction priorities.
template <int P, bool TAU, bool TAIL>
__device__ __forceinline__ void fast2_group(const SweepArgs& a, const FastConsts<P>& k, const float* __restrict__ Sj, const float* __restrict__ Sp,
                                            float* __restrict__ V, int64_t i0, int mode, float2 (&la)[(3 * P + 2) > 2 ? (3 * P + 2) : 2], int& cnt) {
    const F8 s8 = ld8_stream(Sj + i0);
    F8 p8 = s8, v8 = s8;
    if (a.do_moments) p8 = ld8_keep(Sp + i0);
    if (mode != 2) v8 = ld8_stream(V + i0);
    fast2_compute<P, TAU, TAIL>(a, a.g, k, s8, p8, v8, i0, i0 + 4, mode, la, cnt);
    st8_stream(V + i0, v8);
}

template <int P>
__device__ __forceinline__ void fast2_load_consts(const SweepArgs& a, FastConsts<P>& k) {
#pragma unroll
    for (int m = 0; m <= P; ++m) k.c[m] = splat2((float)a.d.coef[(int64_t)a.j * COEF_LD + m]);
    const int jp = a.j > 0 ? a.j - 1 : 0;
    const double is = a.d.inv_s[a.j], is_p = a.d.inv_s[jp];
    k.is = splat2((float)is);
    k.c0 = splat2((float)(-a.d.mu[a.j] * is));
    k.is_p = splat2((float)is_p);
    k.c0_p = splat2((float)(-a.d.mu[jp] * is_p));
}

template <int P, bool TAU>
__global__ void __launch_bounds__(LSM_NT, 3) lsm_sweep_fast2_kernel(SweepArgs a) {
    constexpr int NM = 3 * P + 2;
    constexpr int NV = NM > 2 ? NM : 2;
    constexpr int FLUSH = 8;  // groups (x8 paths) between fp32 -> fp64 folds
    __shared__ double sacc[NV][LSM_NT];
    const float* __restrict__ Sj = reinterpret_cast<const float*>(a.S) + (int64_t)a.j * a.ld;
    const float* __restrict__ Sp = reinterpret_cast<const float*>(a.S) + (int64_t)(a.j > 0 ? a.j - 1 : 0) * a.ld;
    float* __restrict__ V = reinterpret_cast<float*>(a.V);
    const int mode = a.terminal ? 2 : a.d.kind[a.j];

    FastConsts<P> k;
    fast2_load_consts<P>(a, k);

    float2 la[NV];
#pragma unroll
    for (int m = 0; m < NV; ++m) { la[m] = make_float2(0.f, 0.f); sacc[m][threadIdx.x] = 0.0; }
    int since = 0, cnt = 0;

    const int64_t ngroup = (a.n + 7) >> 3, nfull = a.n >> 3, gstride = (int64_t)gridDim.x * LSM_NT;
    for (int64_t ig = (int64_t)blockIdx.x * LSM_NT + threadIdx.x; ig < ngroup; ig += gstride) {
        const int64_t g = (a.j & 1) ? (ngroup - 1 - ig) : ig;  // serpentine: what the previous sweep touched last is read first
        if (g < nfull) fast2_group<P, TAU, false>(a, k, Sj, Sp, V, g * 8, mode, la, cnt);
        else fast2_group<P, TAU, true>(a, k, Sj, Sp, V, g * 8, mode, la, cnt);
        if (++since == FLUSH) {
#pragma unroll
            for (int m = 0; m < NV; ++m) {
                sacc[m][threadIdx.x] += (double)la[m].x + (double)la[m].y;
                la[m] = make_float2(0.f, 0.f);
            }
            since = 0;
        }
    }
    if (a.do_moments || a.do_final) {
        double acc[NV];
#pragma unroll
        for (int m = 0; m < NV; ++m) acc[m] = sacc[m][threadIdx.x] + ((double)la[m].x + (double)la[m].y);
        acc[0] += (double)cnt;
        sweep_epilogue<NV, P>(a, acc);
    }
}

// ---------------------------------------------------------------------------------------------------------
// sweep(j), THROUGHPUT kernel v3: the same arithmetic fed by a TMA ring.
// ncu on v2 (profiles/r01f): 56% of the stall samples are long-scoreboard waits and DRAM is 65% busy -- with 24
// resident warps per SM the direct loads cannot keep enough bytes in flight.  Here one persistent CTA per SM
// (512 threads) streams tiles of 4096 paths through a ring of shared-memory stages filled by bulk async copies
// (cp.async.bulk / UBLKCP, completion on an mbarrier): memory-level parallelism is set by the ring depth
// (up to 3 x 48 KB in flight per SM), not by registers or occupancy.  S_j and V tiles are fetched with an
// L2 evict-first policy (touched once per sweep), S_{j-1} with the default policy (it is next sweep's S_j).
// Thread t owns paths [4t, 4t+4) and [2048+4t, 2048+4t+4) of a tile: conflict-free LDS.128, coalesced STG.128.
// ---------------------------------------------------------------------------------------------------------
constexpr int TMA_NT = 512;
constexpr int TMA_TILE = 4096;                                  // paths per tile
constexpr int TMA_STAGE_BYTES = 3 * TMA_TILE * 4;               // S_j | S_{j-1} | V

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// blocking wait with a suspend-time hint: the thread sleeps in hardware until the phase flips (or the hint expires) instead
// of spinning through issue slots that the other warps of the scheduler could use
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(1000000u)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void stg4_keep(float* p, float2 a, float2 b) {
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y) : "memory");
}
__device__ __forceinline__ void stg4_stream(float* p, float2 a, float2 b) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y) : "memory");
}

template <int P, bool TAU>
__global__ void __launch_bounds__(TMA_NT, 1) lsm_sweep_tma_kernel(SweepArgs a, int n_stages) {
    constexpr int NM = 3 * P + 2;
    constexpr int NV = NM > 2 ? NM : 2;
    constexpr int FLUSH = 8;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);
    unsigned char* after_ring = smem_raw + (size_t)n_stages * TMA_STAGE_BYTES + MCP_DBG_CANARY_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(after_ring);  // "stage filled" barriers (TMA completes them)
    uint64_t* empty = full + 8;                                // "stage consumed" barriers (one arrival per warp)
    double* sacc = reinterpret_cast<double*>(after_ring + 128);  // [NV][TMA_NT]
#ifdef MCP_DEBUG_BOUNDS
    if (threadIdx.x < MCP_DBG_CANARY_BYTES / 4) reinterpret_cast<unsigned int*>(after_ring - MCP_DBG_CANARY_BYTES)[threadIdx.x] = MCP_DBG_CANARY;
#endif
    const float* __restrict__ Sj = reinterpret_cast<const float*>(a.S) + (int64_t)a.j * a.ld;
    const float* __restrict__ Sp = reinterpret_cast<const float*>(a.S) + (int64_t)(a.j > 0 ? a.j - 1 : 0) * a.ld;
    float* __restrict__ V = reinterpret_cast<float*>(a.V);
    const int tid = threadIdx.x;

    // Programmatic dependent launch: this grid may start while the previous sweep is still in its epilogue.  Everything
    // up to griddepcontrol.wait touches only what the previous sweep never writes (the path slab, our own shared
    // memory): the barriers are set up and the S_j / S_{j-1} parts of the first ring stages are already in flight
    // when the wait returns; the carry tiles and the regression coefficients are fetched after it.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int64_t ntile = (a.n + TMA_TILE - 1) / TMA_TILE;
    const int64_t my_tiles = blockIdx.x < ntile ? (ntile - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    // streamed tiles leave L2 first -- unless the whole working set (two slab rows + the carry) fits in L2, where the
    // next sweep finds S_{j-1} (as its S_j) and the carry still resident (multi-GPU shards of config 3)
    const uint64_t pol = a.l2_resident ? l2_policy_evict_normal() : l2_policy_evict_first();
    const bool want_v = !a.terminal;  // mode != 2
    // tile `it` of this CTA, in serpentine order (what the previous sweep touched last is read first)
    auto tile_of = [&](int64_t it) -> int64_t {
        const int64_t t = (int64_t)blockIdx.x + it * gridDim.x;
        return (a.j & 1) ? (ntile - 1 - t) : t;
    };
    // one elected thread: arm the stage's barrier with the bytes of all its copies, start the slab copies (part 1) and
    // the carry copy (part 2)
    auto issue_at = [&](int64_t it, int st, bool part_s, bool part_v) {
        const int64_t i0 = tile_of(it) * TMA_TILE;
        const int64_t cnt = a.ld - i0 < TMA_TILE ? a.ld - i0 : TMA_TILE;  // rows are padded to ld (multiple of 128)
        const uint32_t bytes = (uint32_t)cnt * 4u;
        float* dst = ring + (size_t)st * (3 * TMA_TILE);
        MCP_DBG_CHECK(st >= 0 && st < n_stages && it >= 0 && it < my_tiles, DBG_RING_STAGE);
        MCP_DBG_CHECK(i0 >= 0 && cnt > 0 && cnt <= TMA_TILE && i0 + cnt <= a.ld && (bytes & 15u) == 0u, DBG_RING_ISSUE);
        if (part_s) {
            mbar_expect_tx(full + st, bytes * (1u + (a.do_moments ? 1u : 0u) + (want_v ? 1u : 0u)));
            bulk_g2s_hint(dst, Sj + i0, bytes, full + st, pol);
            if (a.do_moments) bulk_g2s(dst + TMA_TILE, Sp + i0, bytes, full + st);
        }
        if (part_v && want_v) bulk_g2s_hint(dst + 2 * TMA_TILE, V + i0, bytes, full + st, pol);
    };
    if (tid == 0) {
        for (int st = 0; st < n_stages; ++st) { mbar_init(full + st, 1); mbar_init(empty + st, TMA_NT / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int64_t it = 0; it < my_tiles && it < n_stages; ++it) issue_at(it, (int)it, true, false);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the previous sweep (carry, coefficients, ticket) is complete and visible
    const int mode = a.terminal ? 2 : a.d.kind[a.j];
    if (tid == 0)
        for (int64_t it = 0; it < my_tiles && it < n_stages; ++it) issue_at(it, (int)it, false, true);
    FastConsts<P> k;
    fast2_load_consts<P>(a, k);
    float2 la[NV];
#pragma unroll
    for (int m = 0; m < NV; ++m) { la[m] = make_float2(0.f, 0.f); sacc[m * TMA_NT + tid] = 0.0; }
    __syncthreads();

    int since = 0, cnt = 0;
    auto run_tiles = [&](auto kind_tag) {
    constexpr int KIND = decltype(kind_tag)::value;
    const bool dm = KIND == 0 ? true : (a.do_moments != 0), wv = KIND == 0 ? true : (mode != 2);
    int st = 0;
    uint32_t parity = 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
        while (!mbar_try_wait(full + st, parity)) {}
        const float* buf = ring + (size_t)st * (3 * TMA_TILE);
        F8 s8, p8, v8;
        {
            const float4 x0 = *reinterpret_cast<const float4*>(buf + 4 * tid), x1 = *reinterpret_cast<const float4*>(buf + 2048 + 4 * tid);
            s8.q[0] = make_float2(x0.x, x0.y); s8.q[1] = make_float2(x0.z, x0.w); s8.q[2] = make_float2(x1.x, x1.y); s8.q[3] = make_float2(x1.z, x1.w);
        }
        p8 = s8; v8 = s8;
        if (dm) {
            const float4 x0 = *reinterpret_cast<const float4*>(buf + TMA_TILE + 4 * tid), x1 = *reinterpret_cast<const float4*>(buf + TMA_TILE + 2048 + 4 * tid);
            p8.q[0] = make_float2(x0.x, x0.y); p8.q[1] = make_float2(x0.z, x0.w); p8.q[2] = make_float2(x1.x, x1.y); p8.q[3] = make_float2(x1.z, x1.w);
        }
        if (wv) {
            const float4 x0 = *reinterpret_cast<const float4*>(buf + 2 * TMA_TILE + 4 * tid), x1 = *reinterpret_cast<const float4*>(buf + 2 * TMA_TILE + 2048 + 4 * tid);
            v8.q[0] = make_float2(x0.x, x0.y); v8.q[1] = make_float2(x0.z, x0.w); v8.q[2] = make_float2(x1.x, x1.y); v8.q[3] = make_float2(x1.z, x1.w);
        }
        // this warp holds its part of the stage in registers; once all 16 warps have said so the slot is refilled.  No
        // block-wide barrier: warps drift apart by up to the ring depth
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(empty + st);
        if (tid == 0 && it + n_stages < my_tiles) {
            while (!mbar_try_wait(empty + st, parity)) {}
            issue_at(it + n_stages, st, true, true);
        }
        if (++st == n_stages) { st = 0; parity ^= 1u; }

        const int64_t i0 = tile_of(it) * TMA_TILE, ia = i0 + 4 * tid, ib = i0 + 2048 + 4 * tid;
        const bool whole = i0 + TMA_TILE <= a.n;
        if (whole) fast2_compute<P, TAU, false, KIND>(a, a.g, k, s8, p8, v8, ia, ib, mode, la, cnt);
        else fast2_compute<P, TAU, true, KIND>(a, a.g, k, s8, p8, v8, ia, ib, mode, la, cnt);
        MCP_DBG_CHECK(ia >= 0 && (!(whole || ia < a.ld) || ia + 4 <= a.ld) && (!(whole || ib < a.ld) || ib + 4 <= a.ld), DBG_CARRY_STORE);
        if (a.l2_resident) {
            if (whole || ia < a.ld) stg4_keep(V + ia, v8.q[0], v8.q[1]);
            if (whole || ib < a.ld) stg4_keep(V + ib, v8.q[2], v8.q[3]);
        } else {
            if (whole || ia < a.ld) stg4_stream(V + ia, v8.q[0], v8.q[1]);
            if (whole || ib < a.ld) stg4_stream(V + ib, v8.q[2], v8.q[3]);
        }
        if (++since == FLUSH) {
#pragma unroll
            for (int m = 0; m < NV; ++m) {
                sacc[m * TMA_NT + tid] += (double)(la[m].x + la[m].y);
                la[m] = make_float2(0.f, 0.f);
            }
            since = 0;
        }
    }
    };
    if (mode == 0 && a.do_moments && !a.do_final) run_tiles(std::integral_constant<int, 0>{});
    else run_tiles(std::integral_constant<int, 1>{});
#ifdef MCP_DEBUG_BOUNDS
    if (tid < MCP_DBG_CANARY_BYTES / 4) MCP_DBG_CHECK(reinterpret_cast<unsigned int*>(after_ring - MCP_DBG_CANARY_BYTES)[tid] == MCP_DBG_CANARY, DBG_RING_CANARY);
#endif
    if (a.do_moments || a.do_final) {
        double acc[NV];
#pragma unroll
        for (int m = 0; m < NV; ++m) acc[m] = sacc[m * TMA_NT + tid] + (double)(la[m].x + la[m].y);
        acc[0] += (double)cnt;
        sweep_epilogue<NV, P, TMA_NT>(a, acc);
    }
}

// ---------------------------------------------------------------------------------------------------------
// sweep(j), PARITY arithmetic on the TMA ring (fp32 slab, fp64 carry): every decision and every moment in fp64 on the
// stored values -- the arithmetic of lsm_sweep_kernel, bit for bit per path -- but fed like the throughput kernel: one
// persistent 512-thread CTA per SM, tiles of S_j | S_{j-1} (fp32) | V (fp64) = 64 KB per 4096 paths through a three-stage
// ring of bulk async copies, programmatic dependent launch.  The grid-stride parity kernel ran at 0.48 of the 20 B/path-step
// roofline (long-scoreboard stalls: not enough bytes in flight at 24 warps/SM); the DFMA work (~25 per path-step) is a
// third of the HBM time on B200, so this kernel is HBM-bound.  Moments stay in per-thread fp64 registers.
// ---------------------------------------------------------------------------------------------------------
constexpr int TMA64_STAGE_BYTES = (2 * 4 + 8) * TMA_TILE;  // S_j | S_{j-1} | V(double)

__device__ __forceinline__ void stg2d_stream(double* p, double a, double b) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

template <int P>
__global__ void __launch_bounds__(TMA_NT, 1) lsm_sweep_tma64_kernel(SweepArgs a, int n_stages) {
    constexpr int NM = 3 * P + 2;
    constexpr int NV = NM > 2 ? NM : 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)n_stages * TMA64_STAGE_BYTES);
    uint64_t* empty = full + 8;
    const float* __restrict__ Sj = reinterpret_cast<const float*>(a.S) + (int64_t)a.j * a.ld;
    const float* __restrict__ Sp = reinterpret_cast<const float*>(a.S) + (int64_t)(a.j > 0 ? a.j - 1 : 0) * a.ld;
    double* __restrict__ V = reinterpret_cast<double*>(a.V);
    const int tid = threadIdx.x;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int64_t ntile = (a.n + TMA_TILE - 1) / TMA_TILE;
    const int64_t my_tiles = blockIdx.x < ntile ? (ntile - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const uint64_t pol = a.l2_resident ? l2_policy_evict_normal() : l2_policy_evict_first();
    const bool want_v = !a.terminal;
    auto tile_of = [&](int64_t it) -> int64_t {
        const int64_t t = (int64_t)blockIdx.x + it * gridDim.x;
        return (a.j & 1) ? (ntile - 1 - t) : t;
    };
    auto issue_at = [&](int64_t it, int st, bool part_s, bool part_v) {
        const int64_t i0 = tile_of(it) * TMA_TILE;
        const int64_t cnt = a.ld - i0 < TMA_TILE ? a.ld - i0 : TMA_TILE;
        const uint32_t bytes = (uint32_t)cnt * 4u;
        unsigned char* dst = smem_raw + (size_t)st * TMA64_STAGE_BYTES;
        if (part_s) {
            mbar_expect_tx(full + st, bytes * (1u + (a.do_moments ? 1u : 0u) + (want_v ? 2u : 0u)));
            bulk_g2s_hint(dst, Sj + i0, bytes, full + st, pol);
            if (a.do_moments) bulk_g2s(dst + 4 * TMA_TILE, Sp + i0, bytes, full + st);
        }
        if (part_v && want_v) bulk_g2s_hint(dst + 8 * TMA_TILE, V + i0, 2u * bytes, full + st, pol);
    };
    if (tid == 0) {
        for (int st = 0; st < n_stages; ++st) { mbar_init(full + st, 1); mbar_init(empty + st, TMA_NT / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int64_t it = 0; it < my_tiles && it < n_stages; ++it) issue_at(it, (int)it, true, false);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int mode = a.terminal ? 2 : a.d.kind[a.j];  // 0 regress/decide, 1 discount only, 2 terminal payoff
    if (tid == 0)
        for (int64_t it = 0; it < my_tiles && it < n_stages; ++it) issue_at(it, (int)it, false, true);
    double c[P + 1];
#pragma unroll
    for (int k = 0; k <= P; ++k) c[k] = a.d.coef[(int64_t)a.j * COEF_LD + k];
    const double mu = a.d.mu[a.j], inv_s = a.d.inv_s[a.j];
    const double mu_p = a.d.mu[a.j > 0 ? a.j - 1 : 0], inv_s_p = a.d.inv_s[a.j > 0 ? a.j - 1 : 0];
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
    int cnt = 0;
    __syncthreads();

    int st = 0;
    uint32_t parity = 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
        while (!mbar_try_wait(full + st, parity)) {}
        const unsigned char* buf = smem_raw + (size_t)st * TMA64_STAGE_BYTES;
        // thread t owns paths [4t, 4t+4) and [2048 + 4t, 2048 + 4t + 4) of the tile
        float4 sA = *reinterpret_cast<const float4*>(buf + 16 * tid), sB = *reinterpret_cast<const float4*>(buf + 8192 + 16 * tid);
        float4 pA = sA, pB = sB;
        if (a.do_moments) {
            pA = *reinterpret_cast<const float4*>(buf + 4 * TMA_TILE + 16 * tid);
            pB = *reinterpret_cast<const float4*>(buf + 4 * TMA_TILE + 8192 + 16 * tid);
        }
        double v[8];
        if (mode != 2) {
            const double2* vb = reinterpret_cast<const double2*>(buf + 8 * TMA_TILE);
            const double2 v0 = vb[2 * tid], v1 = vb[2 * tid + 1], v2 = vb[1024 + 2 * tid], v3 = vb[1024 + 2 * tid + 1];
            v[0] = v0.x; v[1] = v0.y; v[2] = v1.x; v[3] = v1.y; v[4] = v2.x; v[5] = v2.y; v[6] = v3.x; v[7] = v3.y;
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(empty + st);
        if (tid == 0 && it + n_stages < my_tiles) {
            while (!mbar_try_wait(empty + st, parity)) {}
            issue_at(it + n_stages, st, true, true);
        }
        if (++st == n_stages) { st = 0; parity ^= 1u; }

        const int64_t i0 = tile_of(it) * TMA_TILE, ia = i0 + 4 * tid, ib = ia + 2048;
        const float sv[8] = {sA.x, sA.y, sA.z, sA.w, sB.x, sB.y, sB.z, sB.w};
        const float pv[8] = {pA.x, pA.y, pA.z, pA.w, pB.x, pB.y, pB.z, pB.w};
        // The ALU pipe (selects, 64-bit integer compares, software float->double widening) was what bound the first version of
        // this kernel (ncu profiles/r02d: 133 instructions per path, ALU 68%, issue 70%, DRAM 73% of the copy peak), not the
        // fp64 pipe (33 operations per path): whole tiles therefore run without any bounds arithmetic, the widening is the
        // hardware conversion, and the in-the-money filter of the moments is a 0/1 multiplier on the fp64 pipe.
        auto body = [&](auto tail_tag) {
            constexpr bool TAILT = decltype(tail_tag)::value;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const bool live = !TAILT || ((e < 4 ? ia : ib) + (e & 3) < a.n);
                const double s = (double)sv[e];
                const double pay = payoff_fn(a.is_call, s, a.K);
                if (mode == 2) {
                    v[e] = pay;  // LSMPricer.cpp:37-40
                } else if (mode == 1) {
                    v[e] = v[e] * a.disc;  // LSMPricer.cpp:43-49
                } else {
                    const double x = (s - mu) * inv_s;
                    double cont = c[P];
#pragma unroll
                    for (int k = P - 1; k >= 0; --k) cont = fma(cont, x, c[k]);
                    const bool itm = pay > 1e-14;   // LSMPricer.cpp:55
                    const bool ex = !(pay < cont);  // std::max(immediate, cont) returns immediate (LSMPricer.cpp:85)
                    const double carried = pay < 1e-14 ? v[e] * a.disc : 0.0;  // LSMPricer.cpp:89-94; == 1e-14 keeps the initial 0 (:35)
                    v[e] = itm ? (ex ? pay : cont) : carried;
                    if (a.tau && itm && ex && live) a.tau[(e < 4 ? ia : ib) + (e & 3)] = a.j;
                }
                if (TAILT && !live) v[e] = 0.0;  // pad lanes: never NaN / Inf in the carry
            }
            if (!TAILT || ia < a.ld) { stg2d_stream(V + ia, v[0], v[1]); stg2d_stream(V + ia + 2, v[2], v[3]); }
            if (!TAILT || ib < a.ld) { stg2d_stream(V + ib, v[4], v[5]); stg2d_stream(V + ib + 2, v[6], v[7]); }
            if (a.do_moments) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const bool live = !TAILT || ((e < 4 ? ia : ib) + (e & 3) < a.n);
                    const double sp = live ? (double)pv[e] : 0.0;
                    const bool in = live && payoff_fn(a.is_call, sp, a.K) > 1e-14;  // LSMPricer.cpp:51-58 for step j-1
                    const double msk = in ? 1.0 : 0.0;
                    const double x = ((sp - mu_p) * inv_s_p) * msk, y = (v[e] * a.disc) * msk;  // LSMPricer.cpp:69; 0 kills every moment
                    cnt += in ? 1 : 0;
                    // power sums as FMA products: x^k = x^ceil(k/2) * x^floor(k/2) goes straight into the accumulating DFMA
                    double xp[P + 1];
                    xp[0] = x;
                    if (P >= 1) xp[1] = x;
#pragma unroll
                    for (int k = 2; k <= P; ++k) xp[k] = xp[k - 1] * x;
                    acc[2 * P + 1] += y;
                    if (P >= 1) acc[1] += x;
#pragma unroll
                    for (int k = 2; k <= 2 * P; ++k) acc[k] = fma(xp[(k + 1) / 2], xp[k / 2], acc[k]);
#pragma unroll
                    for (int k = 1; k <= P; ++k) acc[2 * P + 1 + k] = fma(xp[k], y, acc[2 * P + 1 + k]);
                }
            }
            if (a.do_final) {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (!TAILT || (e < 4 ? ia : ib) + (e & 3) < a.n) acc[0] += v[e];
            }
        };
        if (i0 + TMA_TILE <= a.n) body(std::false_type{});
        else body(std::true_type{});
    }
    if (a.do_moments) acc[0] = (double)cnt;
    if (a.do_moments || a.do_final) sweep_epilogue<NV, P, TMA_NT>(a, acc);
}

typedef void (*SweepFn2Fwd)(SweepArgs, int);
inline SweepFn2Fwd pick_sweep_tma64(int p) {
    switch (p) {
        case 0: return lsm_sweep_tma64_kernel<0>;
        case 1: return lsm_sweep_tma64_kernel<1>;
        case 2: return lsm_sweep_tma64_kernel<2>;
        case 3: return lsm_sweep_tma64_kernel<3>;
        case 4: return lsm_sweep_tma64_kernel<4>;
        case 5: return lsm_sweep_tma64_kernel<5>;
        default: return lsm_sweep_tma64_kernel<6>;
    }
}

// ---------------------------------------------------------------------------------------------------------
// sweep(j) for SEVERAL CONTRACTS ON THE SAME PATHS (strike ladder of a surface, BASELINE config 5): one warp per
// contract.  Per path-step that is 8 B of slab shared by the whole ladder + 8 B of carry per contract instead of 16 B per
// contract, and the per-launch fixed cost (launch, fold, solve) is paid once for the ladder.  Same packed fp32 arithmetic
// and the same per-contract standardisation as the single-contract throughput kernels.
//
// Data movement (r02; the first version moved whole 18-array stages of 72 KB, two of which fit: 4.8 TB/s):
//   * slab tiles S_j | S_{j-1} (1024 paths) stream through a deep shared ring, read by all sixteen contract warps from
//     shared memory; a stage is refilled by WHICHEVER WARP LEAVES IT LAST (a shared-memory ticket per stage), so nobody
//     ever waits for a hand-back;
//   * every contract warp owns a PRIVATE ring of carry tiles, which it refills itself the moment it has its tile in
//     registers -- no warp ever waits for another warp's carry, and a stage is held for half a tile of arithmetic only.
// Warps drift apart by up to the slab ring's depth; nothing in the tile loop is block-wide.
// The last CTA folds the per-CTA moment rows of every contract in a fixed order; warp c solves contract c.
// ---------------------------------------------------------------------------------------------------------
constexpr int MULTI_MAXC = 16;                 // contracts per launch = consumer warps per CTA
constexpr int MULTI_NT = MULTI_MAXC * 32;
#ifndef MCP_MULTI_TILE
#define MCP_MULTI_TILE 1024
#define MCP_MULTI_NS 4
#endif
constexpr int MULTI_TILE = MCP_MULTI_TILE;     // paths per tile
constexpr int MULTI_NS = MCP_MULTI_NS;         // slab ring depth (tiles)
constexpr int MULTI_SLAB_STAGE_BYTES = 2 * MULTI_TILE * 4;
constexpr int MULTI_CARRY_SLOT_BYTES = MULTI_TILE * 4;
constexpr int MULTI_MAXD = 8;                  // carry ring depth per warp (upper bound; the host picks what fits)
constexpr int MULTI_BAR_BYTES = (2 * MULTI_NS + MULTI_MAXC * MULTI_MAXD) * 8;  // fullS | tickets | fullV
constexpr int MULTI_ACC_LANES = MULTI_MAXC * 8;  // fp64 accumulator columns: one per group of four lanes

struct MultiArgs {
    const float* S;
    int64_t ld, n;
    float* V;            // [C][ld] carries
    double* coef;        // [C][M][COEF_LD]
    const double* mu;    // [C][M]  per-contract standardisation (mean / 1/std of the contract's in-the-money sample)
    const double* inv_s; // [C][M]
    double* partial;     // [grid][C][MOM_LD]
    double* fin;         // [C][4]: sum V0
    const int* kind;     // [M]
    unsigned int* counter;
    double K[MULTI_MAXC];
    double disc;
    int C, M, is_call, j, terminal, do_moments, do_final, ref_rank;
};

template <int P>
__global__ void __launch_bounds__(MULTI_NT, 1) lsm_multi_kernel(MultiArgs a, int depth) {
    constexpr int NM = 3 * P + 2;
    constexpr int NV = NM > 2 ? NM : 2;
    constexpr int FLUSH_TILES = 8 / (MULTI_TILE / 256) > 0 ? 8 / (MULTI_TILE / 256) : 1;  // fp32 partials cover <= 64 paths per lane
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* slab = reinterpret_cast<float*>(smem_raw);                                                   // [MULTI_NS][2][TILE]
    float* carry = reinterpret_cast<float*>(smem_raw + MULTI_NS * MULTI_SLAB_STAGE_BYTES);              // [MAXC][depth][TILE]
    unsigned char* after = smem_raw + MULTI_NS * MULTI_SLAB_STAGE_BYTES + (size_t)MULTI_MAXC * depth * MULTI_CARRY_SLOT_BYTES + MCP_DBG_CANARY_BYTES;
    uint64_t* fullS = reinterpret_cast<uint64_t*>(after);   // slab stage filled (bulk copies complete it)
    unsigned int* ticket = reinterpret_cast<unsigned int*>(fullS + MULTI_NS);  // warps that have left the stage (monotone; 16 per use)
    uint64_t* fullV = fullS + 2 * MULTI_NS;                  // [MAXC][MAXD] carry slot filled
    double* sacc = reinterpret_cast<double*>(after + MULTI_BAR_BYTES);  // [NV][MULTI_ACC_LANES]
#ifdef MCP_DEBUG_BOUNDS
    if (threadIdx.x < MCP_DBG_CANARY_BYTES / 4) reinterpret_cast<unsigned int*>(after - MCP_DBG_CANARY_BYTES)[threadIdx.x] = MCP_DBG_CANARY;
#endif
    const int tid = threadIdx.x, lane = tid & 31, c = tid >> 5;  // warp = contract
    const bool active = c < a.C;
    const int cc = active ? c : 0;
    const float* __restrict__ Sj = a.S + (int64_t)a.j * a.ld;
    const float* __restrict__ Sp = a.S + (int64_t)(a.j > 0 ? a.j - 1 : 0) * a.ld;
    float* __restrict__ V = a.V + (int64_t)cc * a.ld;
    const int mode = a.terminal ? 2 : a.kind[a.j];
    const bool want_v = mode != 2;

    const int64_t ntile = (a.n + MULTI_TILE - 1) / MULTI_TILE;
    const int my_tiles = blockIdx.x < ntile ? (int)((ntile - 1 - blockIdx.x) / gridDim.x + 1) : 0;
    const uint64_t pol = l2_policy_evict_first();
    // first path of tile `it` of this CTA (serpentine over launches); all ring bookkeeping below is 32-bit counters --
    // the loop runs once per 512 paths per warp, a 64-bit division there is a visible share of the instruction stream
    const int64_t t_first = (a.j & 1) ? (ntile - 1 - (int64_t)blockIdx.x) : (int64_t)blockIdx.x;
    const int64_t t_step = (a.j & 1) ? -(int64_t)gridDim.x : (int64_t)gridDim.x;
    auto tile_of = [&](int it) -> int64_t { return t_first + (int64_t)it * t_step; };
    auto tile_bytes = [&](int64_t i0) -> uint32_t {
        const int64_t cnt = a.ld - i0 < MULTI_TILE ? a.ld - i0 : MULTI_TILE;  // rows are padded to ld (multiple of 128)
        MCP_DBG_CHECK(i0 >= 0 && cnt > 0 && cnt <= MULTI_TILE && i0 + cnt <= a.ld && ((cnt * 4) & 15) == 0, DBG_RING_ISSUE);
        return (uint32_t)cnt * 4u;
    };
    auto issue_slab = [&](int it, int st) {  // one elected lane
        const int64_t i0 = tile_of(it) * MULTI_TILE;
        const uint32_t bytes = tile_bytes(i0);
        float* dst = slab + (size_t)st * (2 * MULTI_TILE);
        MCP_DBG_CHECK(it >= 0 && it < my_tiles, DBG_RING_STAGE);
        mbar_expect_tx(fullS + st, bytes * (a.do_moments ? 2u : 1u));
        bulk_g2s_hint(dst, Sj + i0, bytes, fullS + st, pol);
        if (a.do_moments) bulk_g2s(dst + MULTI_TILE, Sp + i0, bytes, fullS + st);  // read again as the next launch's S_j: default policy
    };
    auto issue_carry = [&](int it, int d) {  // lane 0 of an active contract warp: its own carry tile into its own ring
        const int64_t i0 = tile_of(it) * MULTI_TILE;
        const uint32_t bytes = tile_bytes(i0);
        MCP_DBG_CHECK(it >= 0 && it < my_tiles && d < MULTI_MAXD && a.C >= 1 && a.C <= MULTI_MAXC, DBG_RING_STAGE);
        mbar_expect_tx(fullV + c * MULTI_MAXD + d, bytes);
        bulk_g2s_hint(carry + ((size_t)c * depth + d) * MULTI_TILE, V + i0, bytes, fullV + c * MULTI_MAXD + d, pol);
    };
    if (tid == 0) {
        for (int st = 0; st < MULTI_NS; ++st) { mbar_init(fullS + st, 1); ticket[st] = 0u; }
        for (int q = 0; q < MULTI_MAXC * MULTI_MAXD; ++q) mbar_init(fullV + q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int it = 0; it < my_tiles && it < MULTI_NS; ++it) issue_slab(it, it);
    }
    for (int q = tid; q < NV * MULTI_ACC_LANES; q += MULTI_NT) sacc[q] = 0.0;
    __syncthreads();

    // lane 0 of a warp that has everything it needs from slab stage `st` in registers: the sixteenth warp to say so refills it
    // (release / acquire on the ticket orders the sixteen warps' reads of the stage before the refill is issued; like the
    // mbarrier hand-back of the single-contract ring, no proxy fence is needed for a read-then-bulk-write hand-over)
    auto leave_stage = [&](int it, int st) {
        unsigned int old;
        asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(ticket + st)) : "memory");
        if ((old & (MULTI_MAXC - 1)) == MULTI_MAXC - 1 && it + MULTI_NS < my_tiles) issue_slab(it + MULTI_NS, st);
    };
    {
        // per-warp constants: this contract's strike, coefficients and standardisation
        SweepArgs w;  // the view fast2_compute / fast2_load_consts expect
        memset(&w, 0, sizeof(w));
        w.n = a.n; w.j = a.j; w.do_moments = a.do_moments; w.do_final = a.do_final; w.is_call = a.is_call; w.disc = a.disc;
        w.K = a.K[cc];
        w.d.coef = a.coef + (int64_t)cc * a.M * COEF_LD;
        w.d.mu = const_cast<double*>(a.mu) + (int64_t)cc * a.M;
        w.d.inv_s = const_cast<double*>(a.inv_s) + (int64_t)cc * a.M;
        w.g = make_stepk(w.K, w.disc, w.is_call);  // per warp = per contract, once per launch
        FastConsts<P> k;
        fast2_load_consts<P>(w, k);
        float2 la[NV];
        int cnt = 0, since = 0;
#pragma unroll
        for (int m = 0; m < NV; ++m) la[m] = make_float2(0.f, 0.f);
        if (active && want_v && lane == 0)
            for (int it = 0; it < my_tiles && it < depth; ++it) issue_carry(it, it);

        int st = 0, d = 0;
        uint32_t parS = 0, parV = 0;
        const bool common = mode == 0 && a.do_moments && !a.do_final;
        for (int it = 0; it < my_tiles; ++it) {
            mbar_wait(fullS + st, parS);
            if (active) {
                if (want_v) mbar_wait(fullV + c * MULTI_MAXD + d, parV);
                const float* buf = slab + (size_t)st * (2 * MULTI_TILE);
                const float* vbuf = carry + ((size_t)c * depth + d) * MULTI_TILE;
                const int64_t i0 = tile_of(it) * MULTI_TILE;
                // one tile = MULTI_TILE / 256 straight-line iterations of 256 paths (lane -> 4 + 4 paths); the common step
                // (decision + moments, KIND 0) and whole tiles are compile-time cases, so nothing in the body branches
                auto tile_body = [&](auto kind_tag, auto tail_tag) {
                    constexpr int KIND = decltype(kind_tag)::value;
                    constexpr bool TAIL = decltype(tail_tag)::value;
                    const bool dm = KIND == 0 ? true : (a.do_moments != 0), wv = KIND == 0 ? true : want_v;
#pragma unroll
                    for (int h = 0; h < MULTI_TILE / 256; ++h) {
                        const int oa = h * 256 + 4 * lane, ob = oa + 128;
                        const int64_t ia = i0 + oa, ib = i0 + ob;
                        F8 s8, p8, v8;
                        {
                            const float4 x0 = *reinterpret_cast<const float4*>(buf + oa), x1 = *reinterpret_cast<const float4*>(buf + ob);
                            s8.q[0] = make_float2(x0.x, x0.y); s8.q[1] = make_float2(x0.z, x0.w); s8.q[2] = make_float2(x1.x, x1.y); s8.q[3] = make_float2(x1.z, x1.w);
                        }
                        p8 = s8; v8 = s8;
                        if (dm) {
                            const float4 x0 = *reinterpret_cast<const float4*>(buf + MULTI_TILE + oa), x1 = *reinterpret_cast<const float4*>(buf + MULTI_TILE + ob);
                            p8.q[0] = make_float2(x0.x, x0.y); p8.q[1] = make_float2(x0.z, x0.w); p8.q[2] = make_float2(x1.x, x1.y); p8.q[3] = make_float2(x1.z, x1.w);
                        }
                        if (wv) {
                            const float4 x0 = *reinterpret_cast<const float4*>(vbuf + oa), x1 = *reinterpret_cast<const float4*>(vbuf + ob);
                            v8.q[0] = make_float2(x0.x, x0.y); v8.q[1] = make_float2(x0.z, x0.w); v8.q[2] = make_float2(x1.x, x1.y); v8.q[3] = make_float2(x1.z, x1.w);
                        }
                        if (h == MULTI_TILE / 256 - 1) {
                            // the warp holds the rest of its tile in registers: hand the slab stage back and refill the own carry slot
                            __syncwarp();
                            if (lane == 0) {
                                leave_stage(it, st);
                                if (wv && it + depth < my_tiles) issue_carry(it + depth, d);
                            }
                        }
                        fast2_compute<P, false, TAIL, KIND>(w, w.g, k, s8, p8, v8, ia, ib, mode, la, cnt);
                        MCP_DBG_CHECK(ia >= 0 && (ia >= a.ld || ia + 4 <= a.ld) && (ib >= a.ld || ib + 4 <= a.ld), DBG_CARRY_STORE);
                        if (!TAIL || ia < a.ld) stg4_stream(V + ia, v8.q[0], v8.q[1]);
                        if (!TAIL || ib < a.ld) stg4_stream(V + ib, v8.q[2], v8.q[3]);
                    }
                };
                const bool whole = i0 + MULTI_TILE <= a.n;
                if (common) {
                    if (whole) tile_body(std::integral_constant<int, 0>{}, std::false_type{});
                    else tile_body(std::integral_constant<int, 0>{}, std::true_type{});
                } else {
                    if (whole) tile_body(std::integral_constant<int, 1>{}, std::false_type{});
                    else tile_body(std::integral_constant<int, 1>{}, std::true_type{});
                }
                if (++since == FLUSH_TILES) {  // fp32 partials of <= 64 paths per lane, four lanes folded, then fp64
#pragma unroll
                    for (int m = 0; m < NV; ++m) {
                        float v = la[m].x + la[m].y;
                        v += __shfl_xor_sync(0xffffffffu, v, 1);
                        v += __shfl_xor_sync(0xffffffffu, v, 2);
                        if ((lane & 3) == 0) sacc[m * MULTI_ACC_LANES + c * 8 + (lane >> 2)] += (double)v;
                        la[m] = make_float2(0.f, 0.f);
                    }
                    since = 0;
                }
            } else if (lane == 0) {
                leave_stage(it, st);  // an idle warp (fewer than 16 strikes) only keeps the slab ring turning
            }
            if (++st == MULTI_NS) { st = 0; parS ^= 1u; }
            if (++d == depth) { d = 0; parV ^= 1u; }
        }
        if (a.do_moments || a.do_final) {
            // per-contract (per-warp) partial row of this CTA, in a fixed lane order
            double acc[NV];
#pragma unroll
            for (int m = 0; m < NV; ++m) acc[m] = ((lane & 3) == 0 ? sacc[m * MULTI_ACC_LANES + c * 8 + (lane >> 2)] : 0.0) + (double)(la[m].x + la[m].y);
            acc[0] += (double)cnt;
            double* prow = a.partial + ((int64_t)blockIdx.x * MULTI_MAXC + c) * MOM_LD;
#pragma unroll
            for (int m = 0; m < NV; ++m) {
                const double sres = warp_sum(acc[m]);
                if (lane == 0) prow[m] = sres;
            }
        }
    }
#ifdef MCP_DEBUG_BOUNDS
    __syncthreads();
    if (tid < MCP_DBG_CANARY_BYTES / 4) MCP_DBG_CHECK(reinterpret_cast<unsigned int*>(after - MCP_DBG_CANARY_BYTES)[tid] == MCP_DBG_CANARY, DBG_RING_CANARY);
#endif
    if (!(a.do_moments || a.do_final)) return;
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // warp c folds contract c: lane m < NV sums column m over the CTAs in order; lane 0 solves
    __shared__ double tot[MULTI_MAXC][32];
    if (active) {
        double sres = 0.0;
        if (lane < NV)
            for (int b = 0; b < (int)gridDim.x; ++b) sres += __ldcg(a.partial + ((int64_t)b * MULTI_MAXC + c) * MOM_LD + lane);
        tot[c][lane] = sres;
        __syncwarp();
        if (lane == 0) {
            if (a.do_final) a.fin[c * 4] = tot[c][0];
            else {
                const RefRank rr{a.mu[(int64_t)c * a.M + (a.j - 1)], a.inv_s[(int64_t)c * a.M + (a.j - 1)]};
                solve_normal_equations<P>(&tot[c][0], a.coef + ((int64_t)c * a.M + (a.j - 1)) * COEF_LD, a.ref_rank ? &rr : nullptr);
            }
        }
    }
    if (tid == 0) *a.counter = 0u;
}

typedef void (*MultiFn)(MultiArgs, int);
MultiFn pick_multi(int p) {
    switch (p) {
        case 0: return lsm_multi_kernel<0>;
        case 1: return lsm_multi_kernel<1>;
        case 2: return lsm_multi_kernel<2>;
        case 3: return lsm_multi_kernel<3>;
        case 4: return lsm_multi_kernel<4>;
        case 5: return lsm_multi_kernel<5>;
        default: return lsm_multi_kernel<6>;
    }
}

// Sum the per-CTA partial rows in a fixed order -> out[0..nv).
__global__ void __launch_bounds__(256) lsm_reduce_kernel(const double* __restrict__ partial, int nblocks, int nv, double* __restrict__ out) {
    __shared__ double red[8][32];
    const int k = threadIdx.x & 31, grp = threadIdx.x >> 5;
    double s = 0.0;
    if (k < nv)
        for (int b = grp; b < nblocks; b += 8) s += partial[(int64_t)b * MOM_LD + k];
    red[grp][k] = s;
    __syncthreads();
    if (threadIdx.x < nv) {
        double t = 0.0;
#pragma unroll
        for (int g = 0; g < 8; ++g) t += red[g][threadIdx.x];
        out[threadIdx.x] = t;
    }
}

// Multi-GPU path: after the NCCL all-reduce of the moments every rank solves the same tiny system.
__global__ void lsm_solve_kernel(const double* __restrict__ mom, int p, double* __restrict__ coef_row, const double* __restrict__ mu,
                                 const double* __restrict__ inv_s, int ref_rank) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const RefRank rr{*mu, *inv_s};
        solve_dispatch(mom, p, coef_row, ref_rank ? &rr : nullptr);
    }
}

// All-reduce (sum) of `vals[0..NV)` across GPUs through the peer-memory mailboxes, executed by ONE CTA per rank (the
// last one of its sweep launch).  One-shot all-gather in the style of NCCL's LL protocol: every 8-byte word carries
// 32 bits of payload and a 32-bit sequence tag, so one plain P2P store per word both delivers and publishes it --
// one NVLink traversal, no fence, no flag round trip.  Every rank stores its words into the mailbox of every rank,
// polls its own mailbox until all words carry this exchange's tag, and adds the rows IN RANK ORDER: all ranks obtain
// the bitwise identical sum and then solve the identical system.  Two mailbox halves (sequence parity) keep
// exchange s+1 from overwriting words a slow rank is still reading for exchange s.  A poll that exceeds ~20 s raises
// the error flag instead of hanging.
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <int NV, int NT>
__device__ __forceinline__ void xchg_allreduce(const McpXchg& x, unsigned long long seq, double* vals /* shared [NV] */) {
    static_assert(2 * NV <= MCP_XROW, "row too long for the mailbox");
    __shared__ unsigned int got[MCP_XMAX_RANKS][2 * NV];
    const int tid = threadIdx.x, n = x.nranks;
    const size_t half = (size_t)(seq & 1ull) * (size_t)n;
    const unsigned long long tag = (seq & 0xffffffffull) << 32;
    MCP_DBG_CHECK(n >= 1 && n <= MCP_XMAX_RANKS && x.rank >= 0 && x.rank < n && 2 * NV <= MCP_XROW, DBG_XCHG_ROW);
    for (int idx = tid; idx < n * 2 * NV; idx += NT) {
        const int r = idx / (2 * NV), w = idx - r * (2 * NV);
        const unsigned long long bits = (unsigned long long)__double_as_longlong(vals[w >> 1]);
        const unsigned long long word = ((w & 1) ? (bits >> 32) : (bits & 0xffffffffull)) | tag;
        volatile unsigned long long* row = reinterpret_cast<volatile unsigned long long*>(x.peer[r]) + (half + (size_t)x.rank) * MCP_XROW;
        row[w] = word;
    }
    for (int idx = tid; idx < n * 2 * NV; idx += NT) {
        const int r = idx / (2 * NV), w = idx - r * (2 * NV);
        const volatile unsigned long long* row = reinterpret_cast<const volatile unsigned long long*>(x.peer[x.rank]) + (half + (size_t)r) * MCP_XROW;
        const unsigned long long t0 = global_ns();
        unsigned long long word = row[w];
        while ((word & 0xffffffff00000000ull) != tag) {
            if (global_ns() - t0 > 20000000000ull) { *x.err = 1; break; }
            word = row[w];
        }
        got[r][w] = (unsigned int)word;
    }
    __syncthreads();
    if (tid < NV) {
        double s = 0.0;
        for (int r = 0; r < n; ++r)
            s += __longlong_as_double((long long)(((unsigned long long)got[r][2 * tid + 1] << 32) | (unsigned long long)got[r][2 * tid]));
        vals[tid] = s;
    }
    __syncthreads();
}

// Epilogue of a sweep launch: per-CTA partial row, then the LAST CTA to finish (completion ticket) folds all rows
// in a fixed order -- bitwise reproducible whichever CTA is last -- into `moments` (or the running sum of V0),
// all-reduces them across GPUs through the peer-memory mailboxes when those are up, and solves step j-1 right here,
// so one time step is exactly one kernel launch (on one GPU and on many).
template <int NV, int P, int NT>
__device__ __forceinline__ void sweep_epilogue(const SweepArgs& a, double (&acc)[NV]) {
    __shared__ double red[8][32];
    __shared__ bool is_last;
    block_reduce_to_partial<NV, NT>(acc, a.d.partial + (int64_t)blockIdx.x * MOM_LD);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(a.d.counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // fold gridDim.x rows in a FIXED order with as many independent loads in flight as possible: thread (seg, k) sums
    // a contiguous block of rows for moment k, then thread k adds the NSEG block sums in order
    constexpr int NSEG = NT / NV;
    __shared__ double seg_sum[NSEG][NV];
    const int rows = ((int)gridDim.x + NSEG - 1) / NSEG;
    if (threadIdx.x < NSEG * NV) {
        const int seg = threadIdx.x / NV, k = threadIdx.x - seg * NV;
        const int b0 = seg * rows, b1 = min(b0 + rows, (int)gridDim.x);
        double s = 0.0;
#pragma unroll 8
        for (int b = b0; b < b1; ++b) s += __ldcg(a.d.partial + (int64_t)b * MOM_LD + k);
        seg_sum[seg][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0.0;
#pragma unroll
        for (int g = 0; g < NSEG; ++g) t += seg_sum[g][threadIdx.x];
        red[0][threadIdx.x] = t;
    }
    if (a.do_final && threadIdx.x == 1) red[0][1] = (double)a.n;  // the final exchange carries {sum V0, N}
    __syncthreads();
    if (a.x.enabled) xchg_allreduce<NV, NT>(a.x, a.seq, &red[0][0]);
    if (a.do_final) {
        if (threadIdx.x == 0) { a.d.fin[0] = red[0][0]; if (a.x.enabled) a.d.fin[2] = red[0][1]; }
    } else if (threadIdx.x < NV) {
        a.d.moments[threadIdx.x] = red[0][threadIdx.x];
    }
    if (threadIdx.x == 0) {
        *a.d.counter = 0u;
        if (a.do_moments && a.solve_here) {
            const RefRank rr{a.d.mu[a.j - 1], a.d.inv_s[a.j - 1]};
            solve_normal_equations<P>(&red[0][0], a.d.coef + (int64_t)(a.j - 1) * COEF_LD, a.ref_rank ? &rr : nullptr);
        }
    }
}

#include "lsm_persist.cuh"  // the whole induction in one cooperative launch (uses the step arithmetic and bulk-copy helpers above)

typedef void (*PersistFn)(px::Args);
template <bool TAU>
PersistFn pick_persist(int p) {
    switch (p) {
        case 0: return px::lsm_persist_kernel<0, TAU>;
        case 1: return px::lsm_persist_kernel<1, TAU>;
        case 2: return px::lsm_persist_kernel<2, TAU>;
        case 3: return px::lsm_persist_kernel<3, TAU>;
        case 4: return px::lsm_persist_kernel<4, TAU>;
        case 5: return px::lsm_persist_kernel<5, TAU>;
        default: return px::lsm_persist_kernel<6, TAU>;
    }
}

// Sample statistics for the standardisation: CTA j sums cnt / S / S^2 over the in-the-money prices of the
// first `ns` paths of row j.
template <typename ST>
__global__ void __launch_bounds__(256) lsm_scale_sums_kernel(const ST* __restrict__ S, int64_t ld, int ns, double K, int is_call,
                                                            double* __restrict__ ssum) {
    const int j = blockIdx.x;
    double acc[3] = {0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < ns; i += 256) {
        const double s = (double)S[(int64_t)j * ld + i];
        if (payoff_fn(is_call, s, K) > 1e-14) { acc[0] += 1.0; acc[1] += s; acc[2] = fma(s, s, acc[2]); }
    }
    __shared__ double red[8][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < 3; ++k) {
        const double s = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        ssum[(int64_t)j * 4 + threadIdx.x] = s;
    }
}

__global__ void lsm_scale_finalize_kernel(const double* __restrict__ ssum, int M, double K, double* __restrict__ mu, double* __restrict__ inv_s) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    const double cnt = ssum[j * 4], s1 = ssum[j * 4 + 1], s2 = ssum[j * 4 + 2];
    double m = K, sd = fabs(K) > 0.0 ? fabs(K) : 1.0;
    if (cnt >= 2.0) {
        m = s1 / cnt;
        const double var = (s2 - cnt * m * m) / (cnt - 1.0);
        if (var > 1e-12 * m * m) sd = sqrt(var);
        else sd = fabs(m) > 0.0 ? fabs(m) : 1.0;
    }
    mu[j] = m;
    inv_s[j] = 1.0 / sd;
}

template <typename CT>
__global__ void lsm_copy_v0_kernel(const CT* __restrict__ V, int64_t n, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)V[i];
}

// Second pass of the payoff averaging: sum (V0 - mean)^2 with the (global) mean already known -- the
// two-pass form keeps the standard error exact even when every V0 is identical (j = 0 in the money).
template <typename CT>
__global__ void __launch_bounds__(LSM_NT) lsm_sqdev_kernel(const CT* __restrict__ V, int64_t n, const double* __restrict__ fin,
                                                          double* __restrict__ partial) {
    const double mean = fin[0] / fin[2];
    double acc[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * LSM_NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * LSM_NT) {
        const double dlt = (double)V[i] - mean;
        acc[0] = fma(dlt, dlt, acc[0]);
    }
    block_reduce_to_partial<1>(acc, partial + (int64_t)blockIdx.x * MOM_LD);
}

__global__ void lsm_fill_tau_kernel(int32_t* __restrict__ tau, int64_t n, int32_t val) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tau[i] = val;
}

// ---------------------------------------------------------------------------------------------------------
// Whole backward induction in ONE launch for small path sets (the reference's production rows are 250 paths,
// PredictionGen.cpp:719): a single CTA keeps the fp64 carry in shared memory and loops over the time steps with
// block barriers -- regression moments, fold, solve and decision of every step without leaving the kernel.  Same
// arithmetic as the parity kernel (all decisions in fp64 on the stored values).  One launch instead of one per time
// step matters twice for the row loop of the reference: per-row latency, and the launch stream of 16 host threads no
// longer serialises on the driver.
// ---------------------------------------------------------------------------------------------------------
constexpr int SMALL_NT = SB_NT;
constexpr int SMALL_MAX_PATHS = SB_MAX_PATHS;

template <typename ST, int P>
__global__ void __launch_bounds__(SB_NT, 1) lsm_small_kernel(SweepArgs a, int M, double dt, double maturity) {
    extern __shared__ double sV[];  // carry [n]
    sb_lsm<ST, P>(reinterpret_cast<const ST*>(a.S), a.ld, (int)a.n, M, a.K, a.is_call, a.disc, dt, maturity, sV, a.tau, a.d.coef, a.d.mu, a.d.inv_s,
                  reinterpret_cast<double*>(a.V), a.d.fin);
}

typedef void (*SmallFn)(SweepArgs, int, double, double);
template <typename ST>
SmallFn pick_small(int p) {
    switch (p) {
        case 0: return lsm_small_kernel<ST, 0>;
        case 1: return lsm_small_kernel<ST, 1>;
        case 2: return lsm_small_kernel<ST, 2>;
        case 3: return lsm_small_kernel<ST, 3>;
        case 4: return lsm_small_kernel<ST, 4>;
        case 5: return lsm_small_kernel<ST, 5>;
        default: return lsm_small_kernel<ST, 6>;
    }
}

typedef void (*SweepFn)(SweepArgs);
typedef void (*SweepFn2)(SweepArgs, int);

template <typename ST, typename CT>
SweepFn pick_sweep(int p) {
    switch (p) {
        case 0: return lsm_sweep_kernel<ST, CT, 0>;
        case 1: return lsm_sweep_kernel<ST, CT, 1>;
        case 2: return lsm_sweep_kernel<ST, CT, 2>;
        case 3: return lsm_sweep_kernel<ST, CT, 3>;
        case 4: return lsm_sweep_kernel<ST, CT, 4>;
        case 5: return lsm_sweep_kernel<ST, CT, 5>;
        default: return lsm_sweep_kernel<ST, CT, 6>;
    }
}

template <bool TAU>
SweepFn pick_sweep_fast2(int p) {
    switch (p) {
        case 0: return lsm_sweep_fast2_kernel<0, TAU>;
        case 1: return lsm_sweep_fast2_kernel<1, TAU>;
        case 2: return lsm_sweep_fast2_kernel<2, TAU>;
        case 3: return lsm_sweep_fast2_kernel<3, TAU>;
        case 4: return lsm_sweep_fast2_kernel<4, TAU>;
        case 5: return lsm_sweep_fast2_kernel<5, TAU>;
        default: return lsm_sweep_fast2_kernel<6, TAU>;
    }
}

template <bool TAU>
SweepFn2 pick_sweep_tma(int p) {
    switch (p) {
        case 0: return lsm_sweep_tma_kernel<0, TAU>;
        case 1: return lsm_sweep_tma_kernel<1, TAU>;
        case 2: return lsm_sweep_tma_kernel<2, TAU>;
        case 3: return lsm_sweep_tma_kernel<3, TAU>;
        case 4: return lsm_sweep_tma_kernel<4, TAU>;
        case 5: return lsm_sweep_tma_kernel<5, TAU>;
        default: return lsm_sweep_tma_kernel<6, TAU>;
    }
}

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

SweepFn pick_sweep(int slab_dtype, int carry_dtype, int p, bool want_tau) {
    if (slab_dtype == MCP_F32 && carry_dtype == MCP_F32) return want_tau ? pick_sweep_fast2<true>(p) : pick_sweep_fast2<false>(p);
    if (slab_dtype == MCP_F32) return pick_sweep<float, double>(p);
    return carry_dtype == MCP_F32 ? pick_sweep<double, float>(p) : pick_sweep<double, double>(p);
}

// Express  cont(S) = sum_k c_k ((S - mu) inv_s)^k  in the requested output basis.
void convert_coeffs(const double* c, int p, double mu, double inv_s, double K, int basis, double* out) {
    // monomials in S
    double mono[MAXP + 1] = {0}, pw[MAXP + 1] = {0}, tmp[MAXP + 1];
    pw[0] = 1.0;  // ((S - mu) inv_s)^0
    const double a0 = -mu * inv_s, a1 = inv_s;
    for (int k = 0; k <= p; ++k) {
        for (int m = 0; m <= p; ++m) mono[m] += c[k] * pw[m];
        for (int m = 0; m <= p; ++m) tmp[m] = a0 * pw[m] + (m > 0 ? a1 * pw[m - 1] : 0.0);
        for (int m = 0; m <= p; ++m) pw[m] = tmp[m];
    }
    if (basis == MCP_BASIS_MONOMIAL) {
        for (int m = 0; m <= p; ++m) out[m] = mono[m];
        return;
    }
    // Laguerre in u = S/K:  sum_m mono_m K^m u^m = sum_k b_k L_k(u),  L_k(u) = sum_m C(k,m) (-1)^m / m! u^m
    double um[MAXP + 1], Kp = 1.0;
    for (int m = 0; m <= p; ++m) { um[m] = mono[m] * Kp; Kp *= K; }
    for (int k = p; k >= 0; --k) {
        // leading coefficient of L_k is (-1)^k / k!
        double fact = 1.0;
        for (int m = 2; m <= k; ++m) fact *= m;
        const double lead = ((k & 1) ? -1.0 : 1.0) / fact;
        const double b = um[k] / lead;
        out[k] = b;
        double binom = 1.0, mf = 1.0;  // C(k,m), m!
        for (int m = 0; m <= k; ++m) {
            if (m > 0) { binom = binom * (double)(k - m + 1) / (double)m; mf *= m; }
            um[m] -= b * binom * ((m & 1) ? -1.0 : 1.0) / mf;
        }
    }
}

}  // namespace

extern "C" int mcp_lsm_price(mcp_ctx* ctx, const mcp_pathset* ps, const mcp_lsm_params* prm, mcp_lsm_result* res, double* coeffs,
                             int32_t* first_exercise, double* v0) {
    if (!ctx || !prm || !res) return MCP_ERR_INVALID;
    if (!ps || ps->n_paths <= 0) return mcp_fail(ctx, MCP_ERR_EMPTY_PATHS, "LSM::PredictOptionPrice: Empty pricePaths.");
    if (ps->ctx != ctx) return mcp_fail(ctx, MCP_ERR_INVALID, "lsm: pathset belongs to another ctx");
    const int p = prm->poly_order;
    if (p < 0) return mcp_fail(ctx, MCP_ERR_INVALID, "lsm: poly_order %d < 0", p);
    if (p > MAXP) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "lsm: poly_order %d > %d", p, MAXP);
    if (prm->carry != MCP_F32 && prm->carry != MCP_F64) return mcp_fail(ctx, MCP_ERR_INVALID, "lsm: bad carry dtype");
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));

    const int M = ps->n_steps + 1;
    const int64_t N = ps->n_paths;
    // small path sets: the whole induction in one single-CTA launch (always with the fp64 carry)
    const bool small = N <= SMALL_MAX_PATHS && !(ctx->nranks > 1 && ctx->comm) && env_int("MCP_LSM_SMALL", 1) != 0;
    const int carry = small ? (int)MCP_F64 : prm->carry;
    const size_t csz = carry == MCP_F32 ? 4 : 8;
    const uint64_t launches0 = ctx->launches;

    SweepFn sweep = pick_sweep(ps->dtype, carry, p, first_exercise != nullptr);
    int occ = 0;
    MCP_TRY(mcp_kernel_config(ctx, (const void*)sweep, LSM_NT, 0, &occ));
    if (occ < 1) occ = 1;
    int64_t grid = (N + (int64_t)LSM_NT * 4 - 1) / ((int64_t)LSM_NT * 4);
    const int64_t cap = (int64_t)ctx->sm_count * occ;
    if (grid > cap) grid = cap;
    const int64_t grid_aux = grid;  // the small streaming kernels keep the occupancy-sized grid
    // TMA-ring kernel (one persistent 512-thread CTA per SM) once there are enough 4096-path tiles to feed every SM
    SweepFn2 sweep_tma = nullptr;
    const bool tma_pdl = env_int("MCP_SWEEP_PDL", 1) != 0;
    int tma_stages = 0;
    size_t tma_smem = 0;
    const int64_t ntile = (N + TMA_TILE - 1) / TMA_TILE;
    if (ps->dtype == MCP_F32 && carry == MCP_F64 && !small && env_int("MCP_SWEEP64_IMPL", 3) == 3 && ntile >= 2 * (int64_t)ctx->sm_count) {
        // parity arithmetic on the TMA ring: 64 KB stages (S_j | S_{j-1} | fp64 V), moments in registers
        tma_stages = (int)((227u * 1024u - 12288u - 128u) / TMA64_STAGE_BYTES);
        const int want = env_int("MCP_SWEEP_STAGES", 0);
        if (want > 0 && want < tma_stages) tma_stages = want;
        if (tma_stages >= 2) {
            sweep_tma = pick_sweep_tma64(p);
            tma_smem = (size_t)tma_stages * TMA64_STAGE_BYTES + 128;
            MCP_TRY(mcp_kernel_config(ctx, (const void*)sweep_tma, TMA_NT, tma_smem, nullptr));
            grid = ctx->sm_count < ntile ? ctx->sm_count : ntile;
        }
    }
    if (ps->dtype == MCP_F32 && carry == MCP_F32 && env_int("MCP_SWEEP_IMPL", 3) == 3 && ntile >= 2 * (int64_t)ctx->sm_count) {
        const int nv = 3 * p + 2 > 2 ? 3 * p + 2 : 2;
        const size_t fixed = 128 + MCP_DBG_CANARY_BYTES + (size_t)nv * TMA_NT * 8;
        tma_stages = (int)((227u * 1024u - 12288u - fixed) / TMA_STAGE_BYTES);  // 12 KB: the kernel's static shared memory
        const int want = env_int("MCP_SWEEP_STAGES", 0);
        if (want > 0 && want < tma_stages) tma_stages = want;
        if (tma_stages >= 2) {
            sweep_tma = first_exercise ? pick_sweep_tma<true>(p) : pick_sweep_tma<false>(p);
            tma_smem = (size_t)tma_stages * TMA_STAGE_BYTES + fixed;
            MCP_TRY(mcp_kernel_config(ctx, (const void*)sweep_tma, TMA_NT, tma_smem, nullptr));
            grid = ctx->sm_count < ntile ? ctx->sm_count : ntile;
        }
    }

    // Persistent sweep (one cooperative launch for the whole induction, lsm_persist.cuh): always across GPUs with peer-memory
    // mailboxes (every rank must follow the same protocol whatever its shard size), and on one GPU when the working set of a
    // step stays in L2, where the per-launch tail would dominate.  MCP_SWEEP_IMPL=4 forces it, =3 forbids it.
    const bool multi_early = ctx->nranks > 1 && ctx->comm;
    const int impl_env = env_int("MCP_SWEEP_IMPL", 0);
    PersistFn persist = nullptr;
    int px_workers = 0, px_stages = 0;
    size_t px_smem = 0;
    if (!small && ps->dtype == MCP_F32 && carry == MCP_F32 && impl_env != 3 && (!multi_early || ctx->xchg.enabled)) {
        const int nv = 3 * p + 2 > 2 ? 3 * p + 2 : 2;
        const size_t fixed = 128 + MCP_DBG_CANARY_BYTES + (size_t)nv * (px::NT / 2) * 8;
        px_stages = (int)((227u * 1024u - 4096u - fixed) / px::STAGE_BYTES);
        const int want = env_int("MCP_SWEEP_STAGES", 0);
        if (want > 0 && want < px_stages) px_stages = want;
        if (px_stages > 8) px_stages = 8;
        int64_t w = ntile;  // a worker with more tiles than ring stages chains them across steps, one with fewer refills per step
        const int64_t wmax = (ctx->sm_count - 1 < MCP_PX_MAXW ? ctx->sm_count - 1 : MCP_PX_MAXW);
        if (w > wmax) w = wmax;
        const int w_env = env_int("MCP_PX_WORKERS", 0);  // experiments: fewer streaming CTAs
        if (w_env > 0 && w_env < w) w = w_env;
        if (w < 1) w = 1;
        const bool l2_fit = (size_t)N * 12 <= ((size_t)env_int("MCP_L2_RESIDENT_MB", 104) << 20);
        const bool want_px = (multi_early && ctx->xchg.enabled) || impl_env == 4 || (impl_env == 0 && l2_fit && ntile >= 32);
        if (want_px && px_stages >= 2) {
            persist = first_exercise ? pick_persist<true>(p) : pick_persist<false>(p);
            px_workers = (int)w;
            px_smem = (size_t)px_stages * px::STAGE_BYTES + fixed;
            if (px_smem < (size_t)px::RED_SMEM_BYTES) px_smem = px::RED_SMEM_BYTES;
            MCP_TRY(mcp_kernel_config(ctx, (const void*)persist, px::NT, px_smem, nullptr));
        }
    }

    // ---- workspace ----
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_coef = take((size_t)M * COEF_LD * 8), o_mu = take((size_t)M * 8), o_is = take((size_t)M * 8);
    const size_t o_ssum = take((size_t)M * 4 * 8), o_part = take((size_t)(grid > grid_aux ? grid : grid_aux) * MOM_LD * 8), o_mom = take(MOM_LD * 8);
    const size_t o_fin = take(4 * 8), o_kind = take((size_t)M * 4), o_cnt = take(4);
    MCP_TRY(mcp_scratch_reserve(ctx, off));
    const size_t v_bytes = (size_t)mcp_round_up(N, 128) * csz;
    const size_t tau_bytes = first_exercise ? (size_t)mcp_round_up(N, 128) * 4 : 0;
    const size_t v0_bytes = v0 ? (size_t)mcp_round_up(N, 128) * 8 : 0;
    MCP_TRY(mcp_carry_reserve(ctx, v_bytes + tau_bytes + v0_bytes));
    unsigned char* sb = (unsigned char*)ctx->scratch;
    LsmDev d;
    d.coef = (double*)(sb + o_coef); d.mu = (double*)(sb + o_mu); d.inv_s = (double*)(sb + o_is); d.ssum = (double*)(sb + o_ssum);
    d.partial = (double*)(sb + o_part); d.moments = (double*)(sb + o_mom); d.fin = (double*)(sb + o_fin); d.kind = (int*)(sb + o_kind); d.counter = (unsigned int*)(sb + o_cnt);
    void* dV = ctx->carry;
    int32_t* dTau = first_exercise ? (int32_t*)((unsigned char*)ctx->carry + v_bytes) : nullptr;
    double* dV0 = v0 ? (double*)((unsigned char*)ctx->carry + v_bytes + tau_bytes) : nullptr;

    // step kinds: `thisTime = j * dt; if (thisTime > maturity)` evaluated in double exactly as LSMPricer.cpp:43-44
    std::vector<int> kind(M, STEP_NORMAL);
    for (int j = 0; j < M; ++j) kind[j] = ((double)j * prm->dt > prm->maturity) ? STEP_DISCOUNT : STEP_NORMAL;
    const double disc = exp(-prm->r * prm->dt);  // LSMPricer.cpp:46,69,92

    cudaStream_t st = ctx->stream;
    MCP_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    MCP_TRY(mcp_h2d(ctx, d.kind, kind.data(), (size_t)M * 4));
    MCP_CUDA(ctx, cudaMemsetAsync(d.coef, 0, (size_t)M * COEF_LD * 8, st));
    MCP_CUDA(ctx, cudaMemsetAsync(d.counter, 0, 4, st));

    // ---- standardisation tables from a fixed leading sample of this rank's paths (summed over ranks) ----
    // (the single-launch kernel standardises each step itself, from all of its in-the-money prices; the persistent sweep
    // computes and exchanges the same sample sums inside its launch)
    if (!small && !persist) {
        const int ns = (int)(N < SAMPLE_MAX ? N : SAMPLE_MAX);
        if (ps->dtype == MCP_F32) lsm_scale_sums_kernel<float><<<M, 256, 0, st>>>((const float*)ps->data, ps->ld, ns, prm->strike, prm->is_call, d.ssum);
        else lsm_scale_sums_kernel<double><<<M, 256, 0, st>>>((const double*)ps->data, ps->ld, ns, prm->strike, prm->is_call, d.ssum);
        MCP_LAUNCH_CHECK(ctx);
        MCP_TRY(mcp_allreduce_f64(ctx, d.ssum, M * 4));
        lsm_scale_finalize_kernel<<<(M + 127) / 128, 128, 0, st>>>(d.ssum, M, prm->strike, d.mu, d.inv_s);
        MCP_LAUNCH_CHECK(ctx);
    }
    if (dTau) {
        lsm_fill_tau_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(dTau, N, (int32_t)(M - 1));
        MCP_LAUNCH_CHECK(ctx);
    }

    // ---- backward sweep ----
    SweepArgs a;
    memset(&a, 0, sizeof(a));
    a.S = ps->data; a.ld = ps->ld; a.n = N; a.V = dV; a.tau = dTau; a.d = d;
    a.K = prm->strike; a.disc = disc; a.is_call = prm->is_call;
    a.g = make_stepk(prm->strike, disc, prm->is_call);
    // the reference's SVD rank cut (lsm_solve.cuh): always in parity mode and for p >= 5, where it bites; the throughput mode skips
    // the 3-sweep Jacobi of a full-rank 4 x 4 factor on its per-step critical path.  MCP_LSM_REF_RANK=0/1 overrides (tests).
    a.ref_rank = env_int("MCP_LSM_REF_RANK", (carry == MCP_F64 || p >= 5) ? 1 : 0) ? 1 : 0;
    const bool multi = ctx->nranks > 1 && ctx->comm;
    const bool p2p = multi && ctx->xchg.enabled;
    if (p2p) a.x = ctx->xchg;
    a.solve_here = (!multi || p2p) ? 1 : 0;
    a.l2_resident = ((size_t)N * 12 <= ((size_t)env_int("MCP_L2_RESIDENT_MB", 104) << 20)) ? 1 : 0;
    const int nm = 3 * p + 2;
    if (small) {
        SmallFn fn = ps->dtype == MCP_F32 ? pick_small<float>(p) : pick_small<double>(p);
        const size_t smem = (size_t)N * sizeof(double);
        MCP_TRY(mcp_kernel_config(ctx, (const void*)fn, SMALL_NT, (size_t)SMALL_MAX_PATHS * sizeof(double), nullptr));
        if (ctx->profiling) cudaEventRecord(mcp_prof_event(ctx, 0), st);
        fn<<<1, SMALL_NT, smem, st>>>(a, M, prm->dt, prm->maturity);
        MCP_LAUNCH_CHECK(ctx);
        if (ctx->profiling) cudaEventRecord(mcp_prof_event(ctx, 1), st);
    }
    if (persist) {
        px::Args pa;
        memset(&pa, 0, sizeof(pa));
        pa.S = (const float*)ps->data; pa.ld = ps->ld; pa.n = N; pa.V = (float*)dV; pa.tau = dTau;
        pa.coef = d.coef; pa.mu = d.mu; pa.inv_s = d.inv_s; pa.ssum = d.ssum; pa.fin = d.fin; pa.kind = d.kind;
        pa.K = prm->strike; pa.disc = disc; pa.is_call = prm->is_call; pa.M = M; pa.g = a.g; pa.ref_rank = a.ref_rank;
        pa.ns = (int)(N < SAMPLE_MAX ? N : SAMPLE_MAX);
        pa.l2_resident = a.l2_resident; pa.n_workers = px_workers; pa.n_stages = px_stages;
        MCP_TRY(mcp_px_get(ctx, &pa.x));
        if (env_int("MCP_PX_TRACE", 0)) {  // debugging aid: per-step time stamps of every CTA (tools/px_trace.py)
            const size_t tb = (size_t)M * (px_workers + 1) * 4 * 8;
            if (ctx->px_trace_bytes < tb) {
                if (ctx->px_trace) cudaFree(ctx->px_trace);
                ctx->px_trace = nullptr;
                ctx->px_trace_bytes = 0;
                if (cudaMalloc(&ctx->px_trace, tb) == cudaSuccess) ctx->px_trace_bytes = tb;
                else cudaGetLastError();
            }
            if (ctx->px_trace) {
                cudaMemsetAsync(ctx->px_trace, 0, tb, st);
                pa.trace = (unsigned long long*)ctx->px_trace;
                ctx->px_trace_rows = M;
                ctx->px_trace_cols = px_workers + 1;
            }
        }
        pa.seq0 = ctx->xchg_seq + 1;
        ctx->xchg_seq += px::exchanges_per_launch(M, kind.data());
        void* kargs[] = {(void*)&pa};
        if (ctx->profiling) cudaEventRecord(mcp_prof_event(ctx, 0), st);
        MCP_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)persist, dim3((unsigned)(px_workers + 1)), dim3(px::NT), kargs, px_smem, st));
        MCP_LAUNCH_CHECK(ctx);
        if (ctx->profiling) cudaEventRecord(mcp_prof_event(ctx, 1), st);
    }
    for (int j = M - 1; j >= 0 && !small && !persist; --j) {
        a.j = j;
        a.terminal = (j == M - 1);
        a.do_moments = (j > 0 && kind[j - 1] == STEP_NORMAL);
        a.do_final = (j == 0);
        if (p2p && (a.do_moments || a.do_final)) a.seq = ++ctx->xchg_seq;
        if (ctx->profiling) cudaEventRecord(mcp_prof_event(ctx, 2 * (size_t)j), st);
        if (sweep_tma) {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3((unsigned)grid);
            cfg.blockDim = dim3(TMA_NT);
            cfg.dynamicSmemBytes = tma_smem;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = tma_pdl ? 1 : 0;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            MCP_CUDA(ctx, cudaLaunchKernelEx(&cfg, sweep_tma, a, tma_stages));
        } else {
            sweep<<<(unsigned)grid, LSM_NT, 0, st>>>(a);
        }
        MCP_LAUNCH_CHECK(ctx);
        if (ctx->profiling) cudaEventRecord(mcp_prof_event(ctx, 2 * (size_t)j + 1), st);
        if (a.do_moments && !a.solve_here) {  // multi-GPU: global moments, then every rank solves the same system
            MCP_TRY(mcp_allreduce_f64(ctx, d.moments, nm));
            lsm_solve_kernel<<<1, 32, 0, st>>>(d.moments, p, d.coef + (int64_t)(j - 1) * COEF_LD, d.mu + (j - 1), d.inv_s + (j - 1), a.ref_rank);
            MCP_LAUNCH_CHECK(ctx);
        }
    }
    // ---- payoff averaging: sum V0 (+ N) -> global mean -> sum of squared deviations ----
    double fin[3] = {0, 0, 0};  // d.fin[0] = sum V0 was written by the last CTA of sweep(0)
    if (!small && !persist) {
        const double nloc = (double)N;
        if (!p2p) MCP_TRY(mcp_h2d(ctx, d.fin + 2, &nloc, 8));
        if (!p2p) MCP_TRY(mcp_allreduce_f64(ctx, d.fin, 3));  // fin[1] is overwritten below; with mailboxes sweep(0) already left the global {sum V0, N}
        if (carry == MCP_F32) lsm_sqdev_kernel<float><<<(unsigned)grid_aux, LSM_NT, 0, st>>>((const float*)dV, N, d.fin, d.partial);
        else lsm_sqdev_kernel<double><<<(unsigned)grid_aux, LSM_NT, 0, st>>>((const double*)dV, N, d.fin, d.partial);
        MCP_LAUNCH_CHECK(ctx);
        lsm_reduce_kernel<<<1, 256, 0, st>>>(d.partial, (int)grid_aux, 1, d.fin + 1);
        MCP_LAUNCH_CHECK(ctx);
        MCP_TRY(mcp_allreduce_f64(ctx, d.fin + 1, 1));
    }  // the single-launch kernel leaves {sum V0, sum (V0 - mean)^2, N} itself
    MCP_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    double* fin_pin = (double*)mcp_stage_alloc(ctx, 3 * 8);
    MCP_CUDA(ctx, mcp_memcpy_async(ctx, fin_pin ? fin_pin : fin, d.fin, 3 * 8, cudaMemcpyDeviceToHost, st));
    if (dV0) {
        if (carry == MCP_F32) lsm_copy_v0_kernel<float><<<(unsigned)((N + 255) / 256), 256, 0, st>>>((const float*)dV, N, dV0);
        else lsm_copy_v0_kernel<double><<<(unsigned)((N + 255) / 256), 256, 0, st>>>((const double*)dV, N, dV0);
        MCP_LAUNCH_CHECK(ctx);
        MCP_CUDA(ctx, mcp_memcpy_async(ctx, v0, dV0, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
    }
    if (dTau) MCP_CUDA(ctx, mcp_memcpy_async(ctx, first_exercise, dTau, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
    std::vector<double> hcoef, hmu, his;
    if (coeffs) {
        hcoef.resize((size_t)M * COEF_LD); hmu.resize(M); his.resize(M);
        MCP_CUDA(ctx, mcp_memcpy_async(ctx, hcoef.data(), d.coef, (size_t)M * COEF_LD * 8, cudaMemcpyDeviceToHost, st));
        MCP_CUDA(ctx, mcp_memcpy_async(ctx, hmu.data(), d.mu, (size_t)M * 8, cudaMemcpyDeviceToHost, st));
        MCP_CUDA(ctx, mcp_memcpy_async(ctx, his.data(), d.inv_s, (size_t)M * 8, cudaMemcpyDeviceToHost, st));
    }
    int xerr = 0;
    if ((p2p || persist) && ctx->xchg.err) MCP_CUDA(ctx, mcp_memcpy_async(ctx, &xerr, ctx->xchg.err, sizeof(int), cudaMemcpyDeviceToHost, st));
    MCP_CUDA(ctx, cudaStreamSynchronize(st));
    MCP_CUDA(ctx, cudaGetLastError());
    if (fin_pin) memcpy(fin, fin_pin, 3 * 8);
    if (xerr) {
        cudaMemsetAsync(ctx->xchg.err, 0, sizeof(int), st);  // reported once; the next call starts clean
        return mcp_fail(ctx, MCP_ERR_NCCL, "lsm: peer-memory moment exchange timed out (a rank did not reach the same sweep step)");
    }

    const double ng = fin[2];
    res->sum_v0 = fin[0];
    res->sum_sq_dev = fin[1];
    res->n_paths_global = (int64_t)llround(ng);
    res->price = fin[0] / ng;
    const double var = ng > 1.0 ? fin[1] / (ng - 1.0) : 0.0;
    res->std_error = var > 0.0 ? sqrt(var / ng) : 0.0;
    float ms = 0.f;
    MCP_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    res->elapsed_ms = ms;
    ctx->prof.lsm_total_ms = ms;
    ctx->prof.sweep_kernels_ms = 0.f;
    ctx->prof.n_sweep_launches = 0;
    ctx->prof.n_sweep_steps = M;
    if (ctx->profiling) {
        for (int j = 0; j < ((small || persist) ? 1 : M); ++j) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, mcp_prof_event(ctx, 2 * (size_t)j), mcp_prof_event(ctx, 2 * (size_t)j + 1)) == cudaSuccess)
                ctx->prof.sweep_kernels_ms += t;
            ctx->prof.n_sweep_launches++;
        }
    }
    res->n_kernel_launches = (int)(ctx->launches - launches0);
    if (coeffs && prm->basis == MCP_BASIS_STANDARDISED) {  // rows [c_0 .. c_p, mu, 1/s]: cont(S) = sum_k c_k ((S - mu) / s)^k
        for (int j = 0; j + 1 < M; ++j) {
            double* out = coeffs + (size_t)j * (p + 3);
            for (int k = 0; k <= p; ++k) out[k] = hcoef[(size_t)j * COEF_LD + k];
            out[p + 1] = hmu[j];
            out[p + 2] = his[j];
        }
    } else if (coeffs) {
        for (int j = 0; j + 1 < M; ++j) {
            double* out = coeffs + (size_t)j * (p + 1);
            bool any = false;
            for (int k = 0; k <= p; ++k) any = any || hcoef[(size_t)j * COEF_LD + k] != 0.0;
            if (!any) { for (int k = 0; k <= p; ++k) out[k] = 0.0; continue; }
            convert_coeffs(&hcoef[(size_t)j * COEF_LD], p, hmu[j], his[j], prm->strike, prm->basis, out);
        }
    }
    return MCP_OK;
}

// Several strikes on the SAME path set in one sweep (surface ladders, config 5): throughput mode only (fp32 slab,
// fp32 carry, one GPU).  Other combinations price the strikes one after the other through mcp_lsm_price.
// Debugging aid (not part of the reference boundary): the time stamps of the last persistent sweep run with MCP_PX_TRACE=1,
// [rows][cols][4] uint64 nanoseconds; returns rows * cols * 4 or a negative status.
extern "C" int64_t mcp_debug_px_trace(mcp_ctx* ctx, unsigned long long* out, int64_t max_words, int* rows, int* cols) {
    if (!ctx || !ctx->px_trace) return MCP_ERR_INVALID;
    const int64_t n = (int64_t)ctx->px_trace_rows * ctx->px_trace_cols * 4;
    if (rows) *rows = ctx->px_trace_rows;
    if (cols) *cols = ctx->px_trace_cols;
    if (out && max_words >= n) {
        cudaStreamSynchronize(ctx->stream);
        if (cudaMemcpy(out, ctx->px_trace, (size_t)n * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return MCP_ERR_CUDA;
    }
    return n;
}

extern "C" int mcp_lsm_price_multi(mcp_ctx* ctx, const mcp_pathset* ps, const mcp_lsm_params* prm, const double* strikes, int n_strikes,
                                   mcp_lsm_result* res) {
    if (!ctx || !prm || !strikes || !res || n_strikes <= 0) return MCP_ERR_INVALID;
    if (!ps || ps->n_paths <= 0) return mcp_fail(ctx, MCP_ERR_EMPTY_PATHS, "LSM::PredictOptionPrice: Empty pricePaths.");
    if (ps->ctx != ctx) return mcp_fail(ctx, MCP_ERR_INVALID, "lsm: pathset belongs to another ctx");
    const int p = prm->poly_order;
    if (p < 0) return mcp_fail(ctx, MCP_ERR_INVALID, "lsm: poly_order %d < 0", p);
    if (p > MAXP) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "lsm: poly_order %d > %d", p, MAXP);
    const int64_t N = ps->n_paths;
    const bool batched = ps->dtype == MCP_F32 && prm->carry == MCP_F32 && !(ctx->nranks > 1 && ctx->comm) && n_strikes >= 2 && N > SMALL_MAX_PATHS &&
                         env_int("MCP_LSM_MULTI", 1) != 0;
    if (!batched) {
        for (int k = 0; k < n_strikes; ++k) {
            mcp_lsm_params q = *prm;
            q.strike = strikes[k];
            MCP_TRY(mcp_lsm_price(ctx, ps, &q, &res[k], nullptr, nullptr, nullptr));
        }
        return MCP_OK;
    }
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));
    const int M = ps->n_steps + 1;
    const int64_t ld = ps->ld;
    const int nv = 3 * p + 2 > 2 ? 3 * p + 2 : 2;
    MultiFn fn = pick_multi(p);
    const size_t fixed = (size_t)MULTI_NS * MULTI_SLAB_STAGE_BYTES + MCP_DBG_CANARY_BYTES + MULTI_BAR_BYTES + (size_t)nv * MULTI_ACC_LANES * 8;
    int depth = (int)((227u * 1024u - 12288u - fixed) / ((size_t)MULTI_MAXC * MULTI_CARRY_SLOT_BYTES));  // carry ring depth per contract warp
    if (depth > MULTI_MAXD) depth = MULTI_MAXD;
    if (depth < 2) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "lsm multi: shared memory");
    const size_t smem = fixed + (size_t)depth * MULTI_MAXC * MULTI_CARRY_SLOT_BYTES;
    MCP_TRY(mcp_kernel_config(ctx, (const void*)fn, MULTI_NT, smem, nullptr));
    const int64_t ntile = (N + MULTI_TILE - 1) / MULTI_TILE;
    const int grid = (int)(ctx->sm_count < ntile ? ctx->sm_count : ntile);
    const int grid_aux = (int)((N + (int64_t)LSM_NT * 4 - 1) / ((int64_t)LSM_NT * 4) < (int64_t)ctx->sm_count * 3 ? (N + (int64_t)LSM_NT * 4 - 1) / ((int64_t)LSM_NT * 4)
                                                                                                                   : (int64_t)ctx->sm_count * 3);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_coef = take((size_t)MULTI_MAXC * M * COEF_LD * 8), o_mu = take((size_t)MULTI_MAXC * M * 8), o_is = take((size_t)MULTI_MAXC * M * 8);
    const size_t o_ssum = take((size_t)M * 4 * 8);
    const size_t o_part = take((size_t)(grid > grid_aux ? grid : grid_aux) * MULTI_MAXC * MOM_LD * 8), o_fin = take((size_t)MULTI_MAXC * 4 * 8);
    const size_t o_kind = take((size_t)M * 4), o_cnt = take(4);
    MCP_TRY(mcp_scratch_reserve(ctx, off));
    MCP_TRY(mcp_carry_reserve(ctx, (size_t)MULTI_MAXC * ld * 4));
    unsigned char* sb = (unsigned char*)ctx->scratch;
    MultiArgs a;
    memset(&a, 0, sizeof(a));
    a.S = (const float*)ps->data; a.ld = ld; a.n = N; a.V = (float*)ctx->carry;
    a.coef = (double*)(sb + o_coef); a.mu = (double*)(sb + o_mu); a.inv_s = (double*)(sb + o_is);
    a.partial = (double*)(sb + o_part); a.fin = (double*)(sb + o_fin); a.kind = (int*)(sb + o_kind); a.counter = (unsigned int*)(sb + o_cnt);
    a.disc = exp(-prm->r * prm->dt); a.M = M; a.is_call = prm->is_call; a.ref_rank = p >= 5 ? 1 : 0;
    double* d_ssum = (double*)(sb + o_ssum);
    std::vector<int> kind(M, STEP_NORMAL);
    for (int j = 0; j < M; ++j) kind[j] = ((double)j * prm->dt > prm->maturity) ? STEP_DISCOUNT : STEP_NORMAL;
    cudaStream_t st = ctx->stream;
    MCP_TRY(mcp_h2d(ctx, (void*)a.kind, kind.data(), (size_t)M * 4));
    MCP_CUDA(ctx, cudaMemsetAsync(a.counter, 0, 4, st));
    const int ns = (int)(N < SAMPLE_MAX ? N : SAMPLE_MAX);
    const uint64_t launches0 = ctx->launches;
    for (int k0 = 0; k0 < n_strikes; k0 += MULTI_MAXC) {
        const int C = n_strikes - k0 < MULTI_MAXC ? n_strikes - k0 : MULTI_MAXC;
        a.C = C;
        for (int c = 0; c < MULTI_MAXC; ++c) a.K[c] = strikes[k0 + (c < C ? c : 0)];
        MCP_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
        MCP_CUDA(ctx, cudaMemsetAsync(a.coef, 0, (size_t)MULTI_MAXC * M * COEF_LD * 8, st));
        for (int c = 0; c < MULTI_MAXC; ++c) {  // standardisation of every contract from its own in-the-money sample (as mcp_lsm_price)
            lsm_scale_sums_kernel<float><<<M, 256, 0, st>>>((const float*)ps->data, ld, ns, a.K[c], prm->is_call, d_ssum);
            MCP_LAUNCH_CHECK(ctx);
            lsm_scale_finalize_kernel<<<(M + 127) / 128, 128, 0, st>>>(d_ssum, M, a.K[c], (double*)a.mu + (size_t)c * M, (double*)a.inv_s + (size_t)c * M);
            MCP_LAUNCH_CHECK(ctx);
        }
        for (int j = M - 1; j >= 0; --j) {
            a.j = j;
            a.terminal = (j == M - 1);
            a.do_moments = (j > 0 && kind[j - 1] == STEP_NORMAL);
            a.do_final = (j == 0);
            fn<<<grid, MULTI_NT, smem, st>>>(a, depth);
            MCP_LAUNCH_CHECK(ctx);
        }
        // per contract: mean known -> sum of squared deviations (two-pass standard error)
        const double nloc = (double)N;
        for (int c = 0; c < C; ++c) {
            MCP_TRY(mcp_h2d(ctx, a.fin + c * 4 + 2, &nloc, 8));
            lsm_sqdev_kernel<float><<<grid_aux, LSM_NT, 0, st>>>(a.V + (int64_t)c * ld, N, a.fin + c * 4, a.partial);
            MCP_LAUNCH_CHECK(ctx);
            lsm_reduce_kernel<<<1, 256, 0, st>>>(a.partial, grid_aux, 1, a.fin + c * 4 + 1);
            MCP_LAUNCH_CHECK(ctx);
        }
        MCP_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
        double fin[MULTI_MAXC * 4];
        double* fp = (double*)mcp_stage_alloc(ctx, sizeof(fin));
        MCP_CUDA(ctx, mcp_memcpy_async(ctx, fp ? fp : fin, a.fin, sizeof(fin), cudaMemcpyDeviceToHost, st));
        MCP_CUDA(ctx, cudaStreamSynchronize(st));
        MCP_CUDA(ctx, cudaGetLastError());
        if (fp) memcpy(fin, fp, sizeof(fin));
        float ms = 0.f;
        MCP_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        for (int c = 0; c < C; ++c) {
            mcp_lsm_result& r = res[k0 + c];
            r.sum_v0 = fin[c * 4];
            r.sum_sq_dev = fin[c * 4 + 1];
            r.n_paths_global = N;
            r.price = fin[c * 4] / nloc;
            const double var = nloc > 1.0 ? fin[c * 4 + 1] / (nloc - 1.0) : 0.0;
            r.std_error = var > 0.0 ? sqrt(var / nloc) : 0.0;
            r.elapsed_ms = ms / (float)C;
            r.n_kernel_launches = (int)(ctx->launches - launches0);
        }
    }
    return MCP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Out-of-sample value of a fitted exercise policy [new: the reference has no error estimate at all].
// LSM::PredictOptionPrice carries FITTED continuation values backwards (LSMPricer.cpp:78-86), so the spread of V[:,0] says
// nothing about the noise of the regressions: measured seed-to-seed scatter is 2-4x mcp_lsm_result::std_error.
// The honest companion number: take the coefficient tables of one run (MCP_BASIS_STANDARDISED rows), generate an INDEPENDENT
// path set, stop every path at the first date where the same rule the reference applies says "exercise"
// (in the money by > 1e-14 and !(immediate < fitted continuation), :55,:85; the last column always pays off, :37-40; dates
// past maturity cannot be exercised, :43-49) and average the discounted realised payoffs.  Paths are independent given the
// coefficients, so sqrt(var / N) IS the standard error of this mean; as the value of a feasible stopping rule it is a lower
// bound of the true price in expectation, where the in-sample value-iteration number is biased high.
// One streaming pass, thread per path, 4 B per path-step until the path has stopped.
// ---------------------------------------------------------------------------------------------------------
template <typename ST, int P>
__global__ void __launch_bounds__(LSM_NT) lsm_policy_kernel(const ST* __restrict__ S, int64_t ld, int64_t n, int M, const double* __restrict__ tab /*[M][COEF_LD + 2]*/,
                                                          const int* __restrict__ kind, double K, int is_call, double disc, double* __restrict__ partial) {
    // A CTA owns tiles of POL_PPT * LSM_NT consecutive paths and walks DOWN the rows with them (row j of the whole tile, then row
    // j + 1): every thread has POL_PPT independent loads in flight and a row's 8 KB come from one page -- a thread that walks one
    // path through 253 rows that lie 268 MB apart is latency-bound (82 ms at 2^26 paths; 37 ms this way, with the row's coefficients in registers).
    constexpr int POL_PPT = 8;
    double acc[3] = {0.0, 0.0, 0.0};  // sum of discounted payoffs, sum of squares, sum of stopping indices
    const int64_t tile_paths = (int64_t)POL_PPT * LSM_NT;
    for (int64_t t0 = (int64_t)blockIdx.x * tile_paths; t0 < n; t0 += (int64_t)gridDim.x * tile_paths) {
        double val[POL_PPT];
        int tau[POL_PPT];
        unsigned alive = 0;
#pragma unroll
        for (int q = 0; q < POL_PPT; ++q) {
            val[q] = 0.0;
            tau[q] = M - 1;
            if (t0 + q * LSM_NT + threadIdx.x < n) alive |= 1u << q;
        }
        const unsigned mine = alive;
        double df = 1.0;
        const ST* col = S + t0 + threadIdx.x;
        ST nxt[POL_PPT];  // row j + 1 is in flight while row j is evaluated: two rows of independent loads per thread
#pragma unroll
        for (int q = 0; q < POL_PPT; ++q) nxt[q] = (alive >> q) & 1u ? col[q * LSM_NT] : (ST)0;
        for (int j = 0; j < M; ++j, df *= disc) {
            if (__all_sync(0xffffffffu, alive == 0u)) break;  // every path of this warp has stopped
            ST buf[POL_PPT];
#pragma unroll
            for (int q = 0; q < POL_PPT; ++q) buf[q] = nxt[q];
            if (j + 1 < M) {
                const ST* row = col + (int64_t)(j + 1) * ld;
#pragma unroll
                for (int q = 0; q < POL_PPT; ++q) nxt[q] = (alive >> q) & 1u ? row[q * LSM_NT] : (ST)0;
            }
            const bool last = j == M - 1, normal = __ldg(kind + j) == STEP_NORMAL;
            if (!last && !normal) continue;
            const double* c = tab + (size_t)j * (COEF_LD + 2);
            const double mu = __ldg(c + COEF_LD), is = __ldg(c + COEF_LD + 1);
            double cj[P + 1];  // the row's coefficients once per row, in registers (one uniform load each instead of one per path)
#pragma unroll
            for (int k = 0; k <= P; ++k) cj[k] = __ldg(c + k);
#pragma unroll
            for (int q = 0; q < POL_PPT; ++q) {
                if (!((alive >> q) & 1u)) continue;
                const double s = (double)buf[q];
                const double pay = payoff_fn(is_call, s, K);
                bool stop = last;
                if (!last && pay > 1e-14) {
                    const double x = (s - mu) * is;
                    double cont = cj[P];
#pragma unroll
                    for (int k = P - 1; k >= 0; --k) cont = fma(cont, x, cj[k]);
                    stop = !(pay < cont);
                }
                if (stop) { val[q] = df * pay; tau[q] = j; alive &= ~(1u << q); }
            }
        }
#pragma unroll
        for (int q = 0; q < POL_PPT; ++q)
            if ((mine >> q) & 1u) {
                acc[0] += val[q];
                acc[1] = fma(val[q], val[q], acc[1]);
                acc[2] += (double)tau[q];
            }
    }
    block_reduce_to_partial<3>(acc, partial + (int64_t)blockIdx.x * MOM_LD);
}

extern "C" int mcp_lsm_policy_value(mcp_ctx* ctx, const mcp_pathset* ps, const mcp_lsm_params* prm, const double* coeffs_std, mcp_lsm_result* res,
                                    double* mean_stop_index) {
    if (!ctx || !prm || !coeffs_std || !res) return MCP_ERR_INVALID;
    if (!ps || ps->n_paths <= 0) return mcp_fail(ctx, MCP_ERR_EMPTY_PATHS, "LSM::PredictOptionPrice: Empty pricePaths.");
    if (ps->ctx != ctx) return mcp_fail(ctx, MCP_ERR_INVALID, "lsm: pathset belongs to another ctx");
    const int p = prm->poly_order;
    if (p < 0 || p > MAXP) return mcp_fail(ctx, MCP_ERR_UNSUPPORTED, "lsm: poly_order %d outside [0, %d]", p, MAXP);
    MCP_CUDA(ctx, cudaSetDevice(ctx->device));
    const int M = ps->n_steps + 1;
    const int64_t N = ps->n_paths;
    int64_t grid = (N + (int64_t)8 * LSM_NT - 1) / ((int64_t)8 * LSM_NT);  // POL_PPT paths per thread
    if (grid > (int64_t)ctx->sm_count * 8) grid = (int64_t)ctx->sm_count * 8;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_tab = take((size_t)M * (COEF_LD + 2) * 8), o_kind = take((size_t)M * 4), o_part = take((size_t)grid * MOM_LD * 8), o_fin = take(4 * 8);
    MCP_TRY(mcp_scratch_reserve(ctx, off));
    unsigned char* sb = (unsigned char*)ctx->scratch;
    std::vector<double> tab((size_t)M * (COEF_LD + 2), 0.0);
    for (int j = 0; j + 1 < M; ++j) {  // caller's rows: [c_0 .. c_p, mu, 1/s]
        const double* row = coeffs_std + (size_t)j * (p + 3);
        for (int k = 0; k <= p; ++k) tab[(size_t)j * (COEF_LD + 2) + k] = row[k];
        tab[(size_t)j * (COEF_LD + 2) + COEF_LD] = row[p + 1];
        tab[(size_t)j * (COEF_LD + 2) + COEF_LD + 1] = row[p + 2];
    }
    std::vector<int> kind(M, STEP_NORMAL);
    for (int j = 0; j < M; ++j) kind[j] = ((double)j * prm->dt > prm->maturity) ? STEP_DISCOUNT : STEP_NORMAL;  // LSMPricer.cpp:43-44
    cudaStream_t st = ctx->stream;
    const uint64_t launches0 = ctx->launches;
    MCP_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    MCP_TRY(mcp_h2d(ctx, sb + o_tab, tab.data(), tab.size() * 8));
    MCP_TRY(mcp_h2d(ctx, sb + o_kind, kind.data(), (size_t)M * 4));
    const double disc = exp(-prm->r * prm->dt);
    double* d_part = (double*)(sb + o_part);
    double* d_fin = (double*)(sb + o_fin);
    {
        const double* d_tab = (const double*)(sb + o_tab);
        const int* d_kind = (const int*)(sb + o_kind);
        auto launch = [&](auto tag) {
            constexpr int P = decltype(tag)::value;
            if (ps->dtype == MCP_F32)
                lsm_policy_kernel<float, P><<<(unsigned)grid, LSM_NT, 0, st>>>((const float*)ps->data, ps->ld, N, M, d_tab, d_kind, prm->strike, prm->is_call, disc, d_part);
            else
                lsm_policy_kernel<double, P><<<(unsigned)grid, LSM_NT, 0, st>>>((const double*)ps->data, ps->ld, N, M, d_tab, d_kind, prm->strike, prm->is_call, disc, d_part);
        };
        switch (p) {
            case 0: launch(std::integral_constant<int, 0>{}); break;
            case 1: launch(std::integral_constant<int, 1>{}); break;
            case 2: launch(std::integral_constant<int, 2>{}); break;
            case 3: launch(std::integral_constant<int, 3>{}); break;
            case 4: launch(std::integral_constant<int, 4>{}); break;
            case 5: launch(std::integral_constant<int, 5>{}); break;
            default: launch(std::integral_constant<int, 6>{}); break;
        }
    }
    MCP_LAUNCH_CHECK(ctx);
    lsm_reduce_kernel<<<1, 256, 0, st>>>(d_part, (int)grid, 3, d_fin);
    MCP_LAUNCH_CHECK(ctx);
    const double nloc = (double)N;
    MCP_TRY(mcp_h2d(ctx, d_fin + 3, &nloc, 8));
    MCP_TRY(mcp_allreduce_f64(ctx, d_fin, 4));  // path-sharded ranks: global sums (no-op without a communicator)
    MCP_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    double fin[4] = {0, 0, 0, 0};
    double* fp = (double*)mcp_stage_alloc(ctx, sizeof(fin));
    MCP_CUDA(ctx, mcp_memcpy_async(ctx, fp ? fp : fin, d_fin, sizeof(fin), cudaMemcpyDeviceToHost, st));
    MCP_CUDA(ctx, cudaStreamSynchronize(st));
    MCP_CUDA(ctx, cudaGetLastError());
    if (fp) memcpy(fin, fp, sizeof(fin));
    const double ng = fin[3], mean = fin[0] / ng;
    res->price = mean;
    res->sum_v0 = fin[0];
    res->sum_sq_dev = fin[1] - ng * mean * mean;  // one pass: sum of squares minus N mean^2 (payoffs are O(K), N mean^2 / sum sq ~ 1e-1: no cancellation trouble in fp64)
    if (res->sum_sq_dev < 0.0) res->sum_sq_dev = 0.0;
    res->n_paths_global = (int64_t)llround(ng);
    res->std_error = ng > 1.0 ? sqrt(res->sum_sq_dev / (ng - 1.0) / ng) : 0.0;
    float ms = 0.f;
    MCP_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    res->elapsed_ms = ms;
    res->n_kernel_launches = (int)(ctx->launches - launches0);
    if (mean_stop_index) *mean_stop_index = fin[2] / ng;
    return MCP_OK;
}

extern "C" int mcp_lsm_price_host_rows(mcp_ctx* ctx, const double* const* rows, int64_t n_paths, int n_cols, double r, double strike,
                                       double maturity, double dt, int is_call, int poly_order, double* price) {
    if (!ctx || !price) return MCP_ERR_INVALID;
    if (!rows || n_paths <= 0 || n_cols <= 0) return mcp_fail(ctx, MCP_ERR_EMPTY_PATHS, "LSM::PredictOptionPrice: Empty pricePaths.");
    mcp_pathset* ps = nullptr;
    MCP_TRY(mcp_pathset_create(ctx, n_paths, n_cols - 1, MCP_F64, &ps));
    int rc = mcp_pathset_upload_rows_f64(ps, rows);
    if (rc == MCP_OK) {
        mcp_lsm_params prm;
        prm.r = r; prm.strike = strike; prm.maturity = maturity; prm.dt = dt;
        prm.is_call = is_call; prm.poly_order = poly_order; prm.basis = MCP_BASIS_MONOMIAL; prm.carry = MCP_F64;
        mcp_lsm_result res;
        rc = mcp_lsm_price(ctx, ps, &prm, &res, nullptr, nullptr, nullptr);
        if (rc == MCP_OK) *price = res.price;
    }
    mcp_pathset_destroy(ps);
    return rc;
}

extern "C" int mcp_price_rbergomi_lsm(mcp_ctx* ctx, const mcp_rbergomi_params* model, const mcp_lsm_params* lsm, int64_t n_paths, int n_steps,
                                      uint64_t seed, uint64_t path_offset, mcp_lsm_result* res, float* gen_ms) {
    if (!ctx || !model || !lsm || !res) return MCP_ERR_INVALID;
    mcp_pathset* ps = ctx->cached_ps;
    if (!ps || ps->n_paths != n_paths || ps->n_steps != n_steps || ps->dtype != MCP_F32) {
        if (ps) mcp_pathset_destroy(ps);
        ctx->cached_ps = nullptr;
        MCP_TRY(mcp_pathset_create(ctx, n_paths, n_steps, MCP_F32, &ps));
        ctx->cached_ps = ps;
    }
    // events from the ctx pool (created once, slots far above the per-step profiling events): no create / destroy per call
    cudaEvent_t g0 = mcp_prof_event(ctx, 4094), g1 = mcp_prof_event(ctx, 4095);
    if (!g0 || !g1) return mcp_fail(ctx, MCP_ERR_CUDA, "cudaEventCreate failed");
    MCP_CUDA(ctx, cudaEventRecord(g0, ctx->stream));
    int rc = mcp_gen_rbergomi(ctx, ps, model, seed, path_offset, nullptr, nullptr);
    if (rc == MCP_OK) {
        cudaEventRecord(g1, ctx->stream);
        rc = mcp_lsm_price(ctx, ps, lsm, res, nullptr, nullptr, nullptr);
        if (rc == MCP_OK && gen_ms) cudaEventElapsedTime(gen_ms, g0, g1);
    }
    return rc;
}
