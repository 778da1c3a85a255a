/*
 * oracle/port/mcp_oracle.c -- TEST INFRASTRUCTURE: plain-C, double-precision CPU restatement of the
 * reference's Monte-Carlo hot path.  It is the CHECKER for the CUDA path (tests/, __graft_entry__.smoke(),
 * bench.py's cpu_baseline leg) and must never be imported, linked or executed by the product package.
 *
 * Pinning: the reference ships no tests/golden vectors (CMakeLists.txt:70-82 are `cmake -E echo`), so this
 * port is pinned against the reference ITSELF compiled here (oracle/_ref/libmcp_ref.so, see
 * oracle/Makefile) on injected draws -- tests/test_oracle_vs_ref.py -- and against committed fixtures
 * generated from that build (tests/golden/).
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 * Pieces marked [new] are required by BASELINE.json but have no reference counterpart (Philox streams,
 * GBM generator, standard error, exercise indices).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../lstsq_svd.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------------
 * [new] Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11).
 * Known-answer vectors from the Random123 distribution are checked in tests/test_philox.py.
 * ------------------------------------------------------------------------------------------------ */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    int r;
    for (r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0, p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* [new] Box-Muller on two 32-bit words: u1=(a+0.5)/2^32 in (0,1), u2=(b+0.5)/2^32;
 * z0 = sqrt(-2 ln u1) cos(2 pi u2), z1 = sqrt(-2 ln u1) sin(2 pi u2).  The GPU evaluates the same
 * formula in fp32 with hardware approximations (documented tolerance in tests/test_gpu_philox.py). */
void orc_box_muller(uint32_t a, uint32_t b, double *z0, double *z1)
{
    double u1 = ((double)a + 0.5) * (1.0 / 4294967296.0);
    double u2 = ((double)b + 0.5) * (1.0 / 4294967296.0);
    double rad = sqrt(-2.0 * log(u1));
    *z0 = rad * cos(2.0 * M_PI * u2);
    *z1 = rad * sin(2.0 * M_PI * u2);
}

/* [new] Native stream layout, rough-vol generator.  g = GLOBAL path id, key = (seed_lo, seed_hi):
 *   Z_k (complex):  x = Philox(ctr = (g_lo, g_hi, k>>1, 0));  k even: (x0,x1) -> (Zre_k, Zim_k), k odd: (x2,x3)
 *   W_k (real)   :  x = Philox(ctr = (g_lo, g_hi, k>>2, 2));  pair (x0,x1) serves k&3 in {0,1}, (x2,x3) serves {2,3};
 *                   within a pair the cosine branch is the even k, the sine branch the odd k.
 * The reference mixes two independent normals, dW = rho W1 + sqrt(1-rho^2) W2 (RoughVolatility.cpp:356-358), which
 * is again N(0,1): the native stream draws that ONE normal W_k directly (3 normals per path-step instead of 4; the
 * law of the paths is unchanged).  For replay through the reference's 4-slot interface the draws are written in
 * its consumption order (RoughVolatility.cpp:346-352) as
 *   draws[p][2k]=Zre_k, [2k+1]=Zim_k, [2n+k]=rho W_k, [3n+k]=sqrt(1-rho^2) W_k      (so that rho W1 + rho_c W2 = W_k). */
void orc_rbergomi_draws(uint64_t seed, uint64_t path0, long n_paths, int n, double rho, double *draws)
{
    long p;
    uint32_t key[2];
    const double rho_c = sqrt(1.0 - rho * rho);
    key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
    for (p = 0; p < n_paths; ++p) {
        uint64_t g = path0 + (uint64_t)p;
        double *d = draws + (size_t)p * 4 * (size_t)n;
        int k;
        for (k = 0; k < n; ++k) {
            uint32_t ctr[4], x[4];
            double a, b, w;
            ctr[0] = (uint32_t)g; ctr[1] = (uint32_t)(g >> 32); ctr[2] = (uint32_t)(k >> 1); ctr[3] = 0u;
            orc_philox4x32_10(ctr, key, x);
            if (k & 1) orc_box_muller(x[2], x[3], &d[2 * k], &d[2 * k + 1]);
            else orc_box_muller(x[0], x[1], &d[2 * k], &d[2 * k + 1]);
            ctr[2] = (uint32_t)(k >> 2); ctr[3] = 2u;
            orc_philox4x32_10(ctr, key, x);
            if (k & 2) orc_box_muller(x[2], x[3], &a, &b);
            else orc_box_muller(x[0], x[1], &a, &b);
            w = (k & 1) ? b : a;
            d[2 * n + k] = rho * w;
            d[3 * n + k] = rho_c * w;
        }
    }
}

/* [new] Native stream layout, GBM generator: one Philox call per (g, q) yields the normals of steps
 * 4q..4q+3:  ctr = (g_lo, g_hi, q, 1);  (x0,x1)->(z_{4q}, z_{4q+1}), (x2,x3)->(z_{4q+2}, z_{4q+3}). */
void orc_gbm_draws(uint64_t seed, uint64_t path0, long n_paths, int n, double *draws)
{
    long p;
    uint32_t key[2];
    key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
    for (p = 0; p < n_paths; ++p) {
        uint64_t g = path0 + (uint64_t)p;
        double *d = draws + (size_t)p * (size_t)n;
        int q;
        for (q = 0; 4 * q < n; ++q) {
            uint32_t ctr[4], x[4];
            double z[4];
            int i;
            ctr[0] = (uint32_t)g; ctr[1] = (uint32_t)(g >> 32); ctr[2] = (uint32_t)q; ctr[3] = 1u;
            orc_philox4x32_10(ctr, key, x);
            orc_box_muller(x[0], x[1], &z[0], &z[1]);
            orc_box_muller(x[2], x[3], &z[2], &z[3]);
            for (i = 0; i < 4 && 4 * q + i < n; ++i) d[4 * q + i] = z[i];
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Spectral "fGn" machinery.
 * ------------------------------------------------------------------------------------------------ */
static size_t next_pow2(size_t n) /* RoughVolatility.cpp:204-210 */
{
    size_t p = 1;
    while (p < n) p <<= 1;
    return p;
}

/* Radix-2 DFT, sign = +1: sum x e^{+i theta}, unscaled; sign = -1: e^{-i theta}, divided by n
 * (the conventions of RoughVolatility.cpp:171-202; twiddles evaluated directly instead of by the
 * reference's running product -- differences are O(1e-15)). */
static void dft_radix2(double *re, double *im, size_t n, int sign)
{
    size_t i, j, len;
    for (i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { double t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
    }
    for (len = 2; len <= n; len <<= 1) {
        size_t half = len >> 1, blk, q;
        for (q = 0; q < half; ++q) {
            double ang = (sign < 0 ? -1.0 : 1.0) * 2.0 * M_PI * (double)q / (double)len;
            double wr = cos(ang), wi = sin(ang);
            for (blk = 0; blk < n; blk += len) {
                size_t a = blk + q, b = a + half;
                double vr = re[b] * wr - im[b] * wi, vi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - vr; im[b] = im[a] - vi;
                re[a] += vr; im[a] += vi;
            }
        }
    }
    if (sign < 0)
        for (i = 0; i < n; ++i) { re[i] /= (double)n; im[i] /= (double)n; }
}

/* lambda_i = 0.5 t_i^{2H} (RoughVolatility.cpp:227-236), phi = DFT+(zero-pad(lambda) to nextPow2(n+1))
 * (RoughVolatility.cpp:212-225).  Returns M = length of phi; phi_re/phi_im must hold nextPow2(n+1). */
int orc_rbergomi_phi(int n, double H, double dt, double *phi_re, double *phi_im)
{
    size_t M = next_pow2((size_t)n + 1), i;
    for (i = 0; i < M; ++i) { phi_re[i] = 0.0; phi_im[i] = 0.0; }
    for (i = 0; i <= (size_t)n; ++i) phi_re[i] = 0.5 * pow((double)i * dt, 2.0 * H);
    dft_radix2(phi_re, phi_im, M, +1);
    return (int)M;
}

/* Rough-vol price paths with explicit parameters.  draws [P][4n] in reference order; out [P][n+1].
 *   A_k = phi_k Z_k (k<n), zero-padded to M'=nextPow2(n), DFT-, /M'   RoughVolatility.cpp:264-275
 *   X_k = sqrt(2H) eta Re(A_k), k<n                                    RoughVolatility.cpp:277-291
 *   v_k = xi exp(X_k - 0.5 eta^2 t_k^{2H})                             RoughVolatility.cpp:294-309
 *   S_j = S_{j-1} exp((r - v_{j-1}/2) dt + sqrt(max(0,v_{j-1})) sqrt(dt) (rho W1 + sqrt(1-rho^2) W2))
 *                                                                      RoughVolatility.cpp:354-364 */
int orc_rbergomi_paths(double S0, double r, double xi, double H, double eta, double rho, double dt, int n,
                       long n_paths, const double *draws, double *out, double *X_out, double *v_out)
{
    size_t M = next_pow2((size_t)n + 1), Mp = next_pow2((size_t)n);
    double *phi_re = (double *)malloc(sizeof(double) * M), *phi_im = (double *)malloc(sizeof(double) * M);
    double *a_re = (double *)malloc(sizeof(double) * Mp), *a_im = (double *)malloc(sizeof(double) * Mp);
    double *comp = (double *)malloc(sizeof(double) * (size_t)n);
    const double scale = sqrt(2.0 * H) * eta, sq_dt = sqrt(dt), rho_c = sqrt(1.0 - rho * rho);
    long p;
    int k;
    if (!phi_re || !phi_im || !a_re || !a_im || !comp) return -1;
    orc_rbergomi_phi(n, H, dt, phi_re, phi_im);
    for (k = 0; k < n; ++k) comp[k] = -0.5 * eta * eta * pow((double)k * dt, 2.0 * H);
    for (p = 0; p < n_paths; ++p) {
        const double *d = draws + (size_t)p * 4 * (size_t)n;
        double *S = out + (size_t)p * ((size_t)n + 1);
        for (k = 0; k < (int)Mp; ++k) { a_re[k] = 0.0; a_im[k] = 0.0; }
        for (k = 0; k < n; ++k) {
            double zr = d[2 * k], zi = d[2 * k + 1];
            a_re[k] = phi_re[k] * zr - phi_im[k] * zi;
            a_im[k] = phi_re[k] * zi + phi_im[k] * zr;
        }
        dft_radix2(a_re, a_im, Mp, -1);
        S[0] = S0;
        for (k = 0; k < n; ++k) {
            double X = scale * a_re[k];
            double v = xi * exp(X + comp[k]);
            double dW = rho * (sq_dt * d[2 * n + k]) + rho_c * (sq_dt * d[3 * n + k]);
            S[k + 1] = S[k] * exp((r - 0.5 * v) * dt + sqrt(v > 0.0 ? v : 0.0) * dW);
            if (X_out) X_out[(size_t)p * n + k] = X;
            if (v_out) v_out[(size_t)p * n + k] = v;
        }
    }
    free(phi_re); free(phi_im); free(a_re); free(a_im); free(comp);
    return 0;
}

/* [new] GBM generator (BASELINE config 1; the reference has none): the rough-vol recursion of
 * RoughVolatility.cpp:354-364 with constant variance v = sigma^2 and a single driving normal.
 * draws [P][n]; out [P][n+1]. */
int orc_gbm_paths(double S0, double r, double sigma, double dt, int n, long n_paths, const double *draws,
                  double *out)
{
    const double drift = (r - 0.5 * sigma * sigma) * dt, vol = sigma * sqrt(dt);
    long p;
    int j;
    for (p = 0; p < n_paths; ++p) {
        const double *d = draws + (size_t)p * (size_t)n;
        double *S = out + (size_t)p * ((size_t)n + 1);
        S[0] = S0;
        for (j = 1; j <= n; ++j) S[j] = S[j - 1] * exp(drift + vol * d[j - 1]);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * LSM (the reference's value-iteration variant), instrumented.
 * ------------------------------------------------------------------------------------------------ */
static double payoff_fn(int is_call, double S, double K) /* include/core/common.h:8-14 */
{
    double x = is_call ? S - K : K - S;
    return x > 0.0 ? x : 0.0;
}

/* LSMPricer.cpp:19-102.  paths [N][M] row-major (M = steps+1).  Outputs (all nullable except price):
 *   price      mean_i V[i][0]                                        LSMPricer.cpp:97-101
 *   stderr_out [new] sample std of V[:,0] / sqrt(N)
 *   coeffs     [M-1][p+1] raw-monomial min-norm coefficients c_j      LSMPricer.cpp:76  (zeros if no ITM / past maturity)
 *   first_ex   [new] tau_i = min{ j : payoff>1e-14 and !(payoff < cont) }, M-1 if never  
 *   ex_mask    [new] [N][M] bytes, 1 where the max() at LSMPricer.cpp:85 returned the immediate payoff
 *   V0         [N] V[i][0]
 *   min_gap    [new] min |payoff - cont| over all ITM decisions (near-tie diagnostic)
 *   n_itm      [M-1] ITM counts per step
 */
int orc_lsm(const double *paths, long N, long M, double r, double K, double T, double dt, int is_call, int p,
            double *price, double *stderr_out, double *coeffs, int32_t *first_ex, uint8_t *ex_mask, double *V0,
            double *min_gap, long *n_itm)
{
    double *V, *A, *b, c[ORC_LSQ_MAXN], gap = INFINITY;
    long *idx, i, j;
    int q;
    if (N <= 0 || M <= 0 || p < 0 || p + 1 > ORC_LSQ_MAXN) return -1; /* LSMPricer.cpp:28-30 throws on empty */
    V = (double *)malloc(sizeof(double) * (size_t)N);
    A = (double *)malloc(sizeof(double) * (size_t)N * (size_t)(p + 1));
    b = (double *)malloc(sizeof(double) * (size_t)N);
    idx = (long *)malloc(sizeof(long) * (size_t)N);
    if (!V || !A || !b || !idx) return -2;
    for (i = 0; i < N; ++i) { /* LSMPricer.cpp:37-40 */
        V[i] = payoff_fn(is_call, paths[i * M + (M - 1)], K);
        if (first_ex) first_ex[i] = (int32_t)(M - 1);
        if (ex_mask) memset(ex_mask + i * M, 0, (size_t)M);
    }
    if (coeffs) memset(coeffs, 0, sizeof(double) * (size_t)(M - 1) * (size_t)(p + 1));
    for (j = M - 2; j >= 0; --j) { /* LSMPricer.cpp:42 */
        const double disc = exp(-r * dt);
        long n = 0, kk;
        if ((double)j * dt > T) { /* LSMPricer.cpp:43-49 */
            for (i = 0; i < N; ++i) V[i] = V[i] * disc;
            if (n_itm) n_itm[j] = 0;
            continue;
        }
        for (i = 0; i < N; ++i) /* LSMPricer.cpp:51-58 */
            if (payoff_fn(is_call, paths[i * M + j], K) > 1e-14) idx[n++] = i;
        if (n_itm) n_itm[j] = n;
        if (n > 0) {
            for (kk = 0; kk < n; ++kk) { /* LSMPricer.cpp:61-74 (raw monomials, LSMPricer.cpp:9-17) */
                double S = paths[idx[kk] * M + j], pw = 1.0;
                b[kk] = V[idx[kk]] * disc;
                for (q = 0; q <= p; ++q) { A[kk * (p + 1) + q] = pw; pw *= S; }
            }
            orc_lstsq_minnorm(A, n, p + 1, b, c, NULL); /* LSMPricer.cpp:76 */
            if (coeffs) memcpy(coeffs + j * (p + 1), c, sizeof(double) * (size_t)(p + 1));
        }
        for (i = 0; i < N; ++i) {
            double S = paths[i * M + j], im = payoff_fn(is_call, S, K);
            if (im > 1e-14) { /* LSMPricer.cpp:78-86 */
                double cont = 0.0, pw = 1.0, g;
                for (q = 0; q <= p; ++q) { cont += pw * c[q]; pw *= S; }
                g = fabs(im - cont);
                if (g < gap) gap = g;
                if (!(im < cont)) { /* std::max(immediate, cont) returns immediate */
                    V[i] = im;
                    if (first_ex) first_ex[i] = (int32_t)j;
                    if (ex_mask) ex_mask[i * M + j] = 1;
                } else {
                    V[i] = cont;
                }
            } else if (im < 1e-14) { /* LSMPricer.cpp:89-94 */
                V[i] = V[i] * disc;
            } else {
                V[i] = 0.0; /* payoff == 1e-14 exactly: Values[i][j] keeps its initial 0 (LSMPricer.cpp:35) */
            }
        }
    }
    {
        double s = 0.0, s2 = 0.0, mean;
        for (i = 0; i < N; ++i) { s += V[i]; if (V0) V0[i] = V[i]; }
        mean = s / (double)N;
        for (i = 0; i < N; ++i) s2 += (V[i] - mean) * (V[i] - mean);
        *price = mean;
        if (stderr_out) *stderr_out = N > 1 ? sqrt(s2 / (double)(N - 1) / (double)N) : 0.0;
    }
    if (min_gap) *min_gap = gap;
    free(V); free(A); free(b); free(idx);
    return 0;
}

/* Time-major fp32 slab variant used by the large parity tests: S is [M][N] float (the GPU's own layout,
 * downloaded), widened to double element-wise -- i.e. the oracle prices exactly the values the GPU holds. */
int orc_lsm_timemajor_f32(const float *slab, long N, long M, double r, double K, double T, double dt, int is_call,
                          int p, double *price, double *stderr_out, double *coeffs, int32_t *first_ex, double *V0,
                          double *min_gap, long *n_itm)
{
    double *pm = (double *)malloc(sizeof(double) * (size_t)N * (size_t)M);
    long i, j;
    int rc;
    if (!pm) return -2;
    for (j = 0; j < M; ++j)
        for (i = 0; i < N; ++i) pm[i * M + j] = (double)slab[j * N + i];
    rc = orc_lsm(pm, N, M, r, K, T, dt, is_call, p, price, stderr_out, coeffs, first_ex, NULL, V0, min_gap, n_itm);
    free(pm);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * The other three pricers and the generator's parameter estimators, restated.
 * ------------------------------------------------------------------------------------------------ */

/* AsymptoticAnalysisPricer.cpp:8-36 -- closed-form early-exercise boundary */
static double asym_boundary(int is_call, double t, double T, double K, double r, double D, double sigma)
{
    double eps = T - t, c0, b;
    if (eps < 1e-10) return K;                          /* :10-11, :25-26 */
    c0 = 0.5 * sigma * sqrt(eps * log(1.0 / eps));      /* :13, :28 (NaN when eps > 1) */
    if (is_call) { b = K - c0; if (eps < 0.01) b += 0.5 * (D - r) * eps; }   /* :29-34 */
    else         { b = K + c0; if (eps < 0.01) b -= 0.5 * (r - D) * eps; }   /* :14-19 */
    return b;
}

/* AsymptoticAnalysisPricer.cpp:38-113.  Returns 0.0 on empty input (:48-50); sigma <= 0 is the caller's
 * std::runtime_error (:51-53) -> rc -3 here. */
int orc_asymptotic(const double *paths, long N, long M, double r, double K, double T, double dt, int is_call,
                   double sigma, double dividend, double *price)
{
    long i, j, valid = 0;
    double sum = 0.0;
    *price = 0.0;
    if (N <= 0 || M <= 0) return 0;
    if (!(sigma > 0.0)) return -3;
    for (i = 0; i < N; ++i) {
        double best = 0.0;
        for (j = 0; j < M; ++j) {                       /* :69-96 */
            double t = (double)j * dt, S, b, pay, d;
            int in;
            if (t > T) break;                           /* :71 */
            S = paths[i * M + j];
            if (isnan(S) || isinf(S)) continue;         /* :74 */
            b = asym_boundary(is_call, t, T, K, r, dividend, sigma);
            in = is_call ? (S > b) : (S < b);           /* :80-85 */
            if (!in) continue;
            pay = payoff_fn(is_call, S, K);
            if (isnan(pay) || isinf(pay)) continue;     /* :89 */
            d = exp(-r * t) * pay;                      /* :90 */
            if (d > best) best = d;
        }
        if (!isnan(best) && !isinf(best)) { sum += best; ++valid; }   /* :101-106 */
    }
    *price = valid > 0 ? sum / (double)valid : 0.0;     /* :108 */
    return 0;
}

static double mart_df(long j, double dt, double T, double r) /* MartingaleOptimizationPricer.h:44-49 */
{
    double t = (double)j * dt;
    if (t > T) t = T;
    return exp(-r * t);
}

static double mart_eval(const double *c, int p, double S) /* MartingaleOptimizationPricer.cpp:180-188 */
{
    double v = 0.0, pw = 1.0;
    int k;
    for (k = 0; k <= p; ++k) { v += c[k] * pw; pw *= S; }
    return v;
}

/* MartingaleOptimizationPricer.cpp:21-178, iteration by iteration exactly as written (primal, dual with the
 * previous martingale, refit).  primal/dual (nullable) receive the last iteration's bounds. */
int orc_martingale(const double *paths, long N, long M, double r, double K, double T, double dt, int is_call, int p,
                   int max_iter, double *price, double *primal_out, double *dual_out)
{
    double c[ORC_LSQ_MAXN] = {0}, offset = 0.0, lower = 0.0, upper = 0.0;
    long *stop, i, j;
    double *A, *b;
    int it, q;
    if (N <= 0 || M <= 0) return -1;                    /* :31-33 */
    if (max_iter <= 0) return -3;                       /* :34-36 */
    if (p < 0 || p + 1 > ORC_LSQ_MAXN) return -1;
    stop = (long *)malloc(sizeof(long) * (size_t)N);
    A = (double *)malloc(sizeof(double) * (size_t)(2 * N) * (size_t)(p + 1));
    b = (double *)malloc(sizeof(double) * (size_t)(2 * N));
    if (!stop || !A || !b) return -2;
    for (it = 1; it <= max_iter; ++it) {                /* :56-61 */
        double sp = 0.0, sd = 0.0, s0 = 0.0;
        for (i = 0; i < N; ++i) {                       /* :72-94 */
            double best = 0.0;
            long bi = 0;
            for (j = 0; j < M; ++j) {
                double dp;
                if ((double)j * dt > T) break;
                dp = payoff_fn(is_call, paths[i * M + j], K) * mart_df(j, dt, T, r);
                if (dp > best) { best = dp; bi = j; }
            }
            stop[i] = bi;
            sp += best;
        }
        for (i = 0; i < N; ++i) {                       /* :96-117 */
            double best = 0.0;
            for (j = 0; j < M; ++j) {
                double S, cand;
                if ((double)j * dt > T) break;
                S = paths[i * M + j];
                cand = payoff_fn(is_call, S, K) * mart_df(j, dt, T, r) - (mart_eval(c, p, S) - offset);
                if (cand > best) best = cand;
            }
            sd += best;
        }
        lower = sp / (double)N;
        upper = sd / (double)N;
        if (2 * N >= p + 1) {                           /* :122-178 UpdateMartingale */
            for (i = 0; i < N; ++i) {
                long js = stop[i], jo = (js + M / 2) % M;
                double Ss = paths[i * M + js], So = paths[i * M + jo], pw;
                b[2 * i] = 0.5 * (payoff_fn(is_call, Ss, K) * mart_df(js, dt, T, r));
                b[2 * i + 1] = 0.2 * (payoff_fn(is_call, So, K) * mart_df(jo, dt, T, r));
                for (pw = 1.0, q = 0; q <= p; ++q) { A[(2 * i) * (p + 1) + q] = pw; pw *= Ss; }
                for (pw = 1.0, q = 0; q <= p; ++q) { A[(2 * i + 1) * (p + 1) + q] = pw; pw *= So; }
            }
            orc_lstsq_minnorm(A, 2 * N, p + 1, b, c, NULL);   /* :166 */
            for (i = 0; i < N; ++i) s0 += mart_eval(c, p, paths[i * M]);
            offset = s0 / (double)N;                    /* :172-177 */
        }
    }
    *price = 0.5 * (lower + upper);                     /* :63 */
    if (primal_out) *primal_out = lower;
    if (dual_out) *dual_out = upper;
    free(stop); free(A); free(b);
    return 0;
}

/* BranchingProcessPricer.cpp:13-134 with the resampled path indices INJECTED (the reference draws them from a
 * std::mt19937 seeded by std::random_device, :84-86, shared between OpenMP threads).
 *   rp: [n_visit][N][num_branches], n_visit = exercise dates visited before the first t > maturity.
 * lower/upper nullable. */
int orc_branching(const double *paths, long N, long M, double r, double K, double T, double dt, int is_call,
                  int num_branches, const int *ex, int n_ex, const int32_t *rp, double *price, double *lower_out,
                  double *upper_out)
{
    long i;
    double sl = 0.0, su = 0.0;
    int e, b;
    if (N <= 0 || M <= 0) return -1;                    /* :22-24 */
    if (n_ex <= 0) return -3;                           /* :25-27 */
    if (!(K > 0.0)) return -4;                          /* :28-30 */
    for (i = 0; i < N; ++i) {                           /* lower bound :41-71 */
        double best = 0.0;
        for (e = 0; e < n_ex; ++e) {
            double t = (double)ex[e] * dt, d;
            if (t > T) break;
            d = exp(-r * t) * payoff_fn(is_call, paths[i * M + ex[e]], K);
            if (d > best) { best = d; break; }
        }
        sl += best;
    }
    for (i = 0; i < N; ++i) {                           /* upper bound :73-134 */
        double best = 0.0;
        for (e = 0; e < n_ex; ++e) {
            double t = (double)ex[e] * dt, now, cont = 0.0, better;
            if (t > T) break;
            now = exp(-r * t) * payoff_fn(is_call, paths[i * M + ex[e]], K);
            if (ex[e] < ex[n_ex - 1]) {                 /* :103 */
                double sf = 0.0;
                for (b = 0; b < num_branches; ++b) {
                    long q = rp[((long)e * N + i) * num_branches + b], k;
                    double bf = 0.0;
                    for (k = ex[e] + 1; k < M; ++k) {   /* :110-121 */
                        double tk = (double)k * dt, d;
                        if (tk > T) break;
                        d = exp(-r * (tk - t)) * payoff_fn(is_call, paths[q * M + k], K);
                        if (d > bf) bf = d;
                    }
                    sf += bf;
                }
                cont = (sf / (double)num_branches) * exp(-r * t);   /* :123 */
            }
            better = now < cont ? cont : now;           /* :126 */
            if (better > best) best = better;
        }
        su += best;
    }
    if (lower_out) *lower_out = sl / (double)N;
    if (upper_out) *upper_out = su / (double)N;
    *price = 0.5 * (sl / (double)N + su / (double)N);   /* :38 */
    return 0;
}

/* RoughVolatility.cpp:20-42 */
static double est_mean(const double *v, long n) { double s = 0.0; long i; for (i = 0; i < n; ++i) s += v[i]; return n ? s / (double)n : 0.0; }
static double est_var(const double *v, long n)
{
    double m, a = 0.0; long i;
    if (n < 2) return 0.0;
    m = est_mean(v, n);
    for (i = 0; i < n; ++i) a += (v[i] - m) * (v[i] - m);
    return a / (double)(n - 1);
}

/* RoughVolatility.cpp:72-169, :324-331 -> out7 = S0, r, xi, H, eta, rho, dt */
int orc_estimate_params(const double *hist, long n_hist, double *out7)
{
    long n = n_hist - 1, i, w, nw = 0;
    double *ret, *sq, *prof, *seg, lw[64], lf[64], var, m, mx, my, cov = 0.0, rho, H = 0.5;
    if (n_hist < 2) return -3;                          /* :317-319 */
    ret = (double *)malloc(sizeof(double) * (size_t)n * 4);
    if (!ret) return -2;
    sq = ret + n; prof = sq + n; seg = prof + n;
    for (i = 0; i < n; ++i) { ret[i] = log(hist[i + 1] / hist[i]); sq[i] = ret[i] * ret[i]; }   /* :126-133 */
    var = est_var(ret, n);
    mx = est_mean(ret, n); my = est_mean(sq, n);
    if (n >= 2) { for (i = 0; i < n; ++i) cov += (ret[i] - mx) * (sq[i] - my); cov /= (double)(n - 1); } else cov = 0.0;
    rho = cov / sqrt(var * est_var(sq, n));             /* :157-164 */
    if (rho > 0.0) rho = -0.3;                          /* :165-167 */
    if (n >= 2) {                                       /* DFA :72-122 */
        m = est_mean(ret, n);
        for (i = 0; i < n; ++i) prof[i] = ret[i] - m;
        for (i = 1; i < n; ++i) prof[i] += prof[i - 1];
        for (w = 4; w <= n / 4; w *= 2) {
            long start, cnt = 0;
            double fs = 0.0, tm = 0.0;
            for (i = 0; i < w; ++i) tm += (double)(i + 1);
            tm /= (double)w;
            for (start = 0; start + w <= n; start += w) {
                double ym = 0.0, num = 0.0, den = 0.0, ss = 0.0;
                for (i = 0; i < w; ++i) { seg[i] = prof[start + i]; ym += seg[i]; }
                ym /= (double)w;
                for (i = 0; i < w; ++i) { num += ((double)(i + 1) - tm) * (seg[i] - ym); den += ((double)(i + 1) - tm) * ((double)(i + 1) - tm); }
                if (!(fabs(den) < 1e-14)) {
                    double sl = num / den, ic = ym - sl * tm;
                    for (i = 0; i < w; ++i) seg[i] -= sl * (double)(i + 1) + ic;
                }
                for (i = 0; i < w; ++i) ss += seg[i] * seg[i];
                fs += sqrt(ss / (double)w);
                ++cnt;
            }
            if (cnt > 0 && fs / (double)cnt > 0.0 && nw < 64) { lw[nw] = log((double)w); lf[nw] = log(fs / (double)cnt); ++nw; }
        }
        if (nw >= 2) {
            double sx = 0, sy = 0, sxx = 0, sxy = 0;
            for (i = 0; i < nw; ++i) { sx += lw[i]; sy += lf[i]; sxx += lw[i] * lw[i]; sxy += lw[i] * lf[i]; }
            H = ((double)nw * sxy - sx * sy) / ((double)nw * sxx - sx * sx);
        }
    }
    out7[0] = hist[n_hist - 1]; out7[1] = 0.04; out7[2] = var / (1.0 / 252.0); out7[3] = H;
    out7[4] = 2.0 * sqrt(var); out7[5] = rho; out7[6] = 1.0 / 252.0;
    free(ret);
    return 0;
}
