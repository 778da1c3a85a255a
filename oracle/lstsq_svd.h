/*
 * oracle/lstsq_svd.h -- TEST INFRASTRUCTURE (CPU oracle), never shipped or timed as product.
 *
 * Minimum-norm least-squares solve  x = argmin ||A x - b||_2  (min ||x||_2 among minimisers),
 * restating the arithmetic the reference obtains from the third-party call
 *
 *     A.bdcSvd(Eigen::ComputeThinU | Eigen::ComputeThinV).solve(b)
 *
 * at /root/reference/src/models/LSMPricer.cpp:76 and
 *    /root/reference/src/models/MartingaleOptimizationPricer.cpp:166.
 *
 * Eigen is NOT vendored in the reference (CMakeLists.txt:17 `find_package(Eigen3 REQUIRED)`, CI
 * resolves it to the distro libeigen3-dev == Eigen 3.4.0, .github/workflows/ci.yml:19) and is absent
 * from this image, so the published algorithm is restated here:
 *   - Eigen 3.4 BDCSVD delegates to JacobiSVD when cols < 16 (always true here: cols = polyOrder+1);
 *   - JacobiSVD preconditions a tall matrix with a column-pivoting Householder QR, then runs Jacobi
 *     sweeps on the small square factor;
 *   - SVDBase::solve() keeps singular values  sigma_i >= max(sigma_max * min(rows,cols) * eps, DBL_MIN)
 *     and applies  V * Sigma^+ * U^T * b  on that rank.
 * We use column-pivoted Householder QR followed by one-sided (Hestenes) Jacobi on R.  The singular
 * values/vectors of A are unique up to sign/rotation inside equal-sigma subspaces, so the min-norm
 * solution is the same real vector up to rounding (~ eps * cond(A)).
 *
 * Plain C99, usable from C and C++.  Row-major A (m x n), n <= ORC_LSQ_MAXN.
 */
#ifndef ORC_LSTSQ_SVD_H
#define ORC_LSTSQ_SVD_H

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_LSQ_MAXN 16

/* Returns the numerical rank used; writes x[n]; optional sv[n] (descending singular values). */
static inline int orc_lstsq_minnorm(const double *A_in, long m, int n, const double *b_in, double *x,
                                    double *sv_out)
{
    long mp = m < n ? n : m; /* zero-pad short systems: same singular values + zeros, same min-norm x */
    double *A = (double *)calloc((size_t)mp * (size_t)n, sizeof(double));
    double *b = (double *)calloc((size_t)mp, sizeof(double));
    int perm[ORC_LSQ_MAXN];
    double R[ORC_LSQ_MAXN][ORC_LSQ_MAXN], V[ORC_LSQ_MAXN][ORC_LSQ_MAXN], W[ORC_LSQ_MAXN][ORC_LSQ_MAXN];
    double qtb[ORC_LSQ_MAXN];
    int i, j, k, rank;
    memcpy(A, A_in, (size_t)m * (size_t)n * sizeof(double));
    memcpy(b, b_in, (size_t)m * sizeof(double));
    for (j = 0; j < n; ++j) perm[j] = j;

    /* --- column-pivoted Householder QR, applied to b on the fly --- */
    for (k = 0; k < n; ++k) {
        /* pivot: remaining column with the largest trailing norm (recomputed: O(m n^2) is fine here) */
        int piv = k;
        double best = -1.0;
        for (j = k; j < n; ++j) {
            double s = 0.0;
            long r;
            for (r = k; r < mp; ++r) s += A[r * n + j] * A[r * n + j];
            if (s > best) { best = s; piv = j; }
        }
        if (piv != k) {
            long r;
            for (r = 0; r < mp; ++r) { double t = A[r * n + k]; A[r * n + k] = A[r * n + piv]; A[r * n + piv] = t; }
            i = perm[k]; perm[k] = perm[piv]; perm[piv] = i;
        }
        {
            double normx = sqrt(best), alpha, vnorm2, x0 = A[(long)k * n + k];
            long r;
            if (normx == 0.0) continue; /* whole trailing block is zero */
            alpha = x0 > 0.0 ? -normx : normx;
            /* v = x - alpha e1, stored in place of column k (rows k..mp-1) */
            A[(long)k * n + k] = x0 - alpha;
            vnorm2 = 0.0;
            for (r = k; r < mp; ++r) vnorm2 += A[r * n + k] * A[r * n + k];
            if (vnorm2 > 0.0) {
                for (j = k + 1; j < n; ++j) {
                    double s = 0.0, f;
                    for (r = k; r < mp; ++r) s += A[r * n + k] * A[r * n + j];
                    f = 2.0 * s / vnorm2;
                    for (r = k; r < mp; ++r) A[r * n + j] -= f * A[r * n + k];
                }
                {
                    double s = 0.0, f;
                    for (r = k; r < mp; ++r) s += A[r * n + k] * b[r];
                    f = 2.0 * s / vnorm2;
                    for (r = k; r < mp; ++r) b[r] -= f * A[r * n + k];
                }
            }
            /* finalise column k of R */
            A[(long)k * n + k] = alpha;
            for (r = k + 1; r < mp; ++r) A[r * n + k] = 0.0;
        }
    }
    for (i = 0; i < n; ++i) {
        for (j = 0; j < n; ++j) R[i][j] = (j >= i) ? A[(long)i * n + j] : 0.0;
        qtb[i] = b[i];
    }
    free(A);
    free(b);

    /* --- one-sided Jacobi on W = R: rotate column pairs until mutually orthogonal; R = U S V^T --- */
    for (i = 0; i < n; ++i)
        for (j = 0; j < n; ++j) { W[i][j] = R[i][j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
    {
        int sweep, rotated = 1;
        for (sweep = 0; sweep < 60 && rotated; ++sweep) {
            int p, q;
            rotated = 0;
            for (p = 0; p < n - 1; ++p)
                for (q = p + 1; q < n; ++q) {
                    double a = 0.0, bb = 0.0, c = 0.0, zeta, t, cs, sn;
                    for (i = 0; i < n; ++i) { a += W[i][p] * W[i][p]; bb += W[i][q] * W[i][q]; c += W[i][p] * W[i][q]; }
                    if (c == 0.0 || fabs(c) <= DBL_EPSILON * sqrt(a * bb)) continue;
                    rotated = 1;
                    zeta = (bb - a) / (2.0 * c);
                    t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    cs = 1.0 / sqrt(1.0 + t * t);
                    sn = cs * t;
                    for (i = 0; i < n; ++i) {
                        double wp = W[i][p], wq = W[i][q], vp = V[i][p], vq = V[i][q];
                        W[i][p] = cs * wp - sn * wq; W[i][q] = sn * wp + cs * wq;
                        V[i][p] = cs * vp - sn * vq; V[i][q] = sn * vp + cs * vq;
                    }
                }
        }
    }
    {
        double sig[ORC_LSQ_MAXN], utb[ORC_LSQ_MAXN], z[ORC_LSQ_MAXN], smax = 0.0, thr;
        int order[ORC_LSQ_MAXN];
        int diag = (int)(m < n ? m : n);
        for (j = 0; j < n; ++j) {
            double s = 0.0;
            for (i = 0; i < n; ++i) s += W[i][j] * W[i][j];
            sig[j] = sqrt(s);
            if (sig[j] > smax) smax = sig[j];
            order[j] = j;
        }
        /* Eigen SVDBase::rank(): keep sigma_i >= max(sigma_max * diagSize * eps, DBL_MIN) */
        thr = smax * (double)diag * DBL_EPSILON;
        if (thr < DBL_MIN) thr = DBL_MIN;
        rank = 0;
        for (j = 0; j < n; ++j) {
            double s = 0.0;
            if (sig[j] >= thr && sig[j] > 0.0) {
                for (i = 0; i < n; ++i) s += (W[i][j] / sig[j]) * qtb[i]; /* u_j^T (Q^T b) */
                utb[j] = s / sig[j];
                ++rank;
            } else {
                utb[j] = 0.0;
            }
        }
        for (i = 0; i < n; ++i) {
            double s = 0.0;
            for (j = 0; j < n; ++j) s += V[i][j] * utb[j];
            z[i] = s;
        }
        for (i = 0; i < n; ++i) x[perm[i]] = z[i];
        if (sv_out) {
            /* descending insertion sort */
            for (i = 1; i < n; ++i) {
                int oi = order[i];
                j = i - 1;
                while (j >= 0 && sig[order[j]] < sig[oi]) { order[j + 1] = order[j]; --j; }
                order[j + 1] = oi;
            }
            for (i = 0; i < n; ++i) sv_out[i] = sig[order[i]];
        }
    }
    return rank;
}

#endif /* ORC_LSTSQ_SVD_H */
