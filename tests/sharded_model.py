"""TEST INFRASTRUCTURE: a numpy model of the DEVICE algorithm (moment-based LSM with standardised basis, sharded
over ranks with an all-reduce of the moments) -- used by the gloo world_size-2 test to validate the multi-GPU
host logic on CPU, against the single-process oracle.  Mirrors montecarlooptionspricer_b200/csrc/lsm.cu."""
import numpy as np

SAMPLE_MAX = 16384


def solve_moments(mom, p):
    """lsm_solve_kernel: equilibrated Gram matrix; eigen pseudo-inverse with a relative cut (the Jacobi branch)."""
    n = p + 1
    if not mom[0] > 0:
        return np.zeros(n)
    d = np.array([1 / np.sqrt(mom[2 * a]) if mom[2 * a] > 0 else 0.0 for a in range(n)])
    G = np.array([[mom[a + b] * d[a] * d[b] for b in range(n)] for a in range(n)])
    rhs = np.array([mom[2 * p + 1 + a] * d[a] for a in range(n)])
    lam, Q = np.linalg.eigh(G)
    thr = lam.max() * n * 64 * np.finfo(float).eps
    z = sum(Q[:, e] * (Q[:, e] @ rhs) / lam[e] for e in range(n) if lam[e] > thr)
    return z * d


def sharded_lsm(paths_local, r, K, T, dt, is_call, p, allreduce):
    """paths_local [N_loc][M]; allreduce(np.ndarray) -> summed over ranks.  Returns price, stderr, first_ex (local)."""
    N, M = paths_local.shape
    pay = (lambda S: np.maximum(S - K, 0.0)) if is_call else (lambda S: np.maximum(K - S, 0.0))
    disc = np.exp(-r * dt)
    # standardisation from a fixed leading sample of every rank's shard, summed over ranks
    ns = min(N, SAMPLE_MAX)
    ssum = np.zeros((M, 3))
    for j in range(M):
        s = paths_local[:ns, j]
        m = pay(s) > 1e-14
        ssum[j] = [m.sum(), s[m].sum(), (s[m] ** 2).sum()]
    ssum = allreduce(ssum)
    mu, inv_s = np.full(M, float(K)), np.full(M, 1.0 / abs(K))
    for j in range(M):
        cnt, s1, s2 = ssum[j]
        if cnt >= 2:
            m = s1 / cnt
            var = (s2 - cnt * m * m) / (cnt - 1)
            sd = np.sqrt(var) if var > 1e-12 * m * m else (abs(m) if m != 0 else 1.0)
            mu[j], inv_s[j] = m, 1.0 / sd
    V = pay(paths_local[:, M - 1])
    first = np.full(N, M - 1, dtype=np.int32)
    coef = np.zeros(p + 1)
    for j in range(M - 1, -1, -1):
        if j < M - 1:
            if j * dt > T:
                V = V * disc
            else:
                S = paths_local[:, j]
                im = pay(S)
                x = (S - mu[j]) * inv_s[j]
                cont = np.polyval(coef[::-1], x)
                itm, otm = im > 1e-14, im < 1e-14
                ex = itm & ~(im < cont)
                Vn = np.zeros(N)
                Vn[itm] = np.where(ex[itm], im[itm], cont[itm])
                Vn[otm] = V[otm] * disc
                first[ex] = j
                V = Vn
        if j > 0 and not ((j - 1) * dt > T):
            Sp = paths_local[:, j - 1]
            m = pay(Sp) > 1e-14
            x, y = (Sp[m] - mu[j - 1]) * inv_s[j - 1], V[m] * disc
            mom = np.array([np.sum(x ** k) for k in range(2 * p + 1)] + [np.sum(x ** k * y) for k in range(p + 1)])
            mom = allreduce(mom)
            coef = solve_moments(mom, p)
    tot = allreduce(np.array([V.sum(), float(N)]))
    mean = tot[0] / tot[1]
    sq = allreduce(np.array([np.sum((V - mean) ** 2)]))[0]
    n = tot[1]
    return mean, np.sqrt(sq / (n - 1) / n) if n > 1 else 0.0, first
