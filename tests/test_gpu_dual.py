"""BASELINE config 4 [new -- the reference has no such algorithm]: nested-simulation duality (Andersen-Broadie) under
GBM.  Parity is unpinned by the reference; the oracle below is an independent restatement of the definition in
include/mcp_b200.h that replays the same Philox streams, plus the properties any valid pair of bounds has."""
import numpy as np
import pytest

import montecarlooptionspricer_b200 as m

pytestmark = pytest.mark.gpu
MASK64 = (1 << 64) - 1


def oracle_dual(engine, port, S0, r, sigma, dt, K, is_call, n, p, n_policy, n_outer, n_inner, seed):
    f32 = np.float32
    # policy: the same LSM fit the entry point runs internally (independent path set, seed ^ 1)
    pp = engine.pathset(n_policy, n)
    engine.gen_gbm(pp, S0, r, sigma, dt, seed=seed ^ 1)
    fit = engine.lsm_price(pp, r, K, n * dt + dt, dt, is_call, p, basis=m.capi.MCP_BASIS_STANDARDISED, carry=m.MCP_F64, want_coeffs=True)
    pp.close()
    co = fit.coeffs.astype(f32)  # [n][p+3]
    po = engine.pathset(n_outer, n)
    engine.gen_gbm(po, S0, r, sigma, dt, seed=seed)
    S = po.download_timemajor()  # [n+1][n_outer] fp32, exactly what the device sees
    po.close()
    disc = np.array([np.exp(-r * j * dt) for j in range(n + 1)]).astype(f32)
    log2e = 1.4426950408889634074
    drift2, vol2 = f32((r - 0.5 * sigma * sigma) * dt * log2e), f32(sigma * np.sqrt(dt) * log2e)
    Kf = f32(K)

    def payoff(s):
        return max(f32(s - Kf) if is_call else f32(Kf - s), f32(0))

    def exercises(j, s, pay):
        if j == n:
            return pay > 0
        if not pay > f32(1e-14):
            return False
        x = f32(f32(s - co[j, p + 1]) * co[j, p + 2])
        cont = co[j, p]
        for k in range(p - 1, -1, -1):
            cont = f32(np.float32(cont) * x + co[j, k])  # fmaf on the device: the difference is below the decision noise
        return not pay < cont

    seed_in = (seed ^ 0x9E3779B97F4A7C15) & MASK64
    key = [seed_in & 0xffffffff, seed_in >> 32]
    Q = np.zeros((n, n_outer))
    for i in range(n_outer):
        for j in range(n):
            tot = 0.0
            for k in range(n_inner):
                s, val, q = f32(S[j, i]), f32(0), 0
                for mstep in range(j + 1, n + 1):
                    if q % 4 == 0:
                        u = port.philox([i & 0xffffffff, i >> 32, j * n_inner + k, 0x30000 + q // 4], key)
                        z = list(port.box_muller(int(u[0]), int(u[1]))) + list(port.box_muller(int(u[2]), int(u[3])))
                    s = f32(s * f32(2.0 ** float(f32(vol2 * f32(z[q % 4]) + drift2))))
                    pay = payoff(s)
                    ex = exercises(mstep, s, pay)
                    q += 1
                    if ex or mstep == n:
                        val = f32(pay * disc[mstep]) if ex else f32(0)
                        break
                tot = f32(tot + val)
            Q[j, i] = float(f32(tot / f32(n_inner)))
    lower, dual = np.zeros(n_outer), np.zeros(n_outer)
    for i in range(n_outer):
        M, best, qprev, stopped = 0.0, -1e300, 0.0, False
        for j in range(n + 1):
            pay = payoff(f32(S[j, i]))
            h = float(pay) * float(disc[j])
            ex = exercises(j, f32(S[j, i]), pay)
            q = Q[j, i] if j < n else 0.0
            L = (h if ex else 0.0) if (ex or j == n) else q
            if j > 0:
                M += L - qprev
            best = max(best, h - M)
            if not stopped and (ex or j == n):
                lower[i], stopped = (h if ex else 0.0), True
            qprev = q
        dual[i] = best
    return lower.mean(), dual.mean()


def test_nested_dual_matches_independent_restatement(engine, port):
    args = dict(S0=100.0, r=0.05, sigma=0.2, dt=1.0 / 6, strike=100.0, is_call=False, n_steps=6, poly_order=2,
                n_policy_paths=4000, n_outer=48, n_inner=8, seed=21)
    got = engine.gbm_nested_dual(**args)
    lo, up = oracle_dual(engine, port, 100.0, 0.05, 0.2, 1.0 / 6, 100.0, False, 6, 2, 4000, 48, 8, 21)
    assert got["n_outer_global"] == 48
    assert got["lower"] == pytest.approx(lo, rel=2e-3, abs=1e-4)   # fp32 normals (SFU) vs fp64 restatement: rare near-tie flips
    assert got["upper"] == pytest.approx(up, rel=2e-3, abs=1e-4)


def test_nested_dual_bounds_bracket_the_bermudan_value(engine):
    """Config 1 model (S0 = K = 100, r = 5 %, sigma = 20 %, T = 1, 50 dates): the Bermudan put is worth 6.0786.
    The policy's lower bound sits below it, the dual upper bound above, the duality gap is small and shrinks with more
    inner paths."""
    res = engine.gbm_nested_dual(S0=100.0, r=0.05, sigma=0.2, dt=0.02, strike=100.0, is_call=False, n_steps=50, poly_order=3,
                                 n_policy_paths=1 << 18, n_outer=1 << 14, n_inner=256, seed=5)
    assert res["lower"] - 4 * res["lower_se"] < 6.0786 < res["upper"] + 4 * res["upper_se"]
    assert res["lower"] < res["upper"]
    assert res["upper"] - res["lower"] < 0.3   # the nested estimate biases the upper bound up by O(1 / n_inner)
    assert res["lower"] > 5.9 and res["nested_ms"] > 0
    fine = engine.gbm_nested_dual(S0=100.0, r=0.05, sigma=0.2, dt=0.02, strike=100.0, is_call=False, n_steps=50, poly_order=3,
                                  n_policy_paths=1 << 18, n_outer=1 << 14, n_inner=2048, seed=5)
    assert fine["upper"] < res["upper"] and fine["upper"] + 4 * fine["upper_se"] > 6.0786
    assert fine["lower"] == res["lower"]  # same outer paths, same policy
    print(f"nested dual 2^14 outer x 256 inner x 50 dates: lower {res['lower']:.4f} +- {res['lower_se']:.4f}, upper {res['upper']:.4f} +- "
          f"{res['upper_se']:.4f}; nested kernel {res['nested_ms']:.1f} ms")
