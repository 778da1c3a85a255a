"""Out-of-sample value of the fitted exercise policy (mcp_lsm_policy_value) [new: the reference computes no error estimate].

  * against a numpy restatement of the stopping rule the reference's LSM applies (LSMPricer.cpp:37-49,55,85) on the same
    device paths and the same coefficient table: mean payoff to 1e-12, identical mean stopping column;
  * properties: on an independent sample the policy value is a LOWER bound in expectation (config 1: below the known
    Bermudan-50 value 6.0786 within its own -- exact -- standard error, and close to it), the in-sample value-iteration price of
    the reference's estimator lies above it, and the honest standard error is wider than the in-sample one."""
import numpy as np
import pytest

import montecarlooptionspricer_b200 as m
from conftest import CFG2

pytestmark = pytest.mark.gpu


def policy_value_numpy(slab, tab, r, K, T, dt, is_call, p):
    """slab [M][N] (time-major), tab [M-1][p+3] standardised rows [c_0..c_p, mu, 1/s]."""
    M, N = slab.shape
    S = slab.astype(np.float64)
    disc = np.exp(-r * dt)
    val = np.zeros(N)
    tau = np.full(N, M - 1)
    alive = np.ones(N, dtype=bool)
    for j in range(M):
        pay = np.maximum(S[j] - K, 0.0) if is_call else np.maximum(K - S[j], 0.0)
        if j == M - 1:
            val[alive] = disc ** j * pay[alive]
            break
        if j * dt > T:
            continue
        x = (S[j] - tab[j, p + 1]) * tab[j, p + 2]
        cont = np.full(N, tab[j, p])
        for k in range(p - 1, -1, -1):
            cont = cont * x + tab[j, k]
        ex = alive & (pay > 1e-14) & ~(pay < cont)
        val[ex] = disc ** j * pay[ex]
        tau[ex] = j
        alive &= ~ex
    return val, tau


@pytest.mark.parametrize("n_steps,maturity", [(50, 1.0), (37, 0.5)])
def test_policy_value_matches_the_rule_restated_in_numpy(engine, n_steps, maturity):
    n, dt = 1 << 14, 1.0 / 50
    fit = engine.pathset(n, n_steps)
    engine.gen_gbm(fit, 100.0, 0.05, 0.2, dt, seed=1)
    co = engine.lsm_price(fit, 0.05, 100.0, maturity, dt, False, 3, basis=m.MCP_BASIS_STANDARDISED, carry=m.MCP_F64, want_coeffs=True).coeffs
    fit.close()
    new = engine.pathset(n + 77, n_steps)  # ragged on purpose
    engine.gen_gbm(new, 100.0, 0.05, 0.2, dt, seed=2)
    got, stop = engine.lsm_policy_value(new, co, 0.05, 100.0, maturity, dt, False, 3)
    val, tau = policy_value_numpy(new.download_timemajor(), co, 0.05, 100.0, maturity, dt, False, 3)
    new.close()
    assert abs(got.price - val.mean()) <= 1e-12 * val.mean()
    assert abs(stop - tau.mean()) <= 1e-9 * tau.mean()
    assert abs(got.std_error - val.std(ddof=1) / np.sqrt(val.size)) <= 1e-8 * got.std_error
    assert got.n_paths_global == n + 77


def test_policy_value_is_a_lower_bound_with_an_honest_error_bar_config1(engine):
    n, n_steps, dt = 1 << 20, 50, 1.0 / 50
    fit = engine.pathset(n, n_steps)
    engine.gen_gbm(fit, 100.0, 0.05, 0.2, dt, seed=11)
    ins = engine.lsm_price(fit, 0.05, 100.0, 1.0, dt, False, 3, basis=m.MCP_BASIS_STANDARDISED, carry=m.MCP_F32, want_coeffs=True)
    engine.gen_gbm(fit, 100.0, 0.05, 0.2, dt, seed=12)  # an independent sample in the same slab
    oos, stop = engine.lsm_policy_value(fit, ins.coeffs, 0.05, 100.0, 1.0, dt, False, 3)
    fit.close()
    bermudan50 = 6.0786  # SURVEY 8c known answer
    assert oos.price < bermudan50 + 4 * oos.std_error, (oos.price, oos.std_error)
    assert oos.price > bermudan50 - 0.03, "a cubic policy fitted on 2^20 paths loses well under 3 cents"
    assert ins.price > oos.price - 4 * oos.std_error, "the value-iteration estimate is biased high, the policy value low"
    assert oos.std_error > ins.std_error, "realised cash flows scatter more than fitted values"
    assert 0 < stop < n_steps


def test_policy_value_config3_shape(engine):
    """rBergomi, 252 steps.  Measured: in-sample 2.565 (std_error 0.0009) against an out-of-sample policy value of 2.443 +- 0.0033.
    The reference's estimator takes max(immediate, fitted) at 252 dates with a regression on S alone, although the model's
    state also holds the (rough) variance -- its value-iteration price sits ~5 % above what the fitted policy actually earns.
    The test pins that picture: the policy value is the lower of the two, by less than 8 %."""
    n, n_steps = 1 << 20, CFG2["n"]
    args = (CFG2["S0"], CFG2["r"], CFG2["xi"], CFG2["H"], CFG2["eta"], CFG2["rho"], CFG2["dt"])
    ps = engine.pathset(n, n_steps)
    engine.gen_rbergomi(ps, *args, seed=21)
    ins = engine.lsm_price(ps, 0.05, 100.0, 1.0, CFG2["dt"], False, 3, basis=m.MCP_BASIS_STANDARDISED, carry=m.MCP_F32, want_coeffs=True)
    engine.gen_rbergomi(ps, *args, seed=22)
    oos, _ = engine.lsm_policy_value(ps, ins.coeffs, 0.05, 100.0, 1.0, CFG2["dt"], False, 3)
    ps.close()
    assert oos.price <= ins.price + 4 * oos.std_error
    assert 0.0 < ins.price - oos.price < 0.08 * ins.price, (ins.price, oos.price)
    assert oos.std_error > ins.std_error
