"""GPU parity of the three non-LSM plugins and of the exact-signature generator against the CPU oracle.

All three pricers are deterministic functions of the path values (Branching once its resampled indices are
injected), evaluated in fp64 on both sides: tolerances are rounding-level.  The golden numbers are outputs of the
reference's own translation units (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

import montecarlooptionspricer_b200 as m

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    g = np.load(os.path.join(G, "pricers_ref.npz"))
    return g, g["paths_f32"].astype(np.float64)


def test_asymptotic_vs_golden_and_oracle(engine, port, gold):
    g, P = gold
    for dtype in (m.MCP_F64, m.MCP_F32):  # the values are fp32-representable: both slabs hold them exactly
        ps = engine.upload_paths(P, dtype=dtype)
        assert engine.asymptotic_price(ps, 0.05, 100.0, 1.0, 0.02, False, 0.2, 0.0) == pytest.approx(float(g["asym_put"]), rel=1e-13)
        assert engine.asymptotic_price(ps, 0.05, 100.0, 1.0, 0.02, True, 0.2, 0.01) == pytest.approx(float(g["asym_call"]), rel=1e-13)
        for K, T, sig, div, call in [(95.0, 0.5, 0.3, 0.02, False), (105.0, 1.0, 0.15, 0.0, True), (100.0, 2.0, 0.2, 0.0, False),
                                     (100.0, 0.013, 0.2, 0.0, False)]:
            want = port.asymptotic(P, 0.05, K, T, 0.02, call, sig, div)
            assert engine.asymptotic_price(ps, 0.05, K, T, 0.02, call, sig, div) == pytest.approx(want, rel=1e-13, abs=1e-15)
        ps.close()


def test_asymptotic_nan_inf_rules_and_errors(engine, port):
    rng = np.random.default_rng(11)
    P = 100.0 * np.exp(np.cumsum(0.02 * rng.standard_normal((777, 30)), axis=1))
    P[5, 3] = np.nan
    P[9, 7] = np.inf
    P[11, :] = np.nan
    ps = engine.upload_paths(P, dtype=m.MCP_F64)
    want = port.asymptotic(P, 0.03, 100.0, 0.5, 1.0 / 52, False, 0.2, 0.0)
    assert engine.asymptotic_price(ps, 0.03, 100.0, 0.5, 1.0 / 52, False, 0.2, 0.0) == pytest.approx(want, rel=1e-13)
    with pytest.raises(m.McpError, match="Volatility must be positive"):
        engine.asymptotic_price(ps, 0.03, 100.0, 0.5, 1.0 / 52, False, 0.0, 0.0)
    ps.close()
    assert m.AsymptoticAnalysis(engine).PredictOptionPrice([], 0.03, 100.0, 0.5, 0.02, False, 0.2, 0.0) == 0.0
    with pytest.raises(RuntimeError, match="AsymptoticAnalysis: Volatility must be positive."):
        m.AsymptoticAnalysis(engine).PredictOptionPrice(P, 0.03, 100.0, 0.5, 0.02, False, -1.0, 0.0)


@pytest.mark.parametrize("p", [1, 2, 3])
def test_martingale_vs_golden_and_oracle(engine, port, gold, p):
    g, P = gold
    ps = engine.upload_paths(P, dtype=m.MCP_F64)
    if p == 2:
        assert engine.martingale_price(ps, 0.05, 100.0, 1.0, 0.02, False, 2, 5) == pytest.approx(float(g["mart_put_p2"]), rel=1e-9)
        assert engine.martingale_price(ps, 0.05, 100.0, 1.0, 0.02, True, 2, 5) == pytest.approx(float(g["mart_call_p2"]), rel=1e-9)
    for K, T, call, iters in [(100.0, 1.0, False, 5), (95.0, 0.5, True, 2), (110.0, 1.0, False, 1), (100.0, 3.0, False, 3)]:
        want = port.martingale(P, 0.05, K, T, 0.02, call, p, iters)
        price, lo, up = engine.martingale_price(ps, 0.05, K, T, 0.02, call, p, iters, want_bounds=True)
        assert lo == pytest.approx(want["primal"], rel=1e-13)
        # the dual evaluates a fitted polynomial: the reference solves raw monomials by SVD (cond ~1e9 at p=3), the
        # device a standardised basis by Cholesky -> agreement to ~cond*eps
        assert up == pytest.approx(want["dual"], rel=2e-7 if p == 3 else 1e-8)
        assert price == pytest.approx(want["price"], rel=2e-7 if p == 3 else 1e-8)
    ps.close()


def test_martingale_errors_and_plugin(engine, port, gold):
    _, P = gold
    with pytest.raises(RuntimeError, match="MartingaleOptimization: Empty pricePaths."):
        m.MartingaleOptimization(engine).PredictOptionPrice([], 0.05, 100.0, 1.0, 0.02, False, 2)
    with pytest.raises(RuntimeError, match="MartingaleOptimization: maxIterations must be positive."):
        m.MartingaleOptimization(engine).PredictOptionPrice(P[:10], 0.05, 100.0, 1.0, 0.02, False, 2, 0)
    got = m.MartingaleOptimization(engine).PredictOptionPrice(P[:250], 0.04, 100.0, 0.9, 0.02, False, 2)
    assert got == pytest.approx(port.martingale(P[:250], 0.04, 100.0, 0.9, 0.02, False, 2, 5)["price"], rel=1e-8)
    # one path, two samples, p = 2: fewer samples than basis functions -> the martingale stays zero (:152-154)
    one = P[:1]
    assert m.MartingaleOptimization(engine).PredictOptionPrice(one, 0.04, 100.0, 1.0, 0.02, False, 2) == pytest.approx(
        port.martingale(one, 0.04, 100.0, 1.0, 0.02, False, 2, 5)["price"], rel=1e-12)


@pytest.mark.parametrize("ex,T", [(np.arange(0, 50), 1.0), (np.arange(0, 50, 3), 0.61), (np.array([5, 6, 40]), 1.0), (np.array([50]), 1.0)])
def test_branching_with_injected_indices_matches_oracle(engine, port, gold, ex, T):
    _, P = gold
    N, B = 1000, 10
    rng = np.random.default_rng(3)
    n_visit = int(np.sum(ex * 0.02 <= T))
    rp = rng.integers(0, N, size=(max(n_visit, 1), N, B)).astype(np.int32)
    want = port.branching(P[:N], 0.05, 100.0, T, 0.02, False, B, ex, rp)
    ps = engine.upload_paths(P[:N], dtype=m.MCP_F64)
    price, lo, up = engine.branching_price(ps, 0.05, 100.0, T, 0.02, False, B, ex, injected_rp=rp, want_bounds=True)
    ps.close()
    assert lo == pytest.approx(want["lower"], rel=1e-13)
    assert up == pytest.approx(want["upper"], rel=1e-12)
    assert price == pytest.approx(want["price"], rel=1e-12)


def test_branching_native_resampling_is_statistically_the_reference(engine, ref, gold):
    """With its own (Philox) resampling the upper bound is a random variable, like the reference's (mt19937 seeded
    from random_device): compare the two samples of prices."""
    _, P = gold
    N, B = 2000, 10
    ex = np.arange(0, 50)
    ps = engine.upload_paths(P[:N], dtype=m.MCP_F64)
    mine = np.array([engine.branching_price(ps, 0.05, 100.0, 1.0, 0.02, False, B, ex, seed=s) for s in range(12)])
    ps.close()
    theirs = np.array([ref.branching_price(P[:N], 0.05, 100.0, 1.0, 0.02, False, B, ex) for _ in range(12)])
    se = np.sqrt(mine.var(ddof=1) / mine.size + theirs.var(ddof=1) / theirs.size)
    assert abs(mine.mean() - theirs.mean()) < 4 * se + 1e-9, (mine.mean(), theirs.mean(), se)
    assert mine.std() > 0 and len(set(mine)) == mine.size  # different seeds, different resampling


def test_branching_error_contract(engine, gold):
    _, P = gold
    bp = m.BranchingProcesses(engine, seed=1)
    with pytest.raises(RuntimeError, match="BranchingProcesses: Empty pricePaths."):
        bp.PredictOptionPrice([], 0.05, 100.0, 1.0, 0.02, False, 10, [0, 1])
    with pytest.raises(RuntimeError, match="BranchingProcesses: No exercise times."):
        bp.PredictOptionPrice(P[:10], 0.05, 100.0, 1.0, 0.02, False, 10, [])
    with pytest.raises(RuntimeError, match="BranchingProcesses: Strike must be positive."):
        bp.PredictOptionPrice(P[:10], 0.05, 0.0, 1.0, 0.02, False, 10, [0, 1])
    assert np.isfinite(bp.PredictOptionPrice(P[:250], 0.05, 100.0, 1.0, 0.02, False, 10, list(range(50))))


def test_generate_stock_price_paths_exact_signature(engine, port):
    """RoughVolatility().GenerateStockPricePaths(hist, steps, n): host estimators + device generation.  The native
    normals are dumped by an explicit-parameter run with the same seed and replayed through the oracle."""
    g = np.load(os.path.join(G, "generate_paths_ref.npz"))
    hist = g["hist"]
    rv = m.RoughVolatility(engine, seed=99)
    paths = rv.GenerateStockPricePaths(hist, 63, 500)
    assert paths.shape == (500, 64) and np.all(paths[:, 0] == hist[-1]) and np.all(np.isfinite(paths))
    est = port.estimate_params(hist)
    ps = engine.pathset(500, 63)
    used = engine.gen_rbergomi(ps, est["S0"], est["r"], est["xi"], est["H"], est["eta"], est["rho"], est["dt"], seed=99, dump=True)
    want = port.rbergomi_paths(est["S0"], est["r"], est["xi"], est["H"], est["eta"], est["rho"], est["dt"], 63, used.astype(np.float64))
    ps.close()
    assert np.max(np.abs(paths - want) / want) < 1e-5  # stated fp32 path tolerance
    again = rv.GenerateStockPricePaths(hist, 63, 500)
    assert not np.array_equal(again, paths)  # successive calls draw fresh streams, like the reference's re-seeding
    with pytest.raises(RuntimeError, match="Historical prices vector too small."):
        rv.GenerateStockPricePaths([100.0], 10, 10)


def test_prediction_gen_row_through_the_python_plugins(engine, port):
    """One PredictionGen row (src/core/PredictionGen.cpp:700-719, :736-737, :780-791): 250 paths, four pricers."""
    rng = np.random.default_rng(8)
    hist = 100.0 * np.exp(np.cumsum(0.0126 * rng.standard_normal(300)))
    r, dt, dte = 0.04, 1.0 / 252.0, 120
    T = dte / 365.0
    steps = int(np.floor(T * 252))
    K = hist[-1] * 0.98
    paths = m.RoughVolatility(engine, seed=5).GenerateStockPricePaths(hist, steps, 250)
    ex = list(range(steps))
    aa = m.AsymptoticAnalysis(engine).PredictOptionPrice(paths, r, K, T, dt, False, 0.2, 0.0)
    lsm = m.LSM(engine).PredictOptionPrice(paths, r, K, T, dt, False, 2)
    mo = m.MartingaleOptimization(engine).PredictOptionPrice(paths, r, K, T, dt, False, 2)
    bp = m.BranchingProcesses(engine, seed=3).PredictOptionPrice(paths, r, K, T, dt, False, 10, ex)
    assert aa == pytest.approx(port.asymptotic(paths, r, K, T, dt, False, 0.2, 0.0), rel=1e-12)
    assert lsm == pytest.approx(port.lsm(paths, r, K, T, dt, False, 2)["price"], rel=1e-9)
    assert mo == pytest.approx(port.martingale(paths, r, K, T, dt, False, 2, 5)["price"], rel=1e-8)
    assert 0.0 < bp < 3 * max(aa, lsm)


def test_branching_single_launch_equals_per_date_kernels(engine, gold, monkeypatch):
    """<= 4096 paths: both bounds in one single-CTA launch; same resampling stream, same values as the per-date path."""
    _, P = gold
    ex = np.arange(0, 50)
    for n_paths, dtype in ((250, m.MCP_F64), (3000, m.MCP_F32)):
        ps = engine.upload_paths(P[:n_paths], dtype=dtype)
        l0 = engine.launch_count
        one = engine.branching_price(ps, 0.05, 100.0, 0.9, 0.02, False, 10, ex, seed=4, want_bounds=True)
        assert engine.launch_count - l0 <= 2
        monkeypatch.setenv("MCP_BRANCH_PER_DATE", "1")
        l0 = engine.launch_count
        many = engine.branching_price(ps, 0.05, 100.0, 0.9, 0.02, False, 10, ex, seed=4, want_bounds=True)
        assert engine.launch_count - l0 > 40
        monkeypatch.delenv("MCP_BRANCH_PER_DATE")
        ps.close()
        assert one[1] == pytest.approx(many[1], rel=1e-14) and one[2] == pytest.approx(many[2], rel=1e-13)
