"""Batched row driver (mcp_price_rows): the reference's per-row block (src/core/PredictionGen.cpp:700-791) for a whole
batch of rows in three launches.  Every row must equal what the per-row entry points give on the same paths, and the
oracle on those paths."""
import numpy as np
import pytest

import montecarlooptionspricer_b200 as m

pytestmark = pytest.mark.gpu


def make_rows(rng, n_rows):
    rows = []
    for k in range(n_rows):
        dte = int(rng.choice([0, 30, 61, 91, 150, 240]))
        T = dte / 365.0
        S0 = float(rng.uniform(20, 300))
        model = dict(S0=S0, r=0.04, xi=float(rng.uniform(0.01, 0.09)), H=float(rng.uniform(0.05, 0.6)), eta=float(rng.uniform(0.02, 1.9)),
                     rho=float(rng.uniform(-0.9, 0.0)), dt=1.0 / 252.0)
        rows.append(dict(model=model, n_steps=int(np.floor(T * 252.0)), is_call=bool(k % 2), r=0.04, strike=S0 * (1.0 - float(rng.choice([-0.05, 0.0, 0.03]))),
                         maturity=T, dt=1.0 / 252.0, sigma=0.2, dividend=0.01))
    return rows


def test_batched_rows_equal_per_row_calls_and_oracle(engine, port, monkeypatch):
    # the batch generates with the generic kernel (rows differ in length); pin the per-row calls to the same kernel so that
    # both sides see bit-identical paths (the specialised 256-point kernels agree with it to fp32 rounding only)
    monkeypatch.setenv("MCP_GEN_IMPL", "0")
    rng = np.random.default_rng(12)
    rows = make_rows(rng, 14)
    n_paths, seed = 250, 77
    out, gen_ms, price_ms = engine.price_rows(rows, n_paths=n_paths, poly_order=2, num_branches=10, max_iterations=5, seed=seed)
    assert out.shape == (14, 5) and np.all(np.isfinite(out))
    for k, row in enumerate(rows):
        n = row["n_steps"]
        if n < 1:
            assert np.all(out[k] == 0.0)  # PredictionGen.cpp:720-733 writes zeros for such rows
            continue
        md = row["model"]
        ps = engine.pathset(n_paths, n)
        engine.gen_rbergomi(ps, md["S0"], md["r"], md["xi"], md["H"], md["eta"], md["rho"], md["dt"], seed=seed, path_offset=k * n_paths)
        args = (row["r"], row["strike"], row["maturity"], row["dt"], row["is_call"])
        aa = engine.asymptotic_price(ps, *args, row["sigma"], row["dividend"])
        bp = engine.branching_price(ps, *args, 10, np.arange(n), seed=seed ^ 0x5bd1e995, path_offset=k * n_paths)
        lsm = engine.lsm_price(ps, *args, 2, carry=m.MCP_F64)
        mo = engine.martingale_price(ps, *args, 2, 5)
        slab = ps.download_timemajor()
        ps.close()
        assert out[k, 0] == pytest.approx(aa, rel=1e-13, abs=1e-15), (k, "asymptotic")
        assert out[k, 1] == pytest.approx(bp, rel=1e-13, abs=1e-15), (k, "branching")
        assert out[k, 2] == pytest.approx(lsm.price, rel=1e-13, abs=1e-15), (k, "lsm")
        assert out[k, 4] == pytest.approx(lsm.std_error, rel=1e-10, abs=1e-15), (k, "lsm stderr")
        assert out[k, 3] == pytest.approx(mo, rel=1e-12, abs=1e-15), (k, "martingale")
        paths = slab.T.astype(np.float64)
        assert out[k, 0] == pytest.approx(port.asymptotic(paths, *args, row["sigma"], row["dividend"]), rel=1e-12, abs=1e-14)
        assert out[k, 2] == pytest.approx(port.lsm(paths, *args, 2)["price"], rel=1e-8, abs=1e-10)
        assert out[k, 3] == pytest.approx(port.martingale(paths, *args, 2, 5)["price"], rel=1e-7, abs=1e-10)


def test_batched_rows_many_rows_and_error_contract(engine):
    rng = np.random.default_rng(5)
    rows = make_rows(rng, 600)
    a, _, _ = engine.price_rows(rows, n_paths=250, seed=1)
    b, gen_ms, price_ms = engine.price_rows(rows, n_paths=250, seed=1)
    np.testing.assert_array_equal(a, b)  # deterministic
    live = np.array([r["n_steps"] >= 1 for r in rows])
    assert np.all(a[live, :4] >= 0.0) and np.all(a[~live] == 0.0)
    assert gen_ms > 0 and price_ms > 0
    bad = dict(rows[1]); bad["sigma"] = 0.0; bad["n_steps"] = 30
    with pytest.raises(m.McpError, match="Volatility must be positive"):
        engine.price_rows([bad])
    with pytest.raises(m.McpError, match="n_paths"):
        engine.price_rows(rows[:2], n_paths=5000)


def test_a_degenerate_row_spoils_only_itself(engine):
    """The reference's unclamped DFA slope can go negative on real histories (RoughVolatility.cpp:120-149): it then writes NaN
    paths for THAT row and PredictionGen rejects the row (:753-777).  The batched driver must not abort the batch."""
    rng = np.random.default_rng(5)
    rows = make_rows(rng, 12)
    good, _, _ = engine.price_rows(rows, n_paths=250, seed=3)
    bad = [dict(r, model=dict(r["model"])) for r in rows]
    live = [k for k, r in enumerate(rows) if r["n_steps"] >= 1]
    k_h, k_rho = live[0], live[1]
    bad[k_h]["model"]["H"] = -0.07
    bad[k_rho]["model"]["rho"] = float("nan")
    out, _, _ = engine.price_rows(bad, n_paths=250, seed=3)
    for k in range(len(rows)):
        if k in (k_h, k_rho):
            assert np.all(np.isnan(out[k]))
        else:
            assert np.array_equal(out[k], good[k])  # paths are keyed by (seed, row, path): untouched by the neighbours
