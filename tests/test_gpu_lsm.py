"""GPU parity: Longstaff-Schwartz sweep vs the CPU oracle on IDENTICAL path values.

Protocol: both sides price the same fp32-rounded path values (the GPU from its slab, the
oracle from those values widened to double).  With the fp64 carry the exercise indices must be bit-exact
(up to reported exact near-ties, expected 0) and the price agrees to ~1e-9; with the fp32 carry the price
is within the stated 1e-5 relative tolerance."""
import numpy as np
import pytest

import montecarlooptionspricer_b200 as m
from conftest import CFG1, CFG2, f32_draws

pytestmark = pytest.mark.gpu


def gbm_paths(port, n_paths, n, seed, S0=100.0, r=0.05, sigma=0.2, T=1.0):
    rng = np.random.default_rng(seed)
    z = f32_draws(rng, (n_paths, n)).astype(np.float64)
    p = port.gbm_paths(S0, r, sigma, T / n, n, z)
    return p.astype(np.float32).astype(np.float64)  # the values both sides see


def check_parity(engine, port, paths, r, K, T, dt, is_call, p, price_tol=1e-9):
    ps = engine.upload_paths(paths, dtype=m.MCP_F32)
    got = engine.lsm_price(ps, r, K, T, dt, is_call, p, carry=m.MCP_F64, want_coeffs=True, want_first_exercise=True, want_v0=True)
    want = port.lsm(paths, r, K, T, dt, is_call, p)
    ps.close()
    assert abs(got.price - want["price"]) <= price_tol * max(1.0, abs(want["price"])), (got.price, want["price"])
    assert abs(got.std_error - want["stderr"]) <= 1e-7 * want["stderr"] + 1e-12 * max(1.0, abs(want["price"]))
    mism = np.nonzero(got.first_exercise != want["first_ex"])[0]
    # a mismatch is only tolerated at an exact near-tie |payoff - cont| < 1e-9 K (none expected)
    assert mism.size == 0 or want["min_gap"] < 1e-9 * K, f"{mism.size} exercise-index mismatches, min gap {want['min_gap']:.3e}"
    assert mism.size <= 2
    if mism.size == 0:
        assert np.max(np.abs(got.v0 - want["V0"])) <= 1e-8 * max(1.0, np.max(np.abs(want["V0"])))
    return got, want


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_lsm_config1_put_parity(engine, port, p):
    """BASELINE config 1: American put under GBM, S0=K=100, r=.05, sigma=.2, T=1, 100k paths x 50 steps."""
    paths = gbm_paths(port, 100_000, CFG1["n"], seed=1)
    got, want = check_parity(engine, port, paths, CFG1["r"], CFG1["K"], CFG1["T"], CFG1["T"] / CFG1["n"], False, p)
    if p == 3:  # value-iteration LSM is high-biased vs the Bermudan-50 value 6.0786
        assert 6.0 < got.price < 6.2


@pytest.mark.parametrize("p,n_paths,n", [(5, 100_000, 50), (6, 100_000, 50), (5, 20_000, 50), (6, 20_000, 50), (6, 250, 62), (5, 300_000, 20)])
def test_lsm_high_order_reproduces_the_reference_rank_cut(engine, port, monkeypatch, p, n_paths, n):
    """For polyOrder >= 5 the reference's RAW monomial design [1,S,..,S^p] with S~100 is numerically rank deficient by Eigen's
    own rule (sigma_min < min(rows, cols) eps sigma_max), so bdcSvd().solve() drops directions (LSMPricer.cpp:76).  The device
    regresses in a standardised basis, where nothing is deficient, and re-creates that cut from the Cholesky factor and the
    exact change of basis (lsm_solve.cuh): same exercise indices, same price.  Without the cut the full-rank fit is a different
    (better conditioned) estimator that agrees with the reference only statistically."""
    paths = gbm_paths(port, n_paths, n, seed=p + n_paths)
    got, want = check_parity(engine, port, paths, 0.05, 100.0, 1.0, 1.0 / n, False, p, price_tol=2e-9)
    if n_paths >= 20_000:
        S = paths[:, n // 5]
        sv = np.linalg.svd(np.vander(S[S < 100.0], p + 1, increasing=True), compute_uv=False)
        if sv[-1] < (p + 1) * np.finfo(float).eps * sv[0]:          # the reference truncates at this step: the cut must matter
            monkeypatch.setenv("MCP_LSM_REF_RANK", "0")
            ps = engine.upload_paths(paths, dtype=m.MCP_F32)
            full = engine.lsm_price(ps, 0.05, 100.0, 1.0, 1.0 / n, False, p, carry=m.MCP_F64)
            ps.close()
            assert abs(full.price - want["price"]) > 1e-7 * want["price"]          # not the reference's estimator ...
            assert abs(full.price - want["price"]) < 4 * want["stderr"]             # ... but a statistically equivalent one
    if n_paths > 4096:  # throughput mode takes the same cut for p >= 5
        monkeypatch.delenv("MCP_LSM_REF_RANK", raising=False)
        ps = engine.upload_paths(paths, dtype=m.MCP_F32)
        fast = engine.lsm_price(ps, 0.05, 100.0, 1.0, 1.0 / n, False, p, carry=m.MCP_F32)
        ps.close()
        assert abs(fast.price - want["price"]) < 1e-5 * want["price"]


def test_lsm_fitted_continuation_matches_oracle(engine, port):
    """Coefficient tables are in different bases (device: standardised; oracle: raw monomials, min-norm) but the
    fitted continuation FUNCTION must coincide on the in-the-money range wherever the design has full rank."""
    paths = gbm_paths(port, 50_000, 50, seed=3)
    got, want = check_parity(engine, port, paths, 0.05, 100.0, 1.0, 0.02, False, 3)
    S = np.linspace(80, 99.5, 40)
    for j in range(5, 49):
        a = np.polyval(got.coeffs[j][::-1], S)
        b = np.polyval(want["coeffs"][j][::-1], S)
        assert np.max(np.abs(a - b)) < 1e-6 * np.max(np.abs(b)), j
    lag = engine_lsm_coeffs(engine, paths, basis=m.MCP_BASIS_LAGUERRE)
    from numpy.polynomial import laguerre
    for j in (10, 30, 48):
        a = laguerre.lagval(S / 100.0, lag[j])
        b = np.polyval(want["coeffs"][j][::-1], S)
        assert np.max(np.abs(a - b)) < 1e-6 * np.max(np.abs(b)), j


def engine_lsm_coeffs(engine, paths, basis):
    ps = engine.upload_paths(paths, dtype=m.MCP_F32)
    out = engine.lsm_price(ps, 0.05, 100.0, 1.0, 0.02, False, 3, basis=basis, want_coeffs=True)
    ps.close()
    return out.coeffs


def test_lsm_call_and_strikes(engine, port):
    paths = gbm_paths(port, 20_000, 40, seed=4, sigma=0.35)
    for is_call, K in [(True, 100.0), (True, 90.0), (False, 110.0), (False, 70.0), (True, 140.0)]:
        check_parity(engine, port, paths, 0.03, K, 1.0, 1.0 / 40, is_call, 2)


def test_lsm_rbergomi_paths_cubic(engine, port):
    """Config-3 shape at oracle-friendly size: rough-vol paths, 252 steps, cubic basis."""
    n_paths, n = 20_000, 252
    ps = engine.pathset(n_paths, n)
    engine.gen_rbergomi(ps, CFG2["S0"], CFG2["r"], CFG2["xi"], CFG2["H"], CFG2["eta"], CFG2["rho"], CFG2["dt"], seed=3)
    slab = ps.download_timemajor()
    got = engine.lsm_price(ps, 0.05, 100.0, 1.0, CFG2["dt"], False, 3, carry=m.MCP_F64, want_first_exercise=True)
    want = port.lsm_timemajor_f32(slab, 0.05, 100.0, 1.0, CFG2["dt"], False, 3)
    assert abs(got.price - want["price"]) < 1e-9 * want["price"]
    assert np.array_equal(got.first_exercise, want["first_ex"]) or want["min_gap"] < 1e-7
    got32 = engine.lsm_price(ps, 0.05, 100.0, 1.0, CFG2["dt"], False, 3, carry=m.MCP_F32)
    assert abs(got32.price - want["price"]) < 1e-5 * want["price"]
    ps.close()


def test_lsm_fp32_carry_within_stated_tolerance(engine, port):
    paths = gbm_paths(port, 100_000, 50, seed=6)
    ps = engine.upload_paths(paths, dtype=m.MCP_F32)
    got = engine.lsm_price(ps, 0.05, 100.0, 1.0, 0.02, False, 3, carry=m.MCP_F32)
    want = port.lsm(paths, 0.05, 100.0, 1.0, 0.02, False, 3)
    assert abs(got.price - want["price"]) < 1e-5 * want["price"]  # north_star tolerance
    ps.close()


def test_lsm_maturity_cut(engine, port):
    """More columns than maturity/dt: steps with j*dt > maturity only discount (LSMPricer.cpp:43-49)."""
    paths = gbm_paths(port, 5_000, 60, seed=7)
    for T in (0.5, 0.5 + 1e-12, 0.3333, 5.0, 0.0):
        check_parity(engine, port, paths, 0.05, 100.0, T, 1.0 / 60, False, 2)


def test_lsm_rank_deficient_and_edge_cases(engine, port):
    rng = np.random.default_rng(8)
    # j = 0 in the money: every path at the same S0 < K -> rank-1 design, min-norm fit = mean(y)
    paths = gbm_paths(port, 4_000, 20, seed=9, S0=90.0)
    check_parity(engine, port, paths, 0.05, 100.0, 1.0, 0.05, False, 3)
    # fewer in-the-money paths than basis functions, and none at all
    base = gbm_paths(port, 64, 10, seed=10, S0=100.0, sigma=0.05)
    check_parity(engine, port, base, 0.05, 93.0, 1.0, 0.1, False, 3)
    check_parity(engine, port, base, 0.05, 50.0, 1.0, 0.1, False, 3)
    # tiny shapes: one path; one column; two columns
    check_parity(engine, port, base[:1], 0.05, 101.0, 1.0, 0.1, False, 2)
    check_parity(engine, port, base[:, :1], 0.05, 101.0, 1.0, 0.1, False, 2)
    check_parity(engine, port, base[:7, :2], 0.05, 101.0, 1.0, 0.1, False, 2)
    # duplicated prices (ties in S) and a ragged count that is not a multiple of the vector width
    dup = np.repeat(gbm_paths(port, 333, 12, seed=11), 3, axis=0)[:997]
    check_parity(engine, port, dup, 0.05, 100.0, 1.0, 1.0 / 12, False, 3)
    _ = rng


def test_lsm_fp64_slab_matches_reference_on_double_paths(engine, ref, port):
    """The drop-in call shape: host double paths in (kept fp64 on the device), mean out."""
    rng = np.random.default_rng(12)
    z = rng.standard_normal((30_000, 50))
    paths = port.gbm_paths(100.0, 0.05, 0.2, 0.02, 50, z)  # full double precision, NOT rounded
    want = ref.lsm_price(paths, 0.05, 100.0, 1.0, 0.02, False, 2)
    got = m.LSM(engine).PredictOptionPrice(paths, 0.05, 100.0, 1.0, 0.02, False, 2)
    assert abs(got - want) < 1e-9 * want


def test_lsm_reference_production_shape(engine, ref):
    """The reference's real workload: 250 paths x floor(dte/365*252) steps, polyOrder 2 (PredictionGen.cpp:718-719,790)."""
    rng = np.random.default_rng(13)
    hist = 100 * np.exp(np.cumsum(0.012 * rng.standard_normal(400)))
    for steps in (5, 21, 63, 126):
        draws = rng.standard_normal(250 * 4 * steps)
        paths, used = ref.generate_paths(hist, steps, 250, draws)
        assert used == draws.size
        K = hist[-1] * 1.02
        want = ref.lsm_price(paths, 0.04, K, steps / 252.0, 1 / 252.0, False, 2)
        got = m.LSM(engine).PredictOptionPrice(paths, 0.04, K, steps / 252.0, 1 / 252.0, False, 2)
        assert abs(got - want) < 1e-8 * max(1.0, want), (steps, got, want)


def test_lsm_empty_paths_error_contract(engine):
    with pytest.raises(RuntimeError, match="Empty pricePaths"):
        m.LSM(engine).PredictOptionPrice([], 0.05, 100.0, 1.0, 0.02, False, 2)
    with pytest.raises(RuntimeError, match="Empty pricePaths"):
        m.LSM(engine).PredictOptionPrice([[]], 0.05, 100.0, 1.0, 0.02, False, 2)


def test_price_rbergomi_lsm_native_within_3se_of_oracle(engine, port):
    """Native Philox end to end (config 3 shape) vs the oracle on an INDEPENDENT sample of the same size
    (the value-iteration LSM has an in-sample fitting bias that depends on N, so sizes must match)."""
    model = dict(S0=100.0, r=0.05, xi=0.04, H=0.1, eta=1.9, rho=-0.9, dt=1 / 252)
    lsm = dict(r=0.05, strike=100.0, maturity=1.0, dt=1 / 252, is_call=False, poly_order=3, carry=m.MCP_F32)
    out, gen_ms = engine.price_rbergomi_lsm(model, lsm, 1 << 15, 252, seed=2024)
    d = port.rbergomi_draws(777, 0, 1 << 15, 252, -0.9)
    paths = port.rbergomi_paths(100.0, 0.05, 0.04, 0.1, 1.9, -0.9, 1 / 252, 252, d)
    want = port.lsm(paths, 0.05, 100.0, 1.0, 1 / 252, False, 3)
    se = np.hypot(out.std_error, want["stderr"])
    assert abs(out.price - want["price"]) < 3 * se, (out.price, want["price"], se)
    assert out.n_paths_global == 1 << 15 and gen_ms > 0


def test_lsm_tma_ring_kernel_matches_direct_kernel_and_oracle(engine, port, monkeypatch):
    """The TMA-ring sweep (persistent 512-thread CTAs, bulk async copies into a shared-memory ring) takes over once
    there are >= 2 tiles of 4096 paths per SM.  Ragged path count (tail tile), put and call, with first-exercise
    output: same price as the direct-load kernel to fp32-accumulation noise, and both within the stated 1e-5 of the
    fp64 oracle on the same fp32 path values."""
    n_paths, n = 2 * 148 * 4096 + 4096 * 3 + 77, 12
    ps = engine.pathset(n_paths, n)
    engine.gen_gbm(ps, 100.0, 0.05, 0.2, 1.0 / n, seed=5)
    slab = ps.download_timemajor()
    for is_call, K in ((False, 100.0), (True, 97.0)):
        want = port.lsm_timemajor_f32(slab, 0.05, K, 1.0, 1.0 / n, is_call, 3)
        got = {}
        for impl in ("3", "2"):
            monkeypatch.setenv("MCP_SWEEP_IMPL", impl)
            got[impl] = engine.lsm_price(ps, 0.05, K, 1.0, 1.0 / n, is_call, 3, carry=m.MCP_F32, want_first_exercise=(impl == "3"), want_v0=True)
            assert abs(got[impl].price - want["price"]) < 1e-5 * want["price"], (impl, got[impl].price, want["price"])
            assert abs(got[impl].std_error - want["stderr"]) < 1e-4 * want["stderr"] + 1e-9 * want["price"]  # call: V0 is one constant
        assert abs(got["3"].price - got["2"].price) < 2e-7 * want["price"]
        assert np.max(np.abs(got["3"].v0 - got["2"].v0)) < 1e-4  # same per-path values up to decisions at fp32 near-ties
        agree = np.mean(got["3"].first_exercise == want["first_ex"])
        assert agree > 0.9999, agree  # fp32 decisions differ from the fp64 oracle at near-ties only
    ps.close()


@pytest.mark.parametrize("n_paths,n,p,is_call,K", [(250, 62, 2, False, 100.0), (250, 165, 2, True, 98.0), (1, 10, 2, False, 100.0),
                                                    (4096, 50, 3, False, 100.0), (777, 30, 4, False, 104.0), (3000, 20, 0, False, 101.0)])
def test_lsm_single_launch_kernel_for_small_path_sets(engine, port, monkeypatch, n_paths, n, p, is_call, K):
    """Up to 4096 paths (the reference's production rows are 250, PredictionGen.cpp:719) the whole backward induction
    is ONE launch.  Parity with the oracle as for the per-step kernels, and agreement with them on the same input."""
    paths = gbm_paths(port, n_paths, n, seed=n_paths + n)
    got, want = check_parity(engine, port, paths, 0.04, K, 1.0, 1.0 / n, is_call, p)
    assert got.n_kernel_launches < 12          # scale tables, tau fill, the induction, output copies
    monkeypatch.setenv("MCP_LSM_SMALL", "0")
    ps = engine.upload_paths(paths, dtype=m.MCP_F32)
    big = engine.lsm_price(ps, 0.04, K, 1.0, 1.0 / n, is_call, p, carry=m.MCP_F64, want_first_exercise=True)
    ps.close()
    assert big.n_kernel_launches > n
    assert abs(big.price - got.price) <= 1e-10 * max(1.0, abs(got.price))
    assert np.array_equal(big.first_exercise, got.first_exercise)


def test_config3_full_size_properties(engine):
    """BASELINE config 3 at its full single-GPU size (2^26 paths x 252 steps, cubic basis, 67.9 GB slab) through
    size-independent properties: determinism (same seed, same bits), agreement with independent 2^22-path prices of the
    same estimator within their measured seed-to-seed scatter, standard error shrinking like 1/sqrt(N), the American put
    above its European value on the same paths' law, and the global path count.

    The reported std_error is sqrt(sample variance of V_0 / N): the error of the final average given the fitted
    regressions.  The value-iteration estimator (LSMPricer.cpp:78-86 carries FITTED continuation values) also inherits
    the noise of its 252 regressions, so independent runs scatter 2-4x wider than that (tools/stream_dispersion.py);
    the comparison below therefore uses the scatter of eight independent small runs, not the reported figure."""
    free = engine.device_info()["free_bytes"]
    if free < 90e9:
        pytest.skip("needs ~70 GB of free HBM")
    model = dict(S0=100.0, r=0.05, xi=0.04, H=0.1, eta=1.9, rho=-0.9, dt=1 / 252)
    lsm = dict(r=0.05, strike=100.0, maturity=1.0, dt=1 / 252, is_call=False, poly_order=3, carry=m.MCP_F32)
    big, _ = engine.price_rbergomi_lsm(model, lsm, 1 << 26, 252, seed=3)
    again, _ = engine.price_rbergomi_lsm(model, lsm, 1 << 26, 252, seed=3)
    assert big.price == again.price and big.std_error == again.std_error and big.n_paths_global == 1 << 26
    smalls = [engine.price_rbergomi_lsm(model, lsm, 1 << 22, 252, seed=4 + k)[0] for k in range(8)]
    small = smalls[0]
    p = np.array([o.price for o in smalls])
    sd = p.std(ddof=1)
    assert sd > small.std_error  # the reported figure is a lower bound of the real scatter
    # mean of 8 runs (sd / sqrt 8) against one run with 16x the paths (sd / 4); 5 sigma on a 7-dof estimate of sd
    assert abs(big.price - p.mean()) < 5 * sd * np.sqrt(1 / 8 + 1 / 16), (big.price, p.mean(), sd)
    assert big.std_error == pytest.approx(small.std_error / 4.0, rel=0.02)  # 16x the paths
    euro = dict(lsm, maturity=0.0)  # every step past "maturity" is discount-only: the European value of the terminal payoff
    eu, _ = engine.price_rbergomi_lsm(model, euro, 1 << 22, 252, seed=4)
    assert small.price > eu.price > 0.0


def _poison_padding(ps):
    """Fill the pad columns [n_paths, ld) of every slab row with 0xFF bytes (a NaN bit pattern in fp32)."""
    from cuda.bindings import runtime as rt
    info = ps.info()
    pad = info["ld"] - info["n_paths"]
    if pad <= 0:
        return 0
    esz = 4 if info["dtype"] == m.MCP_F32 else 8
    (err,) = rt.cudaMemset2D(info["device_ptr"] + info["n_paths"] * esz, info["ld"] * esz, 0xFF, pad * esz, info["n_steps"] + 1)
    assert int(err) == 0, err
    (err,) = rt.cudaDeviceSynchronize()
    assert int(err) == 0, err
    return pad


@pytest.mark.parametrize("impl", ["2", "3", "4"])
def test_lsm_throughput_kernels_ignore_nan_in_the_padding(engine, port, monkeypatch, impl):
    """Pad lanes of a slab row (and stale ring bytes behind a short tail tile) may hold any bit pattern; a NaN there must
    not reach the regression moments (NaN * 0 = NaN).  The three throughput sweeps price a ragged path set whose padding
    was filled with NaN on purpose exactly as they price the clean one."""
    n_paths, n = 2 * 148 * 4096 + 4096 + 77, 8   # 77 live lanes in the last 128-lane line, short tail tile
    ps = engine.pathset(n_paths, n)
    engine.gen_gbm(ps, 100.0, 0.05, 0.2, 1.0 / n, seed=15)
    monkeypatch.setenv("MCP_SWEEP_IMPL", impl)
    clean = engine.lsm_price(ps, 0.05, 100.0, 1.0, 1.0 / n, False, 3, carry=m.MCP_F32)
    assert _poison_padding(ps) == 128 - 77
    dirty = engine.lsm_price(ps, 0.05, 100.0, 1.0, 1.0 / n, False, 3, carry=m.MCP_F32)
    assert np.isfinite(dirty.price) and dirty.price == clean.price and dirty.std_error == clean.std_error
    multi = engine.lsm_price_multi(ps, [95.0, 100.0, 105.0], 0.05, 1.0, 1.0 / n, False, 3)
    assert all(np.isfinite(o.price) and o.price > 0 for o in multi)
    ps.close()


@pytest.mark.parametrize("n_paths,n,p,is_call,K,T", [
    (5 * 147 * 4096 + 3 * 4096 + 77, 12, 3, False, 100.0, 1.0),   # every worker busy, uneven tile counts, ragged tail
    (5 * 147 * 4096 + 3 * 4096 + 77, 12, 3, True, 97.0, 1.0),
    (40 * 4096 + 1, 30, 2, False, 103.0, 1.0),                    # a few workers, 1-path tail tile
    (9000, 20, 3, False, 100.0, 1.0),                             # fewer tiles than ring stages: per-step refill mode
    (300_000, 24, 4, False, 100.0, 0.5),                          # maturity cut: discount-only steps before the first regression
    (300_000, 10, 0, False, 101.0, 1.0), (300_000, 10, 1, False, 101.0, 1.0), (200_000, 10, 5, False, 101.0, 1.0),
    (200_000, 10, 6, False, 101.0, 1.0),
])
def test_lsm_persistent_sweep_matches_per_step_kernels_and_oracle(engine, port, monkeypatch, n_paths, n, p, is_call, K, T):
    """The persistent cooperative sweep (whole induction in one launch: worker CTAs on a TMA ring that runs across step
    boundaries, a reducer CTA that folds / solves / broadcasts) against the per-step kernels on the same slab and, for
    p <= 4, against the fp64 oracle within the stated 1e-5."""
    ps = engine.pathset(n_paths, n)
    engine.gen_gbm(ps, 100.0, 0.05, 0.2, 1.0 / n, seed=n_paths % 1000 + p)
    monkeypatch.setenv("MCP_LSM_SMALL", "0")
    monkeypatch.setenv("MCP_SWEEP_IMPL", "4")
    got = engine.lsm_price(ps, 0.05, K, T, 1.0 / n, is_call, p, carry=m.MCP_F32, want_first_exercise=True, want_v0=True, want_coeffs=True)
    assert got.n_kernel_launches <= 4, got.n_kernel_launches   # tau fill, the induction, V0 copy
    again = engine.lsm_price(ps, 0.05, K, T, 1.0 / n, is_call, p, carry=m.MCP_F32)
    assert again.price == got.price and again.std_error == got.std_error   # fixed fold order: bitwise reproducible
    monkeypatch.setenv("MCP_SWEEP_IMPL", "2")
    ref2 = engine.lsm_price(ps, 0.05, K, T, 1.0 / n, is_call, p, carry=m.MCP_F32, want_first_exercise=True, want_v0=True, want_coeffs=True)
    assert ref2.n_kernel_launches > n
    assert abs(got.price - ref2.price) < 3e-7 * ref2.price, (got.price, ref2.price)
    assert abs(got.std_error - ref2.std_error) < 1e-4 * ref2.std_error + 1e-9 * ref2.price
    assert np.mean(got.first_exercise == ref2.first_exercise) > 0.9999
    assert np.max(np.abs(got.v0 - ref2.v0)) < 1e-4 * K
    if p <= 4:
        slab = ps.download_timemajor()
        want = port.lsm_timemajor_f32(slab, 0.05, K, T, 1.0 / n, is_call, p)
        assert abs(got.price - want["price"]) < 1e-5 * want["price"], (got.price, want["price"])
        assert np.mean(got.first_exercise == want["first_ex"]) > 0.9999
    ps.close()


def test_lsm_persistent_sweep_is_the_default_for_l2_resident_problems(engine, monkeypatch):
    monkeypatch.delenv("MCP_SWEEP_IMPL", raising=False)
    ps = engine.pathset(1 << 20, 16)
    engine.gen_gbm(ps, 100.0, 0.05, 0.2, 1.0 / 16, seed=2)
    out = engine.lsm_price(ps, 0.05, 100.0, 1.0, 1.0 / 16, False, 3, carry=m.MCP_F32)
    assert out.n_kernel_launches == 1 and out.n_paths_global == 1 << 20 and 5.5 < out.price < 6.5
    ps.close()


def test_lsm_parity_arithmetic_on_the_tma_ring_is_bit_exact(engine, port, monkeypatch):
    """The fp64-carry sweep fed by the TMA ring (>= 2 tiles of 4096 paths per SM) takes the same decisions as the
    grid-stride parity kernel and as the oracle: exercise indices bit-exact, price to 1e-9; ragged path count, put and call."""
    n_paths, n = 2 * 148 * 4096 + 4096 * 3 + 77, 12
    ps = engine.pathset(n_paths, n)
    engine.gen_gbm(ps, 100.0, 0.05, 0.2, 1.0 / n, seed=25)
    slab = ps.download_timemajor()
    _poison_padding(ps)
    for is_call, K, p in ((False, 100.0, 3), (True, 97.0, 2)):
        want = port.lsm_timemajor_f32(slab, 0.05, K, 1.0, 1.0 / n, is_call, p)
        got = {}
        for impl in ("3", "0"):
            monkeypatch.setenv("MCP_SWEEP64_IMPL", impl)
            got[impl] = engine.lsm_price(ps, 0.05, K, 1.0, 1.0 / n, is_call, p, carry=m.MCP_F64, want_first_exercise=True, want_v0=True)
            assert abs(got[impl].price - want["price"]) <= 1e-9 * want["price"], (impl, got[impl].price, want["price"])
            mism = np.count_nonzero(got[impl].first_exercise != want["first_ex"])
            assert mism == 0 or want["min_gap"] < 1e-9 * K, (impl, mism)
        assert np.array_equal(got["3"].first_exercise, got["0"].first_exercise)
        assert np.max(np.abs(got["3"].v0 - got["0"].v0)) <= 1e-9 * K
        assert abs(got["3"].price - got["0"].price) <= 1e-12 * want["price"]
    ps.close()
