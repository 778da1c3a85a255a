"""BASELINE config 5 (strike x maturity surface under rough-vol LSM), at sizes the CPU oracle finishes in seconds, plus
the size-independent properties the domain offers on a larger grid (put prices increase with strike; the rank split
by maturity reproduces the unsplit surface entry for entry)."""
import numpy as np
import pytest

import montecarlooptionspricer_b200 as m
from conftest import CFG2

pytestmark = pytest.mark.gpu
MODEL = dict(S0=100.0, r=0.05, xi=0.04, H=0.1, eta=1.9, rho=-0.9, dt=1.0 / 252.0)


def test_surface_entries_match_the_oracle(engine, port):
    strikes, mats = [90.0, 100.0, 110.0], [21 / 252.0, 0.25]
    n_paths, seed = 6000, 17
    px, se, _, _ = engine.price_surface_rbergomi_lsm(MODEL, strikes, mats, n_paths, r=0.05, poly_order=3, carry=m.MCP_F64, seed=seed)
    assert px.shape == (2, 3) and np.all(np.isfinite(px)) and np.all(se >= 0)
    for mi, T in enumerate(mats):
        n = int(np.floor(T * 252))
        # the slab of maturity mi: same generator call the surface routine makes (seed derivation is part of the ABI contract)
        ps = engine.pathset(n_paths, n)
        engine.gen_rbergomi(ps, **{k: MODEL[k] for k in ("S0", "r", "xi", "H", "eta", "rho", "dt")},
                            seed=(seed + 0x9E3779B97F4A7C15 * (mi + 1)) % (1 << 64))
        slab = ps.download_timemajor()
        ps.close()
        for ki, K in enumerate(strikes):
            want = port.lsm_timemajor_f32(slab, 0.05, K, T, MODEL["dt"], False, 3)
            assert px[mi, ki] == pytest.approx(want["price"], rel=1e-9, abs=1e-9), (mi, ki)  # parity bar: 1e-9 of max(1, price)
            assert se[mi, ki] == pytest.approx(want["stderr"], rel=1e-6, abs=1e-9)  # K = 110: in the money at j = 0, V0 is one constant


def test_surface_monotone_in_strike_and_rank_split_is_exact(engine):
    strikes = np.arange(70.0, 131.0, 4.0)            # config 5: 16 strikes K = 70, 74, ..., 130
    mats = np.arange(1, 9) / 16.0                     # first 8 of the 16 maturities m / 16 y
    n_paths = 1 << 16
    full, se, gen_ms, lsm_ms = engine.price_surface_rbergomi_lsm(MODEL, strikes, mats, n_paths, r=0.05, seed=3)
    assert np.all(np.diff(full, axis=1) >= 0)         # a put is worth more at a higher strike (same paths per maturity)
    assert np.all(np.diff(full[-1, 5:]) > 0)          # ... strictly, wherever some path is in the money
    assert np.all(full[:, -1] >= 130.0 - 100.0 - 1e-6)  # deep ITM: at least intrinsic value
    parts = [engine.price_surface_rbergomi_lsm(MODEL, strikes, mats, n_paths, r=0.05, seed=3, mat_first=g, mat_stride=4)[0] for g in range(4)]
    merged = np.full_like(full, np.nan)
    for g, p in enumerate(parts):
        own = np.arange(len(mats)) % 4 == g
        assert np.all(np.isnan(p[~own])) and not np.any(np.isnan(p[own]))
        merged[own] = p[own]
    np.testing.assert_array_equal(merged, full)       # zero-collective split: bit-identical to the single-rank surface


def test_surface_config5_full_size(engine):
    """BASELINE config 5 at full size on one GPU: 16 strikes (70..130) x 16 maturities (m/16 y), 2^22 paths per contract,
    cubic basis, native Philox -- 256 American puts.  Size-independent properties: non-decreasing in strike (same paths
    per maturity), the at-the-money put gains value with maturity, deep in-the-money puts are worth at least intrinsic."""
    if engine.device_info()["free_bytes"] < 12e9:
        pytest.skip("needs ~5 GB of free HBM")
    strikes = np.arange(70.0, 131.0, 4.0)
    mats = np.arange(1, 17) / 16.0
    px, se, gen_ms, lsm_ms = engine.price_surface_rbergomi_lsm(MODEL, strikes, mats, 1 << 22, r=0.05, poly_order=3, seed=9)
    assert px.shape == (16, 16) and np.all(np.isfinite(px))
    assert np.all(np.diff(px, axis=1) >= 0)
    atm = px[:, 8]  # K = 102
    assert np.all(np.diff(atm) > -3 * np.hypot(se[1:, 8], se[:-1, 8]))
    assert np.all(px[:, -1] >= 30.0 - 1e-6)
    print(f"config 5 on one B200: generation {gen_ms:.0f} ms + LSM {lsm_ms:.0f} ms for 256 contracts x 2^22 paths")


def test_strike_ladder_in_one_sweep_matches_one_strike_at_a_time(engine, port, monkeypatch):
    """mcp_lsm_price_multi: up to 16 strikes share one sweep (one warp per contract, slab tiles read once).  Ragged path
    count, 19 strikes (two launches groups), against the single-contract throughput kernel and the fp64 oracle."""
    n_paths, n = 300_000 + 77, 16
    ps = engine.pathset(n_paths, n)
    engine.gen_gbm(ps, 100.0, 0.05, 0.25, 1.0 / n, seed=8)
    strikes = np.linspace(82.0, 118.0, 19)
    multi = engine.lsm_price_multi(ps, strikes, 0.05, 1.0, 1.0 / n, False, 3)
    assert multi[0].n_kernel_launches < 200  # 2 groups x 17 sweeps + the per-contract standard-error passes
    monkeypatch.setenv("MCP_LSM_MULTI", "0")
    single = engine.lsm_price_multi(ps, strikes, 0.05, 1.0, 1.0 / n, False, 3)
    slab = ps.download_timemajor()
    ps.close()
    for k, K in enumerate(strikes):
        assert multi[k].price == pytest.approx(single[k].price, rel=1e-5)          # same estimator, different standardisation
        assert multi[k].std_error == pytest.approx(single[k].std_error, rel=1e-4, abs=1e-7)  # K > S0: V0 is one constant
        assert multi[k].n_paths_global == n_paths
    for k in (0, 9, 18):
        want = port.lsm_timemajor_f32(slab, 0.05, float(strikes[k]), 1.0, 1.0 / n, False, 3)
        assert multi[k].price == pytest.approx(want["price"], rel=1e-5)            # the stated fp32 tolerance
    assert np.all(np.diff([x.price for x in multi]) > 0)


@pytest.mark.parametrize("n_paths,is_call,poly,maturity,strikes", [
    (5_000, False, 2, 1.0, [90.0, 97.5, 100.0, 104.0, 111.0]),
    (4_097 + 1024 * 151, True, 3, 0.6, [90.0, 97.5, 100.0, 104.0, 111.0]),
    # order 4: contracts that are NOT in the money at inception.  (In the money at j = 0 every path sits at S0, the regression has rank
    # 1, and from order 4 on the oracle's SVD answer there is off by 4e-4 .. 7e-4 (either sign) from the exact min-norm fit, which is
    # what the device returns in every mode.  Not pinnable: DESIGN 5.)
    (70_003, False, 4, 1.0, [84.0, 90.0, 95.0, 97.5, 100.0]),
])
def test_strike_ladder_edge_shapes(engine, port, n_paths, is_call, poly, maturity, strikes):
    """Fewer tiles than ring slots / than SMs, one more tile than SMs, calls, dates past maturity (discount-only steps), other
    polynomial orders -- every contract of the ladder against the fp64 oracle on the same fp32 paths (stated fp32 tolerance)."""
    n = 20
    ps = engine.pathset(n_paths, n)
    engine.gen_gbm(ps, 100.0, 0.05, 0.3, 1.0 / n, seed=n_paths)
    strikes = np.array(strikes)
    multi = engine.lsm_price_multi(ps, strikes, 0.05, maturity, 1.0 / n, is_call, poly)
    slab = ps.download_timemajor()
    ps.close()
    for k, K in enumerate(strikes):
        want = port.lsm_timemajor_f32(slab, 0.05, float(K), maturity, 1.0 / n, is_call, poly)
        assert multi[k].price == pytest.approx(want["price"], rel=2e-5), (k, K)
        assert multi[k].n_paths_global == n_paths
