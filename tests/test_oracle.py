"""CPU tests of the checker itself: the plain-C port vs (a) the reference compiled here (oracle/_ref), (b) the
committed golden fixtures that build produced, (c) numpy, (d) analytic / lattice known answers, (e) Philox KATs."""
import os

import numpy as np
import pytest

from conftest import CFG1, CFG2

G = os.path.join(os.path.dirname(__file__), "golden")


# ---------------------------------------------------------------------------------------------- Philox
def test_philox4x32_10_known_answer_vectors(port):
    """Random123 kat_vectors, philox4x32 10 rounds."""
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kats:
        assert list(port.philox(ctr, key)) == want


def test_native_stream_is_standard_normal_and_layout(port):
    n, P, rho = 64, 4000, -0.9
    d = port.rbergomi_draws(5, 10, P, n, rho)
    z = d[:, :2 * n]
    w = rho * d[:, 2 * n:3 * n] + np.sqrt(1 - rho * rho) * d[:, 3 * n:]
    for x in (z, w):
        assert abs(x.mean()) < 5 / np.sqrt(x.size) and abs(x.std() - 1) < 5 / np.sqrt(2 * x.size)
    # a shard is a slice of the whole (global path ids key the counter)
    assert np.array_equal(port.rbergomi_draws(5, 10 + 100, 50, n, rho), d[100:150])
    assert np.array_equal(port.gbm_draws(5, 7 + 33, 20, 50), port.gbm_draws(5, 7, 100, 50)[33:53])
    # Box-Muller pair: documented formula
    z0, z1 = port.box_muller(0, 0)
    u1, u2 = 0.5 / 2 ** 32, 0.5 / 2 ** 32
    assert np.isclose(z0, np.sqrt(-2 * np.log(u1)) * np.cos(2 * np.pi * u2)) and np.isclose(z1, np.sqrt(-2 * np.log(u1)) * np.sin(2 * np.pi * u2))


# ------------------------------------------------------------------------------------ port vs reference
def test_spot_values_of_the_survey(ref):
    """n=252, dt=1/252, H=0.1, eta=1.9, xi=0.04, Z_k = sin(1+k) + i cos(k/2)."""
    n = 252
    phi = ref.rbergomi_phi(n, 0.1, 1 / 252)
    assert phi.size == 256
    assert np.isclose(phi[0], 105.192177702129, rtol=1e-13)
    assert np.isclose(phi[1], -5.19276869654139 - 8.08027417054438j, rtol=1e-13)
    k = np.arange(n)
    d = np.zeros((1, 4 * n))
    d[0, 0:2 * n:2], d[0, 1:2 * n:2] = np.sin(1 + k), np.cos(k / 2)
    _, X, v = ref.rbergomi_paths(100, 0.05, 0.04, 0.1, 1.9, -0.9, 1 / 252, n, d, want_xv=True)
    for idx, val in [(0, 0.299453573882088), (1, 0.300148132531547), (10, 0.311290952337686), (251, 0.295110893882153)]:
        assert np.isclose(X[0, idx], val, rtol=1e-13)
    assert np.isclose(v[0, 0], 0.0539648564381114, rtol=1e-13) and np.isclose(v[0, 251], 0.00885006821571217, rtol=1e-13)


@pytest.mark.parametrize("n", [1, 2, 3, 8, 50, 63, 252, 256, 300])
def test_port_paths_match_reference(ref, port, orc, n):
    rng = np.random.default_rng(n)
    d = rng.standard_normal((40, 4 * n)).astype(np.float32).astype(np.float64)
    args = (CFG2["S0"], CFG2["r"], CFG2["xi"], CFG2["H"], CFG2["eta"], CFG2["rho"], CFG2["dt"], n, d)
    a, Xa, va = ref.rbergomi_paths(*args, want_xv=True)
    b, Xb, vb = port.rbergomi_paths(*args, want_xv=True)
    c = orc.np_rbergomi_paths(*args)
    assert np.max(np.abs(a - b) / a) < 1e-13 and np.max(np.abs(Xa - Xb)) < 1e-13 and np.max(np.abs(va - vb) / va) < 1e-12
    assert np.max(np.abs(a - c) / a) < 1e-12
    assert np.max(np.abs(port.rbergomi_phi(n, CFG2["H"], CFG2["dt"]) - ref.rbergomi_phi(n, CFG2["H"], CFG2["dt"]))) < 1e-11


def test_unmodified_generate_paths_consumes_4n_draws_in_documented_order(ref):
    rng = np.random.default_rng(3)
    hist = 100 * np.exp(np.cumsum(0.01 * rng.standard_normal(300)))
    P, n = 3, 9
    d = rng.standard_normal(P * 4 * n)
    paths, used = ref.generate_paths(hist, n, P, d)
    assert used == P * 4 * n
    e = ref.estimate_params(hist)
    again = ref.rbergomi_paths(e["S0"], 0.04, e["xi"], e["H"], e["eta"], e["rho"], 1 / 252, n, d.reshape(P, 4 * n))
    assert np.array_equal(paths, again)  # explicit-parameter door == the reference's own API
    with pytest.raises(RuntimeError, match="too small"):
        ref.generate_paths(hist[:1], n, P, d)
    with pytest.raises(RuntimeError, match="exhausted"):
        ref.generate_paths(hist, n, P, d[:-1])


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_port_lsm_matches_reference_and_numpy(ref, port, orc, p):
    rng = np.random.default_rng(10 + p)
    paths = port.gbm_paths(100, 0.05, 0.2, 0.02, 50, rng.standard_normal((6000, 50)))
    for is_call, K, T in [(False, 100.0, 1.0), (True, 97.0, 1.0), (False, 104.0, 0.61)]:
        a = ref.lsm_price(paths, 0.05, K, T, 0.02, is_call, p)
        b = port.lsm(paths, 0.05, K, T, 0.02, is_call, p)["price"]
        c = orc.np_lsm(paths, 0.05, K, T, 0.02, is_call, p)
        assert a == b  # same solver, same order of operations: bitwise
        # LAPACK gelsd vs QR + one-sided Jacobi differ by rounding * cond(A); the raw monomial design of the
        # reference reaches cond ~1e14 at p = 4 (see test_gpu_lsm: high orders leave its parity domain)
        assert abs(a - c) < (1e-7 if p <= 3 else 1e-4) * a


def test_lsm_error_contract_and_rank_deficiency(ref, port):
    with pytest.raises(RuntimeError, match="Empty pricePaths"):
        ref.lsm_price(np.zeros((0, 0)), 0.05, 100, 1, 0.02, False, 2)
    # j = 0 in the money: rank-1 design -> min-norm solution = mean of the discounted carry
    rng = np.random.default_rng(0)
    paths = port.gbm_paths(90, 0.05, 0.2, 0.05, 20, rng.standard_normal((2000, 20)))
    o = port.lsm(paths, 0.05, 100.0, 1.0, 0.05, False, 3)
    assert o["price"] == ref.lsm_price(paths, 0.05, 100.0, 1.0, 0.05, False, 3)
    assert o["stderr"] < 1e-12 and np.ptp(o["V0"]) < 1e-9  # every V0 identical: max(K - S0, mean(...))
    c0 = o["coeffs"][0]
    S0p = np.array([1.0, 90.0, 90.0 ** 2, 90.0 ** 3])
    assert np.isclose(c0 @ S0p, max(10.0, o["V0"][0]), rtol=1e-9) or o["V0"][0] == 10.0


# ------------------------------------------------------------------------------------- golden fixtures
def test_port_vs_golden_rbergomi(port):
    g = np.load(os.path.join(G, "rbergomi_ref.npz"))
    S0, r, xi, H, eta, rho, dt = g["rb_params"]
    for tag in ("small", "cfg2", "pow2", "n50"):
        d = g[f"rb_{tag}_draws"].astype(np.float64)
        n = d.shape[1] // 4
        paths, X, v = port.rbergomi_paths(S0, r, xi, H, eta, rho, dt, n, d, want_xv=True)
        assert np.max(np.abs(paths - g[f"rb_{tag}_paths"]) / g[f"rb_{tag}_paths"]) < 1e-13
        assert np.max(np.abs(X - g[f"rb_{tag}_X"])) < 1e-13
        assert np.max(np.abs(v - g[f"rb_{tag}_v"]) / g[f"rb_{tag}_v"]) < 1e-12
    assert np.max(np.abs(port.rbergomi_phi(252, 0.1, 1 / 252) - g["phi_252"])) < 1e-11


def test_port_vs_golden_pricers(port):
    g = np.load(os.path.join(G, "pricers_ref.npz"))
    paths = g["paths_f32"].astype(np.float64)
    for p in (1, 2, 3):
        assert port.lsm(paths, 0.05, 100.0, 1.0, 0.02, False, p)["price"] == float(g[f"lsm_put_p{p}"])
        assert port.lsm(paths, 0.05, 95.0, 1.0, 0.02, True, p)["price"] == float(g[f"lsm_call_p{p}"])
    assert port.lsm(paths, 0.05, 100.0, 0.5, 0.02, False, 2)["price"] == float(g["lsm_put_p2_cut"])
    assert port.lsm(paths, 0.05, 110.0, 1.0, 0.02, False, 3)["price"] == float(g["lsm_put_itm0"])
    o = port.lsm(paths, 0.05, 100.0, 1.0, 0.02, False, 3)
    assert np.array_equal(o["first_ex"], g["port_first_ex_p3"])


def test_reference_rebuild_reproduces_golden(ref):
    """Guards the oracle/_ref build recipe (shims, compiler flags): the rebuilt reference gives the committed outputs."""
    g = np.load(os.path.join(G, "generate_paths_ref.npz"))
    P, cols = g["paths"].shape
    paths, used = ref.generate_paths(g["hist"], cols - 1, P, g["draws"])
    assert used == g["draws"].size and np.array_equal(paths, g["paths"])
    e = ref.estimate_params(g["hist"])
    assert np.allclose([e[k] for k in ("xi", "H", "eta", "rho", "S0")], g["est"], rtol=1e-14)
    q = np.load(os.path.join(G, "pricers_ref.npz"))
    assert ref.lsm_price(q["paths_f32"].astype(np.float64), 0.05, 100.0, 1.0, 0.02, False, 2) == float(q["lsm_put_p2"])


# -------------------------------------------------------------------------------------- known answers
def crr_put(S0, K, r, sigma, T, steps, exercise_every=1):
    dt = T / steps
    u = np.exp(sigma * np.sqrt(dt))
    d = 1 / u
    q = (np.exp(r * dt) - d) / (u - d)
    S = S0 * u ** np.arange(steps, -1, -1) * d ** np.arange(0, steps + 1)
    V = np.maximum(K - S, 0)
    for i in range(steps - 1, -1, -1):
        S = S[:-1] / u
        V = np.exp(-r * dt) * (q * V[:-1] + (1 - q) * V[1:])
        if i % exercise_every == 0:
            V = np.maximum(V, K - S)
    return V[0]


def test_config1_known_answers(port):
    """BASELINE config 1 (S0=K=100, r=.05, sigma=.2, T=1): Black-Scholes European 5.5735, Bermudan-50 lattice 6.0786.
    The reference's value-iteration LSM is high-biased: it must sit above the European value and within
    a few percent above the Bermudan value."""
    from math import erf, exp, log, sqrt
    N = lambda x: 0.5 * (1 + erf(x / sqrt(2)))
    d1 = (log(1.0) + (0.05 + 0.02) * 1.0) / 0.2
    bs = 100 * exp(-0.05) * N(-(d1 - 0.2)) - 100 * N(-d1)
    assert abs(bs - 5.5735) < 1e-3
    berm = crr_put(100, 100, 0.05, 0.2, 1.0, 5000, exercise_every=100)
    assert abs(berm - 6.0786) < 5e-3
    rng = np.random.default_rng(1)
    paths = port.gbm_paths(100, 0.05, 0.2, 0.02, 50, rng.standard_normal((100_000, 50)))
    o = port.lsm(paths, CFG1["r"], CFG1["K"], CFG1["T"], 0.02, False, 3)
    assert bs < o["price"] and berm - 3 * o["stderr"] < o["price"] < berm * 1.02


# ------------------------------------------------------------------------------------------------------
# the other three pricers and the generator's estimators -- port vs the compiled reference
# ------------------------------------------------------------------------------------------------------
def _golden_paths():
    g = np.load(os.path.join(G, "pricers_ref.npz"))
    return g, g["paths_f32"].astype(np.float64)


def test_port_asymptotic_and_martingale_vs_golden(port):
    g, P = _golden_paths()
    assert port.asymptotic(P, 0.05, 100.0, 1.0, 0.02, False, 0.2, 0.0) == pytest.approx(float(g["asym_put"]), rel=1e-13)
    assert port.asymptotic(P, 0.05, 100.0, 1.0, 0.02, True, 0.2, 0.01) == pytest.approx(float(g["asym_call"]), rel=1e-13)
    assert port.martingale(P, 0.05, 100.0, 1.0, 0.02, False, 2, 5)["price"] == pytest.approx(float(g["mart_put_p2"]), rel=1e-10)
    assert port.martingale(P, 0.05, 100.0, 1.0, 0.02, True, 2, 5)["price"] == pytest.approx(float(g["mart_call_p2"]), rel=1e-10)


@pytest.mark.parametrize("is_call,K,T,p,iters", [(False, 100.0, 1.0, 2, 5), (True, 95.0, 1.0, 3, 2), (False, 105.0, 0.5, 1, 1),
                                                (False, 100.0, 1.0, 2, 3)])
def test_port_pricers_match_reference(ref, port, is_call, K, T, p, iters):
    _, P = _golden_paths()
    P = P[:1500]
    assert port.asymptotic(P, 0.05, K, T, 0.02, is_call, 0.25, 0.01) == pytest.approx(
        ref.asymptotic_price(P, 0.05, K, T, 0.02, is_call, 0.25, 0.01), rel=1e-13, abs=1e-15)
    assert port.martingale(P, 0.05, K, T, 0.02, is_call, p, iters)["price"] == pytest.approx(
        ref.martingale_price(P, 0.05, K, T, 0.02, is_call, p, iters), rel=1e-9)


def test_martingale_fit_does_not_depend_on_the_iteration_count(ref):
    """Stopping indices and regression samples depend only on the paths (MartingaleOptimizationPricer.cpp:72-94,
    :130-150), so every maxIterations >= 2 returns the same value -- the fact the two-pass device kernel relies on."""
    _, P = _golden_paths()
    v = [ref.martingale_price(P[:800], 0.05, 100.0, 1.0, 0.02, False, 2, k) for k in (1, 2, 3, 5)]
    assert v[1] == v[2] == v[3] and v[0] != v[1]


def test_port_branching_matches_reference_with_injected_indices(ref, port):
    """The reference's Branching TU compiled serially with oracle/shim_uniform.h consumes the injected path indices in
    the order [path][date with continuation][branch]; the port takes them as [date][path][branch]."""
    _, P = _golden_paths()
    N, B = 300, 10
    rng = np.random.default_rng(3)
    for ex, T in ((np.arange(0, 50), 1.0), (np.arange(0, 50, 3), 0.61), (np.array([5, 6, 40]), 1.0)):
        n_visit = int(np.sum(ex * 0.02 <= T))
        rp = rng.integers(0, N, size=(n_visit, N, B)).astype(np.int32)
        got = port.branching(P[:N], 0.05, 100.0, T, 0.02, False, B, ex, rp)
        n_cont = int(np.sum(ex[:n_visit] < ex[-1]))
        want, used = ref.branching_price(P[:N], 0.05, 100.0, T, 0.02, False, B, ex,
                                         rp=np.ascontiguousarray(rp[:n_cont].transpose(1, 0, 2)), want_used=True)
        assert used == N * n_cont * B
        assert got["price"] == pytest.approx(want, rel=1e-12)
        assert got["lower"] <= got["upper"]


def test_port_estimators_match_reference_and_golden(ref, port):
    g = np.load(os.path.join(G, "generate_paths_ref.npz"))
    est = port.estimate_params(g["hist"])
    for k, v in zip(("xi", "H", "eta", "rho", "S0"), g["est"]):
        assert est[k] == pytest.approx(float(v), rel=1e-12), k
    rng = np.random.default_rng(5)
    for n in (2, 3, 17, 64, 500, 1826):
        hist = 50.0 * np.exp(np.cumsum(0.02 * rng.standard_normal(n)))
        a, b = port.estimate_params(hist), ref.estimate_params(hist)
        for k in ("xi", "H", "eta", "rho", "S0"):
            assert (np.isnan(a[k]) and np.isnan(b[k])) or a[k] == pytest.approx(b[k], rel=1e-11, abs=1e-300), (n, k)


def test_product_estimators_on_the_host_match_the_oracle(port):
    """mcp_estimate_rbergomi_params is pure host arithmetic (no device): check it here, on the CPU box."""
    import montecarlooptionspricer_b200 as m
    g = np.load(os.path.join(G, "generate_paths_ref.npz"))
    got = m.Engine.estimate_rbergomi_params(g["hist"])
    for k, v in zip(("xi", "H", "eta", "rho", "S0"), g["est"]):
        assert got[k] == pytest.approx(float(v), rel=1e-12), k
    assert got["r"] == 0.04 and got["dt"] == 1.0 / 252.0
    rng = np.random.default_rng(6)
    for n in (2, 5, 40, 700):
        hist = 80.0 * np.exp(np.cumsum(0.015 * rng.standard_normal(n)))
        a, b = m.Engine.estimate_rbergomi_params(hist), port.estimate_params(hist)
        for k in b:
            assert (np.isnan(a[k]) and np.isnan(b[k])) or a[k] == pytest.approx(b[k], rel=1e-11, abs=1e-300), (n, k)
    with pytest.raises(m.McpError, match="Historical prices vector too small."):
        m.Engine.estimate_rbergomi_params([100.0])


def test_numpy_philox_restatement_matches_the_port_and_the_known_answers(port):
    """tests/philox_np.py (used by the GPU known-answer test on a million counters) is itself pinned here."""
    from philox_np import philox4x32_10
    kats = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
            ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
            ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kats:
        assert [int(x) for x in philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0]] == want
    rng = np.random.default_rng(3)
    ctr = rng.integers(0, 2**32, size=(200, 4), dtype=np.uint64).astype(np.uint32)
    key = [0x9abcdef0, 0x12345678]
    got = philox4x32_10(ctr, key)
    for i in range(200):
        assert list(port.philox(ctr[i], key)) == [int(x) for x in got[i]]
