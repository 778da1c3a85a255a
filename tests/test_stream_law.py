"""CPU: the "pair" native stream of the 256-point generator has the reference's law.

The reference draws X = Re DFT(phi (.) Z) per path (RoughVolatility.cpp:264-292).  The B200 native stream feeds ONE
complex transform with sqrt(w_m) G_m (w = symmetrised |phi|^2) and takes Re and Im as the X of two paths.  Both are
linear maps of iid normals, so equality in law is equality of covariance matrices -- checked here exactly (to rounding),
together with the identity the kernel's dump mode relies on (draws in reference order that reproduce a pair-stream path)."""
import numpy as np
import pytest

import pair_stream as ps


@pytest.fixture(scope="module")
def port():
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    from oracle import oracle
    oracle.build(ref=False)
    return oracle.port()


@pytest.mark.parametrize("n,H,eta", [(252, 0.1, 1.9), (129, 0.3, 1.0), (200, 0.05, 2.5), (256, 0.5, 0.5), (64, 0.1, 1.9)])
def test_pair_stream_covariance_equals_reference(port, n, H, eta):
    phis, w, Mp = ps.spectrum(port, n, H, eta, 1.0 / 252.0)
    k = np.arange(n)
    F = np.exp(-2j * np.pi * np.outer(k, np.arange(Mp)) / Mp)  # [n][Mp]
    # reference: X = Re(F[:, :n] diag(phis) (zr + i zi)) = [Re B, -Im B] [zr; zi]
    B = F[:, :n] * phis
    A_ref = np.hstack([B.real, -B.imag])
    C_ref = A_ref @ A_ref.T
    # pair stream: Y = F diag(sqrt w) (gr + i gi);  X_A = Re Y,  X_B = Im Y
    D = F * np.sqrt(w)
    A_a = np.hstack([D.real, -D.imag])
    A_b = np.hstack([D.imag, D.real])
    scale = np.abs(C_ref).max()
    assert np.abs(A_a @ A_a.T - C_ref).max() < 1e-12 * scale
    assert np.abs(A_b @ A_b.T - C_ref).max() < 1e-12 * scale
    assert np.abs(A_a @ A_b.T).max() < 1e-12 * scale  # the two paths of a transform are uncorrelated => independent
    assert np.isclose(w.sum(), (np.abs(phis) ** 2).sum(), rtol=1e-14)


@pytest.mark.parametrize("n", [252, 129, 200, 256])
def test_equivalent_reference_draws_replay_the_pair_stream(port, n):
    H, eta = 0.1, 1.9
    phis, w, Mp = ps.spectrum(port, n, H, eta, 1.0 / 252.0)
    rng = np.random.default_rng(n)
    G = rng.standard_normal((7, Mp)) + 1j * rng.standard_normal((7, Mp))
    XA, XB = ps.pair_X(G, w, n)
    ZA, ZB = ps.equivalent_reference_draws(G, phis, w, n)
    assert np.allclose(ps.reference_X(ZA, phis, Mp), XA, rtol=0, atol=1e-12)
    assert np.allclose(ps.reference_X(ZB, phis, Mp), XB, rtol=0, atol=1e-12)


def test_pair_stream_w_counter_map_is_a_bijection():
    seen = set()
    for k in range(256):
        seen.add((4 * (k & 15) + (k >> 6), (k >> 4) & 3))
    assert len(seen) == 256 and max(c for c, _ in seen) == 63


@pytest.mark.parametrize("n", [1, 2, 7, 20, 63, 64, 65, 129, 252, 256, 300, 1000, 4096])
@pytest.mark.parametrize("H,eta,xi", [(0.1, 1.9, 0.04), (0.5, 0.5, 0.09), (0.0, 1.0, 0.02)])
def test_product_host_tables_match_the_oracle_phi(port, n, H, eta, xi):
    """The product's host-side table builder (cached unit roots, one exp per grid point, direct sum for short rows and a
    radix-2 transform for long ones) is pure host arithmetic: check it here against the oracle's phi (the reference's
    rbergomiPhi, RoughVolatility.cpp:212-236) and the closed forms of the other tables."""
    import montecarlooptionspricer_b200 as m
    dt = 1.0 / 252.0
    t = m.Engine.rbergomi_host_tables(n, dict(S0=100.0, r=0.05, xi=xi, H=H, eta=eta, rho=-0.9, dt=dt))
    Mp, log2e = t["Mp"], 1.4426950408889634
    assert Mp >= n and Mp < 2 * max(n, 1) + 1
    want = np.zeros(Mp, dtype=complex)
    want[:n] = port.rbergomi_phi(n, H, dt)[:n] * np.sqrt(2 * H) * eta / Mp * log2e
    scale = np.abs(want).max() if np.abs(want).max() > 0 else 1.0
    assert np.abs(t["phis"] - want).max() <= 2e-7 * scale  # fp32 rounding of an fp64 result
    k = np.arange(n)
    comp = -0.5 * eta * eta * (k * dt) ** (2 * H) * log2e + np.log2(xi)
    assert np.allclose(t["comp2"][:n], comp, rtol=3e-7, atol=1e-7) and np.all(t["comp2"][n:] == 0)
    c2 = np.abs(t["phis"]) ** 2
    w = 0.5 * (c2 + c2[(-np.arange(Mp)) % Mp])
    assert np.allclose(t["sw"], np.sqrt(w), rtol=2e-7, atol=1e-12 * scale)
