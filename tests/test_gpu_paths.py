"""GPU parity: path generation vs the CPU oracle on identical (injected) normal draws.

Tolerance (north_star): path values within 1e-5 relative in fp32."""
import numpy as np
import pytest

from conftest import CFG2, f32_draws

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5


def _rb(engine, port, n_paths, n, seed, prm=None):
    prm = dict(CFG2) if prm is None else prm
    dt = prm.get("dt", 1.0 / 252.0)
    rng = np.random.default_rng(seed)
    d = f32_draws(rng, (n_paths, 4 * n))
    ps = engine.pathset(n_paths, n)
    engine.gen_rbergomi(ps, prm["S0"], prm["r"], prm["xi"], prm["H"], prm["eta"], prm["rho"], dt, injected=d)
    got = ps.download()
    want = port.rbergomi_paths(prm["S0"], prm["r"], prm["xi"], prm["H"], prm["eta"], prm["rho"], dt, n, d.astype(np.float64))
    ps.close()
    return got, want


@pytest.mark.parametrize("n_paths,n", [(64, 8), (1, 8), (31, 50), (33, 63), (1000, 252), (257, 256), (100, 300), (40, 1000),
                                       (20, 2047), (7, 1), (5, 2), (9, 3), (4096, 252), (33, 252), (1, 200), (50, 130), (999, 129), (64, 255)])
def test_rbergomi_injected_matches_oracle(engine, port, n_paths, n):
    got, want = _rb(engine, port, n_paths, n, seed=n_paths * 1000 + n)
    assert got.shape == want.shape
    assert np.all(np.isfinite(got))
    rel = np.max(np.abs(got - want) / np.abs(want))
    assert rel < REL_TOL, f"max rel err {rel:.3e}"
    assert np.all(got[:, 0] == np.float32(CFG2["S0"]))


@pytest.mark.parametrize("H,eta,rho,xi", [(0.1, 1.9, -0.9, 0.04), (0.5, 0.5, 0.0, 0.09), (0.3, 1.0, 0.7, 0.02), (0.05, 2.5, -1.0, 0.04)])
def test_rbergomi_parameter_sweep(engine, port, H, eta, rho, xi):
    prm = dict(S0=57.25, r=0.04, xi=xi, H=H, eta=eta, rho=rho, dt=1.0 / 252.0)
    got, want = _rb(engine, port, 300, 126, seed=int(H * 1000), prm=prm)
    rel = np.max(np.abs(got - want) / np.abs(want))
    assert rel < REL_TOL, f"max rel err {rel:.3e}"


def test_rbergomi_config2_chunked_64k(engine, port):
    """BASELINE config 2 shape (n=252, H=0.1, eta=1.9, rho=-0.9) on 65 536 injected paths."""
    got, want = _rb(engine, port, 65536, 252, seed=2)
    rel = np.abs(got - want) / np.abs(want)
    assert rel.max() < REL_TOL, f"max rel err {rel.max():.3e}"
    assert rel.mean() < 1e-6


@pytest.mark.parametrize("n_paths,n", [(1000, 50), (33, 7), (1, 1), (5000, 252), (100, 1001)])
def test_gbm_injected_matches_oracle(engine, port, n_paths, n):
    rng = np.random.default_rng(n_paths + n)
    d = f32_draws(rng, (n_paths, n))
    ps = engine.pathset(n_paths, n)
    engine.gen_gbm(ps, 100.0, 0.05, 0.2, 1.0 / n, injected=d)
    got = ps.download()
    want = port.gbm_paths(100.0, 0.05, 0.2, 1.0 / n, n, d.astype(np.float64))
    rel = np.max(np.abs(got - want) / want)
    assert rel < REL_TOL, f"max rel err {rel:.3e}"


def check_normals(used, ref_draws):
    """fp32 Box-Muller vs the fp64 stream spec.  SFU approximations give ~1e-6 (1+|z|); on top of that u1 is
    resolved to 2^-24 near 1, i.e. |dz| <= 3e-8 / radius for the (rare: P(radius<1e-3) = 5e-7) tiny radii,
    bounded by 2.5e-4 when u1 rounds to 1.  Statistically immaterial; bounded here explicitly."""
    err = np.abs(used - ref_draws)
    assert np.quantile(err, 0.9999) < 5e-6
    assert err.max() < 5e-4
    assert np.mean(err > 2e-5) < 1e-5


@pytest.mark.parametrize("n,impl", [(252, "2"), (252, "0"), (100, None), (300, None)])
def test_native_philox_dump_replays_through_oracle(engine, port, monkeypatch, n, impl):
    """Two-way parity on the per-path stream (generic kernel; x2 kernel with MCP_GEN_IMPL=2): the normals the GPU actually
    used, fed to the oracle, reproduce the GPU paths, and they are the documented Philox / Box-Muller stream."""
    if impl is not None:
        monkeypatch.setenv("MCP_GEN_IMPL", impl)
    n_paths = 2000
    ps = engine.pathset(n_paths, n)
    used = engine.gen_rbergomi(ps, CFG2["S0"], CFG2["r"], CFG2["xi"], CFG2["H"], CFG2["eta"], CFG2["rho"], CFG2["dt"],
                               seed=1234, path_offset=7, dump=True)
    got = ps.download()
    want = port.rbergomi_paths(CFG2["S0"], CFG2["r"], CFG2["xi"], CFG2["H"], CFG2["eta"], CFG2["rho"], CFG2["dt"], n,
                               used.astype(np.float64))
    assert np.max(np.abs(got - want) / want) < REL_TOL
    # and the dumped normals are the documented Philox/Box-Muller stream (fp32 SFU approximations: 5e-6 abs)
    ref_draws = port.rbergomi_draws(1234, 7, n_paths, n, CFG2["rho"])
    check_normals(used, ref_draws)
    # statistical sanity of the native stream: Z slots are N(0,1); the W slots carry rho*W and sqrt(1-rho^2)*W
    z = used[:, :2 * n].astype(np.float64)
    w = CFG2["rho"] * used[:, 2 * n:3 * n].astype(np.float64) + np.sqrt(1 - CFG2["rho"] ** 2) * used[:, 3 * n:].astype(np.float64)
    for x in (z, w):
        assert abs(x.mean()) < 5 / np.sqrt(x.size)
        assert abs(x.std() - 1) < 5 / np.sqrt(2 * x.size)
    assert abs(np.corrcoef(z[:, 0::2].ravel(), z[:, 1::2].ravel())[0, 1]) < 5 / np.sqrt(z.size / 2)
    assert abs(np.corrcoef(w[:, :-1].ravel(), w[:, 1:].ravel())[0, 1]) < 5 / np.sqrt(w.size)
    ps.close()


@pytest.mark.parametrize("n,n_paths,offset", [(252, 2000, 7), (252, 4096, 1 << 33), (129, 333, 64), (200, 70, 31), (256, 130, 0)])
def test_pair_stream_dump_replays_through_oracle(engine, port, n, n_paths, offset):
    """The default native stream for 128 < n <= 256 drives two paths with one transform (gen_rbergomi_pair.cuh).  Its dump
    is, per path, a set of draws in the reference's order that reproduces the path through the reference's formula."""
    ps = engine.pathset(n_paths, n)
    used = engine.gen_rbergomi(ps, CFG2["S0"], CFG2["r"], CFG2["xi"], CFG2["H"], CFG2["eta"], CFG2["rho"], CFG2["dt"],
                               seed=4321, path_offset=offset, dump=True)
    got = ps.download()
    assert np.all(np.isfinite(got)) and np.all(got[:, 0] == np.float32(CFG2["S0"]))
    want = port.rbergomi_paths(CFG2["S0"], CFG2["r"], CFG2["xi"], CFG2["H"], CFG2["eta"], CFG2["rho"], CFG2["dt"], n,
                               used.astype(np.float64))
    assert np.max(np.abs(got - want) / want) < REL_TOL
    # without the dump the same paths come out (the dump kernel is a separate instantiation)
    ps2 = engine.pathset(n_paths, n)
    engine.gen_rbergomi(ps2, CFG2["S0"], CFG2["r"], CFG2["xi"], CFG2["H"], CFG2["eta"], CFG2["rho"], CFG2["dt"], seed=4321, path_offset=offset)
    assert np.array_equal(ps2.download(), got)
    # W slots: iid N(0,1) split as rho W, sqrt(1-rho^2) W
    w = CFG2["rho"] * used[:, 2 * n:3 * n].astype(np.float64) + np.sqrt(1 - CFG2["rho"] ** 2) * used[:, 3 * n:].astype(np.float64)
    assert abs(w.mean()) < 5 / np.sqrt(w.size) and abs(w.std() - 1) < 5 / np.sqrt(2 * w.size)
    assert abs(np.corrcoef(w[:, :-1].ravel(), w[:, 1:].ravel())[0, 1]) < 5 / np.sqrt(w.size)
    ps.close()
    ps2.close()


def test_pair_stream_matches_its_written_spec(engine, port):
    """Philox counters / Box-Muller lanes / Re-Im pairing as documented in gen_rbergomi_pair.cuh, restated in numpy
    (tests/pair_stream.py): paths 60..67 and 95..97 of a shard starting at global id 1000 (tiles 15, 16, 17)."""
    import pair_stream
    n, seed, offset, n_paths = 252, 77, 1000, 200
    ps = engine.pathset(n_paths, n)
    engine.gen_rbergomi(ps, CFG2["S0"], CFG2["r"], CFG2["xi"], CFG2["H"], CFG2["eta"], CFG2["rho"], CFG2["dt"], seed=seed, path_offset=offset)
    got = ps.download()
    loc = [0, 1, 23, 24, 55, 56, 87, 88, 120, 199]
    d = pair_stream.spec_draws(port, seed, [offset + i for i in loc], n, CFG2["H"], CFG2["eta"], CFG2["rho"], CFG2["dt"])
    want = port.rbergomi_paths(CFG2["S0"], CFG2["r"], CFG2["xi"], CFG2["H"], CFG2["eta"], CFG2["rho"], CFG2["dt"], n, d)
    # fp32 Box-Muller on the SFU (~1e-6 per normal) through a 256-term sum: looser than the injected-draw tolerance
    assert np.max(np.abs(got[loc] - want) / want) < 1e-4
    ps.close()


def test_pair_stream_has_the_law_of_the_per_path_stream(engine, monkeypatch):
    """Same model through both native streams, 2^18 paths: moments of the terminal price and of the realised variance
    agree within Monte-Carlo error, and the two paths that share a transform are uncorrelated."""
    n, N = 252, 1 << 18
    stats = {}
    for impl in ("3", "2"):
        monkeypatch.setenv("MCP_GEN_IMPL", impl)
        ps = engine.pathset(N, n)
        engine.gen_rbergomi(ps, CFG2["S0"], CFG2["r"], CFG2["xi"], CFG2["H"], CFG2["eta"], CFG2["rho"], CFG2["dt"], seed=11)
        S = ps.download_timemajor().astype(np.float64)  # [n+1][N]
        ps.close()
        lr = np.diff(np.log(S), axis=0)
        rv = (lr ** 2).sum(axis=0)
        half = (lr[:n // 2] ** 2).sum(axis=0)
        stats[impl] = dict(ST=S[-1], rv=rv, half=half, lr=lr[[0, 10, 100, 251]])
    a, b = stats["3"], stats["2"]

    def close(x, y, k=5.0):
        se = np.sqrt(x.var() / x.size + y.var() / y.size)
        return abs(x.mean() - y.mean()) < k * se

    assert close(a["ST"], b["ST"])
    assert close(a["rv"], b["rv"]) and close(a["rv"] ** 2, b["rv"] ** 2)
    assert close(a["half"], b["half"]) and close(a["half"] * (a["rv"] - a["half"]), b["half"] * (b["rv"] - b["half"]))
    for i in range(4):
        assert close(a["lr"][i] ** 2, b["lr"][i] ** 2) and close(a["lr"][i] ** 3, b["lr"][i] ** 3)
    # Re / Im partners: global ids 64 T + c and 64 T + 32 + c
    rv = a["rv"].reshape(-1, 2, 32)
    x, y = np.log(rv[:, 0, :].ravel()), np.log(rv[:, 1, :].ravel())
    assert abs(np.corrcoef(x, y)[0, 1]) < 5 / np.sqrt(x.size)
    # neighbours in the packed pair (columns 2 pl, 2 pl + 1) as well
    x, y = np.log(a["rv"][0::2]), np.log(a["rv"][1::2])
    assert abs(np.corrcoef(x, y)[0, 1]) < 5 / np.sqrt(x.size)


def test_gbm_native_dump_matches_stream_spec(engine, port):
    n_paths, n = 3000, 50
    ps = engine.pathset(n_paths, n)
    used = engine.gen_gbm(ps, 100.0, 0.05, 0.2, 1.0 / n, seed=99, path_offset=123456789012, dump=True)
    ref_draws = port.gbm_draws(99, 123456789012, n_paths, n)
    check_normals(used, ref_draws)
    got = ps.download()
    want = port.gbm_paths(100.0, 0.05, 0.2, 1.0 / n, n, used.astype(np.float64))
    assert np.max(np.abs(got - want) / want) < REL_TOL


def test_path_offset_shards_are_slices_of_the_whole(engine):
    """Global path ids key the Philox counter: a rank's shard is bit-identical to its slice of the full set."""
    n_paths, n = 4096, 63
    args = (CFG2["S0"], CFG2["r"], CFG2["xi"], CFG2["H"], CFG2["eta"], CFG2["rho"], CFG2["dt"])
    full = engine.pathset(n_paths, n)
    engine.gen_rbergomi(full, *args, seed=5)
    whole = full.download_timemajor()
    for off, cnt in [(0, 1024), (1024, 1024), (2048, 2048), (100, 77)]:
        ps = engine.pathset(cnt, n)
        engine.gen_rbergomi(ps, *args, seed=5, path_offset=off)
        part = ps.download_timemajor()
        assert np.array_equal(part, whole[:, off:off + cnt])
        ps.close()
    full.close()
    # the pair stream (128 < n <= 256) keys transforms by the global 64-path tile: odd / unaligned shards are slices too
    n = 252
    full = engine.pathset(n_paths, n)
    engine.gen_rbergomi(full, *args, seed=6, path_offset=1 << 40)
    whole = full.download_timemajor()
    for off, cnt in [(0, 1024), (1024, 3072), (100, 77), (33, 95), (63, 2), (1, 4095), (4095, 1)]:
        ps = engine.pathset(cnt, n)
        engine.gen_rbergomi(ps, *args, seed=6, path_offset=(1 << 40) + off)
        part = ps.download_timemajor()
        assert np.array_equal(part, whole[:, off:off + cnt])
        ps.close()


def test_upload_download_roundtrip(engine):
    rng = np.random.default_rng(0)
    for N, M in [(1, 1), (33, 5), (1000, 51), (70000, 9)]:
        x = (100 * np.exp(0.1 * rng.standard_normal((N, M))))
        for dtype, exact in [(1, True), (0, False)]:
            ps = engine.upload_paths(x, dtype=dtype)
            y = ps.download()
            if exact:
                assert np.array_equal(x, y)
            else:
                assert np.array_equal(x.astype(np.float32).astype(np.float64), y)
            ps.close()


def test_martingale_property_large(engine):
    """Size-independent property at scale: E[e^{-rT} S_T] = S0 (4M paths, native Philox)."""
    n_paths, n = 1 << 22, 252
    ps = engine.pathset(n_paths, n)
    engine.gen_rbergomi(ps, CFG2["S0"], CFG2["r"], CFG2["xi"], CFG2["H"], CFG2["eta"], CFG2["rho"], CFG2["dt"], seed=11)
    tm = ps.download_timemajor()
    ST = tm[-1].astype(np.float64)
    disc = np.exp(-CFG2["r"] * n * CFG2["dt"])
    m, se = (disc * ST).mean(), (disc * ST).std() / np.sqrt(n_paths)
    assert abs(m - CFG2["S0"]) < 4 * se, (m, se)
    assert np.all(np.isfinite(tm)) and tm.min() > 0
    ps.close()


def test_rbergomi_config2_full_size_one_million_injected_paths(engine, port):
    """BASELINE config 2 at its full size: 2^20 paths x 252 steps, H=0.1, eta=1.9, rho=-0.9, every normal injected in
    the reference's order (4.2 GB of fp32 draws, streamed in 64k-path chunks), all 253 columns against the oracle."""
    n, chunk = 252, 1 << 16
    worst, mean_sum = 0.0, 0.0
    for c in range(16):
        rng = np.random.default_rng(1000 + c)
        d = f32_draws(rng, (chunk, 4 * n))
        ps = engine.pathset(chunk, n)
        engine.gen_rbergomi(ps, CFG2["S0"], CFG2["r"], CFG2["xi"], CFG2["H"], CFG2["eta"], CFG2["rho"], CFG2["dt"], injected=d,
                            path_offset=c * chunk)
        got = ps.download()
        ps.close()
        want = port.rbergomi_paths(CFG2["S0"], CFG2["r"], CFG2["xi"], CFG2["H"], CFG2["eta"], CFG2["rho"], CFG2["dt"], n, d.astype(np.float64))
        rel = np.abs(got - want) / np.abs(want)
        worst = max(worst, float(rel.max()))
        mean_sum += float(rel.mean())
    assert worst < REL_TOL, f"max rel err over 2^20 paths {worst:.3e}"
    assert mean_sum / 16 < 1e-6
