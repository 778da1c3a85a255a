"""The C++ host plugin classes (montecarlooptionspricer_b200/host/mcp_plugins.hpp -- the reference's class names and
signatures over the C ABI) driven by a PredictionGen-shaped OpenMP row loop (host/plugin_rows_demo.cpp).  The demo
checks the exception contract and re-entrancy itself; here its paths are re-priced by the CPU oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_plugin_rows_match_oracle(port, tmp_path):
    from montecarlooptionspricer_b200 import build
    build.build_all()
    exe = build.DEMO
    assert os.path.exists(exe)
    out = tmp_path / "rows.bin"
    env = dict(os.environ, OMP_NUM_THREADS="4")
    res = subprocess.run([exe, str(out), "8", "250", "91"], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "0 failures" in res.stdout
    raw = out.read_bytes()
    n_rows, n_paths, M, n_hist, p, n_br = struct.unpack_from("6i", raw, 0)
    r, dt, T = struct.unpack_from("3d", raw, 24)
    off = 48
    assert (n_rows, n_paths, M) == (8, 250, int(np.floor(91 / 365.0 * 252)) + 1)
    for idx in range(n_rows):
        K, sigma, aa, bp, lsm, mo = struct.unpack_from("6d", raw, off)
        off += 48
        hist = np.frombuffer(raw, dtype=np.float64, count=n_hist, offset=off)
        off += 8 * n_hist
        paths = np.frombuffer(raw, dtype=np.float64, count=n_paths * M, offset=off).reshape(n_paths, M)
        off += 8 * n_paths * M
        call = idx % 2 == 1
        assert np.all(paths[:, 0] == hist[-1]) and np.all(np.isfinite(paths))
        assert aa == pytest.approx(port.asymptotic(paths, r, K, T, dt, call, sigma, 0.0), rel=1e-12, abs=1e-14)
        assert lsm == pytest.approx(port.lsm(paths, r, K, T, dt, call, p)["price"], rel=1e-8, abs=1e-12)
        assert mo == pytest.approx(port.martingale(paths, r, K, T, dt, call, p, 5)["price"], rel=1e-7, abs=1e-12)
        assert np.isfinite(bp) and bp >= 0.0
    assert off == len(raw)
