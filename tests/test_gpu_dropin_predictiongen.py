"""Drop-in proof on the reference's own driver: src/core/PredictionGen.cpp, UNMODIFIED, compiled once against the
reference's CPU model classes (oracle/_ref/PredictionGen_ref) and once against the B200 plugin header
(oracle/_ref/PredictionGen_b200: mcp_plugins.hpp + libmcp_b200_plugins.so + libmcp_b200.so) by oracle/Makefile `dropin`.
Both read the same synthetic CSVs.  The reference seeds its RNGs from std::random_device, so per-row prices are random
on both sides (250 paths per row, PredictionGen.cpp:719): host-side columns must match exactly, pricer columns in
aggregate."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import dropin_data as D

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "PredictionGen_ref")
B200_EXE = os.path.join(ROOT, "oracle", "_ref", "PredictionGen_b200")


@pytest.mark.skipif(not (os.path.exists(REF_EXE) and os.path.exists(B200_EXE)), reason="drop-in binaries not built (needs /root/reference at build time)")
def test_unmodified_predictiongen_runs_on_the_b200_plugins():
    n_rows = 36
    dirs = [tempfile.mkdtemp(prefix="pg_ref_"), tempfile.mkdtemp(prefix="pg_b200_")]
    for d in dirs:
        rows = D.write_inputs(d, n_rows)
    env = dict(os.environ, OMP_NUM_THREADS="4")
    procs = [subprocess.Popen([exe], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
             for exe, d in zip((REF_EXE, B200_EXE), dirs)]
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, so[-2000:] + se[-2000:]
    ref, b200 = D.read_output(dirs[0]), D.read_output(dirs[1])
    assert ref.shape == b200.shape == (n_rows, 6)
    assert np.all(np.isfinite(b200)) and np.all(b200[:, :4] >= 0.0)
    errlog = open(os.path.join(dirs[1], "error_log.txt")).read() if os.path.exists(os.path.join(dirs[1], "error_log.txt")) else ""
    assert "Row" not in errlog, errlog[:2000]          # no row fell into the driver's catch blocks
    assert np.count_nonzero(b200[:, 2]) == n_rows      # every LSM price was produced
    np.testing.assert_allclose(b200[:, 4:], ref[:, 4:], rtol=1e-12)  # 20-day vol / momentum: pure host code of the driver
    for col, name in enumerate(("asymptotic", "branching", "lsm", "martingale")):
        ratio = b200[:, col].sum() / ref[:, col].sum()
        assert 0.85 < ratio < 1.18, (name, ratio)
    # row by row the two Monte-Carlo estimates (250 paths each) agree within their sampling noise
    rel = np.abs(b200[:, 2] - ref[:, 2]) / np.maximum(ref[:, 2], 1e-9)
    assert np.median(rel) < 0.25
