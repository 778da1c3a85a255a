"""compute-sanitizer is closed on the development pool, so the asynchronous sweep kernels (bulk-copy rings, cross-step
cursors of the persistent sweep, exchange rows) are run in a build with their own index checks and ring canaries
(-DMCP_DEBUG_BOUNDS=1, libmcp_b200_dbg.so) at ragged sizes: no check may fire, and the instrumented build must price
exactly what the release build prices."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(lib):
    env = dict(os.environ)
    if lib:
        env["MCP_B200_LIB"] = lib
    else:
        env.pop("MCP_B200_LIB", None)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "debug_bounds_worker.py")], capture_output=True, text=True, env=env, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    return json.loads(res.stdout.strip().splitlines()[-1])


def test_no_bounds_check_fires_and_the_debug_build_prices_identically():
    from montecarlooptionspricer_b200 import build
    dbg = build.LIB_DBG
    if not os.path.exists(dbg):
        dbg = build.build_debug()
    got = _run(dbg)
    assert got["violations"] is not None and len(got["violations"]) == 8, "not an MCP_DEBUG_BOUNDS build"
    names = ["carry store", "first-exercise store", "ring issue", "ring stage", "ring canary", "exchange row"]
    assert not any(got["violations"]), dict(zip(names, got["violations"]))
    ref = _run(None)
    assert ref["violations"] is None          # the release build carries no checks
    assert got["prices"] == ref["prices"]     # same arithmetic, bit for bit
