"""Child process of tests/test_gpu_debug_bounds.py: loads whichever libmcp_b200 build MCP_B200_LIB names, runs ragged path
counts through every asynchronous sweep kernel and prints one JSON line {prices: [...], violations: [...] | null}."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlooptionspricer_b200 as m  # noqa: E402

eng = m.Engine(0)
L = eng._L
prices = []
cases = [(2 * 148 * 4096 + 4096 + 77, 9), (5 * 147 * 4096 + 3 * 4096 + 1, 7), (40 * 4096 + 1, 11), (9000, 6), (4097, 5), (300_001, 8)]
for n_paths, n in cases:
    ps = eng.pathset(n_paths, n)
    eng.gen_gbm(ps, 100.0, 0.05, 0.2, 1.0 / n, seed=n_paths % 97)
    os.environ["MCP_LSM_SMALL"] = "0"
    for impl in ("2", "3", "4"):
        os.environ["MCP_SWEEP_IMPL"] = impl
        out = eng.lsm_price(ps, 0.05, 100.0, 1.0, 1.0 / n, False, 3, carry=m.MCP_F32, want_first_exercise=(impl != "2"))
        prices.append(out.price)
    os.environ.pop("MCP_SWEEP_IMPL")
    out = eng.lsm_price(ps, 0.05, 100.0, 1.0, 1.0 / n, False, 2, carry=m.MCP_F64, want_first_exercise=True)   # parity arithmetic (TMA ring when large)
    prices.append(out.price)
    multi = eng.lsm_price_multi(ps, [90.0, 95.0, 100.0, 105.0, 110.0], 0.05, 1.0, 1.0 / n, False, 3)
    prices += [o.price for o in multi]
    ps.close()
viol = None
if hasattr(L, "mcp_debug_violations"):
    L.mcp_debug_violations.restype = C.c_int
    buf = (C.c_uint * 8)()
    k = L.mcp_debug_violations(buf, 8)
    viol = [int(x) for x in buf[:k]] if k > 0 else None
eng.close()
print(json.dumps({"prices": prices, "violations": viol}))
