"""Time the parity-mode (fp64 carry) sweep: grid-stride kernel (MCP_SWEEP64_IMPL=0) vs TMA ring (=3).
    python tools/sweep64_timing.py [log2_paths=26]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlooptionspricer_b200 as m  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 26
eng = m.Engine(0)
mdl = dict(S0=100.0, r=0.05, xi=0.04, H=0.1, eta=1.9, rho=-0.9, dt=1.0 / 252.0)
ps = eng.pathset(1 << k, 252)
eng.gen_rbergomi(ps, mdl["S0"], mdl["r"], mdl["xi"], mdl["H"], mdl["eta"], mdl["rho"], mdl["dt"], seed=3)
for impl in ("0", "3"):
    os.environ["MCP_SWEEP64_IMPL"] = impl
    best = 1e30
    for _ in range(3):
        out = eng.lsm_price(ps, 0.05, 100.0, 1.0, mdl["dt"], False, 3, carry=m.MCP_F64)
        best = min(best, out.elapsed_ms)
    us = best * 1e3 / 253
    print(f"2^{k} fp64 carry impl {impl}: lsm {best:8.3f} ms  {us:7.2f} us/step  {20.0 * (1 << k) / (us * 1e-6) / 1e9 / 6544.3:5.3f} of HBM peak (20 B)  price {out.price:.9f}", flush=True)
ps.close()
eng.close()
