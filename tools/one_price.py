"""One pricing pass of BASELINE config 3 (or a 2^k-path cut of it) -- the short command that ncu wraps.
    python tools/one_price.py [log2_paths=26] [repeats=1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlooptionspricer_b200 as m  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 26
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = m.Engine(0)
model = dict(S0=100.0, r=0.05, xi=0.04, H=0.1, eta=1.9, rho=-0.9, dt=1.0 / 252.0)
lsm = dict(r=0.05, strike=100.0, maturity=1.0, dt=1.0 / 252.0, is_call=False, poly_order=3, carry=m.MCP_F32)
for i in range(reps):
    out, gen_ms = eng.price_rbergomi_lsm(model, lsm, 1 << k, 252, seed=1 + i)
    print(f"paths 2^{k}: price {out.price:.6f} +- {out.std_error:.6f}, gen {gen_ms:.2f} ms, lsm {out.elapsed_ms:.2f} ms, "
          f"{out.n_kernel_launches} launches")
eng.close()
