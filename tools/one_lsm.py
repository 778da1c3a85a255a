"""One generate + one LSM pass with explicit carry / sizes -- the short command that ncu wraps.
    python tools/one_lsm.py [log2_paths=26] [carry=f32|f64] [n_steps=252]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlooptionspricer_b200 as m  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 26
carry = m.MCP_F64 if (len(sys.argv) > 2 and sys.argv[2] == "f64") else m.MCP_F32
n = int(sys.argv[3]) if len(sys.argv) > 3 else 252
eng = m.Engine(0)
ps = eng.pathset(1 << k, n)
eng.gen_rbergomi(ps, 100.0, 0.05, 0.04, 0.1, 1.9, -0.9, 1.0 / 252.0, seed=3)
out = eng.lsm_price(ps, 0.05, 100.0, n / 252.0, 1.0 / 252.0, False, 3, carry=carry)
print(f"2^{k} x {n}: price {out.price:.7f} lsm {out.elapsed_ms:.3f} ms launches {out.n_kernel_launches}")
ps.close()
eng.close()
