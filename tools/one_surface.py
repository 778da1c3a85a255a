"""Strike ladder(s) of BASELINE config 5 through mcp_price_surface_rbergomi_lsm -- the short command that ncu wraps.
    python tools/one_surface.py [n_maturities=1 (the longest ones)] [log2_paths=22]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlooptionspricer_b200 as m  # noqa: E402

nm = int(sys.argv[1]) if len(sys.argv) > 1 else 1
k = int(sys.argv[2]) if len(sys.argv) > 2 else 22
eng = m.Engine(0)
model = dict(S0=100.0, r=0.05, xi=0.04, H=0.1, eta=1.9, rho=-0.9, dt=1.0 / 252.0)
strikes, mats = np.arange(70.0, 131.0, 4.0), (np.arange(1, 17) / 16.0)[-nm:]
eng.price_surface_rbergomi_lsm(model, strikes, mats[:1], 1 << 18, r=0.05, seed=1)
t0 = time.perf_counter()
px, se, g, l = eng.price_surface_rbergomi_lsm(model, strikes, mats, 1 << k, r=0.05, poly_order=3, seed=9)
wall = time.perf_counter() - t0
steps = sum(int(np.floor(T * 252)) for T in mats)
print(f"{nm} maturities x 16 strikes x 2^{k} paths: wall {wall * 1e3:.1f} ms, gen {g:.1f} ms, lsm {l:.1f} ms = {l * 1e3 / steps:.1f} us per step of 16 contracts; "
      f"ATM longest {px[-1, 8]:.5f}")
eng.close()
