"""Rows per second of the batched row driver (mcp_price_rows) on PredictionGen-shaped rows: 250 paths, dte-dependent
step counts, four pricers per row.  python tools/rows_throughput.py [n_rows=4096] [n_paths=250]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import montecarlooptionspricer_b200 as m  # noqa: E402
from test_gpu_rows import make_rows  # noqa: E402

n_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n_paths = int(sys.argv[2]) if len(sys.argv) > 2 else 250
eng = m.Engine(0)
rows = make_rows(np.random.default_rng(1), n_rows)
eng.price_rows(rows[:64], n_paths=n_paths, seed=0)  # warm-up
t0 = time.perf_counter()
out, gen_ms, price_ms = eng.price_rows(rows, n_paths=n_paths, seed=1)
wall = time.perf_counter() - t0
steps = sum(r["n_steps"] for r in rows)
print(f"{n_rows} rows x {n_paths} paths (mean {steps / n_rows:.0f} steps): wall {wall * 1e3:.1f} ms = {n_rows / wall:.0f} rows/s "
      f"(device: generation {gen_ms:.1f} ms, four pricers {price_ms:.1f} ms; the rest is host-side table building and marshalling); "
      f"{steps * n_paths / wall:.3e} path-steps/s")
eng.close()
