"""Steady-state rows per second of the batched row driver on the bench's row mix (allocations done by a full-size warm-up):
    python tools/rows_steady.py [n_rows=16384] [repeats=7]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import montecarlooptionspricer_b200 as m  # noqa: E402
from bench import make_rows  # noqa: E402

n_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
eng = m.Engine(0)
arr = m.engine.rows_to_array(make_rows(np.random.default_rng(1), n_rows))
eng.price_rows(arr, n_paths=250, seed=0)
ts, g, p = [], 0, 0
for r in range(reps):
    t0 = time.perf_counter()
    res, g, p = eng.price_rows(arr, n_paths=250, seed=1 + r)
    ts.append(time.perf_counter() - t0)
ts.sort()
print(f"{n_rows} rows: best {1e3 * ts[0]:.1f} ms = {n_rows / ts[0]:.0f} rows/s, median {1e3 * ts[len(ts) // 2]:.1f} ms = {n_rows / ts[len(ts) // 2]:.0f} rows/s "
      f"(device: generation {g:.1f} ms, four pricers {p:.1f} ms; MCP_ROWS_CHUNK={os.environ.get('MCP_ROWS_CHUNK', 'default')})")
eng.close()
