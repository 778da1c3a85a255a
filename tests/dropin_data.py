"""Synthetic inputs for the reference's own driver (src/core/PredictionGen.cpp): option_data.csv (>= 15 columns:
[0] ticker, [1] optionType (1 = call), [2] quote date M/D/YYYY, [3] underlying_last, [4] dte, [5] strike_distance_pct,
[14] dividend -- PredictionGen.cpp:582-654, :709) and nasdaq_stock_data.csv (header Date,<tickers>; one row per
calendar day, M/D/YYYY -- :177-238).  The reference ships no data (.gitignore:23-24)."""
import datetime as dt
import os

import numpy as np


def write_inputs(folder, n_rows=24, seed=7):
    rng = np.random.default_rng(seed)
    tickers = ["aaa", "bbb", "ccc"]
    end = dt.date(2024, 6, 28)
    days = [end - dt.timedelta(days=k) for k in range(1900, -1, -1)]
    px = {t: 100.0 * (1 + 0.3 * i) * np.exp(np.cumsum(0.0126 * rng.standard_normal(len(days)))) for i, t in enumerate(tickers)}
    with open(os.path.join(folder, "nasdaq_stock_data.csv"), "w") as f:
        f.write("Date," + ",".join(tickers) + "\n")
        for k, d in enumerate(days):
            f.write(f"{d.month}/{d.day}/{d.year}," + ",".join(f"{px[t][k]:.6f}" for t in tickers) + "\n")
    rows = []
    with open(os.path.join(folder, "option_data.csv"), "w") as f:
        f.write("ticker,type,quote_date,underlying_last,dte,strike_distance_pct,c6,c7,c8,c9,c10,c11,c12,c13,dividend\n")
        for i in range(n_rows):
            t = tickers[i % 3]
            q = end - dt.timedelta(days=int(rng.integers(0, 40)))
            k = days.index(q)
            dte = int(rng.choice([30, 61, 91, 150, 240]))
            dist = float(rng.choice([-0.05, -0.02, 0.0, 0.02, 0.05]))
            call = int(i % 2)
            rows.append(dict(ticker=t, call=call, S=px[t][k], dte=dte, dist=dist))
            f.write(f"{t},{call},{q.month}/{q.day}/{q.year},{px[t][k]:.6f},{dte},{dist},0,0,0,0,0,0,0,0,0.01\n")
    return rows


def read_output(folder):
    out = []
    with open(os.path.join(folder, "option_data_augmented.csv")) as f:
        for line in f:
            tok = line.strip().split(",")
            if len(tok) >= 21 and tok[0] != "ticker":
                out.append([float(x) for x in tok[15:21]])  # asym, branch, lsm, martingale, 20d vol, 20d momentum
    return np.array(out)
