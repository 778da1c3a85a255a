"""Seed-to-seed dispersion of the config-3 price under the two native streams of the 256-point generator (per-path
transform, MCP_GEN_IMPL=2; one transform per pair of paths, default).  The value-iteration LSM reports a standard
error from the sample variance of V_0, which ignores the error of the 252 fitted regressions; this tool measures the
real dispersion and shows that both streams scatter around the same mean.   python tools/stream_dispersion.py [log2 paths] [seeds]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import montecarlooptionspricer_b200 as m  # noqa: E402


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    n_seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    eng = m.Engine(0)
    model = dict(S0=100.0, r=0.05, xi=0.04, H=0.1, eta=1.9, rho=-0.9, dt=1 / 252)
    lsm = dict(r=0.05, strike=100.0, maturity=1.0, dt=1 / 252, is_call=False, poly_order=3, carry=m.MCP_F32)
    for impl, name in (("2", "per-path"), ("3", "pair")):
        os.environ["MCP_GEN_IMPL"] = impl
        prices, ses = [], []
        for s in range(n_seeds):
            out, _ = eng.price_rbergomi_lsm(model, lsm, 1 << lg, 252, seed=1000 + s)
            prices.append(out.price)
            ses.append(out.std_error)
        p = np.array(prices)
        print(f"{name:9s} 2^{lg} paths x {n_seeds} seeds: mean {p.mean():.6f}  sd over seeds {p.std(ddof=1):.6f}  "
              f"sd of the mean {p.std(ddof=1) / np.sqrt(n_seeds):.6f}  reported std_error {np.mean(ses):.6f}", flush=True)
        print("   ", " ".join(f"{x:.6f}" for x in p), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
