"""Time the LSM sweep implementations against each other on one GPU (or, under torchrun, on the ranks' shards):

    python tools/sweep_timing.py [log2_paths ...]            # default 20 22 23 24 26
    MCP_SWEEP_IMPL: 2 = direct 256-bit loads per step, 3 = TMA ring per step (PDL), 4 = persistent cooperative sweep

Prints, per size and implementation, the best-of-3 device time of mcp_lsm_price (CUDA events inside the library), the
time per step and the fraction of the measured HBM peak on the algorithmic 12 B/path-step."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlooptionspricer_b200 as m  # noqa: E402


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [20, 22, 23, 24, 26]
    peak = 6544.3
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    eng = m.Engine(0)
    mdl = dict(S0=100.0, r=0.05, xi=0.04, H=0.1, eta=1.9, rho=-0.9, dt=1.0 / 252.0)
    for k in sizes:
        n = 1 << k
        ps = eng.pathset(n, 252)
        eng.gen_rbergomi(ps, mdl["S0"], mdl["r"], mdl["xi"], mdl["H"], mdl["eta"], mdl["rho"], mdl["dt"], seed=3)
        for impl in ("2", "3", "4"):
            os.environ["MCP_SWEEP_IMPL"] = impl
            best, price = 1e30, None
            for _ in range(4):
                out = eng.lsm_price(ps, 0.05, 100.0, 1.0, mdl["dt"], False, 3, carry=m.MCP_F32)
                best, price = min(best, out.elapsed_ms), out.price
            us = best * 1e3 / 253
            print(f"2^{k} impl {impl}: lsm {best:8.3f} ms  {us:7.2f} us/step  {12.0 * n / (us * 1e-6) / 1e9 / peak:5.3f} of HBM peak  "
                  f"price {price:.7f} launches {out.n_kernel_launches}", flush=True)
        ps.close()
    eng.close()


if __name__ == "__main__":
    main()
