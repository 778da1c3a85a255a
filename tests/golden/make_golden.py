"""Generate the committed golden fixtures from the REFERENCE ITSELF (oracle/_ref/libmcp_ref.so = the reference's
own translation units compiled by oracle/Makefile).  Run in the authoring container (needs /root/reference):

    python tests/golden/make_golden.py

The reference ships no golden vectors (CMakeLists.txt:70-82 are echo tests), so these pins are created here:
every array below is an OUTPUT OF THE REFERENCE CODE on seeded, injected normal draws.  Files are small (<1 MB).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402


def main():
    O.build(ref=True)
    ref = O.ref()
    port = O.port()
    rng = np.random.default_rng(20261018)

    # (1) rough-vol paths through the reference's private fGn members, explicit parameters (BASELINE config 2 model)
    prm = dict(S0=100.0, r=0.05, xi=0.04, H=0.1, eta=1.9, rho=-0.9, dt=1.0 / 252.0)
    out = {}
    for tag, P, n in [("small", 64, 8), ("cfg2", 48, 252), ("pow2", 16, 256), ("n50", 32, 50)]:
        d = rng.standard_normal((P, 4 * n)).astype(np.float32)
        paths, X, v = ref.rbergomi_paths(prm["S0"], prm["r"], prm["xi"], prm["H"], prm["eta"], prm["rho"], prm["dt"], n,
                                         d.astype(np.float64), want_xv=True)
        out[f"rb_{tag}_draws"] = d
        out[f"rb_{tag}_paths"] = paths
        out[f"rb_{tag}_X"] = X
        out[f"rb_{tag}_v"] = v
    out["rb_params"] = np.array([prm[k] for k in ("S0", "r", "xi", "H", "eta", "rho", "dt")])
    out["phi_252"] = ref.rbergomi_phi(252, 0.1, 1.0 / 252.0)
    np.savez_compressed(os.path.join(HERE, "rbergomi_ref.npz"), **out)

    # (2) the UNMODIFIED GenerateStockPricePaths (estimators + generator) on a synthetic history, injected draws
    hist = 100.0 * np.exp(np.cumsum(0.0126 * rng.standard_normal(300)))
    steps, P = 21, 40
    d = rng.standard_normal(P * 4 * steps)
    paths, used = ref.generate_paths(hist, steps, P, d)
    assert used == d.size
    est = ref.estimate_params(hist)
    np.savez_compressed(os.path.join(HERE, "generate_paths_ref.npz"), hist=hist, draws=d, paths=paths,
                        est=np.array([est[k] for k in ("xi", "H", "eta", "rho", "S0")]))

    # (3) LSM prices from the reference's LSMPricer.cpp on GBM paths (config 1 model; fp32-rounded path values),
    #     plus the other three pricers on the same paths
    z = rng.standard_normal((4096, 50)).astype(np.float32).astype(np.float64)
    gp = port.gbm_paths(100.0, 0.05, 0.2, 0.02, 50, z).astype(np.float32)
    res = {"paths_f32": gp}
    gp64 = gp.astype(np.float64)
    for p in (1, 2, 3):
        res[f"lsm_put_p{p}"] = np.array(ref.lsm_price(gp64, 0.05, 100.0, 1.0, 0.02, False, p))
        res[f"lsm_call_p{p}"] = np.array(ref.lsm_price(gp64, 0.05, 95.0, 1.0, 0.02, True, p))
    res["lsm_put_p2_cut"] = np.array(ref.lsm_price(gp64, 0.05, 100.0, 0.5, 0.02, False, 2))
    res["lsm_put_itm0"] = np.array(ref.lsm_price(gp64, 0.05, 110.0, 1.0, 0.02, False, 3))
    res["asym_put"] = np.array(ref.asymptotic_price(gp64, 0.05, 100.0, 1.0, 0.02, False, 0.2, 0.0))
    res["asym_call"] = np.array(ref.asymptotic_price(gp64, 0.05, 100.0, 1.0, 0.02, True, 0.2, 0.01))
    res["mart_put_p2"] = np.array(ref.martingale_price(gp64, 0.05, 100.0, 1.0, 0.02, False, 2, 5))
    res["mart_call_p2"] = np.array(ref.martingale_price(gp64, 0.05, 100.0, 1.0, 0.02, True, 2, 5))
    o = port.lsm(gp64, 0.05, 100.0, 1.0, 0.02, False, 3)
    res["port_first_ex_p3"] = o["first_ex"]
    res["port_stderr_p3"] = np.array(o["stderr"])
    np.savez_compressed(os.path.join(HERE, "pricers_ref.npz"), **res)
    for f in ("rbergomi_ref.npz", "generate_paths_ref.npz", "pricers_ref.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
