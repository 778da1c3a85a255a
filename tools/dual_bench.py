"""BASELINE config 4 at full size on one GPU: 2^20 outer x 1000 inner paths, 50 exercise dates, GBM (config 1 model).
python tools/dual_bench.py [log2_outer=20] [n_inner=1000]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlooptionspricer_b200 as m  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n_inner = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
eng = m.Engine(0)
for rep in range(2):
    res = eng.gbm_nested_dual(S0=100.0, r=0.05, sigma=0.2, dt=0.02, strike=100.0, is_call=False, n_steps=50, poly_order=3,
                              n_policy_paths=1 << 20, n_outer=1 << k, n_inner=n_inner, seed=5 + rep)
print(f"2^{k} outer x {n_inner} inner x 50 dates: lower {res['lower']:.4f} +- {res['lower_se']:.4f}, upper {res['upper']:.4f} +- {res['upper_se']:.4f} "
      f"(Bermudan-50 value 6.0786); policy {res['policy_ms']:.1f} ms, outer paths {res['outer_ms']:.1f} ms, nested + combine {res['nested_ms']:.1f} ms")
eng.close()
