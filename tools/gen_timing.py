"""Time the native rBergomi generator alone (CUDA events around the kernel, inside the library), best of 4:
    python tools/gen_timing.py [log2_paths=26] [n_steps=252 ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlooptionspricer_b200 as m  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 26
steps = [int(x) for x in sys.argv[2:]] or [252]
eng = m.Engine(0)
args = (100.0, 0.05, 0.04, 0.1, 1.9, -0.9, 1.0 / 252.0)
eng.set_profiling(True)
for n in steps:
    ps = eng.pathset(1 << k, n)
    best = 1e30
    for rep in range(4):
        eng.gen_rbergomi(ps, *args, seed=11)
        best = min(best, eng.profile()["gen_kernel_ms"])
    print(f"2^{k} x {n}: {best:8.3f} ms  ({(1 << k) * n / best / 1e6:.1f} G path-steps/s, {(1 << k) * (n + 1) * 4 / best / 1e6:.0f} GB/s stored)", flush=True)
    ps.close()
eng.close()
