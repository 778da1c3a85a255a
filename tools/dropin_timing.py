"""Run the reference's unmodified PredictionGen driver built against (a) its own CPU classes and (b) the B200 plugins
on the same synthetic CSVs and print the driver's own progress/timing lines (PredictionGen.cpp:850-863)."""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dropin_data as D  # noqa: E402

n_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 96
threads = sys.argv[2] if len(sys.argv) > 2 else "16"
for name in ("PredictionGen_ref", "PredictionGen_b200"):
    exe = os.path.join(ROOT, "oracle", "_ref", name)
    d = tempfile.mkdtemp(prefix=name)
    D.write_inputs(d, n_rows)
    t0 = time.time()
    r = subprocess.run([exe], cwd=d, capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS=threads), timeout=900)
    wall = time.time() - t0
    last = [l for l in r.stdout.replace("\r", "\n").splitlines() if "Progress" in l][-1:]
    print(f"{name}: rc={r.returncode} wall {wall:.1f}s (includes the driver's 30 s keep-alive sleep) | {last[0] if last else r.stdout[-200:]}")
