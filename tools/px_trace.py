"""Where does a step of the persistent sweep go?  MCP_PX_TRACE=1 makes every CTA stamp %globaltimer at its milestones;
this prints, per size, the medians over steps of: streaming time (constants in hand -> thread 0 through its tiles), spread
between the first and the last worker to send its row, and the reducer's chain (rows in -> folded -> exchanged -> broadcast)
and the time from the last row sent to the last worker holding the next constants.

    python tools/px_trace.py [log2_paths ...]
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlooptionspricer_b200 as m  # noqa: E402


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [20, 23, 26]
    os.environ["MCP_PX_TRACE"] = "1"
    os.environ["MCP_SWEEP_IMPL"] = "4"
    eng = m.Engine(0)
    L = eng._L
    L.mcp_debug_px_trace.restype = C.c_int64
    L.mcp_debug_px_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    mdl = dict(S0=100.0, r=0.05, xi=0.04, H=0.1, eta=1.9, rho=-0.9, dt=1.0 / 252.0)
    for k in sizes:
        ps = eng.pathset(1 << k, 252)
        eng.gen_rbergomi(ps, mdl["S0"], mdl["r"], mdl["xi"], mdl["H"], mdl["eta"], mdl["rho"], mdl["dt"], seed=3)
        for _ in range(2):
            out = eng.lsm_price(ps, 0.05, 100.0, 1.0, mdl["dt"], False, 3, carry=m.MCP_F32)
        rows, cols = C.c_int(), C.c_int()
        n = L.mcp_debug_px_trace(eng._h, None, 0, C.byref(rows), C.byref(cols))
        buf = np.zeros(n, dtype=np.uint64)
        L.mcp_debug_px_trace(eng._h, buf.ctypes.data_as(C.c_void_p), n, None, None)
        t = buf.reshape(rows.value, cols.value, 4).astype(np.float64) * 1e-3  # us
        wk, rd = t[:, 1:, :], t[:, 0, :]
        steps = slice(2, rows.value - 2)
        stream = np.median((wk[steps, :, 1] - wk[steps, :, 0]).max(axis=1))
        stream_med = np.median(np.median(wk[steps, :, 1] - wk[steps, :, 0], axis=1))
        send_spread = np.median(wk[steps, :, 2].max(axis=1) - wk[steps, :, 2].min(axis=1))
        last_sent = wk[steps, :, 2].max(axis=1)
        gather = np.median(rd[steps, 0] - last_sent)
        fold = np.median(rd[steps, 1] - rd[steps, 0])
        xchg = np.median(rd[steps, 2] - rd[steps, 1])
        solve = np.median(rd[steps, 3] - rd[steps, 2])
        nxt = wk[3:rows.value - 1, :, 0]
        seen = np.median(nxt.max(axis=1) - rd[2:rows.value - 2, 3])
        period = np.median(np.diff(rd[steps, 3]))
        print(f"2^{k}: lsm {out.elapsed_ms:.3f} ms, step period {period:.2f} us | stream max {stream:.2f} med {stream_med:.2f} | send spread {send_spread:.2f} | "
              f"last row -> gathered {gather:.2f} | fold {fold:.2f} | exchange {xchg:.2f} | solve+bcast {solve:.2f} | bcast -> last worker ready {seen:.2f}", flush=True)
        ps.close()
    eng.close()


if __name__ == "__main__" and not os.environ.get("PX_TRACE_WORKERS"):
    main()


def per_worker(k=26):
    """Which workers are slow, and are they the same ones every step?  Prints the distribution of per-worker mean stream time."""
    os.environ["MCP_PX_TRACE"] = "1"
    os.environ["MCP_SWEEP_IMPL"] = "4"
    eng = m.Engine(0)
    L = eng._L
    L.mcp_debug_px_trace.restype = C.c_int64
    L.mcp_debug_px_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    ps = eng.pathset(1 << k, 252)
    eng.gen_rbergomi(ps, 100.0, 0.05, 0.04, 0.1, 1.9, -0.9, 1.0 / 252.0, seed=3)
    for _ in range(2):
        eng.lsm_price(ps, 0.05, 100.0, 1.0, 1.0 / 252.0, False, 3, carry=m.MCP_F32)
    rows, cols = C.c_int(), C.c_int()
    n = L.mcp_debug_px_trace(eng._h, None, 0, C.byref(rows), C.byref(cols))
    buf = np.zeros(n, dtype=np.uint64)
    L.mcp_debug_px_trace(eng._h, buf.ctypes.data_as(C.c_void_p), n, None, None)
    t = buf.reshape(rows.value, cols.value, 4).astype(np.float64) * 1e-3
    wk = t[2:-2, 1:, :]
    dur = wk[:, :, 1] - wk[:, :, 0]            # [step][worker]
    mean_w = dur.mean(axis=0)
    order = np.argsort(mean_w)
    print(f"2^{k}: per-worker mean stream time: min {mean_w.min():.2f} p10 {np.percentile(mean_w, 10):.2f} median {np.median(mean_w):.2f} "
          f"p90 {np.percentile(mean_w, 90):.2f} max {mean_w.max():.2f} us; per-step max/median {np.median(dur.max(axis=1) / np.median(dur, axis=1)):.3f}")
    print("   slowest workers:", [(int(w), round(float(mean_w[w]), 2)) for w in order[-8:]], " fastest:", [(int(w), round(float(mean_w[w]), 2)) for w in order[:6]])
    # how persistent is the ranking: correlation of a worker's time in even vs odd steps
    a, b = dur[0::2].mean(axis=0), dur[1::2].mean(axis=0)
    print(f"   even/odd-step correlation of per-worker means: {np.corrcoef(a, b)[0, 1]:.3f}; per-step noise (std over steps of one worker, median): "
          f"{np.median(dur.std(axis=0)):.2f} us")
    ps.close()
    eng.close()


if __name__ == "__main__" and os.environ.get("PX_TRACE_WORKERS"):
    per_worker(int(os.environ["PX_TRACE_WORKERS"]))
