#!/usr/bin/env python
"""bench.py -- the reference's headline workload on B200: rBergomi American-put LSM price, 64M paths x 252
steps, cubic basis (BASELINE.json configs[2]), strong-scaled over N GPUs of one node.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own CPU code on ALL host cores of the box

A "step" = one whole pricing pass: Philox normals -> rough-vol paths (time-major fp32 slab in HBM) -> LSM
backward induction (fp64 moments) -> price + standard error.  Prints ONE JSON line.

  value        path-steps/s, device-resident loop (paths + carry + tables live in HBM; CUDA events on the launching stream)
  e2e          same metric through the public host call mcp_price_rbergomi_lsm (host parameter structs in, host result out;
               per-step table H2D + result D2H + all synchronisation inside the wall-clock region); the bytes are COUNTED
               by the library (mcp_copy_counters), not computed here
  roofline     dominant kernel by time, measured live with CUDA events; `bound` names its real limiter; `kernels` lists
               every hot kernel with its fraction of the measured HBM peak on the algorithmic bytes
  parity_mode  the bit-exact (fp64 carry, fp64 decisions) sweep timed on the same slab, against its 20 B/path-step
  shard_parity (N > 1) a 2^20-path fp64-carry problem priced sharded and unsharded: relative price difference and
               first-exercise mismatches -- measured in this run, outside the timed regions
  policy_value what the fitted exercise rule earns on an independent path set (lower bound, exact standard error) next to the
               in-sample value-iteration price the reference's estimator reports
  configs      the other BASELINE configs (1, 2, 4, 5) and the reference's row workload, one number each
  cpu_baseline the reference's unmodified generator + LSM (oracle/_ref) on a bounded sample, all host cores, every N
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_STEPS = 252
MODEL = dict(S0=100.0, r=0.05, xi=0.04, H=0.1, eta=1.9, rho=-0.9, dt=1.0 / 252.0)
STRIKE, MATURITY, POLY = 100.0, 1.0, 3
METRIC, UNIT = "path-steps/sec", "path-steps/s"
GEN_BYTES_PER_PATHSTEP = 4.0    # one fp32 store, time-major
LSM_BYTES_PER_PATHSTEP = 12.0   # S read 4 + carry read 4 + carry write 4 (fp32 carry); 20 with the fp64 carry


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_constants() -> dict:
    """Counter values that only ncu can give (DRAM bytes per launch, pipe utilisation).  They come from the committed
    captures under profiles/ and every one carries the file it was read from; nothing here is typed into bench.py."""
    p = os.path.join(ROOT, "profiles", "ncu_constants.json")
    try:
        with open(p) as f:
            return json.load(f)
    except Exception:
        return {}


def host_threads() -> int:
    """All cores this process may run on -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, dev: int):
        self.dev, self.rows, self.proc = dev, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_history(n=400, seed=20261018):
    import numpy as np
    rng = np.random.default_rng(seed)
    return 100.0 * np.exp(np.cumsum(0.0126 * rng.standard_normal(n)))  # ~20% annualised vol


def run_reference_rows(paths_per_row: int, rows_per_thread: int):
    """Reference CPU arm: its own GenerateStockPricePaths + LSM::PredictOptionPrice, parallel over rows exactly as
    src/core/PredictionGen.cpp:542-546 does, on every core of the box.  Returns (path-steps/s, dict)."""
    from oracle import oracle as O
    cores = host_threads()
    if O.have_ref():
        ref = O.ref()
        hist = synthetic_history()
        n_rows = rows_per_thread * cores
        out = ref.bench_rows(hist, n_rows, paths_per_row, N_STEPS, 0.05, float(hist[-1]), False, POLY, threads=cores)  # omp_set_num_threads(cores)
        ps = n_rows * paths_per_row * N_STEPS
        return ps / out["seconds"], dict(kind="reference", cores=cores, seconds=out["seconds"], rows=n_rows, paths_per_row=paths_per_row,
                                         gen_share=out["gen_seconds_sum"] / max(1e-9, out["gen_seconds_sum"] + out["lsm_seconds_sum"]))
    # port fallback (single thread): oracle restatement
    port = O.port()
    t0 = time.perf_counter()
    d = port.rbergomi_draws(1, 0, paths_per_row, N_STEPS, MODEL["rho"])
    paths = port.rbergomi_paths(MODEL["S0"], MODEL["r"], MODEL["xi"], MODEL["H"], MODEL["eta"], MODEL["rho"], MODEL["dt"], N_STEPS, d)
    port.lsm(paths, 0.05, STRIKE, MATURITY, MODEL["dt"], False, POLY)
    sec = time.perf_counter() - t0
    return paths_per_row * N_STEPS / sec, dict(kind="port", cores=1, seconds=sec, rows=1, paths_per_row=paths_per_row, gen_share=None)


CPU_NOTE = ("omp over rows as PredictionGen.cpp:542-546, RNG as shipped; the LSM's Eigen bdcSvd().solve is the restated SVD of oracle/lstsq_svd.h "
            "(Eigen is absent from this image; its speed relative to Eigen 3.4 is unknown)")


def bench_reference(args, rank, world, emit):
    if rank != 0:
        return
    vals, info = [], None
    for i in range(args.warmup + args.steps):
        v, info = run_reference_rows(paths_per_row=8192, rows_per_thread=6)
        if i >= args.warmup:
            vals.append((v, info["seconds"]))
    value = sum(v for v, _ in vals) / len(vals)
    sample = (f"{info['rows']} independent rows x {info['paths_per_row']} paths x {N_STEPS} steps (generate + LSM p={POLY}) per step, {info['cores']} threads; "
              + CPU_NOTE)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(s for _, s in vals) / len(vals), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "rBergomi American put LSM, 252 steps, cubic basis (BASELINE configs[2]) -- bounded CPU sample", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------ the other configs
def make_rows(rng, n_rows):
    """PredictionGen-shaped rows (src/core/PredictionGen.cpp:700-791): dte-dependent step counts, estimated-looking models."""
    import numpy as np
    rows = []
    for k in range(n_rows):
        dte = int(rng.choice([0, 30, 61, 91, 150, 240]))
        T = dte / 365.0
        S0 = float(rng.uniform(20, 300))
        model = dict(S0=S0, r=0.04, xi=float(rng.uniform(0.01, 0.09)), H=float(rng.uniform(0.05, 0.6)), eta=float(rng.uniform(0.02, 1.9)),
                     rho=float(rng.uniform(-0.9, 0.0)), dt=1.0 / 252.0)
        rows.append(dict(model=model, n_steps=int(np.floor(T * 252.0)), is_call=bool(k % 2), r=0.04, strike=S0 * (1.0 - float(rng.choice([-0.05, 0.0, 0.03]))),
                         maturity=T, dt=1.0 / 252.0, sigma=0.2, dividend=0.01))
    return rows


def other_configs(m, eng, eng_solo, rank, world, tmax, barrier, peak):
    """BASELINE configs 1, 2, 4, 5 and the reference's row workload: one driver-visible number each, with a one-number
    parity spot check (the oracle is the checker here, never the thing timed).  Under torchrun config 4 shards its outer
    paths (one final all-reduce), config 5 splits maturities and the rows split by slices (no collective)."""
    import numpy as np
    out = {}
    try:
        from oracle import oracle as O
        port = O.port()
    except Exception:
        port = None

    # ---- config 1: 100k x 50 GBM American put through the exact reference call shape (host rows in, price out) ----
    if rank == 0:
        n, N = 50, 100_000
        rng = np.random.default_rng(1)
        z = rng.standard_normal((N, n)).astype(np.float32).astype(np.float64)
        dt = 1.0 / n
        paths = 100.0 * np.exp(np.cumsum(np.concatenate([np.zeros((N, 1)), (0.05 - 0.02) * dt + 0.2 * np.sqrt(dt) * z], axis=1), axis=1))
        eng_solo.lsm_price_host_rows(paths, 0.05, 100.0, 1.0, dt, False, 3)  # warm-up at full size: pinned staging, scratch and slab pool are sized once
        h0, d0 = eng_solo.copy_counters()
        t0 = time.perf_counter()
        px = eng_solo.lsm_price_host_rows(paths, 0.05, 100.0, 1.0, dt, False, 3)
        sec = time.perf_counter() - t0
        h1, d1 = eng_solo.copy_counters()
        c1 = {"workload": "American put LSM under GBM, 100k paths x 50 steps, p=3, mcp_lsm_price_host_rows (host vector<vector<double>> rows in, price out)",
              "e2e_ms": 1e3 * sec, "path_steps_per_s": N * n / sec, "h2d_bytes": h1 - h0, "d2h_bytes": d1 - d0, "price": px,
              "bermudan50_known": 6.0786}
        if port is not None:
            want = port.lsm(paths, 0.05, 100.0, 1.0, dt, False, 3)
            c1["parity_rel_vs_oracle"] = abs(px - want["price"]) / want["price"]
        out["cfg1"] = c1

    # ---- config 2: 2^20 x 252 rBergomi path generation on one GPU ----
    if rank == 0:
        n_paths = 1 << 20
        ps = eng_solo.pathset(n_paths, N_STEPS)
        args = (MODEL["S0"], MODEL["r"], MODEL["xi"], MODEL["H"], MODEL["eta"], MODEL["rho"], MODEL["dt"])
        eng_solo.gen_rbergomi(ps, *args, seed=5)
        eng_solo.set_profiling(True)
        eng_solo.gen_rbergomi(ps, *args, seed=6)
        gms = eng_solo.profile()["gen_kernel_ms"]
        eng_solo.set_profiling(False)
        c2 = {"workload": "rBergomi (H=0.1, eta=1.9, rho=-0.9) path generation, 2^20 paths x 252 steps, native Philox", "gen_kernel_ms": gms,
              "path_steps_per_s": n_paths * N_STEPS / (gms * 1e-3)}
        ps.close()
        if port is not None:  # dump -> oracle replay on 2048 paths of the same stream (path i depends on (seed, i) only)
            small = eng_solo.pathset(2048, N_STEPS)
            used = eng_solo.gen_rbergomi(small, *args, seed=6, dump=True)
            slab = small.download_timemajor()
            want = port.rbergomi_paths(*args, N_STEPS, used.astype(np.float64))
            c2["parity_max_rel_vs_oracle_2048_paths"] = float(np.max(np.abs(slab.T - want) / want))
            small.close()
        out["cfg2"] = c2

    # ---- config 5: 16 strikes x 16 maturities x 2^22 paths, maturities round-robin over the ranks ----
    strikes, mats = np.arange(70.0, 131.0, 4.0), np.arange(1, 17) / 16.0
    eng_solo.price_surface_rbergomi_lsm(MODEL, strikes, mats[-1:], 1 << 22, r=0.05, seed=1, mat_first=0, mat_stride=1)  # warm-up: the workspaces of the longest ladder
    barrier()
    t0 = time.perf_counter()
    px, se, gms, lms = eng_solo.price_surface_rbergomi_lsm(MODEL, strikes, mats, 1 << 22, r=0.05, poly_order=3, seed=9, mat_first=rank, mat_stride=world)
    t5 = tmax(time.perf_counter() - t0)
    steps5 = [int(np.floor(T * 252.0)) for T in mats]
    cps = sum(steps5[i] for i in range(rank, len(mats), world)) * len(strikes) * (1 << 22)  # contract-path-steps of this rank
    if rank == 0:
        # ladder kernel accounting: per contract and path-step 8 B of carry + 8 B of slab shared by the 16 strikes
        bytes_cps = 8.0 + 8.0 / len(strikes)
        out["cfg5"] = {"workload": "256 contracts = 16 strikes x 16 maturities, 2^22 paths each, rBergomi LSM p=3, maturities split over ranks (no collective)",
                       "time_s": t5, "gen_ms_rank0": gms, "lsm_ms_rank0": lms, "atm_1y_put": float(px[-1, 8]) if not np.isnan(px[-1, 8]) else None,
                       "lsm_multi_kernel_frac_hbm_rank0": (cps * bytes_cps / (lms * 1e-3) / 1e9 / peak) if lms > 0 else None,
                       "lsm_multi_bytes_per_contract_path_step": bytes_cps, "ncu": ncu_constants().get("strike_ladder")}
        # spot check: the ATM strike of the first maturity this rank owns, priced alone through mcp_price_rbergomi_lsm-shaped calls
        ps = eng_solo.pathset(1 << 22, steps5[0])
        eng_solo.gen_rbergomi(ps, MODEL["S0"], MODEL["r"], MODEL["xi"], MODEL["H"], MODEL["eta"], MODEL["rho"], MODEL["dt"],
                              seed=(9 + 0x9E3779B97F4A7C15) & ((1 << 64) - 1))  # the surface call's seed of maturity 0 (csrc/surface.cu)
        alone = eng_solo.lsm_price(ps, 0.05, float(strikes[8]), float(mats[0]), MODEL["dt"], False, 3, carry=m.MCP_F32)
        ps.close()
        out["cfg5"]["parity_rel_ladder_vs_single_contract"] = abs(px[0, 8] - alone.price) / alone.price if alone.price > 0 else None

    # ---- config 4: nested duality 2^20 outer x 1000 inner x 50 dates (GBM), outer paths sharded ----
    n_outer = (1 << 20) // world
    kw = dict(S0=100.0, r=0.05, sigma=0.2, dt=0.02, strike=100.0, is_call=False, n_steps=50, poly_order=3, n_policy_paths=1 << 20, n_inner=1000)
    eng.gbm_nested_dual(n_outer=1 << 12, seed=1, path_offset=rank << 12, **kw)  # warm-up
    barrier()
    t0 = time.perf_counter()
    d = eng.gbm_nested_dual(n_outer=n_outer, seed=5, path_offset=rank * n_outer, **kw)
    t4 = tmax(time.perf_counter() - t0)
    if rank == 0:
        out["cfg4"] = {"workload": "Andersen-Broadie nested duality under GBM, 2^20 outer x 1000 inner x 50 dates [new: no reference algorithm, parity unpinned]",
                       "time_s": t4, "lower": d["lower"], "lower_se": d["lower_se"], "upper": d["upper"], "upper_se": d["upper_se"],
                       "outer_paths": d["n_outer_global"], "inner_path_steps_per_s": (1 << 20) * 1000 * (50 * 51 // 2) / t4,
                       "brackets_bermudan50_known_6.0786": bool(d["lower"] - 4 * d["lower_se"] < 6.0786 < d["upper"] + 4 * d["upper_se"])}

    # ---- rows: the reference's own workload (250 paths, four pricers per row), slices of the row list ----
    rows = make_rows(np.random.default_rng(1), 16384)
    mine = m.engine.rows_to_array(rows[rank::world])
    eng_solo.price_rows(mine, n_paths=250, seed=0)  # warm-up at full size: slabs, staging and tables of the batch are allocated once
    barrier()
    t0 = time.perf_counter()
    res, gms, pms = eng_solo.price_rows(mine, n_paths=250, seed=1, path_offset=rank * (1 << 32))
    tr = tmax(time.perf_counter() - t0)
    if rank == 0:
        out["rows"] = {"workload": "16384 PredictionGen-shaped rows x 250 paths x 4 pricers (mcp_price_rows), rows split over ranks",
                       "time_s": tr, "rows_per_s": 16384 / tr, "gen_ms_rank0": gms, "price_ms_rank0": pms,
                       "finite": bool(np.all(np.isfinite(res)))}
        k = next(i for i, r in enumerate(rows[0::world]) if r["n_steps"] >= 1)   # spot check: one row through the per-row entry point
        row = rows[0::world][k]
        md = row["model"]
        os.environ["MCP_GEN_IMPL"] = "0"   # the batch uses the generic generator; pin the per-row call to the same kernel
        ps = eng_solo.pathset(250, row["n_steps"])
        eng_solo.gen_rbergomi(ps, md["S0"], md["r"], md["xi"], md["H"], md["eta"], md["rho"], md["dt"], seed=1, path_offset=k * 250)
        one = eng_solo.lsm_price(ps, row["r"], row["strike"], row["maturity"], row["dt"], row["is_call"], 2, carry=m.MCP_F64)
        ps.close()
        os.environ.pop("MCP_GEN_IMPL", None)
        out["rows"]["parity_rel_lsm_batch_vs_single_row"] = abs(res[k][2] - one.price) / max(1e-12, abs(one.price))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--paths-log2", type=int, default=26, help="total paths = 2^k (default 26 = BASELINE config 3)")
    ap.add_argument("--carry", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs (1, 2, 4, 5, rows)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: everything libraries print there (NCCL's version banner ...) is sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    if args.impl == "reference":
        bench_reference(args, rank, world, emit)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import montecarlooptionspricer_b200 as m

    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n_total = 1 << args.paths_log2
    n_loc = n_total // world
    path_offset = rank * n_loc
    carry = m.MCP_F32 if args.carry == "f32" else m.MCP_F64
    lsm_bytes = LSM_BYTES_PER_PATHSTEP if args.carry == "f32" else 20.0

    stream = torch.cuda.current_stream()
    eng = m.Engine(local_rank, stream=stream.cuda_stream)
    if world > 1:
        uid = [m.Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(rank, world, uid[0])
    eng_solo = m.Engine(local_rank, stream=stream.cuda_stream)  # no communicator: unsharded checks, per-rank configs

    peer_mem = world > 1 and eng.comm_uses_peer_memory()
    ps = eng.pathset(n_loc, N_STEPS)
    model = dict(MODEL)
    lsm = dict(r=MODEL["r"], strike=STRIKE, maturity=MATURITY, dt=MODEL["dt"], is_call=False, poly_order=POLY, carry=carry)

    def step_device(seed):
        eng.gen_rbergomi(ps, MODEL["S0"], MODEL["r"], MODEL["xi"], MODEL["H"], MODEL["eta"], MODEL["rho"], MODEL["dt"],
                         seed=seed, path_offset=path_offset)
        return eng.lsm_price(ps, MODEL["r"], STRIKE, MATURITY, MODEL["dt"], False, POLY, carry=carry)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def tmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up ----
    for w in range(max(args.warmup, 0)):
        out = step_device(1000 + w)
    barrier()

    # ---- timed region A: device-resident loop, CUDA events on the launching stream ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    lsm_ms_sum, price, se = 0.0, None, None
    for k in range(args.steps):
        out = step_device(1 + k)
        lsm_ms_sum += out.elapsed_ms
        price, se = out.price, out.std_error
    ev1.record(stream)
    barrier()
    launches = eng.launch_count - launches0
    ms_per_step = tmax(ev0.elapsed_time(ev1)) / args.steps
    value = n_total * N_STEPS / (ms_per_step * 1e-3)

    # ---- parity mode (bit-exact decisions: fp64 carry) on the slab of the last step; outside the timed regions ----
    parity_mode = None
    if args.carry == "f32":
        eng.lsm_price(ps, MODEL["r"], STRIKE, MATURITY, MODEL["dt"], False, POLY, carry=m.MCP_F64)
        barrier()
        t64 = []
        for _ in range(2):
            o64 = eng.lsm_price(ps, MODEL["r"], STRIKE, MATURITY, MODEL["dt"], False, POLY, carry=m.MCP_F64)
            t64.append(o64.elapsed_ms)
        lsm64_ms = tmax(min(t64))
        parity_mode = {"what": "fp64 carry, every decision in fp64 on the stored values (exercise indices bit-exact with the oracle)",
                       "lsm_ms": lsm64_ms, "us_per_sweep_step": 1e3 * lsm64_ms / (N_STEPS + 1), "algorithmic_bytes_per_path_step": 20.0,
                       "frac_hbm": 20.0 * n_loc * (N_STEPS + 1) / (lsm64_ms * 1e-3) / 1e9 / peaks()[0],
                       "price": o64.price, "rel_diff_to_fp32_carry_price": abs(o64.price - price) / o64.price,
                       "ncu": (ncu_constants().get("sweep_parity") if world == 1 else None)}

    # ---- what the fitted policy earns on fresh paths (mcp_lsm_policy_value; outside the timed regions) ----
    policy = None
    try:
        fit = eng.lsm_price(ps, MODEL["r"], STRIKE, MATURITY, MODEL["dt"], False, POLY, basis=m.MCP_BASIS_STANDARDISED,
                            carry=m.MCP_F32 if args.carry == "f32" else m.MCP_F64, want_coeffs=True)
        eng.gen_rbergomi(ps, MODEL["S0"], MODEL["r"], MODEL["xi"], MODEL["H"], MODEL["eta"], MODEL["rho"], MODEL["dt"], seed=4242, path_offset=path_offset)
        oos, stop = eng.lsm_policy_value(ps, fit.coeffs, MODEL["r"], STRIKE, MATURITY, MODEL["dt"], False, POLY)
        policy = {"what": "mean discounted payoff of the fitted exercise rule on an INDEPENDENT path set of the same size: a lower bound in expectation, "
                          "with an exact standard error (the in-sample `price` carries fitted values and is biased high; its std_error ignores regression noise)",
                  "in_sample_price": fit.price, "in_sample_std_error": fit.std_error, "policy_value": oos.price, "policy_value_std_error": oos.std_error,
                  "mean_stopping_step": stop, "pass_ms": oos.elapsed_ms}
    except Exception as e:  # never let a reporting extra take the bench line down
        policy = {"error": str(e)[:200]}

    # ---- timed region B: end to end through the public host call (host structs in, host result out) ----
    ps.close()  # the host call owns (and caches) its own slab
    for w in range(2):
        eng.price_rbergomi_lsm(model, lsm, n_loc, N_STEPS, seed=2000 + w, path_offset=path_offset)
    barrier()
    h2d0, d2h0 = eng.copy_counters()
    t0 = time.perf_counter()
    for k in range(args.steps):
        out_e, gen_ms_e = eng.price_rbergomi_lsm(model, lsm, n_loc, N_STEPS, seed=1 + k, path_offset=path_offset)
    barrier()
    e2e_s = tmax(time.perf_counter() - t0)
    clocks = sampler.stop()   # sampled every 20 ms from the start of timed region A to the end of timed region B
    clocks["window"] = "timed regions A (device loop) and B (host-call loop) and the parity-mode pass between them"
    h2d1, d2h1 = eng.copy_counters()
    e2e_value = n_total * N_STEPS * args.steps / e2e_s

    # ---- per-kernel durations (CUDA events around each launch of the hot kernels; separate, untimed pass) ----
    eng.set_profiling(True)
    eng.price_rbergomi_lsm(model, lsm, n_loc, N_STEPS, seed=77, path_offset=path_offset)
    prof = eng.profile()
    eng.set_profiling(False)
    peak, peak_src = peaks()
    ncu = ncu_constants()
    gen_gbs = GEN_BYTES_PER_PATHSTEP * n_loc * (N_STEPS + 1) / (prof["gen_kernel_ms"] * 1e-3) / 1e9
    sweep_steps = max(1, prof["n_sweep_steps"])
    sweep_step_ms = prof["sweep_kernels_ms"] / sweep_steps
    sweep_gbs = lsm_bytes * n_loc / (sweep_step_ms * 1e-3) / 1e9
    persistent = prof["n_sweep_launches"] == 1 and sweep_steps > 1
    l2_fit = n_loc * 12 <= (104 << 20)  # the library's own rule (MCP_L2_RESIDENT_MB): two slab rows + the carry stay in L2
    g_ncu = ncu.get("generator", {})
    # a capture is quoted only for the regime it was taken in (L2-resident persistent sweep / HBM-streaming per-step sweep)
    s_ncu = ncu.get("sweep_persistent", {}) if (persistent and l2_fit) else ncu.get("sweep_per_step", {}) if not persistent else {}
    kernels = {
        "rbergomi_paths_kernel": {
            "bound": "fp32 pipe / issue slots (Philox IMAD.WIDE + packed fp32 + SFU); HBM fraction reported because the contract asks for it",
            "launches_per_step": 1, "ms": prof["gen_kernel_ms"], "algorithmic_bytes": GEN_BYTES_PER_PATHSTEP * n_loc * (N_STEPS + 1),
            "achieved_gbs": gen_gbs, "frac_hbm": gen_gbs / peak,
            "traffic": (g_ncu["dram_bytes_per_path_step"] * n_loc * (N_STEPS + 1)) if "dram_bytes_per_path_step" in g_ncu else None,
            "ncu": g_ncu or None},
        "lsm_sweep_kernel": {
            "bound": "hbm" if not (persistent and l2_fit) else "fp32 pipe (the step's working set is L2-resident: DRAM traffic = the one new slab row)",
            "kernel": "lsm_persist_kernel (one cooperative launch, all steps)" if persistent else "lsm_sweep_tma_kernel (one launch per step, PDL)",
            "launches_per_step": prof["n_sweep_launches"], "sweep_steps": sweep_steps, "avg_ms_per_sweep_step": sweep_step_ms,
            "total_ms": prof["sweep_kernels_ms"], "algorithmic_bytes_per_sweep_step": lsm_bytes * n_loc,
            "achieved_gbs": sweep_gbs, "frac_hbm": sweep_gbs / peak, "frac_hbm_nominal_8000": sweep_gbs / 8000.0,
            "traffic": (s_ncu["dram_bytes_per_path_step"] * n_loc) if ("dram_bytes_per_path_step" in s_ncu and args.carry == "f32") else None,
            "ncu": s_ncu or None,
            "lsm_total_ms_incl_solves_and_collectives": prof["lsm_total_ms"]},
    }
    dominant = "rbergomi_paths_kernel" if prof["gen_kernel_ms"] >= prof["sweep_kernels_ms"] else "lsm_sweep_kernel"
    dk = kernels[dominant]
    hbm_bound = dominant == "lsm_sweep_kernel" and not (persistent and l2_fit)
    roofline = {"kernel": dominant, "bound": "hbm" if hbm_bound else "fp32-pipe", "achieved": dk["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": dk["achieved_gbs"] / peak, "traffic": dk["traffic"], "peak_source": peak_src,
                "frac_of_bound": (dk["ncu"] or {}).get("bound_pipe_utilisation") if not hbm_bound else dk["achieved_gbs"] / peak,
                "frac_of_bound_source": (dk["ncu"] or {}).get("source") if not hbm_bound else "live (CUDA events)",
                "note": "achieved/peak/frac are the dominant kernel's ALGORITHMIC bytes over its live-timed duration against the measured HBM copy peak, as the "
                        "contract prescribes; `bound` names what actually limits it and frac_of_bound is that resource's measured utilisation "
                        "(ncu capture named in frac_of_bound_source).  The HBM-bound kernel of the step is lsm_sweep_kernel (frac_hbm in `kernels`).",
                "step_share": {"rbergomi_paths_kernel": prof["gen_kernel_ms"] / (prof["gen_kernel_ms"] + prof["lsm_total_ms"]),
                               "lsm_sweep_kernel": prof["sweep_kernels_ms"] / (prof["gen_kernel_ms"] + prof["lsm_total_ms"])},
                "kernels": kernels}

    # ---- sharded vs unsharded parity, measured here (N > 1): 2^20 paths, fp64 carry ----
    shard_parity = None
    if world > 1:
        n_chk = 1 << 20
        n_chk_loc = n_chk // world
        lsm64 = dict(lsm, carry=m.MCP_F64)
        psc = eng.pathset(n_chk_loc, N_STEPS)
        eng.gen_rbergomi(psc, MODEL["S0"], MODEL["r"], MODEL["xi"], MODEL["H"], MODEL["eta"], MODEL["rho"], MODEL["dt"], seed=11, path_offset=rank * n_chk_loc)
        sh = eng.lsm_price(psc, MODEL["r"], STRIKE, MATURITY, MODEL["dt"], False, POLY, carry=m.MCP_F64, want_first_exercise=True)
        sh32 = eng.lsm_price(psc, MODEL["r"], STRIKE, MATURITY, MODEL["dt"], False, POLY, carry=m.MCP_F32)
        psc.close()
        psu = eng_solo.pathset(n_chk, N_STEPS)
        eng_solo.gen_rbergomi(psu, MODEL["S0"], MODEL["r"], MODEL["xi"], MODEL["H"], MODEL["eta"], MODEL["rho"], MODEL["dt"], seed=11)
        un = eng_solo.lsm_price(psu, MODEL["r"], STRIKE, MATURITY, MODEL["dt"], False, POLY, carry=m.MCP_F64, want_first_exercise=True)
        un32 = eng_solo.lsm_price(psu, MODEL["r"], STRIKE, MATURITY, MODEL["dt"], False, POLY, carry=m.MCP_F32)
        psu.close()
        mism = int(np.count_nonzero(sh.first_exercise != un.first_exercise[rank * n_chk_loc:(rank + 1) * n_chk_loc]))
        t = torch.tensor([float(mism), abs(sh.price - un.price) / un.price, abs(sh32.price - un32.price) / un32.price], dtype=torch.float64, device="cuda")
        tm = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        shard_parity = {"paths": n_chk, "carry": "f64", "shard_parity_rel": float(tm[1].item()), "first_exercise_mismatches": int(t[0].item()),
                        "sharded_price": sh.price, "unsharded_price": un.price, "n_paths_global": sh.n_paths_global,
                        "fp32_carry_shard_parity_rel": float(tm[2].item())}
    _ = lsm_bytes

    # ---- the other configs (driver-visible numbers) ----
    configs = None
    if not args.no_configs:
        try:
            configs = other_configs(m, eng, eng_solo, rank, world, tmax, barrier, peak)
        except Exception as e:  # never take the headline down
            configs = {"error": f"{type(e).__name__}: {e}"[:300]}
            barrier()

    # ---- CPU baseline beside it (rank 0, every N): the reference's own code on a bounded sample, every host core ----
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            v, info = run_reference_rows(paths_per_row=8192, rows_per_thread=12)
            cpu = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                   "sample": f"{info['rows']} independent rows x {info['paths_per_row']} paths x {N_STEPS} steps, generate+LSM p={POLY}, "
                             f"{info['cores']} threads, {info['seconds']:.1f} s wall; " + CPU_NOTE}
        except Exception as e:  # the checker must never take the bench down
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)[:200]}
    barrier()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 paths / f64 regression moments / " + ("f32" if args.carry == "f32" else "f64") + " carry",
            "data": "synthetic (native Philox4x32-10 normals, fixed seeds)",
            "config": {"workload": f"rBergomi (H=0.1, eta=1.9, rho=-0.9) American put LSM, {n_total} paths x {N_STEPS} steps, cubic basis "
                                   f"(BASELINE configs[2]); generate + price each step",
                       "paths_total": n_total, "paths_per_gpu": n_loc, "n_steps": N_STEPS, "poly_order": POLY,
                       "l2": "inputs exceed L2 (slab %.1f GB per GPU, written and re-read every step)" % (n_loc * (N_STEPS + 1) * 4 / 1e9),
                       "parallelism": (f"paths sharded x{world}; per-step all-reduce of {3 * POLY + 2} fp64 moments "
                                       + ("inside the persistent sweep kernel over NVLink peer memory (CUDA IPC mailboxes, one hop)" if peer_mem
                                          else "by ncclAllReduce" if world > 1 else "(single GPU: none)"))},
            "time_to_price_s": ms_per_step * 1e-3, "lsm_price_time_s": lsm_ms_sum / args.steps * 1e-3,
            "price": price, "std_error": se,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": (h2d1 - h2d0) / args.steps, "d2h_bytes_per_step": (d2h1 - d2h0) / args.steps,
                    "bytes_counted_by": "mcp_copy_counters (every cudaMemcpy the library issues)",
                    "ms_per_step": 1e3 * e2e_s / args.steps,
                    "call": "mcp_price_rbergomi_lsm (host parameter structs in, host result struct out)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "parity_mode": parity_mode,
        "policy_value": policy,
            "shard_parity": shard_parity,
            "configs": configs,
            "cpu_baseline": cpu,
        }
        emit(line)
    eng_solo.close()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
