"""BASELINE configs 4 and 5 and the batched row driver across the GPUs of one node (run under torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 tools/configs_multi_gpu.py

  config 5  256 contracts (16 strikes x 16 maturities, 2^22 paths each): ranks split the maturities, no collective.
  config 4  2^20 outer x 1000 inner x 50 dates nested duality (GBM): ranks shard the outer paths; one final all-reduce.
  rows      16384 PredictionGen-shaped rows (250 paths, four pricers): ranks take slices of the rows, no collective.
Times are device/host wall per rank, reported as the max over ranks."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import montecarlooptionspricer_b200 as m  # noqa: E402
from test_gpu_rows import make_rows  # noqa: E402


def tmax(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model = dict(S0=100.0, r=0.05, xi=0.04, H=0.1, eta=1.9, rho=-0.9, dt=1.0 / 252.0)
    eng = m.Engine(local)

    # ---- config 5: maturities round-robin over ranks ----
    strikes, mats = np.arange(70.0, 131.0, 4.0), np.arange(1, 17) / 16.0
    eng.price_surface_rbergomi_lsm(model, strikes, mats[:2], 1 << 20, r=0.05, seed=1)  # warm-up
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    px, se, g, l = eng.price_surface_rbergomi_lsm(model, strikes, mats, 1 << 22, r=0.05, poly_order=3, seed=9, mat_first=rank, mat_stride=world)
    t5 = tmax(time.perf_counter() - t0)
    parts = [None] * world
    dist.all_gather_object(parts, px)
    full = np.full_like(px, np.nan)
    for p in parts:
        full[~np.isnan(p)] = p[~np.isnan(p)]

    # ---- config 4: outer paths sharded, final sums all-reduced ----
    uid = [m.Engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    eng4 = m.Engine(local)
    if world > 1:
        eng4.comm_init(rank, world, uid[0])
    n_outer = (1 << 20) // world
    kw = dict(S0=100.0, r=0.05, sigma=0.2, dt=0.02, strike=100.0, is_call=False, n_steps=50, poly_order=3, n_policy_paths=1 << 20, n_inner=1000)
    eng4.gbm_nested_dual(n_outer=1 << 12, seed=1, path_offset=rank << 12, **kw)  # warm-up
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    d = eng4.gbm_nested_dual(n_outer=n_outer, seed=5, path_offset=rank * n_outer, **kw)
    t4 = tmax(time.perf_counter() - t0)

    # ---- rows: slices of the row list ----
    rows = make_rows(np.random.default_rng(1), 16384)
    mine = rows[rank::world]
    eng.price_rows(mine[:64], n_paths=250, seed=0)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, _, _ = eng.price_rows(mine, n_paths=250, seed=1, path_offset=rank * (1 << 32))
    tr = tmax(time.perf_counter() - t0)

    if rank == 0:
        print(f"GPUs {world}: config 5 (256 contracts x 2^22 paths) {t5 * 1e3:.1f} ms, surface complete: {not np.any(np.isnan(full))}, "
              f"ATM 1y put {full[-1, 8]:.4f}")
        print(f"GPUs {world}: config 4 (2^20 outer x 1000 inner x 50 dates) {t4 * 1e3:.1f} ms, lower {d['lower']:.4f} +- {d['lower_se']:.4f}, "
              f"upper {d['upper']:.4f} +- {d['upper_se']:.4f}, outer paths {d['n_outer_global']}")
        print(f"GPUs {world}: rows 16384 x 250 paths x 4 pricers {tr * 1e3:.1f} ms = {16384 / tr:.0f} rows/s")
    eng.close(); eng4.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
