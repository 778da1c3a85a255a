"""world_size-2 `gloo` tests (CPU): the host-side multi-GPU logic -- shard arithmetic, unique-id hand-off, and the
sharded moment-all-reduce formulation of the sweep (numpy model of the device algorithm) against the oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_paths_partition():
    from montecarlooptionspricer_b200.dist import shard_paths
    for n, w in [(1 << 26, 8), (1 << 26, 1), (1000, 3), (7, 8), (0, 2)]:
        parts = [shard_paths(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == n
        for (o0, c0), (o1, _) in zip(parts, parts[1:]):
            assert o0 + c0 == o1
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        shard_paths(10, 2, 2)


def _worker(rank, world, port_no, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from montecarlooptionspricer_b200.dist import broadcast_blob, combine_mean_and_stderr, shard_paths
        from oracle import oracle as O
        from sharded_model import sharded_lsm
        blob = broadcast_blob(bytes(range(128)) if rank == 0 else None)
        assert blob == bytes(range(128))

        port = O.port()
        n_total, n = 6001, 30  # ragged split on purpose
        z = np.random.default_rng(42).standard_normal((n_total, n))  # identical on every rank
        paths = port.gbm_paths(100.0, 0.05, 0.25, 1.0 / n, n, z)
        off, cnt = shard_paths(n_total, rank, world)

        def allreduce(a):
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())
            dist.all_reduce(t)
            return t.numpy()

        out = {}
        for p, is_call, K, T in [(2, False, 100.0, 1.0), (3, False, 104.0, 1.0), (3, True, 98.0, 0.7)]:
            price, se, first = sharded_lsm(paths[off:off + cnt], 0.05, K, T, 1.0 / n, is_call, p, allreduce)
            want = port.lsm(paths, 0.05, K, T, 1.0 / n, is_call, p)
            assert abs(price - want["price"]) < 1e-9 * want["price"], (price, want["price"])
            assert abs(se - want["stderr"]) < 1e-7 * want["stderr"] + 1e-12 * abs(want["price"])
            assert np.array_equal(first, want["first_ex"][off:off + cnt])
            out[(p, is_call)] = price
        m, s = combine_mean_and_stderr(10.0, 4.0, 5)
        assert m == 2.0 and abs(s - (4.0 / 4 / 5) ** 0.5) < 1e-15
        q.put((rank, "ok", out))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_sharded_sweep_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, payload in res:
        assert status == "ok", f"rank {rank}:\n{payload}"
    assert res[0][2] == res[1][2]  # every rank holds the same global price
