"""Multi-GPU consistency check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py [log2_paths=22]

Every rank prices its shard of the same global path set twice -- once with the per-step moment all-reduce inside the
sweep kernel over NVLink peer memory (CUDA IPC mailboxes), once through ncclAllReduce -- and rank 0 also prices the
whole set alone.  The three prices must agree to rounding (the path set is identical: Philox is keyed by the global
path id; only the summation order of the moments differs)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlooptionspricer_b200 as m  # noqa: E402


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_total = 1 << k
    n_loc = n_total // world
    model = dict(S0=100.0, r=0.05, xi=0.04, H=0.1, eta=1.9, rho=-0.9, dt=1.0 / 252.0)
    out = {}
    for impl in ("p2p", "nccl"):
        os.environ["MCP_COMM_IMPL"] = impl
        eng = m.Engine(local)
        uid = [m.Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(rank, world, uid[0])
        peer = eng.comm_uses_peer_memory()
        for carry in (m.MCP_F32, m.MCP_F64):
            lsm = dict(r=0.05, strike=100.0, maturity=1.0, dt=1.0 / 252.0, is_call=False, poly_order=3, carry=carry)
            res, _ = eng.price_rbergomi_lsm(model, lsm, n_loc, 252, seed=11, path_offset=rank * n_loc)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier(); torch.cuda.synchronize()
            t = []
            for _ in range(3):
                res, gen_ms = eng.price_rbergomi_lsm(model, lsm, n_loc, 252, seed=11, path_offset=rank * n_loc)
                t.append(res.elapsed_ms)
            out[(impl, carry)] = (res.price, res.std_error, res.n_paths_global, min(t), peer)
        eng.close()
    if rank == 0:
        eng = m.Engine(local)
        for carry in (m.MCP_F32, m.MCP_F64):
            lsm = dict(r=0.05, strike=100.0, maturity=1.0, dt=1.0 / 252.0, is_call=False, poly_order=3, carry=carry)
            res, _ = eng.price_rbergomi_lsm(model, lsm, n_total, 252, seed=11, path_offset=0)
            out[("single", carry)] = (res.price, res.std_error, res.n_paths_global, res.elapsed_ms, False)
        eng.close()
        ok = True
        for carry in (m.MCP_F32, m.MCP_F64):
            ref = out[("single", carry)]
            for impl in ("p2p", "nccl"):
                got = out[(impl, carry)]
                rel = abs(got[0] - ref[0]) / ref[0]
                tol = 1e-9 if carry == m.MCP_F64 else 2e-6
                good = rel < tol and got[2] == n_total and abs(got[1] - ref[1]) < 1e-6 * ref[1] + 1e-12
                ok = ok and good and (impl != "p2p" or got[4])
                print(f"world {world} carry {'f32' if carry == m.MCP_F32 else 'f64'} {impl:5s}: price {got[0]:.12f} (single {ref[0]:.12f}, rel {rel:.2e}) "
                      f"lsm {got[3]:.2f} ms peer_memory={got[4]} {'OK' if good else 'MISMATCH'}")
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
