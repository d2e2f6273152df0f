"""Shared (column-distributed) factorisation against the replicated solve, same shards, same ranks.

Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/dist_check.py [--workload config3] [--tile 6] [--iters 3]
Rank 0 prints one JSON line; exit code 1 when the two paths differ by more than --tol (relative,
per unknown group).  tests/test_gpu_parity.py runs it when two GPUs are visible.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config3")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--tile", type=int, default=6)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--tol", type=float, default=1e-9)
    ap.add_argument("--packed", action="store_true",
                    help="also run the shared factorisation with the packed exchange (FEBA_PACKED_REDUCE=1)")
    args = ap.parse_args()
    if args.tile:
        os.environ["FEBA_DAG_TILE"] = str(args.tile)
    import torch
    import torch.distributed as dist
    import feba_b200 as fb
    from feba_b200 import shard as sh
    import bench

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    prob, _ = bench.make_workload(args.workload, args.scale)
    shard = sh.shard_problem(prob, rank, world)
    x0 = fb.Buildxhat(shard.prob)[1]
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    out = {}
    for mode in ("0", "1") + (("1p",) if args.packed else ()):
        os.environ["FEBA_DIST_CHOL"] = mode[0]
        os.environ["FEBA_PACKED_REDUCE"] = "1" if mode.endswith("p") else "0"
        h = fb.Handle(shard.prob, plan=-1)       # replicated form: identity row order on every rank
        h.set_stream(stream.cuda_stream)
        adj = sh.ShardedAdjustment(h, shard)
        h.set_xhat(x0)
        trace = [adj.iterate() for _ in range(args.iters)]
        t = h.last_timing()
        out[mode] = (h.get_xhat(), trace, t)
        h.close()
    xa, xb = out["0"][0], out["1"][0]
    u_c = prob.u_c
    err_cam = float(np.max(np.abs(xa[:u_c] - xb[:u_c]) / (np.abs(xa[:u_c]) + 1e-3)))
    err_pts = float(np.max(np.abs(xa[u_c:] - xb[u_c:]) / (np.abs(xa[u_c:]) + 1e-3))) if xa.size > u_c else 0.0
    errs = torch.tensor([err_cam, err_pts], dtype=torch.float64, device="cuda")
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    ok = bool((errs <= args.tol).all().item())
    extra = {}
    if args.packed:
        xp = out["1p"][0]
        ep = torch.tensor([float(np.max(np.abs(xp - xb) / (np.abs(xb) + 1e-3)))], dtype=torch.float64, device="cuda")
        dist.all_reduce(ep, op=dist.ReduceOp.MAX)
        # same numbers, same collective; with more than two ranks the order of the sum may differ per element
        ok = ok and float(ep.item()) <= args.tol
        extra = {"packed_rel_diff_vs_shared": float(ep.item()), "factor_ms_shared_packed": out["1p"][2]["factor_ms"]}
    if rank == 0:
        print(json.dumps({"world": world, "workload": args.workload, "tile": args.tile, "u_c": int(u_c),
                          "err_cam": float(errs[0]), "err_pts": float(errs[1]), "ok": ok,
                          "deltasum_replicated": out["0"][1], "deltasum_shared": out["1"][1],
                          "factor_ms_replicated": out["0"][2]["factor_ms"], "factor_ms_shared": out["1"][2]["factor_ms"], **extra}),
              flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
