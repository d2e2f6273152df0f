"""Group of GPUs on a nested-dissection plan (feba_create_shard) against one GPU and the oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/group_check.py
        [--images 400] [--points 40000] [--mode free|mixed]

Every rank creates its shard of the SAME problem; the whole adjustment runs through the collective calls; rank 0
also runs it alone on one GPU and with the oracle and compares (xhat 1e-9 group-normalised, v 1e-8 max|v|,
sigma02 1e-8, first increment 1e-9, identical iteration counts and bit-identical results on every rank)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import feba_b200 as fb
    from feba_b200 import shard as sh
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=400)
    ap.add_argument("--points", type=int, default=40000)
    ap.add_argument("--rays", type=int, default=8)
    ap.add_argument("--mode", default="both")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    torch.cuda.set_stream(torch.cuda.Stream())
    ok = True
    for mode, kw in (("free", {}), ("mixed", {"n_control": 80})):
        if args.mode not in ("both", mode):
            continue
        prob = fb.synth.make_network(args.images, args.points, args.rays, 4711, mode=mode, **kw)
        err, x0, _ = fb.Buildxhat(prob)
        adj = sh.GroupAdjustment(prob)
        h = adj.h
        h.set_xhat(x0)
        h.iterate()
        d1 = h.get_delta()
        h.set_xhat(x0)
        it, trace = h.solve()
        res = h.residuals()
        xh = h.get_xhat()
        info = h.plan_info()
        # owned-only transfers: the EOP/IOP part + this rank's tie points, bit-identical to the collective result;
        # every tie point is owned by exactly one rank; upload of the owned parts + collective download = identity
        xo = np.full(h.u, np.nan)
        h.get_xhat_owned(xo)
        got = ~np.isnan(xo)
        owned_ok = bool(np.array_equal(xo[got], xh[got])) and int(got.sum()) == prob.u_c + 3 * h.num_owned_ties()
        n_own = torch.tensor([float(h.num_owned_ties())], dtype=torch.float64, device="cuda")
        dist.all_reduce(n_own)
        owned_ok = owned_ok and int(n_own.item()) == prob.numtie
        h.set_xhat_owned(xh)
        owned_ok = owned_ok and bool(np.array_equal(h.get_xhat(), xh))
        # the same through a page-locked buffer (the device addresses it directly: no staging, no host gather)
        pin = torch.full((h.u,), float("nan"), dtype=torch.float64).pin_memory()
        xp = pin.numpy()
        h.get_xhat_owned(xp)
        owned_ok = owned_ok and bool(np.array_equal(xp[got], xh[got])) and bool(np.all(np.isnan(xp[~got])))
        xp[:] = xh
        h.set_xhat_owned(xp)
        owned_ok = owned_ok and bool(np.array_equal(h.get_xhat(), xh))
        flag_own = torch.tensor([1.0 if owned_ok else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(flag_own, op=dist.ReduceOp.MIN)
        owned_ok = flag_own.item() > 0
        # every rank holds the same complete result
        chk = torch.tensor([float(np.sum(xh)), float(np.sum(res["v"])), float(res["sigma02"]), float(it)],
                           dtype=torch.float64, device="cuda")
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(lo, hi))
        h.close()
        if rank == 0:
            from oracle import cport
            from oracle.compare import group_rel
            with fb.Handle(prob) as h1:
                h1.set_xhat(x0)
                h1.iterate()
                e1 = h1.get_delta()
                h1.set_xhat(x0)
                it1, tr1 = h1.solve()
                r1 = h1.residuals()
                x1 = h1.get_xhat()
            ref = cport.CPort(prob).gauss_newton(x0)
            e_step = group_rel(prob, d1, e1)
            e_x1, e_xo = group_rel(prob, xh, x1), group_rel(prob, xh, ref["xhat"])
            e_v = np.max(np.abs(res["v"] - ref["v"])) / np.max(np.abs(ref["v"]))
            e_s = abs(res["sigma02"] - ref["sigma02"]) / ref["sigma02"]
            print(f"[{mode}] world {world} plan {info}\n[{mode}] iterations {it} (one GPU {it1}, oracle {ref['iterations']})  "
                  f"first step vs one GPU {e_step:.2e}  xhat vs one GPU {e_x1:.2e}  vs oracle {e_xo:.2e}  v {e_v:.2e}  "
                  f"sigma02 {e_s:.2e}  identical on all ranks {same}  owned-only transfers {owned_ok}")
            good = (it == it1 == ref["iterations"] and e_step < 1e-9 and e_x1 < 1e-9 and e_xo < 1e-9 and e_v < 1e-8
                    and e_s < 1e-8 and same and owned_ok and info["world"] == world)
            ok = ok and good
        dist.barrier()
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.broadcast(flag, src=0)
    if rank == 0:
        print("group form ok" if ok else "group form FAILED")
    dist.destroy_process_group()
    return 0 if flag.item() > 0 else 1


if __name__ == "__main__":
    sys.exit(main())
