"""The GROUP form of the N>1 path (feba_create_shard, the default of bench.py at N > 1) on CPU: a world_size-2
``gloo`` run of one Gauss-Newton step with the library's own plan.

The CUDA kernels cannot run here.  Every process is one rank: it builds the nested-dissection plan with the
library's host code (csrc/feba_order.h, compiled by tests/test_reduced_plan_host.py), takes the object points
``plan_point_owner`` gives it, forms ITS partial reduced system with the NumPy oracle (what the rank's assembly
leaves in S), and runs the rank's three phases of the solve half (``RankSolve``: border + scaling, elimination of its
own subtree, replicated factorisation of the shared top, backward substitution) with the THREE exchanges of
``feba_iterate`` as real ``torch.distributed.all_reduce`` calls over gloo:

    1. diagonal of S            (small)     2. lower trapezoid of the shared top part   (the large one)
    3. rows of the solution     (small)

The step must equal the bordered solution of main.m:428-437 in extended precision, be identical on both ranks, and
every tie point must be back-substituted by exactly one rank.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import feba_b200 as fb
from feba_b200 import synth
from oracle import sparse
from tests.test_reduced_plan_host import Plan, RankSolve, load_host, reduced_of, truth_step


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _allreduce(a):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        host = load_host()
        prob = synth.make_network(120, 9000, 8, 99, mode="free")
        err, x0, _ = fb.Buildxhat(prob)
        plan = Plan(host, prob, world=world, leaf=16, tile_max=2)
        own_seg = plan.point_owner()
        pt_owner = -np.ones(prob.numPts, dtype=np.int64)
        pt_owner[plan.seg_pt] = own_seg
        mine = pt_owner == rank
        S_r, g_r = reduced_of(prob, x0, mine)                       # this rank's assembly
        Gc = sparse.normal_blocks(prob, x0)["Gc"]                    # inner-constraint rows: EOP part, known to every rank
        rk = RankSolve(host, plan, S_r, g_r, Gc, rank, world)
        dg = _allreduce(rk.diag())                                   # exchange 1
        top = rk.eliminate(dg)
        top_sum = _allreduce(top)                                    # exchange 2: only the shared top part travels
        sol = _allreduce(rk.finish(top_sum))                         # exchange 3
        delta_c = rk.delta(sol)
        n_top = top.shape[0]
        owned = np.zeros(prob.numtie)
        tie_of = prob.pt_tie[np.nonzero(mine)[0]]
        owned[tie_of[tie_of >= 0]] = 1.0
        owned = _allreduce(owned)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), delta_c=delta_c, n_top=n_top, n_pad=plan.n_pad, owned=owned,
                 n_obs_here=int(np.sum(mine[prob.obs_pt])))
        plan.close()
    finally:
        dist.destroy_process_group()


def test_two_rank_group_step_over_gloo(tmp_path):
    world = 2
    load_host()                                                      # build the host library once, before the ranks race for it
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    # both ranks end with the same camera increments
    assert np.array_equal(r[0]["delta_c"], r[1]["delta_c"])
    # only a part of the system crossed the wire in the large exchange
    assert 0 < int(r[0]["n_top"]) - 64 < int(r[0]["n_pad"]) // 2
    # every tie point is back-substituted by exactly one rank; the observations are split roughly evenly
    assert np.all(r[0]["owned"] == 1.0)
    prob = synth.make_network(120, 9000, 8, 99, mode="free")
    n_here = [int(x["n_obs_here"]) for x in r]
    assert sum(n_here) == prob.n_obs and max(n_here) < 0.75 * prob.n_obs
    # and the step is the bordered solution of main.m:428-437
    err, x0, _ = fb.Buildxhat(prob)
    nbk = sparse.normal_blocks(prob, x0)
    S, g, _ = sparse.reduce(prob, nbk)
    truth = truth_step(S, g, nbk["Gc"])
    rel = np.linalg.norm(r[0]["delta_c"] - truth) / np.linalg.norm(truth)
    assert rel < 1e-8, rel
