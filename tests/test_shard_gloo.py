"""The N>1 path on CPU: world_size-2 ``gloo`` run of the point-sharded iteration.

The GPU kernels cannot run here, so each rank forms ITS SHARD's partial reduced camera system with
the NumPy oracle (exactly what ``feba_iterate_assemble`` leaves in the buffer that is all-reduced),
the partial systems are summed with ``torch.distributed.all_reduce`` over gloo, every rank solves the
reduced system, back-substitutes its own points, and the result is compared with the un-sharded
oracle: this covers the partition (``shard.point_owner`` / ``shard_problem``), the additivity of
the Schur-reduced system over point shards, the local<->global unknown maps and the reduction of
``deltasum`` / the variance factor across ranks.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import feba_b200 as fb
from feba_b200 import shard as sh, synth
from oracle import sparse


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prob = synth.make_network(12, 500, 7, 2024, mode=mode, n_control=30 if mode == "mixed" else 0)
        err, xhat0, _ = fb.Buildxhat(prob)
        shd = sh.shard_problem(prob, rank, world)
        loc = shd.prob
        x_loc = shd.local_xhat(xhat0)
        # this rank's partial reduced system
        nb = sparse.normal_blocks(loc, x_loc)
        S, g, Vinv = sparse.reduce(loc, nb)
        red = torch.from_numpy(np.concatenate([S.ravel(), g]))
        dist.all_reduce(red, op=dist.ReduceOp.SUM)
        n = S.shape[0]
        S_all, g_all = red[:n * n].numpy().reshape(n, n), red[n * n:].numpy()
        d_c = sparse.solve_reduced(loc, S_all, g_all, nb.get("Gc"))
        delta_s = sparse.back_substitute(loc, nb, Vinv, d_c)
        delta = sparse.unscale(loc, nb["L"], nb["q"], delta_s)
        d_cam = float(np.sum(np.abs(delta[:shd.u_c])))
        d_pts = torch.tensor([float(np.sum(np.abs(delta[shd.u_c:])))], dtype=torch.float64)
        dist.all_reduce(d_pts, op=dist.ReduceOp.SUM)
        deltasum = d_cam + float(d_pts.item())
        x_new_loc = x_loc + delta
        res = sparse.residuals(loc, nb, delta, x_new_loc)
        ss = torch.tensor([float(np.sum(res["v"][0::2] ** 2)), float(np.sum(res["v"][1::2] ** 2))], dtype=torch.float64)
        parts = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, ss)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x_loc=x_new_loc, deltasum=deltasum, v=res["v"],
                 rows=shd.obs_rows, parts=np.stack([p.numpy() for p in parts]), tie_global=shd.tie_global)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["free", "mixed"])
def test_two_rank_point_sharded_iteration(tmp_path, mode):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, mode, str(tmp_path)), nprocs=world, join=True)
    prob = synth.make_network(12, 500, 7, 2024, mode=mode, n_control=30 if mode == "mixed" else 0)
    err, xhat0, _ = fb.Buildxhat(prob)
    x_ref, ds_ref, st = sparse.iterate(prob, xhat0)
    res_ref = sparse.residuals(prob, st["nb"], st["delta"], x_ref)
    x_glob = np.full(prob.u, np.nan)
    v_glob = np.full(2 * prob.n_obs, np.nan)
    owner = sh.point_owner(prob, world)
    assert set(np.unique(owner)) == {0, 1}
    cnt = np.bincount(owner[prob.obs_pt], minlength=world)
    assert abs(cnt[0] - cnt[1]) < 0.1 * prob.n_obs                     # balanced by observations
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        shd = sh.shard_problem(prob, r, world)
        shd.scatter_xhat(z["x_loc"], x_glob)
        v_glob[2 * z["rows"]] = z["v"][0::2]
        v_glob[2 * z["rows"] + 1] = z["v"][1::2]
        assert abs(float(z["deltasum"]) - ds_ref) < 1e-9 * ds_ref
        stats = sh.combine_stats(z["parts"], prob.n_obs, prob.u, 1 / prob.settings.sigma_x ** 2,
                                 1 / prob.settings.sigma_y ** 2)
        assert abs(stats["sigma02"] - res_ref["sigma02"]) < 1e-9 * res_ref["sigma02"]
    assert not np.isnan(x_glob).any() and not np.isnan(v_glob).any()    # every unknown / row owned once
    assert np.linalg.norm(x_glob - x_ref) < 1e-10 * np.linalg.norm(x_ref)
    assert np.max(np.abs(v_glob - res_ref["v"])) < 1e-8


@pytest.mark.parametrize("world", [3, 4, 8])
def test_partition_covers_every_point_and_observation_once(world):
    """shard_problem for the bench's group sizes (no process group needed): points, observations and
    tie unknowns are split without overlap, the shards are balanced by observation count, local slots
    are consistent, and local <-> global xhat maps invert each other."""
    prob = synth.make_network(16, 1200, 7, 99, mode="mixed", n_control=60)
    err, xhat0, _ = fb.Buildxhat(prob)
    owner = sh.point_owner(prob, world)
    assert owner.shape == (prob.numPts,) and set(np.unique(owner)) == set(range(world))
    assert np.all(np.diff(owner) >= 0)                                   # contiguous CNT ranges
    seen_rows = np.zeros(prob.n_obs, dtype=int)
    seen_pts = np.zeros(prob.numPts, dtype=int)
    seen_tie = np.zeros(prob.numtie, dtype=int)
    x_glob = np.full(prob.u, np.nan)
    counts = []
    for r in range(world):
        shd = sh.shard_problem(prob, r, world)
        loc = shd.prob
        loc.validate()
        seen_rows[shd.obs_rows] += 1
        seen_pts[shd.pts] += 1
        seen_tie[shd.tie_global] += 1
        counts.append(loc.n_obs)
        assert loc.numImg == prob.numImg and loc.numCam == prob.numCam    # full image / camera tables
        assert np.array_equal(shd.pts[loc.obs_pt], prob.obs_pt[shd.obs_rows])
        assert np.array_equal(loc.obs_x, prob.obs_x[shd.obs_rows])
        # local tie slots keep the global TIE order
        assert np.all(np.diff(shd.tie_global) > 0)
        assert np.array_equal(shd.pts[loc.tie_pt], prob.tie_pt[shd.tie_global])
        x_loc = shd.local_xhat(xhat0)
        assert x_loc.size == loc.u and np.array_equal(x_loc[:prob.u_c], xhat0[:prob.u_c])
        shd.scatter_xhat(x_loc, x_glob)
    assert np.all(seen_rows == 1) and np.all(seen_pts == 1) and np.all(seen_tie == 1)
    assert np.array_equal(x_glob, xhat0)
    assert max(counts) - min(counts) < 0.25 * prob.n_obs / world + 50
