"""Point-sharded multi-GPU Gauss-Newton: one process per GPU.

Every observation adds to the normal equations (main.m:424-425), so object points are the shard
unit: all observations of a point live on one rank, which makes V_p, W_p and the point's whole
Schur contribution local; the camera-block sums are additive.  Two forms:

``GroupAdjustment`` (default of bench.py): ``feba_create_shard``.  The library cuts the image block by nested
dissection into one subtree per rank plus shared separators, every rank keeps the points of its subtree,
eliminates its subtree locally and only the shared top part of the reduced system is summed over NVLink; all
collectives (NCCL) are issued inside ``feba_iterate`` on the handle's stream.  Every rank passes the complete
problem and works with the global xhat.

``ShardedAdjustment`` (replicated form, also what the gloo CPU tests exercise): contiguous point ranges, the WHOLE
reduced system is summed by the caller.  Per iteration:

    rank r:  feba_iterate_assemble()            partial S_r, g_r of its points
    all   :  all_reduce(sum) of the reduced-system buffer (NCCL over NVLink; gloo in CPU tests)
    rank r:  feba_iterate_solve()               identical EOP/IOP update everywhere, own points
             (with feba_dist_init the factorisation inside is shared: supertile columns are dealt
              out over the ranks, finished panels are broadcast; see csrc/feba_chol.cu)
    all   :  all_reduce(sum) of sum|delta_points|  -> deltasum of main.m:487

No other exchange exists on the path.  ``shard_problem`` is host-side index bookkeeping only.
"""
from __future__ import annotations

import copy
import os
import sys
from dataclasses import dataclass
from typing import Optional

import numpy as np

from .problem import Problem


@dataclass
class Shard:
    """Local problem of one rank + the maps back to the global unknown vector / PHO rows."""
    prob: Problem
    obs_rows: np.ndarray       # global PHO row of each local observation
    pts: np.ndarray            # global CNT row of each local point
    tie_global: np.ndarray     # global tie index of each local tie (local order)
    u_c: int
    u_global: int

    def local_xhat(self, xhat_global: np.ndarray) -> np.ndarray:
        cols = (self.u_c + 3 * self.tie_global[:, None] + np.arange(3)[None, :]).reshape(-1)
        return np.concatenate([xhat_global[:self.u_c], xhat_global[cols]])

    def scatter_xhat(self, xhat_local: np.ndarray, xhat_global: np.ndarray) -> None:
        """Write this rank's tie coordinates (and the replicated EOP/IOP part) into the global vector."""
        cols = (self.u_c + 3 * self.tie_global[:, None] + np.arange(3)[None, :]).reshape(-1)
        xhat_global[:self.u_c] = xhat_local[:self.u_c]
        xhat_global[cols] = xhat_local[self.u_c:]


def point_owner(prob: Problem, world: int) -> np.ndarray:
    """Contiguous ranges of CNT rows balanced by observation count (fixed, deterministic)."""
    cnt = np.bincount(prob.obs_pt, minlength=prob.numPts).astype(np.int64)
    csum = np.cumsum(cnt)
    total = int(csum[-1]) if csum.size else 0
    # point p goes to the rank whose observation quota its first observation falls in
    first = csum - cnt
    owner = np.minimum((first * world) // max(total, 1), world - 1).astype(np.int32)
    return owner


def shard_problem(prob: Problem, rank: int, world: int) -> Shard:
    if world == 1:
        return Shard(prob=prob, obs_rows=np.arange(prob.n_obs), pts=np.arange(prob.numPts),
                     tie_global=np.arange(prob.numtie), u_c=prob.u_c, u_global=prob.u)
    owner = point_owner(prob, world)
    pts = np.nonzero(owner == rank)[0].astype(np.int64)
    remap = -np.ones(prob.numPts, dtype=np.int64)
    remap[pts] = np.arange(pts.size)
    rows = np.nonzero(owner[prob.obs_pt] == rank)[0]
    tie_g = prob.pt_tie[pts]
    is_tie = tie_g >= 0
    order = np.argsort(tie_g[is_tie], kind="stable")          # local ties keep the global TIE order
    tie_global = tie_g[is_tie][order].astype(np.int64)
    pt_tie_local = -np.ones(pts.size, dtype=np.int32)
    pt_tie_local[np.nonzero(is_tie)[0][order]] = np.arange(tie_global.size, dtype=np.int32)
    tie_pt_local = np.nonzero(is_tie)[0][order].astype(np.int32)
    loc = copy.copy(prob)
    loc.settings = copy.copy(prob.settings)
    loc.obs_x, loc.obs_y = prob.obs_x[rows], prob.obs_y[rows]
    loc.obs_img = prob.obs_img[rows]
    loc.obs_pt = remap[prob.obs_pt[rows]].astype(np.int32)
    loc.xyz0 = prob.xyz0[pts]
    loc.pt_tie, loc.tie_pt = pt_tie_local, tie_pt_local
    loc.point_ids = None if prob.point_ids is None else [prob.point_ids[p] for p in pts]
    return Shard(prob=loc, obs_rows=rows, pts=pts, tie_global=tie_global, u_c=prob.u_c, u_global=prob.u)


class DeviceBuffer:
    """Minimal ``__cuda_array_interface__`` carrier so torch can wrap the library's buffer."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = dict(shape=(count,), typestr="<f8", data=(ptr, False), version=3,
                                             strides=None)


class GroupAdjustment:
    """One rank of a group of GPUs on a nested-dissection plan (feba_create_shard).  ``prob`` is the COMPLETE
    problem on every rank; ``xhat`` vectors are global.  Needs an initialised NCCL process group (the library
    creates its own communicator from an id broadcast through it)."""

    def __init__(self, prob: Problem, group=None, plan: int = 0):
        import torch
        import torch.distributed as dist
        from .lib import Handle, dist_unique_id
        self.dist, self.torch, self.group = dist, torch, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        src = dist.get_global_rank(group, 0) if group is not None else 0
        ids = [dist_unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(ids, src=src, group=group)
        self.h = Handle(prob, plan=plan, group=(self.rank, self.world, ids[0]))
        # kernels, copies and the library's collectives are ordered on torch's current stream (callers make a
        # non-default stream current: the library captures CUDA graphs on it)
        self.h.set_stream(torch.cuda.current_stream().cuda_stream)
        self.shared_factorisation = True

    def iterate(self) -> float:
        return self.h.iterate()

    def iterate_async(self):
        self.h.iterate_async()


class ShardedAdjustment:
    """One rank of a sharded run in the replicated form.  ``handle`` is a ``lib.Handle`` of ``shard.prob`` created
    with ``plan=-1`` (every rank must use the same row order of the reduced system: the identity order)."""

    def __init__(self, handle, shard: Shard, group=None):
        import torch
        import torch.distributed as dist
        self.h, self.shard, self.dist, self.torch, self.group = handle, shard, dist, torch, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._red = None
        self._pk, self._pk_key = None, None
        self.shared_factorisation = False
        # opt-in until it has been timed on the 8-GPU box: exchange the packed lower trapezoids only
        self._packed = os.environ.get("FEBA_PACKED_REDUCE", "0") == "1"
        if self.world > 1 and handle.sparse_info()["active"]:
            # a nested-dissection plan computed from THIS shard's images differs from rank to rank: the summed
            # buffers would not line up
            raise RuntimeError("the replicated form needs the identity row order: create the handle with plan=-1")
        if self.world > 1:
            # kernels, copies and the collective are all ordered on torch's current stream (callers
            # should make a non-default stream current: the library captures CUDA graphs on it)
            handle.set_stream(torch.cuda.current_stream().cuda_stream)
            ptr, count = handle.reduced_dev()
            self._red = torch.as_tensor(DeviceBuffer(ptr, count), device=torch.device("cuda", torch.cuda.current_device()))
            self._scal = torch.zeros(1, dtype=torch.float64, device=self._red.device)
            # factorise the summed system together (feba_dist_init) instead of once per rank;
            # FEBA_DIST_CHOL=0 keeps the replicated solve
            if os.environ.get("FEBA_DIST_CHOL", "1") != "0" and dist.get_backend(group) == "nccl":
                from .lib import FebaError, dist_unique_id
                rank = dist.get_rank(group)
                src = dist.get_global_rank(group, 0) if group is not None else 0
                # feba_dist_init is collective (ncclCommInitRank): make sure EVERY rank can bind NCCL
                # before any rank enters it, otherwise the others would wait for ever
                try:
                    my_id, ok = dist_unique_id(), 1.0
                except FebaError as exc:
                    my_id, ok = None, 0.0
                    print(f"[feba] rank {rank}: {exc.text}", file=sys.stderr)
                flag = torch.tensor([ok], dtype=torch.float64, device=self._red.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
                if float(flag.item()) < 1.0:
                    if rank == 0:
                        print("[feba] NCCL could not be bound on every rank: keeping the replicated solve",
                              file=sys.stderr)
                else:
                    ids = [my_id if rank == 0 else None]
                    dist.broadcast_object_list(ids, src=src, group=group)
                    handle.dist_init(rank, self.world, ids[0])
                    self.shared_factorisation = True

    def _exchange(self):
        """Sum the partial reduced systems over the ranks (between the two halves of an iteration)."""
        if not self._packed:
            self.dist.all_reduce(self._red, op=self.dist.ReduceOp.SUM, group=self.group)
            return
        # FEBA_PACKED_REDUCE=1: only the lower trapezoids the solve half reads (~52 % of the bytes)
        ptr, count = self.h.reduced_pack()
        if self._pk is None or self._pk_key != (ptr, count):
            self._pk = self.torch.as_tensor(DeviceBuffer(ptr, count), device=self._red.device)
            self._pk_key = (ptr, count)
        self.dist.all_reduce(self._pk, op=self.dist.ReduceOp.SUM, group=self.group)
        self.h.reduced_unpack()

    def iterate(self) -> float:
        """One pass of main.m:412-494 over all ranks; returns the global deltasum."""
        if self.world == 1:
            return self.h.iterate()
        self.h.iterate_assemble()
        self._exchange()
        d_cam, d_pts = self.h.iterate_solve()
        self._scal[0] = d_pts
        self.dist.all_reduce(self._scal, op=self.dist.ReduceOp.SUM, group=self.group)
        return d_cam + float(self._scal.item())

    def iterate_async(self):
        """Enqueue one step without reading anything back (bench: inputs resident in HBM)."""
        if self.world == 1:
            self.h.iterate_async()
            return
        self.h.iterate_assemble()
        self._exchange()
        self.h.iterate_solve_async()


def combine_stats(parts, n_obs_total: int, u_total: int, px: float, py: float) -> dict:
    """main.m:594-601 from per-rank (sum vx^2, sum vy^2)."""
    sxx = float(sum(p[0] for p in parts))
    syy = float(sum(p[1] for p in parts))
    rmsx, rmsy = np.sqrt(sxx / n_obs_total), np.sqrt(syy / n_obs_total)
    return dict(RMSx=rmsx, RMSy=rmsy, RMS=np.sqrt(rmsx ** 2 + rmsy ** 2),
                sigma02=(sxx * px + syy * py) / (2 * n_obs_total - u_total))
