"""Plan of the reduced camera system (csrc/feba_order.h) on the CPU: nested-dissection row order, supertiles,
pattern with fill, subtree owners of a group of GPUs -- and a numpy emulation of what the device does with it.

The header is compiled with g++ (tests/host_model/model_host.cpp).  Checked here, without a GPU:
  * structure: every unknown has one row, nodes are padded to whole 64-blocks, images of different subtrees are
    never adjacent (separator property), datum images sit in the root, owners follow the tree;
  * the masked supertile factorisation over the plan's tiles equals the unmasked one (pattern + fill complete) and
    never touches a structurally zero tile;
  * the whole solve half as the library runs it -- sparse-datum border, Jacobi scaling, masked factorisation,
    14x14 border, backward substitution -- gives the bordered solution of main.m:428-437 (extended precision);
  * the GROUP algorithm of feba_create_shard (world = 2, 4): every rank assembles the points it owns, eliminates
    its own subtrees, the shared top part is summed, the top is factorised by everybody, rows of the solution are
    summed -- gives the same step as one rank.
The device kernels and the NCCL exchanges themselves are covered by the GPU suite.
"""
import copy
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg as sla
import scipy.sparse as sp

import feba_b200 as fb
from oracle import exact, sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_model", "model_host.cpp")
HDRS = [os.path.join(ROOT, "fish-eye_bundle_adjustment_b200", "csrc", f)
        for f in ("feba_model.cuh", "feba_sparse.h", "feba_order.h", "feba_chunks.h")]
LIB = os.path.join(ROOT, "tests", "_build", "libfeba_model_host.so")
_pd, _pi, _pb = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_ubyte)


def load_host():
    newest = max(os.path.getmtime(p) for p in [SRC] + HDRS)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-pthread",
                        "-Wno-unknown-pragmas", SRC, "-o", LIB], check=True)
    lib = C.CDLL(LIB)
    lib.feba_host_sparse_border.argtypes = [_pd, _pd]
    lib.feba_host_adjacency.argtypes = [C.c_int, C.c_int, _pi, _pi, _pb, C.c_int, _pi, _pi, C.c_int]
    lib.feba_host_plan_masked.restype = C.c_void_p
    lib.feba_host_plan_masked.argtypes = [C.c_int] * 3 + [_pi, _pi, _pd] + [C.c_int] * 5
    lib.feba_host_plan_free.argtypes = [C.c_void_p]
    lib.feba_host_plan_info.argtypes = [C.c_void_p, _pi, _pd]
    lib.feba_host_plan_arrays.argtypes = [C.c_void_p, _pi, _pi, _pi, _pb, _pi, _pi, _pi, _pi, _pi]
    lib.feba_host_plan_point_owner.argtypes = [C.c_void_p, C.c_int, _pi, _pi, _pi]
    return lib


@pytest.fixture(scope="module")
def host():
    return load_host()


def sorted_by_point(prob):
    """Observations grouped by object point as feba_create does (stable): seg_start, simg, seg_pt."""
    order = np.argsort(prob.obs_pt, kind="stable")
    pts, first = np.unique(prob.obs_pt[order], return_index=True)
    seg_start = np.concatenate([first, [prob.n_obs]]).astype(np.int32)
    return seg_start, np.ascontiguousarray(prob.obs_img[order], dtype=np.int32), pts.astype(np.int32), order


class Plan:
    def __init__(self, lib, prob, world=1, max_depth=-1, leaf=12, tile_max=2):
        self.lib = lib
        s = prob.settings
        self.ui, self.n_img = s.u_perimage, prob.numImg
        self.cam_rows = s.u_percam * prob.numCam
        seg_start, simg, seg_pt, _ = sorted_by_point(prob)
        self.seg_start, self.simg, self.seg_pt = seg_start, simg, seg_pt
        seg_tie = np.ascontiguousarray((prob.pt_tie[seg_pt] >= 0).astype(np.uint8))
        ptr = np.zeros(self.n_img + 1, dtype=np.int32)
        nnz = lib.feba_host_adjacency(self.n_img, len(seg_pt), seg_start.ctypes.data_as(_pi), simg.ctypes.data_as(_pi),
                                      seg_tie.ctypes.data_as(_pb), int(world > 1), ptr.ctypes.data_as(_pi), None, 0)
        idx = np.zeros(max(nnz, 1), dtype=np.int32)
        lib.feba_host_adjacency(self.n_img, len(seg_pt), seg_start.ctypes.data_as(_pi), simg.ctypes.data_as(_pi),
                                seg_tie.ctypes.data_as(_pb), int(world > 1), ptr.ctypes.data_as(_pi),
                                idx.ctypes.data_as(_pi), nnz)
        self.adj_ptr, self.adj_idx = ptr, idx[:nnz]
        pos = np.ascontiguousarray(prob.eop0[:, :3], dtype=np.float64)
        self.p = lib.feba_host_plan_masked(self.n_img, self.ui, self.cam_rows, ptr.ctypes.data_as(_pi),
                                           idx.ctypes.data_as(_pi), pos.ctypes.data_as(_pd),
                                           int(s.Inner_Constraints), world, max_depth, leaf, tile_max)
        info = (C.c_int * 8)()
        fl = (C.c_double * 2)()
        lib.feba_host_plan_info(self.p, info, fl)
        (self.n_pad, self.NT, self.off_cam, self.n_nodes, n_datum, self.world, self.top_tile0,
         self.chain_blocks) = [int(v) for v in info]
        self.flop, self.flop_dense = float(fl[0]), float(fl[1])
        nb = self.n_pad // 64
        self.img_row = np.zeros(self.n_img, dtype=np.int32)
        self.row_ext = np.zeros(self.n_pad, dtype=np.int32)
        self.tile_b0 = np.zeros(self.NT + 1, dtype=np.int32)
        nz = np.zeros((self.NT + 1, self.NT + 1), dtype=np.uint8)
        self.tile_node = np.zeros(self.NT, dtype=np.int32)
        self.datum = np.zeros(max(n_datum, 1), dtype=np.int32)
        self.img_node = np.zeros(self.n_img, dtype=np.int32)
        node_info = np.zeros((self.n_nodes, 6), dtype=np.int32)
        self.first_block = np.zeros(nb, dtype=np.int32)
        lib.feba_host_plan_arrays(self.p, self.img_row.ctypes.data_as(_pi), self.row_ext.ctypes.data_as(_pi),
                                  self.tile_b0.ctypes.data_as(_pi), nz.ctypes.data_as(_pb),
                                  self.tile_node.ctypes.data_as(_pi), self.datum.ctypes.data_as(_pi),
                                  self.img_node.ctypes.data_as(_pi), node_info.ctypes.data_as(_pi),
                                  self.first_block.ctypes.data_as(_pi))
        self.datum = self.datum[:n_datum]
        self.nz = nz.astype(bool)
        self.node_parent, self.node_depth, self.node_owner = node_info[:, 0], node_info[:, 1], node_info[:, 2]
        self.tile_owner = self.node_owner[self.tile_node]
        # rows of the xhat entries of the EOP/IOP part
        n_red = self.ui * self.n_img + self.cam_rows
        self.row_of = np.zeros(n_red, dtype=np.int64)
        used = self.row_ext >= 0
        self.row_of[self.row_ext[used]] = np.nonzero(used)[0]
        assert used.sum() == n_red

    def point_owner(self):
        own = np.zeros(len(self.seg_pt), dtype=np.int32)
        rc = self.lib.feba_host_plan_point_owner(self.p, len(self.seg_pt), self.seg_start.ctypes.data_as(_pi),
                                                 self.simg.ctypes.data_as(_pi), own.ctypes.data_as(_pi))
        assert rc == 0
        return own

    def tile_rows(self, t):
        if t == self.NT:
            return slice(self.n_pad, self.n_pad + 64)
        return slice(64 * self.tile_b0[t], 64 * self.tile_b0[t + 1])

    def close(self):
        self.lib.feba_host_plan_free(self.p)


def masked_tiles_cholesky(A, plan, nz, k_range, owner_rank=None):
    """chol_tiles: tasks of the columns k in k_range whose owner is owner_rank (or shared), on non-zero tiles."""
    NT = plan.NT
    on = (lambda i, j: True) if nz is None else (lambda i, j: bool(nz[i, j]))
    for k in k_range:
        if owner_rank is not None and plan.tile_owner[k] >= 0 and plan.tile_owner[k] != owner_rank:
            continue
        ks = plan.tile_rows(k)
        A[ks, ks] = np.linalg.cholesky(np.tril(A[ks, ks]) + np.tril(A[ks, ks], -1).T)
        rows = [i for i in range(k + 1, NT + 1) if on(i, k)]
        for i in rows:
            sl = plan.tile_rows(i)
            A[sl, ks] = sla.solve_triangular(A[ks, ks], A[sl, ks].T, lower=True).T
        for i in rows:
            for j in rows:
                if j > i or (j == NT and i != NT):
                    continue
                si, sj = plan.tile_rows(i), plan.tile_rows(j)
                A[si, sj] -= A[si, ks] @ A[sj, ks].T
    return A


def local_problem(prob, pts_mask):
    """The observations of the points in pts_mask (all tables kept global): what a rank of a group holds."""
    rows = np.nonzero(pts_mask[prob.obs_pt])[0]
    loc = copy.copy(prob)
    loc.obs_x, loc.obs_y = prob.obs_x[rows], prob.obs_y[rows]
    loc.obs_img, loc.obs_pt = prob.obs_img[rows], prob.obs_pt[rows]
    return loc


def reduced_of(prob, x0, pts_mask=None):
    """S, g (xhat order of the EOP/IOP part) of the points in pts_mask, and G of the whole network."""
    sub = prob if pts_mask is None else local_problem(prob, pts_mask)
    nbk = sparse.normal_blocks(sub, x0)
    if sub.numtie and pts_mask is not None:
        # tie points without observations here: unit V so that the inverse exists (their W, u_p are zero)
        empty = np.ones(sub.numtie, dtype=bool)
        t = sub.pt_tie[sub.obs_pt]
        empty[t[t >= 0]] = False
        nbk["V"][empty] = np.eye(3)
    S, g, _ = sparse.reduce(sub, nbk)
    return S, g


class RankSolve:
    """The solve half of ONE rank of a group as feba_iterate runs it (csrc/feba_api.cu::enqueue_solve), in three
    phases separated by the three exchanges: diag() -> [sum of diag(S)] -> eliminate() -> [sum of the shared top
    part] -> finish() -> [sum of the solution rows].  world = 1: the same code without exchanges."""

    def __init__(self, host, plan, S_r, g_r, Gc, rank, world):
        self.host, self.plan, self.rank, self.world = host, plan, rank, world
        n_pad = plan.n_pad
        self.ro = ro = plan.row_of
        self.blk_owner = np.repeat(plan.tile_owner, np.diff(plan.tile_b0))
        row_owner = np.repeat(self.blk_owner, 64)
        self.mine = ((row_owner == rank) | ((row_owner < 0) & (rank == 0))) if world > 1 else np.ones(n_pad, bool)
        # conditioned G (any non-singular C gives the same bordered solution): unit column norms
        Gn = Gc / np.linalg.norm(Gc, axis=0)[None, :]
        self.G = np.zeros((n_pad, 7))
        self.G[ro] = Gn
        self.E = np.zeros((n_pad, 7))
        for im in plan.datum:
            r0 = plan.img_row[im]
            self.E[r0:r0 + 6] = self.G[r0:r0 + 6]
        self.pad = plan.row_ext < 0
        a = np.zeros((n_pad + 64, n_pad + 64))
        a[np.ix_(ro, ro)] = S_r
        a[n_pad, ro] = g_r
        self.a = a
        self.top = plan.top_tile0 if world > 1 else plan.NT
        self.t0 = 64 * plan.tile_b0[self.top] if self.top < plan.NT else n_pad

    def diag(self):
        """exchange 1 sends this: the diagonal of the rank's partial S"""
        return np.diag(self.a)[:self.plan.n_pad].copy()

    def eliminate(self, dg):
        """border + scaling with the summed diagonal, elimination of the own subtrees; exchange 2 sends the result:
        the rank's contribution to the shared top part (lower trapezoid incl. the augmented rows)"""
        a, E, G, mine, pad, n_pad = self.a, self.E, self.G, self.mine, self.pad, self.plan.n_pad
        self.d = d = 1.0 / np.sqrt(np.where(pad, 1.0, dg + np.sum(E * E, axis=1)))
        a[:n_pad, :n_pad] += (E * mine[:, None]) @ E.T                    # datum term where the ROW is initialised here
        a[np.nonzero(pad & mine)[0], np.nonzero(pad & mine)[0]] = 1.0
        a[:n_pad, :n_pad] *= np.outer(d, d)
        a[n_pad, :n_pad] *= d
        a[n_pad + 1:n_pad + 8, :n_pad] = (G * d[:, None] * mine[:, None]).T
        a[n_pad + 8:n_pad + 15, :n_pad] = (E * d[:, None] * mine[:, None]).T
        a[:] = np.tril(a)
        masked_tiles_cholesky(a, self.plan, self.plan.nz, range(0, self.top), self.rank if self.world > 1 else None)
        return a[self.t0:, self.t0:].copy()

    def finish(self, top_sum):
        """factorisation of the summed top (replicated), 14x14 border, backward substitution of the top and the own
        subtrees; exchange 3 sends the result: the rows of the solution this rank is responsible for"""
        a, plan, n_pad = self.a, self.plan, self.plan.n_pad
        if self.world > 1:
            a[self.t0:, self.t0:] = top_sum
            masked_tiles_cholesky(a, plan, plan.nz, range(self.top, plan.NT), self.rank)
        Tm = a[n_pad:n_pad + 15, n_pad:n_pad + 15]
        Tm = np.tril(Tm) + np.tril(Tm, -1).T
        coef = np.zeros(14)
        assert self.host.feba_host_sparse_border(np.ascontiguousarray(Tm).ctypes.data_as(_pd), coef.ctypes.data_as(_pd)) == 0
        y = a[n_pad, :n_pad] + coef @ a[n_pad + 1:n_pad + 15, :n_pad]
        x = np.zeros(n_pad)
        for k in range(n_pad // 64 - 1, -1, -1):                          # k_backstep, owner filter
            if self.world > 1 and self.blk_owner[k] >= 0 and self.blk_owner[k] != self.rank:
                continue
            ks = slice(64 * k, 64 * k + 64)
            x[ks] = sla.solve_triangular(np.tril(a[ks, ks]), y[ks], lower=True, trans="T")
            y[:64 * k] -= a[ks, :64 * k].T @ x[ks]
        x[~self.mine] = 0.0
        return x

    def delta(self, sol_sum):
        return (-sol_sum * self.d)[self.ro]


def solve_like_the_device(host, plan, parts, Gc, world):
    """parts: per rank (S_r, g_r) in xhat order.  Returns delta_c (xhat order, scaled system undone).  All ranks in
    one process, the exchanges are plain sums (tests/test_group_gloo.py runs the same phases in separate processes
    with torch.distributed all_reduce in between)."""
    ranks = [RankSolve(host, plan, parts[r][0], parts[r][1], Gc, r, world) for r in range(world)]
    dg = sum(rk.diag() for rk in ranks)                                   # exchange 1
    tops = [rk.eliminate(dg) for rk in ranks]
    top_sum = sum(tops) if world > 1 else None                            # exchange 2
    sols = [rk.finish(top_sum) for rk in ranks]
    sol = sum(sols)                                                       # exchange 3: every row exactly once
    return ranks[0].delta(sol), [rk.a for rk in ranks]


def truth_step(S, g, Gc):
    n = len(g)
    K = np.zeros((n + 7, n + 7), dtype=np.longdouble)
    K[:n, :n], K[:n, n:], K[n:, :n] = S, Gc, Gc.T
    return np.asarray(exact.refined_solve(K, np.concatenate([-g, np.zeros(7)]).astype(np.longdouble))[:n], dtype=np.float64)


def check_structure(plan, prob, world):
    ui = plan.ui
    assert plan.n_pad % 64 == 0
    rows = np.concatenate([plan.img_row[:, None] + np.arange(ui)[None, :]]).reshape(-1)
    assert len(np.unique(rows)) == ui * plan.n_img                        # one row per unknown
    assert np.all(plan.row_ext[rows] == np.arange(ui * plan.n_img))
    cam = plan.off_cam + np.arange(plan.cam_rows)
    assert np.all(plan.row_ext[cam] == ui * plan.n_img + np.arange(plan.cam_rows))
    # images of one tile-node occupy that node's tiles only
    tile_of_row = np.repeat(np.arange(plan.NT), 64 * np.diff(plan.tile_b0))
    assert np.all(plan.tile_node[tile_of_row[plan.img_row]] == plan.img_node)
    assert np.all(plan.img_node[plan.datum] == 0)                         # datum images: root node
    # separator property: adjacent images lie on one root-to-leaf path
    def ancestors(n):
        out = []
        while n >= 0:
            out.append(n)
            n = plan.node_parent[n]
        return out
    anc = [set(ancestors(n)) for n in range(plan.n_nodes)]
    for a in range(plan.n_img):
        na = plan.img_node[a]
        for b in plan.adj_idx[plan.adj_ptr[a]:plan.adj_ptr[a + 1]]:
            nb_ = plan.img_node[b]
            assert na in anc[nb_] or nb_ in anc[na], (a, b)
    if world > 1:
        assert plan.world == world
        owners = set(int(o) for o in plan.node_owner if o >= 0)
        assert owners == set(range(world))
        # shared tiles are the tail of the order, owned tiles of different ranks are never coupled
        assert np.all(plan.tile_owner[plan.top_tile0:] < 0) and np.all(plan.tile_owner[:plan.top_tile0] >= 0)
        for i in range(plan.NT):
            for j in range(i):
                if plan.nz[i, j] and plan.tile_owner[i] >= 0 and plan.tile_owner[j] >= 0:
                    assert plan.tile_owner[i] == plan.tile_owner[j], (i, j)


@pytest.mark.parametrize("n_img,n_pts,leaf,tile_max", [(150, 15000, 12, 2), (150, 15000, 30, 3), (320, 20000, 40, 2)])
def test_masked_plan_solves_the_bordered_system(host, n_img, n_pts, leaf, tile_max):
    prob = fb.synth.make_network(n_img, n_pts, 8, 977 + n_img, mode="free")
    err, x0, _ = fb.Buildxhat(prob)
    plan = Plan(host, prob, world=1, leaf=leaf, tile_max=tile_max)
    try:
        check_structure(plan, prob, 1)
        assert plan.n_nodes >= 3 and plan.chain_blocks < plan.n_pad // 64  # the block was actually cut
        nbk = sparse.normal_blocks(prob, x0)
        S, g, _ = sparse.reduce(prob, nbk)
        Gc = nbk["Gc"]
        delta, A = solve_like_the_device(host, plan, [(S, g)], Gc, 1)
        # structurally zero tiles were never touched
        Lm = A[0]
        for i in range(plan.NT):
            for j in range(i):
                if not plan.nz[i, j]:
                    assert not np.any(Lm[plan.tile_rows(i), plan.tile_rows(j)]), (i, j)
        # envelope of the backward substitution
        for k in range(plan.n_pad // 64):
            assert 0 <= plan.first_block[k] <= k
            assert not np.any(Lm[64 * k:64 * k + 64, :64 * plan.first_block[k]]), k
        truth = truth_step(S, g, Gc)
        rel = np.linalg.norm(delta - truth) / np.linalg.norm(truth)
        dense = sparse.solve_reduced(prob, S, g, Gc)
        rel_dense = np.linalg.norm(dense - truth) / np.linalg.norm(truth)
        assert rel <= max(2.0 * rel_dense, 1e-9) and rel < 1e-7, (rel, rel_dense)
        assert np.max(np.abs(Gc.T @ delta)) <= 1e-7 * np.max(np.abs(Gc)) * np.max(np.abs(delta))
    finally:
        plan.close()


@pytest.mark.parametrize("world", [2, 4])
def test_group_algorithm_equals_one_rank(host, world):
    prob = fb.synth.make_network(200, 16000, 8, 4242, mode="free")
    err, x0, _ = fb.Buildxhat(prob)
    plan = Plan(host, prob, world=world, leaf=16, tile_max=2)
    try:
        check_structure(plan, prob, world)
        own_seg = plan.point_owner()
        assert set(np.unique(own_seg)) == set(range(world))
        pt_owner = -np.ones(prob.numPts, dtype=np.int64)
        pt_owner[plan.seg_pt] = own_seg
        # balance: no rank holds more than twice its share of the observations
        cnt = np.bincount(pt_owner[prob.obs_pt], minlength=world)
        assert cnt.max() <= 2.0 * prob.n_obs / world, cnt
        # a rank's points see images of its own subtrees and of shared nodes only
        for r in range(world):
            imgs = np.unique(prob.obs_img[pt_owner[prob.obs_pt] == r])
            own = plan.node_owner[plan.img_node[imgs]]
            assert np.all((own == r) | (own < 0))
        nbk = sparse.normal_blocks(prob, x0)
        S, g, _ = sparse.reduce(prob, nbk)
        Gc = nbk["Gc"]
        parts = [reduced_of(prob, x0, pt_owner == r) for r in range(world)]
        assert np.max(np.abs(sum(p[0] for p in parts) - S)) <= 1e-9 * np.max(np.abs(S))
        delta_group, _ = solve_like_the_device(host, plan, parts, Gc, world)
        truth = truth_step(S, g, Gc)
        rel = np.linalg.norm(delta_group - truth) / np.linalg.norm(truth)
        assert rel < 1e-8, rel
        # and the same plan driven by one rank holding everything
        plan1 = copy.copy(plan)
        delta_one, _ = solve_like_the_device(host, plan1, [(S, g)], Gc, 1)
        assert np.linalg.norm(delta_group - delta_one) / np.linalg.norm(delta_one) < 1e-9
    finally:
        plan.close()


def test_dense_block_is_not_worth_cutting(host):
    # 25 rays per point on 60 images: almost every image pair shares points -- separators are most of the block
    prob = fb.synth.make_network(60, 6000, 25, 3, mode="free")
    plan = Plan(host, prob, world=1, leaf=8, tile_max=2)
    try:
        check_structure(plan, prob, 1)
        nb_id = (prob.u_c + 63) // 64
        dense = (64.0 * nb_id) ** 3 / 3
        pays = plan.flop <= 0.5 * dense and 2 * plan.chain_blocks <= nb_id
        assert not pays                                                   # the library keeps the identity plan
    finally:
        plan.close()
