"""Chunk schedule of the assembly (csrc/feba_chunks.h) on the CPU.

The header is compiled with g++; a numpy emulation of what the device does with the schedule -- per chunk the
image partials and the image-pair partials from per-observation records, then the fixed-order final sums -- must
reproduce the image part of the oracle's reduced system S = N_cc - W V^-1 W' (main.m:424-425 with the points
eliminated).  Also checked: every observation sits in exactly one image slot of its chunk, every ordered
observation pair of a tie point appears exactly once with row(a) >= row(b), a point with more observations than
the chunk cap gets a chunk of its own.  The kernels themselves are covered by the GPU suite (the reduced-system
block tests run through this path)."""
import ctypes as C
import os

import numpy as np
import pytest

import feba_b200 as fb
from oracle import sparse
from tests.test_reduced_plan_host import host as _host_fixture, sorted_by_point  # noqa: F401  (fixture re-use)

host = _host_fixture
_pi, _pb = C.POINTER(C.c_int), C.POINTER(C.c_ubyte)


class Chunks:
    def __init__(self, lib, prob, img_row):
        lib.feba_host_chunks.restype = C.c_void_p
        lib.feba_host_chunks.argtypes = [C.c_int, C.c_int, _pi, _pi, _pb, _pi]
        lib.feba_host_chunks_free.argtypes = [C.c_void_p]
        lib.feba_host_chunks_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        lib.feba_host_chunks_arrays.argtypes = [C.c_void_p] + [_pi] * 9 + [C.POINTER(C.c_uint)] + [_pi] * 6
        self.lib = lib
        self.seg_start, self.simg, self.seg_pt, self.order = sorted_by_point(prob)
        seg_tie = np.ascontiguousarray((prob.pt_tie[self.seg_pt] >= 0).astype(np.uint8))
        row = np.ascontiguousarray(img_row, dtype=np.int32)
        self.p = lib.feba_host_chunks(prob.numImg, len(self.seg_pt), self.seg_start.ctypes.data_as(_pi),
                                      self.simg.ctypes.data_as(_pi), seg_tie.ctypes.data_as(_pb), row.ctypes.data_as(_pi))
        sz = (C.c_longlong * 6)()
        lib.feba_host_chunks_sizes(self.p, sz)
        self.n_chunks, n_is, n_bs, n_pairs, n_tb, ok = [int(v) for v in sz]
        assert ok
        i32 = lambda n: np.zeros(max(n, 1), dtype=np.int32)
        self.obs0, self.img0, self.slot_img, self.slot_obs0 = i32(self.n_chunks + 1), i32(self.n_chunks + 1), i32(n_is), i32(n_is + 1)
        self.slot_obs = i32(prob.n_obs)
        self.blk0, self.bslot_a, self.bslot_b, self.bslot_pair0 = i32(self.n_chunks + 1), i32(n_bs), i32(n_bs), i32(n_bs + 1)
        self.pairs = np.zeros(max(n_pairs, 1), dtype=np.uint32)
        self.timg_ptr, self.timg_slots = i32(prob.numImg + 1), i32(n_is)
        self.tblk_a, self.tblk_b, self.tblk_ptr, self.tblk_slots = i32(n_tb), i32(n_tb), i32(n_tb + 1), i32(n_bs)
        a = lambda v: v.ctypes.data_as(_pi)
        lib.feba_host_chunks_arrays(self.p, a(self.obs0), a(self.img0), a(self.slot_img), a(self.slot_obs0), a(self.slot_obs),
                                    a(self.blk0), a(self.bslot_a), a(self.bslot_b), a(self.bslot_pair0),
                                    self.pairs.ctypes.data_as(C.POINTER(C.c_uint)), a(self.timg_ptr), a(self.timg_slots),
                                    a(self.tblk_a), a(self.tblk_b), a(self.tblk_ptr), a(self.tblk_slots))
        self.n_is, self.n_bs, self.n_pairs, self.n_tb = n_is, n_bs, n_pairs, n_tb

    def close(self):
        self.lib.feba_host_chunks_free(self.p)


@pytest.mark.parametrize("permute_rows", [False, True])
def test_chunk_partials_sum_to_the_reduced_system(host, permute_rows):
    prob = fb.synth.make_network(40, 3000, 7, 99, mode="mixed", n_control=40)
    err, x0, _ = fb.Buildxhat(prob)
    ui = prob.settings.u_perimage
    rng = np.random.default_rng(3)
    img_row = ui * (rng.permutation(prob.numImg) if permute_rows else np.arange(prob.numImg))
    ch = Chunks(host, prob, img_row)
    try:
        assert ch.n_chunks >= prob.numPts // 64
        nbk = sparse.normal_blocks(prob, x0)
        S, g, _ = sparse.reduce(prob, nbk)
        o = ch.order                                            # sorted position -> PHO row
        Je = nbk["Je"][o]                                       # (n, 2, 6)
        Jt = nbk["Jt"][o]
        pw = nbk["pw"]
        tie = prob.pt_tie[prob.obs_pt][o]
        Vinv = np.linalg.inv(nbk["V"])
        # per-observation "records": We = Je' P Jt (6x3); pair term  We_a V^-1 We_b'
        We = np.einsum("nri,nrk->nik", Je * pw[None, :, None], Jt)
        Vi = np.where((tie >= 0)[:, None, None], Vinv[np.maximum(tie, 0)], 0.0)
        # every observation in exactly one image slot of its chunk
        seen = np.zeros(prob.n_obs, dtype=int)
        for c in range(ch.n_chunks):
            for s in range(ch.img0[c], ch.img0[c + 1]):
                loc = ch.slot_obs[ch.slot_obs0[s]:ch.slot_obs0[s + 1]]
                glob = ch.obs0[c] + loc
                assert np.all(ch.simg[glob] == ch.slot_img[s]) and np.all(np.diff(loc) > 0)
                seen[glob] += 1
        assert np.all(seen == 1)
        # image partials -> diagonal blocks; block partials -> off-diagonal (and same-image) Schur terms
        img_part = np.zeros((ch.n_is, 6, 6))
        for c in range(ch.n_chunks):
            for s in range(ch.img0[c], ch.img0[c + 1]):
                t = ch.obs0[c] + ch.slot_obs[ch.slot_obs0[s]:ch.slot_obs0[s + 1]]
                img_part[s] = (np.einsum("nri,nrj->ij", Je[t] * pw[None, :, None], Je[t])
                               - np.einsum("nik,nkl,njl->ij", We[t], Vi[t], We[t]))
        blk_part = np.zeros((ch.n_bs, 6, 6))
        n_seen_pairs = 0
        for c in range(ch.n_chunks):
            for b in range(ch.blk0[c], ch.blk0[c + 1]):
                pr = ch.pairs[ch.bslot_pair0[b]:ch.bslot_pair0[b + 1]]
                ta, tb = ch.obs0[c] + (pr & 0xffff).astype(np.int64), ch.obs0[c] + (pr >> 16).astype(np.int64)
                assert np.all(ch.simg[ta] == ch.bslot_a[b]) and np.all(ch.simg[tb] == ch.bslot_b[b])
                assert img_row[ch.bslot_a[b]] >= img_row[ch.bslot_b[b]]
                assert np.all(tie[ta] == tie[tb]) and np.all(tie[ta] >= 0)
                blk_part[b] = np.einsum("nik,nkl,njl->ij", We[ta], Vi[ta], We[tb])
                n_seen_pairs += len(pr)
        m = np.bincount(prob.obs_pt[prob.pt_tie[prob.obs_pt] >= 0])
        # ordered pairs (o, b), b in an earlier image: m (m - 1) / 2 per tie point (no point is seen twice by one image)
        assert n_seen_pairs == ch.n_pairs == int(np.sum(m * (m - 1) // 2))
        n = ui * prob.numImg
        S_chunks = np.zeros((n, n))
        for im in range(prob.numImg):
            r0 = ui * im
            for q in range(ch.timg_ptr[im], ch.timg_ptr[im + 1]):
                assert ch.slot_img[ch.timg_slots[q]] == im
                S_chunks[r0:r0 + 6, r0:r0 + 6] += img_part[ch.timg_slots[q]]
        for t in range(ch.n_tb):
            ia, ib = ch.tblk_a[t], ch.tblk_b[t]
            for q in range(ch.tblk_ptr[t], ch.tblk_ptr[t + 1]):
                s = ch.tblk_slots[q]
                assert ch.bslot_a[s] == ia and ch.bslot_b[s] == ib
                S_chunks[ui * ia:ui * ia + 6, ui * ib:ui * ib + 6] -= blk_part[s]
                if ia != ib:
                    S_chunks[ui * ib:ui * ib + 6, ui * ia:ui * ia + 6] -= blk_part[s].T
        assert np.max(np.abs(S_chunks - S[:n, :n])) <= 1e-11 * np.max(np.abs(S[:n, :n]))
    finally:
        ch.close()


def test_a_point_with_more_observations_than_the_cap_gets_its_own_chunk(host):
    prob = fb.synth.make_network(720, 900, 8, 5, mode="mixed", n_control=20)
    # one control point observed by every image: 720 observations > kChunkObs = 416
    p = int(np.nonzero(prob.pt_tie < 0)[0][0])
    extra = np.arange(prob.numImg, dtype=np.int32)
    prob.obs_img = np.concatenate([prob.obs_img, extra])
    prob.obs_pt = np.concatenate([prob.obs_pt, np.full(prob.numImg, p, dtype=np.int32)])
    prob.obs_x = np.concatenate([prob.obs_x, np.full(prob.numImg, 1200.0)])
    prob.obs_y = np.concatenate([prob.obs_y, np.full(prob.numImg, 1000.0)])
    ch = Chunks(host, prob, 6 * np.arange(prob.numImg))
    try:
        sizes = np.diff(ch.obs0[:ch.n_chunks + 1])
        assert sizes.max() >= 720 and sizes.sum() == prob.n_obs
        big = int(np.argmax(sizes))
        assert ch.img0[big + 1] - ch.img0[big] == prob.numImg           # every image has a slot in that chunk
        assert np.all(sizes[np.arange(ch.n_chunks) != big] <= 416)
    finally:
        ch.close()
