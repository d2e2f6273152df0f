"""Block-sparse form of the reduced camera system (csrc/feba_sparse.h, opt-in FEBA_SPARSE=1), on the CPU.

The host-compiled helpers the library uses -- supertile pattern with symbolic fill, datum images, the 14x14
border of the sparse-datum form -- drive a numpy emulation of the masked supertile factorisation
(chol_dag with the pattern: TRSM / UPDATE tasks on structurally zero supertiles are skipped), and the step
that comes out is compared with the bordered system of main.m:428-437 solved in extended precision.
Checks: (1) the pattern is complete (the masked factor equals the unmasked one), (2) the algebra of
sparse_border_solve / k_combine gives the bordered solution, (3) the pattern is actually sparse.
The device kernels themselves are covered by the GPU suite (FEBA_SPARSE=1 cases of tests/test_gpu_parity.py).
"""
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
HDRS = [os.path.join(ROOT, "fish-eye_bundle_adjustment_b200", "csrc", f) for f in ("feba_model.cuh", "feba_sparse.h")]
LIB = os.path.join(ROOT, "tests", "_build", "libfeba_model_host.so")
_pd, _pi = C.POINTER(C.c_double), C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def host():
    newest = max(os.path.getmtime(p) for p in [SRC] + HDRS)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall",
                        "-Wno-unknown-pragmas", SRC, "-o", LIB], check=True)
    lib = C.CDLL(LIB)
    lib.feba_host_sparse_border.argtypes = [_pd, _pd]
    lib.feba_host_sparse_datum.argtypes = [C.c_int] * 4 + [_pi]
    lib.feba_host_sparse_pattern.argtypes = [C.c_int] * 7 + [_pi, C.c_int, _pi, C.POINTER(C.c_ubyte)]
    lib.feba_host_sparse_row_first.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_ubyte), _pi]
    return lib


def image_pairs(prob):
    """(a, b <= a) image pairs that share a tie point: the blocks of the pair schedule."""
    tie = prob.pt_tie[prob.obs_pt] >= 0
    A = sp.csr_matrix((np.ones(int(tie.sum()), dtype=np.float32), (prob.obs_pt[tie], prob.obs_img[tie])),
                      shape=(prob.numPts, prob.numImg))
    B = sp.tril((A.T @ A).tocsr(), k=-1).tocoo()
    return np.ascontiguousarray(np.stack([B.row, B.col], axis=1).astype(np.int32))


def pattern(lib, prob, n_red, nb, T):
    ui = prob.settings.u_perimage
    off_cam = ui * prob.numImg
    datum = np.zeros(8, dtype=np.int32)
    nd = lib.feba_host_sparse_datum(prob.numImg, ui, nb, T, datum.ctypes.data_as(_pi))
    datum = datum[:nd]
    blocks = image_pairs(prob)
    NT = (nb + T - 1) // T
    nz = np.zeros((NT + 1, NT + 1), dtype=np.uint8)
    got = lib.feba_host_sparse_pattern(nb, T, ui, prob.numImg, off_cam, n_red, len(blocks),
                                       blocks.ctypes.data_as(_pi), nd, datum.ctypes.data_as(_pi),
                                       nz.ctypes.data_as(C.POINTER(C.c_ubyte)))
    assert got == NT
    return nz.astype(bool), datum


def masked_supertile_cholesky(A, nb, T, nz):
    """chol_dag on the augmented matrix A ((nb+1) 64-blocks per side, lower triangle), skipping tasks on
    structurally zero supertiles.  nz=None: every task.  Returns the factor (in place copy)."""
    A = A.copy()
    NT = (nb + T - 1) // T
    b0 = lambda t: (nb if t == NT else t * T) * 64
    b1 = lambda t: (nb + 1 if t == NT else min(nb, (t + 1) * T)) * 64
    on = (lambda i, j: True) if nz is None else (lambda i, j: bool(nz[i, j]))
    for k in range(NT):
        ks = slice(b0(k), b1(k))
        A[ks, ks] = np.linalg.cholesky(np.tril(A[ks, ks]) + np.tril(A[ks, ks], -1).T)
        Lk = A[ks, ks]
        rows = [i for i in range(k + 1, NT + 1) if on(i, k)]
        for i in rows:
            sl = slice(b0(i), b1(i))
            A[sl, ks] = sla.solve_triangular(Lk, A[sl, ks].T, lower=True).T
        for i in rows:
            for j in rows:
                if j > i or (j == NT and i != NT):
                    continue
                si, sj = slice(b0(i), b1(i)), slice(b0(j), b1(j))
                A[si, sj] -= A[si, ks] @ A[sj, ks].T
    return A


@pytest.mark.parametrize("n_img,n_pts,T", [(60, 5000, 1), (150, 15000, 2), (150, 15000, 3)])
def test_sparse_datum_form_matches_bordered_solution(host, n_img, n_pts, T):
    prob = fb.synth.make_network(n_img, n_pts, 8, 977 + n_img, mode="free")
    err, x0, _ = fb.Buildxhat(prob)
    nbk = sparse.normal_blocks(prob, x0)
    S, g, _ = sparse.reduce(prob, nbk)
    Gc = nbk["Gc"]
    n = S.shape[0]
    nb = (n + 63) // 64
    n_pad = 64 * nb
    nz, datum = pattern(host, prob, n, nb, T)
    NT = nz.shape[0] - 1
    ui = prob.settings.u_perimage
    # --- what the device does before the factorisation (k_datum_split, k_diag_scale, k_border_scale)
    E = np.zeros_like(Gc)
    for im in datum:
        E[ui * im:ui * im + ui] = Gc[ui * im:ui * im + ui]
    Ms = S + E @ E.T
    d = 1.0 / np.sqrt(np.diag(Ms))
    A = np.zeros((n_pad + 64, n_pad + 64))
    A[:n, :n] = Ms * np.outer(d, d)
    A[np.arange(n, n_pad), np.arange(n, n_pad)] = 1.0
    Baug = np.column_stack([g, Gc, E]) * d[:, None]            # 15 columns -> augmented rows 0..14
    A[n_pad:n_pad + 15, :n] = Baug.T
    # --- masked factorisation == unmasked factorisation (the pattern with its fill is complete)
    Lm = masked_supertile_cholesky(A, nb, T, nz)
    Lf = masked_supertile_cholesky(A, nb, T, None)
    low = np.tril(np.ones_like(A, dtype=bool))
    assert np.max(np.abs((Lm - Lf)[low])) <= 1e-12 * np.max(np.abs(Lf[low]))
    # structurally zero supertiles were never touched
    for i in range(NT):
        for j in range(i):
            if not nz[i, j]:
                blk = Lm[i * T * 64:min(nb, (i + 1) * T) * 64, j * T * 64:(j + 1) * T * 64]
                assert not np.any(blk), (i, j)
    # the envelope the backward substitution starts from (k_backstep's c_begin): nothing left of it
    first = np.zeros(nb, dtype=np.int32)
    nz8 = np.ascontiguousarray(nz.astype(np.uint8))
    host.feba_host_sparse_row_first(nb, T, nz8.ctypes.data_as(C.POINTER(C.c_ubyte)), first.ctypes.data_as(_pi))
    for k in range(nb):
        assert 0 <= first[k] <= k
        assert not np.any(Lm[64 * k:64 * k + 64, :64 * first[k]]), k
    if n_img >= 150 and T == 2:
        assert first.max() > 0                                  # the envelope actually cuts something
    # --- border (k_border_solve in its sparse form) and combination (k_combine), backward substitution
    Tm = Lm[n_pad:n_pad + 15, n_pad:n_pad + 15]
    Tm = np.tril(Tm) + np.tril(Tm, -1).T
    coef = np.zeros(14)
    assert host.feba_host_sparse_border(np.ascontiguousarray(Tm).ctypes.data_as(_pd), coef.ctypes.data_as(_pd)) == 0
    Y = Lm[n_pad:n_pad + 15, :n_pad]                            # rows: (L^-1 B)'
    y = Y[0] + coef @ Y[1:15]
    L = np.tril(Lm[:n_pad, :n_pad])
    delta = -(sla.solve_triangular(L, y, lower=True, trans="T"))[:n] * d
    # --- the bordered system of main.m:428-437 in extended precision
    K = np.zeros((n + 7, n + 7), dtype=np.longdouble)
    K[:n, :n], K[:n, n:], K[n:, :n] = S, Gc, Gc.T
    truth = np.asarray(exact.refined_solve(K, np.concatenate([-g, np.zeros(7)]).astype(np.longdouble))[:n],
                       dtype=np.float64)
    rel = np.linalg.norm(delta - truth) / np.linalg.norm(truth)
    dense = sparse.solve_reduced(prob, S, g, Gc)
    rel_dense = np.linalg.norm(dense - truth) / np.linalg.norm(truth)
    # at least as accurate as the oracle's dense form M = S + G G' (and never worse than 1e-7)
    assert rel <= max(2.0 * rel_dense, 1e-9) and rel < 1e-7, (rel, rel_dense)
    # the constraint G' delta = 0 holds
    assert np.max(np.abs(Gc.T @ delta)) <= 1e-7 * np.max(np.abs(Gc)) * np.max(np.abs(delta))


def test_pattern_of_a_banded_block_is_sparse(host):
    # 24 x 24 images, 10 rays per point: the shape of BASELINE configs[3] at a quarter of its side
    prob = fb.synth.make_network(576, 20000, 10, 5, mode="free")
    n = prob.u_c
    nb = (n + 63) // 64
    for T, frac_max in ((2, 0.55), (4, 0.75)):
        nz, datum = pattern(host, prob, n, nb, T)
        NT = nz.shape[0] - 1
        lower = np.tril(np.ones((NT, NT), dtype=bool))
        frac = nz[:NT, :NT][lower].sum() / lower.sum()
        assert frac <= frac_max, (T, frac)
        assert nz[NT].all() and nz[NT - 1, :NT].all()          # augmented and camera rows are dense
        assert len(datum) == 8 and datum[0] == 0 and datum[-1] == prob.numImg - 1
        assert np.all(np.diff(datum) > 0)


def test_sparse_border_reports_a_singular_system(host):
    coef = np.zeros(14)
    T = np.zeros((15, 15))
    T[8:, 8:] = -np.eye(7)                                      # -T_EE - I = 0 and T_GG = 0: singular
    assert host.feba_host_sparse_border(np.ascontiguousarray(T).ctypes.data_as(_pd), coef.ctypes.data_as(_pd)) == 1
